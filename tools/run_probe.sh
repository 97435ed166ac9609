#!/bin/bash
# Runs every umma_probe test in its own process; logs to gpurun_out/probe.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe.log 2>&1
for t in 0 1 2 3 4 5 6 7 8 9 10 11 12 13; do
  timeout 30 ./tools/umma_probe $t >> gpurun_out/probe.log 2>&1
  echo "test $t exit=$?" >> gpurun_out/probe.log
done
cat gpurun_out/probe.log
