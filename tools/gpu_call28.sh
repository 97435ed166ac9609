#!/bin/bash
cd $GRAFT_REPO_ROOT
( time timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | cut -c1-400 ) > gpurun_out/c31_tests.log 2>&1
cat gpurun_out/c31_tests.log
for a in "--size 256 --batch 64 --arch NestedUNet --no-ds" "--size 256 --batch 64 --arch UNet --no-ds" "--size 512 --batch 8 --arch NestedUNet"; do
  for ns in 0 1; do
    if [ $ns = 1 ]; then export NUNET_NO_STRIPS=1; else unset NUNET_NO_STRIPS; fi
    timeout 300 python bench.py $a --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/c31_bench.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$a no_strips=$ns', round(d['value'],1), round(d['ms_per_step'],2), d['roofline'].get('frac'))
except Exception as e: print('$a no_strips=$ns FAILED', e)"
    tail -2 gpurun_out/c31_bench.err | cut -c1-200
  done
done > gpurun_out/c31_cfg.log 2>&1
cat gpurun_out/c31_cfg.log
