#!/bin/bash
# 1 GPU: host trace of the end-to-end loop (full and with the resident batch), 60 steps each
cd $GRAFT_REPO_ROOT
export NUNET_E2E_TRACE=1
for v in "" no_h2d "" no_h2d; do
  NUNET_E2E_VARIANT=$v timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/diag.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant=%-7s value %.3f ms  e2e %.3f ms  diff %.3f' % ('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step']-d['ms_per_step']))"
  grep "host trace\|gc gen" gpurun_out/diag.err | cut -c1-1200
done
