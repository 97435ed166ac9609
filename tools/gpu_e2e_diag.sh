#!/bin/bash
# 1 GPU: does the clock sampler (nvidia-smi -lms 200 beside the run) cause the rare host stalls inside loss.backward()?  600-step loops.
cd $GRAFT_REPO_ROOT
for v in "" 1; do
  NUNET_NO_CLOCKS=$v timeout 300 python bench.py --steps 600 --warmup 5 --no-cpu-baseline 2>gpurun_out/diag_$v.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('noclocks=%-2s value %.3f ms  e2e %.3f ms  ratio %.3f samples %s' % ('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['ms_per_step']/d['e2e']['ms_per_step'], d['clocks']['samples']))"
  grep "e2e iterations\|host trace\|gc gen" gpurun_out/diag_$v.err | cut -c1-1500
done
