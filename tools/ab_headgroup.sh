#!/bin/bash
# A/B inside one call: grouped deep-supervision loss on / off
cd $GRAFT_REPO_ROOT
for rep in 1 2; do
  for v in "" 1; do
    NUNET_NO_HEAD_GROUP=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/hg_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NO_GROUP=$v value %.0f ms %.3f e2e %.0f ratio %.3f clocks %s launches %d loss_fwd %.3f loss_bwd %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value']/d['value'], d['clocks']['sm_mhz'], d['gpu_launches'], d['roofline']['class_ms']['loss_fwd'], d['roofline']['class_ms']['loss_bwd']))"
    tail -2 gpurun_out/hg_bench.err
  done
done
