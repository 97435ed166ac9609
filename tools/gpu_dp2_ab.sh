#!/bin/bash
# 2 GPUs: where does an occasional slow end-to-end run lose its time?  (host trace + garbage-collector log; gc.freeze A/B)
cd $GRAFT_REPO_ROOT
export CUDA_DEVICE_MAX_CONNECTIONS=32
export NUNET_E2E_TRACE=1
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline $2 2> gpurun_out/dp2ab_$1.err > gpurun_out/dp2ab_$1.json
  python -c "import json,sys; d=json.load(open('gpurun_out/dp2ab_$1.json')); print('$3 value %.0f ms %.3f e2e %.0f (%.3f ms) ratio %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['value']/d['value']))"
  grep "rank 0 e2e trace" gpurun_out/dp2ab_$1.err | cut -c1-700
  grep "rank 0 gc" gpurun_out/dp2ab_$1.err | head -8
}
run 29521 "" A
NUNET_GC=freeze run 29522 "" A-freeze
run 29523 "" A
NUNET_GC=freeze run 29524 "" A-freeze
