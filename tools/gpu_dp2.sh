#!/bin/bash
cd $GRAFT_REPO_ROOT
export CUDA_DEVICE_MAX_CONNECTIONS=32
( time timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -x -s 2>&1 | tail -30 | cut -c1-400 ) > gpurun_out/dp2_tests.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/dp2_bench.json 2> gpurun_out/dp2_bench.err
cat gpurun_out/dp2_tests.log; tail -5 gpurun_out/dp2_bench.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/dp2_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])"
