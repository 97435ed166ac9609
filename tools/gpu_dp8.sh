#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/dp8_bench.json 2> gpurun_out/dp8_bench.err
tail -4 gpurun_out/dp8_bench.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/dp8_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks'])"
