#!/bin/bash
# BASELINE configs 3 and 4 on 8 GPUs (SyncBN, captured data-parallel step)
cd $GRAFT_REPO_ROOT
run() { # tag args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 8 --warmup 4 --no-cpu-baseline ${@:3} > gpurun_out/dp8_$2.json 2> gpurun_out/dp8_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/dp8_$2.json')); print('$2', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['roofline'].get('frac'))" || tail -3 gpurun_out/dp8_$2.err
}
run 29521 c4_nested512_k4 --size 512 --batch 8 --arch NestedUNet --classes 4
run 29522 c3_nested256 --size 256 --batch 64 --arch NestedUNet --no-ds
run 29523 c3_unet256 --size 256 --batch 64 --arch UNet --no-ds
