#!/bin/bash
# 2 GPUs: the default bench line, driver-style, three times; slowest host iterations of the end-to-end loop from stderr
cd $GRAFT_REPO_ROOT
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 $2 2> gpurun_out/dp2ab_$1.err > gpurun_out/dp2ab_$1.json
  python -c "import json,sys; d=json.load(open('gpurun_out/dp2ab_$1.json')); print('$3 value %.0f ms %.3f e2e %.0f (%.3f ms) ratio %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['value']/d['value']))"
  grep "e2e iterations\|host trace\|gc gen" gpurun_out/dp2ab_$1.err | cut -c1-900
}
run 29521 "" A
run 29522 "" A
run 29523 "" A
