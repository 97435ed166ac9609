#!/bin/bash
# `ncu --set full` captures of steady-state launches of the kernel classes that have no single-kernel capture yet; only the
# per-kernel metric rows travel back (the reports are tens of MB each)
cd $GRAFT_REPO_ROOT
export NUNET_GRAPH_TRAIN=0
mkdir -p /tmp/cls
run() {  # name regex skip count
  timeout 400 ncu --set full --clock-control none -k regex:$2 --launch-skip $3 --launch-count $4 -o /tmp/cls/$1 -f python bench.py --steps 1 --warmup 2 --kernels-only > gpurun_out/cls_$1.log 2>&1
  echo "$1 rc=$?"
}
run wgrad conv3x3_wgrad_tma 60 2
run activate activate_ 60 3
run bnbwd bn_bwd_apply 60 1
run router grad_router 30 3
python tools/ncu_kernels.py gpurun_out/r02_ncu_kernels_step.csv wgrad=/tmp/cls/wgrad.ncu-rep activate=/tmp/cls/activate.ncu-rep bn_bwd_apply=/tmp/cls/bnbwd.ncu-rep router=/tmp/cls/router.ncu-rep
cut -c1-300 gpurun_out/r02_ncu_kernels_step.csv
