#!/bin/bash
cd $GRAFT_REPO_ROOT
export N=256 H=96 CIN=32 COUT=32 KINDS=stats
timeout 200 python tools/one_conv.py > gpurun_out/c13_one.log 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3x3_igemm -c 1 -o gpurun_out/c13_stats -f python tools/one_conv.py > gpurun_out/c13_ncu.log 2>&1
tail -3 gpurun_out/c13_one.log
