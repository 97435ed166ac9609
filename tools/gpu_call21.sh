#!/bin/bash
cd $GRAFT_REPO_ROOT
export N=256 H=96 CIN=96 COUT=32 KINDS=stats STACKED=1
timeout 200 python tools/one_conv.py > gpurun_out/c21_one.log 2>&1 || { cat gpurun_out/c21_one.log; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3x3_c32 -c 1 -o gpurun_out/c21_c32 -f python tools/one_conv.py > gpurun_out/c21_ncu.log 2>&1
tail -3 gpurun_out/c21_one.log
