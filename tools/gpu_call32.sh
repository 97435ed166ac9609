#!/bin/bash
cd $GRAFT_REPO_ROOT
( for c in "96 32 32" "48 64 64" "24 128 128" "96 192 32" "48 320 64"; do set -- $c; H=$1 CIN=$2 COUT=$3 python tools/one_wgrad.py 2>&1 | tail -2; done ) > gpurun_out/c32_wgrad.log
cat gpurun_out/c32_wgrad.log
H=96 CIN=32 COUT=32 REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad_tma -s 2 -c 1 -o /tmp/wg32 -f python tools/one_wgrad.py > gpurun_out/c32_ncu.log 2>&1
ncu -i /tmp/wg32.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/c32_wg32_src.csv 2>/dev/null
python tools/ncu_kernels.py gpurun_out/c32_wg32_row.csv wgrad_32to32=/tmp/wg32.ncu-rep
H=48 CIN=64 COUT=64 REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad_tma -s 2 -c 1 -o /tmp/wg64 -f python tools/one_wgrad.py > gpurun_out/c32_ncu2.log 2>&1
ncu -i /tmp/wg64.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/c32_wg64_src.csv 2>/dev/null
python tools/ncu_kernels.py gpurun_out/c32_wg64_row.csv wgrad_64to64=/tmp/wg64.ncu-rep
ls -la gpurun_out/
