#!/bin/bash
cd $GRAFT_REPO_ROOT
( time python -m pytest tests/test_model_gpu.py -m gpu -q -s 2>&1 | grep -E "fp32 check|relmax|passed|failed|Error|assert " | cut -c1-700 ) > gpurun_out/c3_tests.log 2>&1
python tools/ncu_kernels.py > gpurun_out/c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv3x3|grad_router|up_transpose|upsample2x" -o gpurun_out/c3_prof -f python tools/ncu_kernels.py > gpurun_out/c3_ncu.log 2>&1
cat gpurun_out/c3_tests.log; tail -3 gpurun_out/c3_plain.log; tail -5 gpurun_out/c3_ncu.log; ls -la gpurun_out/*.ncu-rep
