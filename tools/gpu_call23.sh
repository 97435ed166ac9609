#!/bin/bash
cd $GRAFT_REPO_ROOT
NUNET_NVCC_EXTRA="-DNUNET_EXPERIMENT" python pytorch_nested-unet_b200/build.py --force > /dev/null 2>&1
for nt in 0 64 128; do TAG="nt=$nt" NUNET_CONV_NT=$nt python tools/ab_epilogue.py 2>&1 | tail -1 | tr '|' '\n' | head -9; done > gpurun_out/c23_ab.log
cat gpurun_out/c23_ab.log
