#!/bin/bash
cd $GRAFT_REPO_ROOT
NUNET_NVCC_EXTRA="-DNUNET_EXPERIMENT" python pytorch_nested-unet_b200/build.py --force > /dev/null 2>&1
for d in 0 1 2 3 4 6; do TAG="dbg=$d" NUNET_C32_DBG=$d python tools/ab_epilogue.py 2>&1 | tail -1 | tr '|' '\n' | grep -E "dbg|c32"; done > gpurun_out/c20_ab.log
cat gpurun_out/c20_ab.log
