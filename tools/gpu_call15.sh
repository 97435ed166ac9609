#!/bin/bash
cd $GRAFT_REPO_ROOT
NUNET_NVCC_EXTRA="-DNUNET_EXPERIMENT" python pytorch_nested-unet_b200/build.py --force > /dev/null 2>&1
( TAG="nt256" NUNET_NT256=1 python tools/ab_epilogue.py 2>&1 | tail -1
TAG="nt128" python tools/ab_epilogue.py 2>&1 | tail -1
TAG="nt128T2" NUNET_CONV_T=2 python tools/ab_epilogue.py 2>&1 | tail -1 ) > gpurun_out/c15_ab.log
cat gpurun_out/c15_ab.log
