#!/bin/bash
cd $GRAFT_REPO_ROOT
for v in "16 0" "8 0" "16 1" "8 1" "16 2" "16 4" "16 6" "16 7"; do
  set -- $v
  NUNET_NVCC_EXTRA="-DNUNET_TMA_EPI_WARPS=$1 -DNUNET_DBG_SKIP=$2" python pytorch_nested-unet_b200/build.py --force > /dev/null 2>&1
  TAG="epi=$1 skip=$2" python tools/ab_epilogue.py 2>&1 | tail -1
done > gpurun_out/ab_epilogue.log 2>&1
cat gpurun_out/ab_epilogue.log
