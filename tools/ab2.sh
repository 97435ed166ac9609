#!/bin/bash
cd $GRAFT_REPO_ROOT
for v in "-DNUNET_PROFILE_WAITS" "-DNUNET_PROFILE_WAITS -DNUNET_POLL1" "-DNUNET_PROFILE_WAITS -DNUNET_DBG_SKIP=7"; do
  NUNET_NVCC_EXTRA="$v" python pytorch_nested-unet_b200/build.py --force > /dev/null 2>&1
  echo "=== $v"; python tools/prof_waits.py 2>&1 | tail -12
done > gpurun_out/ab2.log 2>&1
NUNET_NVCC_EXTRA="-DNUNET_POLL1" python pytorch_nested-unet_b200/build.py --force > /dev/null 2>&1
TAG="poll1" python tools/ab_epilogue.py 2>&1 | tail -1 >> gpurun_out/ab2.log
cat gpurun_out/ab2.log
