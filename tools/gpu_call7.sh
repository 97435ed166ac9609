#!/bin/bash
cd $GRAFT_REPO_ROOT
NUNET_NVCC_EXTRA="-DNUNET_PROFILE_WAITS" python pytorch_nested-unet_b200/build.py --force > /dev/null 2>&1
python tools/prof_waits.py 2>&1 | tail -40 > gpurun_out/c25_waits.log
cat gpurun_out/c25_waits.log
