#!/bin/bash
# baseline validation of the round-1 tree + tiling sweep
cd $GRAFT_REPO_ROOT
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/c1_tests.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
cp profiles/layers_latest.json gpurun_out/c1_layers.json
python tools/time_elem.py > gpurun_out/c1_elem.log 2>&1
python tools/time_wgrad.py > gpurun_out/c1_wgrad.log 2>&1
NUNET_NVCC_EXTRA="-DNUNET_EXPERIMENT" python pytorch_nested-unet_b200/build.py --force > gpurun_out/c1_build.log 2>&1
python tools/sweep_conv.py > gpurun_out/c1_sweep.log 2>&1
tail -5 gpurun_out/c1_tests.log; cat gpurun_out/c1_bench.json | head -c 600; tail -30 gpurun_out/c1_sweep.log
