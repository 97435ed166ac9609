#!/bin/bash
cd $GRAFT_REPO_ROOT
( time python -m pytest tests/test_model_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -60 ) > gpurun_out/c2_tests.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
cp profiles/layers_latest.json gpurun_out/c2_layers.json
NUNET_NVCC_EXTRA="-DNUNET_PROFILE_WAITS" python pytorch_nested-unet_b200/build.py --force > gpurun_out/c2_build.log 2>&1
python tools/prof_waits.py > gpurun_out/c2_waits.log 2>&1
tail -40 gpurun_out/c2_tests.log; python -c "
import json; d=json.load(open('gpurun_out/c2_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['class_ms'])"; cat gpurun_out/c2_waits.log
