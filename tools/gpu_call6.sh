#!/bin/bash
cd $GRAFT_REPO_ROOT
TAG="fetch" timeout 300 python tools/ab_epilogue.py 2>&1 | tail -2 > gpurun_out/c26_ab.log
( time timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_conv_fused_gpu.py tests/test_backward_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -15 | cut -c1-400 ) > gpurun_out/c26_tests.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err
cp profiles/layers_latest.json gpurun_out/c26_layers.json
cat gpurun_out/c26_ab.log; cat gpurun_out/c26_tests.log; tail -3 gpurun_out/c26_bench.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/c26_bench.json')); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['class_ms'])"
