#!/bin/bash
python -m pytest tests/test_conv_gpu.py tests/test_backward_gpu.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-layers 2>gpurun_out/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], d['roofline']['class_ms'], 'frac', d['roofline']['frac'])"
tail -3 gpurun_out/bench.err
cp profiles/layers_latest.json gpurun_out/layers_latest.json
