#!/bin/bash
# A/B: conv kernel occupancy bound 1 vs 2 CTAs/SM
python -m pytest tests/test_conv_gpu.py tests/test_backward_gpu.py -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-layers 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('A minblocks=1', d['value'], d['roofline']['class_ms'])"
cp profiles/layers_latest.json gpurun_out/layers_A.json
NUNET_NVCC_EXTRA="-DNUNET_CONV_MINBLOCKS=2" python pytorch_nested-unet_b200/build.py --force > /dev/null
python -m pytest tests/test_conv_gpu.py -m gpu -q -x 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-layers 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B minblocks=2', d['value'], d['roofline']['class_ms'])"
cp profiles/layers_latest.json gpurun_out/layers_B.json
