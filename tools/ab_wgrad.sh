#!/bin/bash
for pw in 8 16; do
  NUNET_NVCC_EXTRA="-DNUNET_WGRAD_PRODUCER_WARPS=$pw" python pytorch_nested-unet_b200/build.py --force > /dev/null
  python -m pytest tests/test_backward_gpu.py -m gpu -q -k wgrad 2>&1 | tail -1
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_$pw.json
  python -c "import json; d=json.load(open('gpurun_out/ab_$pw.json')); print('PW $pw', d['value'], d['roofline']['class_ms']['conv_wgrad'])"
done
