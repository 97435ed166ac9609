#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_backward_gpu.py -m gpu -q -x -k up_transpose 2>&1 | tail -2 | cut -c1-300 > gpurun_out/upt_t.log
cat gpurun_out/upt_t.log
for rb in 16 12 24 32; do echo "rb=$rb"; NUNET_UPT_RB=$rb timeout 200 python tools/one_upt.py; done > gpurun_out/upt.log 2>&1
cat gpurun_out/upt.log
