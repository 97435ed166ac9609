#!/bin/bash
cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | cut -c1-600 ) > gpurun_out/c4_tests.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
cp profiles/layers_latest.json gpurun_out/c4_layers.json
python tools/membw.py > gpurun_out/c4_membw.log 2>&1; cat gpurun_out/c4_membw.log; cat gpurun_out/c4_tests.log; tail -5 gpurun_out/c4_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c4_bench.json')); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['class_ms'])"
