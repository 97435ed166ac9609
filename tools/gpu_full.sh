#!/bin/bash
# full GPU suite, smoke, bench (with the CPU baseline leg), reference arm
cd $GRAFT_REPO_ROOT
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | cut -c1-400 ) > gpurun_out/full_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/full_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/full_bench_reference.json 2> gpurun_out/full_bench_reference.err
cat gpurun_out/full_tests.log; tail -2 gpurun_out/full_smoke.log; cat gpurun_out/full_bench.json | cut -c1-900; tail -4 gpurun_out/full_bench.err | cut -c1-400; cat gpurun_out/full_bench_reference.json | cut -c1-600
