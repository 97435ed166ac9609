#!/bin/bash
# loop tests + default bench line (e2e through DevicePrefetcher + StepScalars)
cd $GRAFT_REPO_ROOT
( timeout 600 python -m pytest tests/test_loop_gpu.py tests/test_dp_gpu.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-400 ) > gpurun_out/e2e_tests.log 2>&1
cat gpurun_out/e2e_tests.log
for rep in 1 2 3; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/e2e_bench_$rep.json 2> gpurun_out/e2e_bench.err
python -c "import json; d=json.load(open('gpurun_out/e2e_bench_$rep.json')); print('value %.0f ms %.3f e2e %.0f (%.3f ms) ratio %.4f clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['value']/d['value'], d['clocks']['sm_mhz'])); print({k: d['roofline'][k] for k in ('bound','achieved','peak','unit','frac')})"
tail -3 gpurun_out/e2e_bench.err
done
