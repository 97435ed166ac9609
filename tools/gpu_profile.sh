#!/bin/bash
cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300 ) > gpurun_out/r02_tests.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --profile-layers > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
cp profiles/layers_latest.json gpurun_out/r02_layers.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
python tools/membw.py > gpurun_out/r02_membw.log 2>&1
NUNET_GRAPH_TRAIN=0 python bench.py --steps 1 --warmup 2 --kernels-only > gpurun_out/r02_plain.log 2>&1 &&
NUNET_GRAPH_TRAIN=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 2 --kernels-only > gpurun_out/r02_ncu.log 2>&1
cat gpurun_out/r02_tests.log; head -c 1500 gpurun_out/r02_bench.json; echo; cat gpurun_out/r02_bench_reference.json | head -c 400; echo; tail -3 gpurun_out/r02_ncu.log; wc -l gpurun_out/r02_launches.csv
