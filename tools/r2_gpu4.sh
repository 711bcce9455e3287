#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_gpu_decode_chain.py -x -q 2>&1 | tail -3
timeout -k 5 90 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_10_bench_chain.json 2> gpurun_out/r2_10_bench_chain.err
python -c "import json;d=json.load(open('gpurun_out/r2_10_bench_chain.json'));print('chain',d['value'],d['ms_per_step'],d['roofline']['frac'])"
timeout -k 5 400 python -m pytest tests/test_gpu_optim2.py tests/test_gpu_mbwq.py tests/test_gpu_configs.py -q > gpurun_out/r2_10_new_tests.log 2>&1
echo "new tests rc=$?"; tail -40 gpurun_out/r2_10_new_tests.log | cut -c1-300
timeout -k 5 420 python tools/bench_configs.py --quick --out gpurun_out/r2_10_configs_quick.json > gpurun_out/r2_10_configs.log 2>&1
echo "configs rc=$?"; tail -5 gpurun_out/r2_10_configs.log | cut -c1-300
