#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 90 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_23_bench_chain.json 2> gpurun_out/r2_23_bench_chain.err
python -c "import json;d=json.load(open('gpurun_out/r2_23_bench_chain.json'));print('chain',d['value'],d['ms_per_step'],d['roofline']['frac'])"
timeout -k 5 600 python -m pytest tests -m gpu -q > gpurun_out/r2_23_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -12 gpurun_out/r2_23_gpu_tests.log | cut -c1-250
timeout -k 5 500 python tools/bench_configs.py --out gpurun_out/r2_23_configs.json > gpurun_out/r2_23_configs.log 2>&1
echo "configs rc=$?"; tail -3 gpurun_out/r2_23_configs.log | cut -c1-250
