#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_gpu_decode_chain.py -x -q 2>&1 | tail -3
timeout -k 5 90 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_22_bench_chain.json 2> gpurun_out/r2_22_bench_chain.err
python -c "import json;d=json.load(open('gpurun_out/r2_22_bench_chain.json'));print('chain',d['value'],d['ms_per_step'],d['roofline']['frac'])"
timeout -k 5 90 python tools/trace_chain.py 4 > gpurun_out/r2_22_chain_timeline.txt 2>&1
sed -n 9,16p gpurun_out/r2_22_chain_timeline.txt | cut -c1-360
timeout -k 5 300 python -m pytest tests/test_gpu_mpq_forward.py tests/test_gpu_configs.py tests/test_gpu_optim2.py -q -x 2>&1 | tail -5 | cut -c1-300
