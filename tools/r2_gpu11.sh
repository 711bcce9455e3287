#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 4 2 1; do
  B200BIT_CHAIN_POLLS=$d timeout -k 5 90 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_24_bench_polls$d.json 2> gpurun_out/r2_24_bench_polls$d.err
  python -c "import json;d=json.load(open('gpurun_out/r2_24_bench_polls$d.json'));print('polls $d',d['value'],d['ms_per_step'],d['roofline']['frac'])"
done
timeout -k 5 600 python -m pytest tests -m gpu -q > gpurun_out/r2_24_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/r2_24_gpu_tests.log | cut -c1-250
