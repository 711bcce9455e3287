#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_gpu_decode_chain.py -q -x 2>&1 | tail -3
for a in 1 0; do
  B200BIT_CHAIN_ADAPTIVE=$a timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_47_bench_adaptive$a.json 2> gpurun_out/r2_47_bench_adaptive$a.err
  echo "adaptive=$a rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2_47_bench_adaptive$a.json'));print(d['value'], d['roofline']['frac'])")"
done
