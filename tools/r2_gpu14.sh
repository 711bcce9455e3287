#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for e in 0 25 50 75 95; do
  B200BIT_CHAIN_EARLY=$e timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_27_bench_early$e.json 2> gpurun_out/r2_27_bench_early$e.err
  echo "early=$e rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2_27_bench_early$e.json'));print(d['value'], d['roofline']['frac'])")"
done
B200BIT_CHAIN_EARLY=50 timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --order consumer-first > gpurun_out/r2_27_bench_order_cf.json 2> gpurun_out/r2_27_bench_order_cf.err
echo "consumer-first rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2_27_bench_order_cf.json'));print(d['value'], d['roofline']['frac'])")"
timeout -k 5 200 python -m pytest tests/test_gpu_decode_chain.py -q -x 2>&1 | tail -3
