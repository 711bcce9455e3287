#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for r in 1 4 8 16; do
  B200BIT_CHAIN_REPLICAS=$r timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_29_bench_rep$r.json 2> gpurun_out/r2_29_bench_rep$r.err
  echo "replicas=$r rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2_29_bench_rep$r.json'));print(d['value'], d['roofline']['frac'])")"
done
timeout -k 5 200 python -m pytest tests/test_gpu_decode_chain.py -q -x 2>&1 | tail -3
timeout -k 5 200 python tools/trace_chain.py 4 > gpurun_out/r2_29_chain_timeline.txt 2>&1
echo "trace rc=$?"; sed -n 16,23p gpurun_out/r2_29_chain_timeline.txt | cut -c1-420
