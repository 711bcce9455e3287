#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pd in 1 2 4; do
  B200BIT_CHAIN_POLLS=$pd timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_44_bench_polls$pd.json 2> gpurun_out/r2_44_bench_polls$pd.err
  echo "polls=$pd rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2_44_bench_polls$pd.json'));print(d['value'], d['roofline']['frac'])")"
done
for e in 35 65; do
  B200BIT_CHAIN_EARLY=$e timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_44_bench_early$e.json 2> gpurun_out/r2_44_bench_early$e.err
  echo "early=$e rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2_44_bench_early$e.json'));print(d['value'], d['roofline']['frac'])")"
done
