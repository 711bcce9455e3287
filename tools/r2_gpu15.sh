#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_28_bench.json 2> gpurun_out/r2_28_bench.err
echo "bench rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2_28_bench.json'));print(d['value'], d['roofline']['frac'])")"
timeout -k 5 200 python tools/trace_chain.py 4 > gpurun_out/r2_28_chain_timeline.txt 2>&1
echo "trace rc=$?"; sed -n 16,23p gpurun_out/r2_28_chain_timeline.txt | cut -c1-420
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:mpq_chain -s 3 -c 1 -o gpurun_out/r2_28_chain_ncu python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_28_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_28_ncu.log
