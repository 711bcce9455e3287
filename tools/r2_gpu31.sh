#!/bin/bash
# last call of the round: the whole per-config table and the ncu capture of the chain kernel on the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 1200 python tools/bench_configs.py --out gpurun_out/r2_55_configs.json > gpurun_out/r2_55_configs.log 2>&1
echo "configs rc=$?"; tail -2 gpurun_out/r2_55_configs.log | cut -c1-200
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:mpq_chain -s 3 -c 1 -o gpurun_out/r2_55_chain_ncu python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_55_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_55_ncu.log | cut -c1-200
timeout -k 5 300 python tools/trace_chain.py 4 > gpurun_out/r2_55_chain_timeline.txt 2>&1
echo "trace rc=$?"
