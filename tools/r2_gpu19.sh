#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mpq_" -c 1000 --csv --log-file gpurun_out/r2_32_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_32_ncu_bench.log 2>&1
echo "launch list rc=$?"; grep -c "mpq_chain" gpurun_out/r2_32_launches.csv
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:mpq_tc -s 4 -c 1 -o gpurun_out/r2_32_tc_ncu_m32 python tools/tc_one.py 32 4096 4096 > gpurun_out/r2_32_ncu_tc32.log 2>&1
echo "ncu tc32 rc=$?"; tail -2 gpurun_out/r2_32_ncu_tc32.log
