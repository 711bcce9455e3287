#!/bin/bash
# final validation B: per-config table, ncu of the tcgen05 kernel, launch list of the bench command, config #5 harness
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 1200 python tools/bench_configs.py --out gpurun_out/r2_31_configs.json > gpurun_out/r2_31_configs.log 2>&1
echo "configs rc=$?"; tail -3 gpurun_out/r2_31_configs.log | cut -c1-300
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mpq_|b200bit" -c 100 --csv --log-file gpurun_out/r2_31_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_31_ncu_bench.log 2>&1
echo "launch list rc=$?"; grep -c "mpq_chain" gpurun_out/r2_31_launches.csv
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:mpq_tc -s 1 -c 1 -o gpurun_out/r2_31_tc_ncu_m2048 python tools/tc_one.py 2048 4096 4096 > gpurun_out/r2_31_ncu_tc.log 2>&1
echo "ncu tc rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:mpq_tc -s 4 -c 1 -o gpurun_out/r2_31_tc_ncu_m32 python tools/tc_one.py 32 4096 4096 > gpurun_out/r2_31_ncu_tc32.log 2>&1
echo "ncu tc32 rc=$?"
timeout -k 5 400 python bench.py --workload llama3_8b_prefill512 --no-cpu-baseline > gpurun_out/r2_31_llama3.json 2> gpurun_out/r2_31_llama3.err
echo "llama3 rc=$?"; cut -c1-500 gpurun_out/r2_31_llama3.json
