#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:mpq_tc -s 1 -c 1 -o gpurun_out/r2_15_tc_ncu python tools/tc_one.py 512 4096 4096 > gpurun_out/r2_15_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_15_ncu.log
