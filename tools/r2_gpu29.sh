#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:exl2_tma -s 2 -c 1 -o gpurun_out/r2_50_exl2_tma_m1 python tools/exl2_one.py 1 4096 11008 > gpurun_out/r2_50_ncu1.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_50_ncu1.log
