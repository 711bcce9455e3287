#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:exl2_gemv -s 2 -c 1 -o gpurun_out/r2_36_exl2_m1 python tools/exl2_one.py 1 11008 4096 > gpurun_out/r2_36_ncu1.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_36_ncu1.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:exl2_gemv -s 2 -c 1 -o gpurun_out/r2_36_exl2_m32 python tools/exl2_one.py 32 4096 4096 > gpurun_out/r2_36_ncu32.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_36_ncu32.log
