#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python tools/tc_check.py quick > gpurun_out/r2_13_tc_check.log 2>&1
echo "tc rc=$?"; tail -60 gpurun_out/r2_13_tc_check.log | cut -c1-330
