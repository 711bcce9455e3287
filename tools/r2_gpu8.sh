#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 200 python tools/tc_check.py quick > gpurun_out/r2_20_tc_check.jsonl 2>&1
echo "tc rc=$?"; grep -v '"asym": true\|"group": 32' gpurun_out/r2_20_tc_check.jsonl | tail -32 | cut -c90-330
