#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 120 python tools/tc_check.py quick > gpurun_out/r2_12_tc_check.log 2>&1
echo "tc rc=$?"; tail -20 gpurun_out/r2_12_tc_check.log | cut -c1-330
timeout -k 5 300 python -m pytest tests/test_gpu_optim2.py tests/test_gpu_mbwq.py tests/test_gpu_mpq_aux.py -q -x 2>&1 | tail -5 | cut -c1-300
timeout -k 5 200 python tools/bench_configs.py --quick --only exl2,functions --out gpurun_out/r2_12_configs_exl2.json > gpurun_out/r2_12_configs.log 2>&1
echo "configs rc=$?"; grep -E "exl2_forward fused|REFERENCE mbwq|functions" gpurun_out/r2_12_configs.log | cut -c1-220
