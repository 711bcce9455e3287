#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_mbwq.py -q -x 2>&1 | tail -4 | cut -c1-300
timeout -k 5 300 python tools/bench_configs.py --only exl2 --out gpurun_out/r2_53_configs_exl2.json > gpurun_out/r2_53_configs.log 2>&1
echo "configs rc=$?"; grep -E "exl2_forward fused|REFERENCE mbwq" gpurun_out/r2_53_configs.log | cut -c1-200
