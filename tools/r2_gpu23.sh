#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 1 2 4 8; do
  B200BIT_EXL2_CSIZE=$c timeout -k 5 300 python tools/bench_configs.py --quick --only exl2 --out gpurun_out/r2_38_configs_exl2_c$c.json > gpurun_out/r2_38_configs_c$c.log 2>&1
  echo "csize=$c rc=$?"; grep -E "CUDA graph" gpurun_out/r2_38_configs_c$c.log | cut -c10-120
done
