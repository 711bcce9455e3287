#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 128 256 512; do
  B200BIT_OPTIM_WIDTH=$w timeout -k 5 300 python tools/bench_configs.py --only optim --out gpurun_out/r2_56_optim_w$w.json > gpurun_out/r2_56_optim_w$w.log 2>&1
  echo "width=$w rc=$?"; grep -E "diodemix_mpq" gpurun_out/r2_56_optim_w$w.log | cut -c10-150
done
