#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_mpq_forward.py tests/test_gpu_configs.py -q -x 2>&1 | tail -3 | cut -c1-300
timeout -k 5 300 python tools/bench_configs.py --quick --only nbit --out gpurun_out/r2_58_configs_nbit.json > gpurun_out/r2_58_configs.log 2>&1
echo "configs rc=$?"; grep -E "M=32 \(CUDA graph|M=512 \(CUDA graph" gpurun_out/r2_58_configs.log | cut -c10-140
