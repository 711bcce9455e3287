#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_mpq_forward.py tests/test_gpu_configs.py tests/test_gpu_mbwq.py -q 2>&1 | tail -6 | cut -c1-250
timeout -k 5 300 python tools/bench_configs.py --quick --only nbit --out gpurun_out/r2_26_configs_nbit.json > gpurun_out/r2_26_configs.log 2>&1
echo "configs rc=$?"; grep -E "M=32 \(CUDA graph|M=32 \(eager" gpurun_out/r2_26_configs.log | cut -c1-200
