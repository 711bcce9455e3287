#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_functions.py tests/test_gpu_mpq_aux.py tests/test_gpu_optim.py tests/test_gpu_optim2.py -q -x 2>&1 | tail -4 | cut -c1-300
timeout -k 5 300 python tools/bench_configs.py --only optim,functions --out gpurun_out/r2_45_configs_optim_fn.json > gpurun_out/r2_45_configs.log 2>&1
echo "configs rc=$?"; grep -E "diodemix|functions_cuda" gpurun_out/r2_45_configs.log | cut -c1-200
timeout -k 5 300 python tools/bench_configs.py --quick --only nbit --out gpurun_out/r2_45_configs_nbit.json > gpurun_out/r2_45_configs_nbit.log 2>&1
echo "nbit rc=$?"; grep -E "mpq_pack_weight|mpq_dequant" gpurun_out/r2_45_configs_nbit.log | cut -c1-200
