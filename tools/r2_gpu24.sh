#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -3
B200BIT_TEST_TP=1 timeout -k 5 300 python -m pytest tests/test_gpu_tensor_parallel.py -q -x 2>&1 | tail -5 | cut -c1-300
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2_39_bench_2gpu.json 2> gpurun_out/r2_39_bench_2gpu.err
echo "bench2 rc=$?"; cut -c1-300 gpurun_out/r2_39_bench_2gpu.json
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload llama3_8b_prefill512 > gpurun_out/r2_39_llama3_2gpu.json 2> gpurun_out/r2_39_llama3_2gpu.err
echo "llama3 x2 rc=$?"; cut -c1-300 gpurun_out/r2_39_llama3_2gpu.json
