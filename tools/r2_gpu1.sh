#!/bin/bash
# round-2 GPU call 1: chain parity, bench (chain vs per-layer), chain timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_chain.py -x -q > gpurun_out/r2_01_chain_tests.log 2>&1
echo "chain tests rc=$?" | tee -a gpurun_out/r2_01_chain_tests.log
tail -15 gpurun_out/r2_01_chain_tests.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_01_bench_chain.json 2> gpurun_out/r2_01_bench_chain.err
echo "bench chain rc=$?"; cat gpurun_out/r2_01_bench_chain.json; tail -3 gpurun_out/r2_01_bench_chain.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --chain 0 > gpurun_out/r2_01_bench_perlayer.json 2> gpurun_out/r2_01_bench_perlayer.err
echo "bench per-layer rc=$?"; cat gpurun_out/r2_01_bench_perlayer.json
timeout 300 python tools/trace_chain.py 4 > gpurun_out/r2_01_chain_timeline.txt 2>&1
echo "trace rc=$?"; cat gpurun_out/r2_01_chain_timeline.txt
