#!/bin/bash
# final validation A: whole GPU suite, headline bench (+ reference arm), ncu launch list of the bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2_59_gpu_tests.log
timeout -k 5 400 python bench.py > gpurun_out/r2_59_bench.json 2> gpurun_out/r2_59_bench.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_59_bench.json
timeout -k 5 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_59_bench_reference.json 2> gpurun_out/r2_59_bench_reference.err
echo "ref rc=$?"; cut -c1-400 gpurun_out/r2_59_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
