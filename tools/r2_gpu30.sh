#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_decode_chain.py tests/test_gpu_mbwq.py tests/test_chain_plan_cpu.py -q -x 2>&1 | tail -3
timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_57_bench.json 2> gpurun_out/r2_57_bench.err
echo "bench rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2_57_bench.json'));print(d['value'], d['roofline']['frac'])")"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
