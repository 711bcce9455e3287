#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_configs.py -q -x 2>&1 | tail -2 | cut -c1-300
timeout -k 5 300 python bench.py --workload llama3_8b_prefill512 --no-cpu-baseline > gpurun_out/r2_60_llama3.json 2> gpurun_out/r2_60_llama3.err
echo "llama3 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2_60_llama3.json'));print(d['value'], d['prefill'], d['decode'])"
timeout -k 5 200 python tools/tc_one.py 512 4096 1024 2>&1 | tail -2
