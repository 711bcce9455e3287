#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_gpu_decode_chain.py -x -q 2>&1 | tail -3
[ ${PIPESTATUS[0]} -ne 0 ] && { echo "chain tests failed/hung: stop"; exit 1; }
for S in 6 4; do
  B200BIT_CHAIN_SLOTS=$S timeout -k 5 90 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_08_bench_S$S.json 2> gpurun_out/r2_08_bench_S$S.err
  echo "S=$S rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2_08_bench_S$S.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'])"
done
timeout -k 5 90 python tools/trace_chain.py 4 > gpurun_out/r2_08_chain_timeline.txt 2>&1
sed -n 9,16p gpurun_out/r2_08_chain_timeline.txt | cut -c1-360
