#!/bin/bash
# round-2 GPU call: chain A/B (hint vs direct polling), timeline, ncu full of the chain kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_chain.py -x -q 2>&1 | tail -3
for mode in 0 1; do
  B200BIT_CHAIN_POLL=$mode timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_07_bench_poll$mode.json 2> gpurun_out/r2_07_bench_poll$mode.err
  echo "poll=$mode rc=$?"; python -c "import json;d=json.load(open('gpurun_out/r2_07_bench_poll$mode.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'])"
done
B200BIT_CHAIN_POLL=1 timeout 600 python -m pytest tests/test_gpu_decode_chain.py -x -q 2>&1 | tail -3
timeout 300 python tools/trace_chain.py 4 > gpurun_out/r2_07_chain_timeline_hint.txt 2>&1
B200BIT_CHAIN_POLL=1 timeout 300 python tools/trace_chain.py 4 > gpurun_out/r2_07_chain_timeline_direct.txt 2>&1
sed -n 9,23p gpurun_out/r2_07_chain_timeline_direct.txt | cut -c1-330
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpq_chain -s 3 -c 1 -o gpurun_out/r2_07_chain_ncu python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_07_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_07_ncu.log; ls -la gpurun_out/*.ncu-rep
