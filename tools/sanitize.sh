#!/bin/bash
# compute-sanitizer (memcheck + synccheck; racecheck on the shared-memory hand-overs) over the three hot kernels at small sizes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool command...
  local name=$1 tool=$2; shift 2
  timeout -k 5 400 $S --tool $tool --error-exitcode 7 "$@" > gpurun_out/r2_43_sanitizer_${name}_${tool}.log 2>&1
  echo "$name $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_43_sanitizer_${name}_${tool}.log | tail -1)"
}
run tc memcheck python tools/tc_one.py 64 1024 512
run tc synccheck python tools/tc_one.py 64 1024 512
run exl2 memcheck python tools/exl2_one.py 4 1024 512
run exl2 racecheck python tools/exl2_one.py 4 1024 512
run chain memcheck python tools/chain_one.py
run chain synccheck python tools/chain_one.py
run imma memcheck python tools/chain_one.py layers
