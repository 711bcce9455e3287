#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 200 python tools/tc_check.py quick > gpurun_out/r2_25_tc_check.jsonl 2>&1
echo "tc rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_25_tc_check.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    if 'tc_us' in d: print(d['K'],d['N'],d['M'],'tc',d['tc_us'],'deq+cublas',d['dequant_cublas_us'],'small',d.get('small_batch_kernel_us'),'TF',d['tc_TFLOPs'],'err %.2e'%d['rel_err_vs_fp16W_fp64'])
PY
timeout -k 5 120 python -m pytest tests/test_gpu_mpq_forward.py -q -k "row_counts or config" 2>&1 | tail -3
