#!/bin/bash
cd "$(dirname "$0")/.."
timeout -k 5 500 python -m pytest tests/test_gpu_mpq_forward.py -q -x 2>&1 | tail -3 | cut -c1-300
