#!/usr/bin/env python
"""N back-to-back calls of the decode kernel on ONE layer (weights stay in L2): with `ncu --cache-control none -s N-1 -c 1`
the profiled launch sees warm data, i.e. the kernel's compute side rather than HBM latency."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_mpq_inputs
from bitorch_engine_b200 import _cabi
from bitorch_engine_b200.extensions import q_linear_cuda
K, N, path, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lib = _cabi.lib(); lib.b200bit_set_path(path, 0)
inp = make_mpq_inputs(K, N, 4, 128, "f16", False, M=1, seed=1, device="cuda")
for _ in range(reps):
    y = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, False)
torch.cuda.synchronize()
print("ok", y[0, :4].tolist())
