#!/usr/bin/env python
"""The three Llama-7B decode shapes through the default path, twice each (for `ncu --set full`: launches 1, 3, 5 are the
warmed-up ones; ncu's default cache control flushes L2 in front of every replay, so DRAM traffic is the cold-cache one)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_mpq_inputs
from bitorch_engine_b200.extensions import q_linear_cuda
for K, N in ((4096, 4096), (4096, 11008), (11008, 4096)):
    inp = make_mpq_inputs(K, N, 4, 128, "f16", False, M=1, seed=1, device="cuda")
    for _ in range(2):
        y = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, False)
    torch.cuda.synchronize()
    print("ok", K, N, y[0, :2].tolist())
