#!/usr/bin/env python
"""Structured-input diagnosis of the pipelined decode GEMV (path 6)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bitorch_engine_b200 import _cabi
from bitorch_engine_b200.extensions import q_linear_cuda
lib = _cabi.lib()
dev = "cuda"
K, N, g = 512, 64, 128
G = K // g
gi = torch.arange(K, dtype=torch.int32, device=dev) // g

def run(x, qw, sc, zr, path):
    lib.b200bit_set_path(path, 0)
    y = q_linear_cuda.mpq_forward(x, qw, sc, zr, gi, 16, 4, False)
    torch.cuda.synchronize()
    return y.float().cpu()

ones_q = torch.full((K // 8, N), 0x11111111, dtype=torch.int32, device=dev)
s1 = torch.ones((G, N), dtype=torch.float16, device=dev)
z0 = torch.zeros((G, N), dtype=torch.float16, device=dev)
xo = torch.ones((1, K), dtype=torch.float16, device=dev)
for path in (1, 6):
    print("path", path)
    y = run(xo, ones_q, s1, z0, path); print(" q=1,s=1,z=0,x=1 -> expect", K, ":", y[0, :8].tolist(), "uniq", y.unique().tolist()[:8])
    y = run(xo, ones_q * 0, s1 * 0, z0 + 1, path); print(" s=0,z=1,x=1 -> expect", -K, ":", y[0, :8].tolist(), "uniq", y.unique().tolist()[:8])
    # one-hot x: y[n] = q[k, n]
    q = torch.randint(0, 16, (K, N), device=dev)
    qw = torch.zeros((K // 8, N), dtype=torch.int64, device=dev)
    for j in range(8):
        qw |= q[j::8].long() << (4 * j)
    qw = torch.where(qw >= 2 ** 31, qw - 2 ** 32, qw).int()
    bad = []
    for k in list(range(0, 40)) + [127, 128, 255, 256, 300, 511]:
        x = torch.zeros((1, K), dtype=torch.float16, device=dev); x[0, k] = 1
        y = run(x, qw, s1, z0, path)
        if not torch.equal(y[0], q[k].float().cpu()):
            bad.append((k, y[0, :6].tolist(), q[k, :6].tolist()))
    print(" one-hot mismatches:", len(bad)); [print("   ", b) for b in bad[:12]]
    # column identity: s varies per column
    sc = (torch.arange(N, device=dev).half() + 1).repeat(G, 1)
    y = run(xo, ones_q, sc, z0, path); print(" s=n+1 -> expect", K, "*(n+1):", (y[0, :8] / K).tolist())
    sg = (torch.arange(G, device=dev).half() + 1)[:, None].repeat(1, N)
    y = run(xo, ones_q, sg, z0, path); print(" s=g+1 -> expect", g * sum(range(1, G + 1)), ":", y[0, :4].tolist())
lib.b200bit_set_path(0, 0)
