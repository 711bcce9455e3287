#!/usr/bin/env python
"""Timeline of the cross-kernel pipelined decode GEMV (mpq_pipe.cuh) from its in-kernel %globaltimer stamps:
a CUDA graph of a Llama-7B decoder layer's 7 linears (q,k,v,o 4096x4096; gate,up 4096x11008; down 11008x4096) repeated
`--layers` times, PDL-chained; per node: when CTAs start, pass griddepcontrol.wait, see their first stage, finish."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bitorch_engine_b200 import _cabi
from bitorch_engine_b200.extensions import q_linear_cuda

ap = argparse.ArgumentParser()
ap.add_argument("--pdl", type=int, default=1)
ap.add_argument("--tune", default="0:0:0")
ap.add_argument("--layers", type=int, default=3)
ap.add_argument("--shapes", default="4096x4096,4096x4096,4096x4096,4096x4096,4096x11008,4096x11008,11008x4096")
args = ap.parse_args()
dev = torch.device("cuda:0")
lib = _cabi.lib()
lib.b200bit_set_path(6, 0)
L, wp, sk = (int(v) for v in args.tune.split(":"))
lib.b200bit_set_gemv_tuning(L, wp, sk)
g = torch.Generator(device=dev).manual_seed(0)
shapes = [tuple(int(v) for v in s.split("x")) for s in args.shapes.split(",")] * args.layers
ws, xs = [], {}
for K, N in shapes:
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), dtype=torch.int32, device=dev, generator=g)
    sc = (torch.rand((K // 128, N), device=dev, generator=g) * 0.01 + 0.005).half()
    ws.append((K, N, qw, sc, (sc.float() * 8).half(), torch.arange(K, dtype=torch.int32, device=dev) // 128))
    xs.setdefault(K, torch.randn((1, K), device=dev, generator=g).half())
MAXCTA = 4096
trace = torch.zeros((len(ws), MAXCTA * 8), dtype=torch.int64, device=dev)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    def run(tr):
        for i, (K, N, qw, sc, zr, gi) in enumerate(ws):
            if tr:
                lib.b200bit_set_trace_buffer(trace[i].data_ptr())
            q_linear_cuda.mpq_forward(xs[K], qw, sc, zr, gi, 16, 4, False, pdl=bool(args.pdl))
    run(False); stream.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        run(True)
    lib.b200bit_set_trace_buffer(None)
    for _ in range(3):
        graph.replay()
    stream.synchronize()
t = trace.cpu().view(len(ws), MAXCTA, 8)
names = ["start", "waited", "xsums", "stage0", "loopend", "written", "ticketed"]
first = len(shapes) // args.layers          # skip the first layer (graph start-up)
base = int(t[first][:, 0][t[first][:, 0] > 0].min())
print(f"pdl {args.pdl} tune {args.tune}  (ns relative to node {first} first CTA start; min / median / max over CTAs)")
prev_wait = None
for node in range(first, len(ws)):
    K, N = ws[node][0], ws[node][1]
    tn = t[node]
    line = [f"node {node:3d} {K}x{N}:"]
    for si, nm in enumerate(names):
        v = tn[:, si]; v = v[v > 0] - base
        if v.numel():
            line.append(f"{nm} {int(v.min())}/{int(v.median())}/{int(v.max())}")
    w = tn[:, 1]; w = w[w > 0] - base
    if prev_wait is not None and w.numel():
        line.append(f"| period {int(w.median()) - prev_wait}")
    prev_wait = int(w.median()) if w.numel() else None
    print("  ".join(line))
