#!/usr/bin/env python
"""Timeline of the TMA-streamed kernel from its in-kernel %globaltimer stamps (diagnostic)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bitorch_engine_b200 import _cabi
from bitorch_engine_b200.extensions import q_linear_cuda

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="4096x4096")
ap.add_argument("--pdl", type=int, default=1)
ap.add_argument("--tune", default="0:0:0")
ap.add_argument("--chain", type=int, default=6)
ap.add_argument("--path", type=int, default=4)
ap.add_argument("--M", type=int, default=1)
args = ap.parse_args()
K, N = (int(v) for v in args.shape.split("x"))
dev = torch.device("cuda:0")
lib = _cabi.lib()
lib.b200bit_set_path(args.path, 1)
L, wp, sk = (int(v) for v in args.tune.split(":"))
lib.b200bit_set_gemv_tuning(L, wp, sk)
g = torch.Generator(device=dev).manual_seed(0)
ws = []
for i in range(args.chain):
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), dtype=torch.int32, device=dev, generator=g)
    sc = (torch.rand((K // 128, N), device=dev, generator=g) * 0.01 + 0.005).half()
    ws.append((qw, sc, (sc.float() * 8).half()))
gi = torch.arange(K, dtype=torch.int32, device=dev) // 128
x = torch.randn((args.M, K), device=dev, generator=g).half()
trace = torch.zeros((args.chain, 148 * 16 * 8), dtype=torch.int64, device=dev)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    def run(tr):
        for i, (qw, sc, zr) in enumerate(ws):
            if tr:
                lib.b200bit_set_trace_buffer(trace[i].data_ptr())
            q_linear_cuda.mpq_forward(x, qw, sc, zr, gi, 16, 4, False, pdl=bool(args.pdl))
    run(False); stream.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        run(True)
    lib.b200bit_set_trace_buffer(None)
    for _ in range(3):
        graph.replay()
    stream.synchronize()
t = trace.cpu().view(args.chain, 148, 16, 8)
names = ["start", "init", "depwait", "tile0", "run0", "runs", "bar", "end"]
base = int(t[2][t[2] > 0].min())
print(f"shape {args.shape} pdl {args.pdl} tune {args.tune}")
for node in range(2, args.chain):
    tn = t[node]
    print(f"node {node}: (ns relative to node 2 first stamp; min / median / max over CTAs x warps that wrote the stamp)")
    for si, nm in enumerate(names):
        for wsel, wname in ((slice(0, 4), "w0-3"), (slice(4, 8), "w4-7"), (slice(8, 12), "w8-11"), (slice(12, 14), "w12-13"), (slice(14, 15), "w14/prod"), (slice(15, 16), "w15/mma")):
            v = tn[:, wsel, si]; v = v[v > 0] - base
            if v.numel():
                print(f"   {nm:8s} {wname:5s} min {int(v.min()):7d}  med {int(v.median()):7d}  max {int(v.max()):7d}   n={v.numel()}")
