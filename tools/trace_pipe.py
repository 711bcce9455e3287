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
ap.add_argument("--dataflow", default="llama")
ap.add_argument("--path", type=int, default=7)
ap.add_argument("--dump", default="", help="comma-separated node indices: per-CTA table (block, SM, stamps) of those nodes")
ap.add_argument("--shapes", default="4096x4096,4096x4096,4096x4096,4096x4096,4096x11008,4096x11008,11008x4096")
args = ap.parse_args()
dev = torch.device("cuda:0")
lib = _cabi.lib()
lib.b200bit_set_path(args.path, 0)
L, wp, sk = (int(v) for v in args.tune.split(":"))
lib.b200bit_set_gemv_tuning(L, wp, sk)
g = torch.Generator(device=dev).manual_seed(0)
shapes = [tuple(int(v) for v in s.split("x")) for s in args.shapes.split(",")] * args.layers
ws, xs = [], {}
for K, N in shapes:
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), dtype=torch.int32, device=dev, generator=g)
    sc = ((torch.rand((K // 128, N), device=dev, generator=g) * 0.5 + 0.75) / (K ** 0.5 * 4.61)).half()
    ws.append((K, N, qw, sc, (sc.float() * 7.5).half(), torch.arange(K, dtype=torch.int32, device=dev) // 128))
    xs.setdefault(K, torch.randn((1, K), device=dev, generator=g).half())
MAXCTA = 4096
trace = torch.zeros((len(ws), MAXCTA * 8), dtype=torch.int64, device=dev)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    def run(tr):
        # bench.py's dataflow when the shape list is the decoder block's (q,k,v <- hidden; o <- v; gate,up <- o;
        # down <- up; next block <- down); otherwise one fixed activation per K
        llama = len(ws) % 7 == 0 and args.dataflow == "llama"
        hid = xs[ws[0][0]]
        outs = []
        for i, (K, N, qw, sc, zr, gi) in enumerate(ws):
            if tr:
                lib.b200bit_set_trace_buffer(trace[i].data_ptr())
            if llama:
                j = i % 7
                x = hid if j < 3 else outs[-1] if j in (3, 4, 6) else outs[-2]
            else:
                x = xs[K]
            outs.append(q_linear_cuda.mpq_forward(x, qw, sc, zr, gi, 16, 4, False, pdl=bool(args.pdl)))
            if llama and j == 6:
                hid = outs[-1]
    run(False); stream.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        run(True)
    lib.b200bit_set_trace_buffer(None)
    for _ in range(3):
        graph.replay()
    stream.synchronize()
t = trace.cpu().view(len(ws), MAXCTA, 8)
names = ["start", "waited", "xsums", "stage0", "loopend", "written", "ticketed", "outwait"]
first = len(shapes) // args.layers          # skip the first layer (graph start-up)
base = int(t[first][:, 0][t[first][:, 0] > 0].min())
print(f"pdl {args.pdl} tune {args.tune}  (ns relative to node {first} first CTA start; min / median / max over CTAs)")
prev_wait = None
for node in range(first, len(ws)):
    K, N = ws[node][0], ws[node][1]
    tn = t[node]
    line = [f"node {node:3d} {K}x{N}:"]
    for si, nm in enumerate(names):
        if si == 6 and int(tn[:, 6].max()) < (1 << 20):
            continue                                   # SM ids, not a stamp
        v = tn[:, si]; v = v[v > 0] - base
        if v.numel():
            line.append(f"{nm} {int(v.min())}/{int(v.median())}/{int(v.max())}")
    w = tn[:, 5]; w = w[w > 0] - base          # "written": the node's last CTA finishing is what the chain sees
    if prev_wait is not None and w.numel():
        line.append(f"| end-to-end period {int(w.max()) - prev_wait}")
    prev_wait = int(w.max()) if w.numel() else None
    print("  ".join(line))

for node in [int(v) for v in args.dump.split(",") if v]:
    tn = t[node]
    rows = [(int(tn[b, 4]) - base, b) for b in range(MAXCTA) if int(tn[b, 0]) > 0]
    rows.sort()
    print(f"--- node {node} {ws[node][0]}x{ws[node][1]}: per CTA, sorted by loop end (ns relative to the same base)")
    print("block  sm  start waited xsums stage0 loopend written  loop_ns  ctas_on_sm")
    sms = [int(tn[b, 6]) - 1 for _, b in rows]
    for le, b in rows:
        st = [int(tn[b, i]) - base for i in range(6)]
        sm = int(tn[b, 6]) - 1
        print(f"{b:5d} {sm:3d} {st[0]:6d} {st[1]:6d} {st[2]:6d} {st[3]:6d} {st[4]:6d} {st[5]:6d} {st[4] - st[3]:7d} {sms.count(sm):3d}")
