"""In-kernel %globaltimer stamps of the decode chain (diagnostics build of mpq_chain_kernel): per node, per CTA
{x of the first tile staged (warp 0), last own strip's loop end (warp 0), all 16 warps' partial sums in (epilogue warp), last own strip written}; prints min / median / max over
CTAs relative to node 0's first stamp.  Usage: python tools/trace_chain.py [blocks=4] [skip_nodes=14]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bitorch_engine_b200  # noqa: E402,F401
from bitorch_engine_b200 import _cabi  # noqa: E402
from bitorch_engine_b200.decode_chain import DecodeChain  # noqa: E402
from bitorch_engine_b200.extensions import q_linear_cuda  # noqa: E402
import bench  # noqa: E402

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
cfg = dict(bench.LLAMA7B, layers=blocks)
bench.LLAMA7B.update(layers=blocks)
layers = bench.build_model(dev, 1)
x = torch.randn((1, cfg["hidden"]), device=dev).half()


def token_pass():
    hid = x
    for li in range(blocks):
        lq, lk, lv, lo, lg, lu, ld = layers[li * 7:(li + 1) * 7]
        f = lambda xx, l: q_linear_cuda.mpq_forward(xx, l[3], l[4], l[5], l[6], 16, 4, False)
        q, k, v = f(hid, lq), f(hid, lk), f(hid, lv)
        o = f(v, lo)
        g, u = f(o, lg), f(o, lu)
        hid = f(u, ld)
    return hid


chain = DecodeChain.capture(token_pass)
grid = chain.grid
for _ in range(3):
    chain.launch()
chain.check()
tr = torch.zeros((grid, 32, 8), dtype=torch.int64, device=dev)
_cabi.check(_cabi.lib().b200bit_set_trace_buffer(tr.data_ptr()))
chain.launch()
chain.check()
_cabi.check(_cabi.lib().b200bit_set_trace_buffer(None))
t = tr.cpu().numpy().astype(np.int64)
n = min(32, len(chain.nodes))
t0 = t[:, 0, 0][t[:, 0, 0] > 0].min()
names = ["q", "k", "v", "o", "gate", "up", "down"]
print(f"grid {grid} ring slots {chain.ring_slots}; ns relative to node 0 (min/median/max over CTAs)")
prev_written = None
for i in range(n):
    row = []
    for s in range(8):
        v = t[:, i, s]
        v = v[v > 0] - t0
        row.append((int(v.min()), int(np.median(v)), int(v.max())) if v.size else (0, 0, 0))
    K, N = chain.nodes[i][5], chain.nodes[i][6]
    extra = f" | period {row[3][2] - prev_written}" if prev_written is not None else ""
    prev_written = row[3][2]
    print(f"node {i:3d} {names[i % 7]:>4} {K}x{N}: enter {row[6]} xhere {row[7]} staged {row[1]} deppoll {row[4]} pair0 {row[5]} loopend(w0) {row[2]} partials {row[0]} written {row[3]}{extra}")
