"""Quick parity + timing check of one forward path (default 5 = tcgen05 kernel) against the numpy oracle (GPU box only)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bitorch_engine_b200 import _cabi
from bitorch_engine_b200.extensions import q_linear_cuda
from oracle import nbit
from helpers import make_mpq_inputs, to_np_f32

ap = argparse.ArgumentParser(); ap.add_argument("--path", type=int, default=5); ap.add_argument("--quick", action="store_true"); ap.add_argument("--skip", type=int, default=0)
args = ap.parse_args()
lib = _cabi.lib()
cases = [(1024, 512, 128, False, 1), (4096, 4096, 128, False, 1), (4096, 4096, 128, True, 1), (4096, 4096, 64, False, 3),
         (4096, 11008, 128, False, 1), (11008, 4096, 128, False, 1), (11008, 4096, 128, True, 4), (2048, 1024, 1024, False, 2),
         (2048, 1024, 256, True, 1), (4096, 4096, 128, False, 7), (1280, 96, 128, False, 1), (4096, 4096, 128, False, 16),
         (4096, 4096, 128, True, 32), (11008, 4096, 128, False, 32), (4096, 11008, 64, False, 13), (2048, 1024, 512, False, 70)]
if args.quick: cases = cases[:2]
cases = cases[args.skip:]
for K, N, g, asym, M in cases:
    inp = make_mpq_inputs(K, N, 4, g, "f16", asym, M=M, seed=K + N + M, device="cuda")
    _cabi.check(lib.b200bit_set_path(args.path, 1))
    y = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, asym)
    torch.cuda.synchronize()
    _cabi.check(lib.b200bit_set_path(0, 1))
    zeros = inp["zeros"].cpu().numpy() if asym else to_np_f32(inp["zeros"])
    ye = nbit.mpq_forward_exact(to_np_f32(inp["x"]), inp["qweight"].cpu().numpy(), to_np_f32(inp["scales"]), zeros, None, 4, asym)
    yn = to_np_f32(y)
    rel = float(np.linalg.norm(yn - ye) / (np.linalg.norm(ye) + 1e-30))
    print(json.dumps({"K": K, "N": N, "g": g, "asym": asym, "M": M, "rel_err": rel, "max_abs": float(np.abs(yn - ye).max()),
                      "finite": bool(np.isfinite(yn).all())}), flush=True)
