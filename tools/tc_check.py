"""Parity + timing of the tcgen05 batched kernel (b200bit_mpq_forward_tc) against dequantise + cuBLAS (the reference's
large-batch path) on the Llama shapes.  python tools/tc_check.py [quick]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bitorch_engine_b200  # noqa
from bitorch_engine_b200 import _cabi
from bitorch_engine_b200.extensions import q_linear_cuda
from helpers import make_mpq_inputs

lib = _cabi.lib()
WS = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")
quick = len(sys.argv) > 1


def tc(x, inp, asym):
    M, K = x.shape
    N = inp["qweight"].shape[1]
    y = torch.empty((M, N), dtype=torch.float16, device=x.device)
    _cabi.check(lib.b200bit_mpq_forward_tc(x.data_ptr(), inp["qweight"].data_ptr(), inp["scales"].data_ptr(), inp["zeros"].data_ptr(),
                                           y.data_ptr(), M, K, N, inp["scales"].shape[0], 4, int(asym), _cabi.F16,
                                           WS.data_ptr(), WS.numel(), torch.cuda.current_stream().cuda_stream))
    return y


def t_us(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


shapes = [(4096, 4096), (4096, 11008), (11008, 4096)] + ([] if quick else [(4096, 14336), (14336, 4096), (4096, 1024)])
for K, N in shapes:
    for group, asym in ((128, False), (128, True), (32, False)):
        if quick and (asym or group != 128): continue
        for M in (16, 32, 33, 64, 128, 512, 2048):
            inp = make_mpq_inputs(K, N, 4, group, "f16", asym, M=M, seed=K + N + M, device="cuda")
            x = inp["x"]
            W = q_linear_cuda.mpq_dequant(inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 4, asym, fused=True)
            ref = (x.double() @ W.double())
            try:
                y = tc(x, inp, asym)
            except NotImplementedError as e:
                print(json.dumps({"K": K, "N": N, "M": M, "group": group, "skipped": str(e)[:90]}), flush=True)
                continue
            torch.cuda.synchronize()
            err = float((y.double() - ref).norm() / ref.norm())
            maxabs = float((y.double() - ref).abs().max())
            row = {"K": K, "N": N, "M": M, "group": group, "asym": asym, "rel_err_vs_fp16W_fp64": err, "max_abs": maxabs}
            if group == 128 and not asym:
                row["tc_us"] = round(t_us(lambda: tc(x, inp, asym)), 2)
                row["dequant_cublas_us"] = round(t_us(lambda: torch.matmul(x, q_linear_cuda.mpq_dequant(inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 4, asym))), 2)
                row["cublas_only_us"] = round(t_us(lambda: torch.matmul(x, W)), 2)
                row["tc_TFLOPs"] = round(2 * M * K * N / row["tc_us"] / 1e6, 1)
                if M <= 32:
                    row["small_batch_kernel_us"] = round(t_us(lambda: q_linear_cuda.mpq_forward(x, inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, asym)), 2)
            print(json.dumps(row), flush=True)
            assert err < 2e-3, "tc kernel parity"
