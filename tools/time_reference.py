#!/usr/bin/env python
"""Times the reference's own CUDA path (oracle/_ref/q_linear_cuda, compiled unmodified from /root/reference for sm_100a)
next to ours on the same inputs, same harness: eager calls back to back over a weight pool larger than L2, CUDA events.
The reference launches on the legacy default stream (mpq_linear_cuda_kernel.cu:563), so it cannot be stream-captured; both
are therefore timed eagerly here (ours additionally graph-captured, which is how bench.py runs it)."""
import importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import algorithmic_bytes, hbm_peak
from bitorch_engine_b200.extensions import q_linear_cuda

spec = importlib.util.spec_from_file_location("_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
br = importlib.util.module_from_spec(spec); spec.loader.exec_module(br)
ref = br.load_ref("q_linear_cuda")
peak, _ = hbm_peak()
dev = torch.device("cuda:0")
for shp in ("4096x4096", "4096x11008", "11008x4096"):
    K, N = (int(v) for v in shp.split("x"))
    pool = max(4, int(400e6 // (K * N // 2)) + 1)
    g = torch.Generator(device=dev).manual_seed(0)
    ws = []
    for i in range(pool):
        qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), dtype=torch.int32, device=dev, generator=g)
        sc = (torch.rand((K // 128, N), device=dev, generator=g) * 0.01 + 0.005).half()
        ws.append((qw, sc, (sc.float() * 8).half()))
    gi = torch.arange(K, dtype=torch.int32, device=dev) // 128
    x = torch.randn((1, K), device=dev, generator=g).half()
    nbytes = algorithmic_bytes(K, N)
    def timeit(fn, reps=5):
        for i in range(pool): fn(*ws[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for i in range(pool): fn(*ws[i])
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * pool)
    res = {"shape": shp}
    if ref is not None:
        us = timeit(lambda qw, sc, zr: ref.mpq_forward(x, qw, sc, zr, gi, 16, 4, False))
        res["reference_cuda_eager_us"] = round(us, 2); res["reference_GBs"] = round(nbytes / us / 1e3, 1)
        y_ref = ref.mpq_forward(x, *ws[0], gi, 16, 4, False)
    us = timeit(lambda qw, sc, zr: q_linear_cuda.mpq_forward(x, qw, sc, zr, gi, 16, 4, False))
    res["ours_eager_us"] = round(us, 2)
    y = q_linear_cuda.mpq_forward(x, *ws[0], gi, 16, 4, False)
    if ref is not None:
        res["rel_diff_vs_reference_kernel"] = float((y.float() - y_ref.float()).norm() / y_ref.float().norm())
        W = q_linear_cuda.mpq_dequant(*ws[0], gi, 4, False)
        exact = x.float() @ W.float()
        res["ours_rel_err_vs_fp32"] = float((y.float() - exact).norm() / exact.norm())
        res["reference_rel_err_vs_fp32"] = float((y_ref.float() - exact).norm() / exact.norm())
    print(json.dumps(res), flush=True)
    del ws; torch.cuda.empty_cache()
