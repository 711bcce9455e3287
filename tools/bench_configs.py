#!/usr/bin/env python
"""Measures every kernel of the path that bench.py's headline does not cover and writes profiles/configs_r02.json:
BASELINE configs #3 (2-bit / 4-bit / exl2 mixed at bs = 32), #4 (binary, ResNet-18 fc + conv-as-GEMM shapes, bs = 128),
#5 (Llama-3-8B shapes, decode M = 1 and prefill M = 512), the format kernels (dequant, pack), grad_input, both fused
optimizer kernels and the functions_cuda streams -- each as {us, algorithmic bytes, GB/s, fraction of the measured HBM
peak} -- with the reference's own CUDA extensions (oracle/_ref, compiled unmodified for sm_100a) and its CPU extension
timed beside them where they exist.  CUDA events, weight pools larger than L2 for the n-bit kernels, eager launches
(the reference launches on the legacy default stream and cannot be graph-captured).

    python tools/bench_configs.py [--quick] [--out profiles/configs_r02.json]
"""
import argparse
import importlib.util
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import hbm_peak  # noqa: E402
from bitorch_engine_b200.extensions import q_linear_cuda, binary_linear_cuda, binary_linear_cpp, functions_cuda  # noqa: E402

spec = importlib.util.spec_from_file_location("_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
br = importlib.util.module_from_spec(spec)
spec.loader.exec_module(br)
PEAK, PEAK_SRC = hbm_peak()
DEV = torch.device("cuda:0")
RESULTS = []


def ev_time(fn, pool, reps):
    """mean microseconds per call of fn(i), i cycling through `pool` distinct argument sets"""
    for i in range(pool):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(pool):
            fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * pool)


def graph_time(fn, pool, reps):
    """same, with the `pool` calls captured once into a CUDA graph: device time per call without the Python / launch
    overhead of the shim (what a graph-capturing caller sees)"""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(pool):
            fn(i)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(pool):
                fn(i)
        g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            g.replay()
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * pool)


def record(name, us, nbytes=None, **extra):
    row = {"name": name, "us": round(us, 3)}
    if nbytes:
        row.update(bytes=int(nbytes), GBs=round(nbytes / us / 1e3, 1), frac_of_hbm_peak=round(nbytes / us / 1e3 / PEAK, 4))
    row.update(extra)
    RESULTS.append(row)
    print(json.dumps(row), flush=True)


def mpq_bytes(K, N, M, w_bit, group):
    return K * N * w_bit // 8 + 2 * (K // group) * N * 2 + 2 * M * K + 2 * M * N


def make_pool(K, N, w_bit, group, budget=300e6):
    n = max(3, int(budget // (K * N * w_bit // 8)) + 1)
    g = torch.Generator(device=DEV).manual_seed(K + N + w_bit)
    nb = 32 // w_bit
    out = []
    for _ in range(n):
        qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // nb, N), dtype=torch.int32, device=DEV, generator=g)
        sc = (torch.rand((K // group, N), device=DEV, generator=g) * 0.01 + 0.005).half()
        out.append((qw, sc, (sc.float() * 2 ** (w_bit - 1)).half()))
    return out


def bench_nbit(quick):
    ref = br.load_ref("q_linear_cuda")
    shapes7 = [(4096, 4096), (4096, 11008), (11008, 4096)]
    shapes3 = [(4096, 1024), (4096, 14336), (14336, 4096)]
    for (K, N) in shapes7 + shapes3:
        for w_bit, group in ((4, 128), (2, 32)):
            if (K, N) in shapes3 and w_bit == 2:
                continue
            pool = make_pool(K, N, w_bit, group)
            gi = torch.arange(K, dtype=torch.int32, device=DEV) // group
            for M in (1, 32, 512):
                if M == 512 and (w_bit == 2 or quick):
                    continue
                x = torch.randn((M, K), device=DEV).half()
                f = lambda i: q_linear_cuda.mpq_forward(x, *pool[i], gi, 16, w_bit, False, pdl=(M == 1))
                us = ev_time(f, len(pool), 3)
                record(f"mpq_forward w{w_bit}g{group} {K}x{N} M={M} (eager, per call)", us, mpq_bytes(K, N, M, w_bit, group),
                       flops=2 * M * K * N, TFLOPs=round(2 * M * K * N / us / 1e6, 1))
                us = graph_time(f, len(pool), 5)
                record(f"mpq_forward w{w_bit}g{group} {K}x{N} M={M} (CUDA graph, per call)", us, mpq_bytes(K, N, M, w_bit, group),
                       flops=2 * M * K * N, TFLOPs=round(2 * M * K * N / us / 1e6, 1))
                if ref is not None and (K, N) in shapes7 and M <= 32 and w_bit == 4:
                    try:
                        us_r = ev_time(lambda i: ref.mpq_forward(x, *pool[i], gi, 16, w_bit, False), len(pool), 1)
                        record(f"REFERENCE q_linear_cuda.mpq_forward w{w_bit}g{group} {K}x{N} M={M}", us_r, mpq_bytes(K, N, M, w_bit, group))
                    except Exception as e:      # noqa: BLE001
                        record(f"REFERENCE mpq_forward {K}x{N} M={M} failed: {str(e)[:80]}", 0.0)
                if M == 512:
                    Wd = q_linear_cuda.mpq_dequant(*pool[0], gi, w_bit, False)
                    us_c = ev_time(lambda i: torch.matmul(x, Wd), 1, 20)
                    record(f"cuBLAS fp16 GEMM alone {K}x{N} M={M}", us_c, flops=2 * M * K * N, TFLOPs=round(2 * M * K * N / us_c / 1e6, 1))
            if (K, N) in shapes7:
                # format kernels + grad_input on the same pool
                us = ev_time(lambda i: q_linear_cuda.mpq_dequant(*pool[i], gi, w_bit, False), len(pool), 2)
                record(f"mpq_dequant w{w_bit}g{group} {K}x{N}", us, K * N * w_bit // 8 + 2 * K * N + 4 * (K // group) * N)
                W = q_linear_cuda.mpq_dequant(*pool[0], gi, w_bit, False)
                us = ev_time(lambda i: q_linear_cuda.mpq_pack_weight(W, pool[i][1], pool[i][2], gi, w_bit, False), len(pool), 2)
                record(f"mpq_pack_weight w{w_bit}g{group} {K}x{N}", us, K * N * w_bit // 8 + 2 * K * N + 4 * (K // group) * N)
                del W
                if w_bit == 4:
                    for M in ((32,) if quick else (32, 2048)):
                        dy = torch.randn((M, N), device=DEV).half()
                        us = ev_time(lambda i: q_linear_cuda.mpq_grad_input(*pool[i], gi, dy, 16, w_bit, False), min(3, len(pool)), 1)
                        record(f"mpq_grad_input w{w_bit}g{group} {K}x{N} M={M}", us, mpq_bytes(K, N, M, w_bit, group),
                               flops=2 * M * K * N, TFLOPs=round(2 * M * K * N / us / 1e6, 2))
                        if M == 2048:
                            Wd = q_linear_cuda.mpq_dequant(*pool[0], gi, w_bit, False)
                            us_c = ev_time(lambda i: torch.matmul(dy, q_linear_cuda.mpq_dequant(*pool[i], gi, w_bit, False).t()), min(3, len(pool)), 2)
                            record(f"dequant + cuBLAS grad_input {K}x{N} M={M}", us_c, flops=2 * M * K * N, TFLOPs=round(2 * M * K * N / us_c / 1e6, 1))
                            del Wd
            del pool
            torch.cuda.empty_cache()


def bench_exl2(quick):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_mbwq import _make_exl2, STRATEGIES
    ref = br.load_ref("q_linear_cuda")
    for (K, N) in [(4096, 4096), (4096, 11008), (11008, 4096)]:
        layers = [_make_exl2(K, N, STRATEGIES[0], seed=s, permute=True) for s in range(8 if not quick else 3)]
        wbytes = layers[0].qweight.numel() * 4 + 4 * layers[0].scales.numel()
        for M in (1, 32):
            x = torch.randn((M, K), device=DEV).half()
            f = lambda i: q_linear_cuda.mbwq_exl2_forward(x, layers[i].qweight.data, layers[i].scales, layers[i].zeros, layers[i].q_perm,
                                                          layers[i].q_group_map, layers[i].rows)
            us = ev_time(f, len(layers), 5)
            record(f"mbwq_exl2_forward fused 4b x0.75 + 2b x0.25 g32 {K}x{N} M={M}", us, wbytes + 2 * M * (K + N))
            record(f"mbwq_exl2_forward fused {K}x{N} M={M} (CUDA graph, per call)", graph_time(f, len(layers), 5), wbytes + 2 * M * (K + N))
            fc = lambda i: q_linear_cuda.mbwq_exl2_forward(x, layers[i].qweight.data, layers[i].scales, layers[i].zeros, layers[i].q_perm,
                                                           layers[i].q_group_map, layers[i].rows, use_cublas=True)
            record(f"exl2 dequant + cuBLAS (round-1 path) {K}x{N} M={M}", ev_time(fc, len(layers), 5), wbytes + 2 * M * (K + N))
            if ref is not None:
                fr = lambda i: ref.mbwq_exl2_forward(x, layers[i].qweight.data, layers[i].scales, layers[i].zeros, layers[i].q_perm,
                                                     layers[i].q_group_map, layers[i].rows[:7], False)
                try:
                    record(f"REFERENCE mbwq_exl2_forward {K}x{N} M={M}", ev_time(fr, len(layers), 2), wbytes + 2 * M * (K + N))
                except Exception as e:      # noqa: BLE001
                    record(f"REFERENCE mbwq_exl2_forward {K}x{N} M={M} failed: {str(e)[:80]}", 0.0)
        del layers
        torch.cuda.empty_cache()


def bench_binary(quick):
    ref = br.load_ref("binary_linear_cuda")
    refc = br.load_ref("binary_linear_cpp")
    shapes = [(128, 512, 1000), (401408, 576, 64), (100352, 1152, 128), (25088, 2304, 256), (6272, 4608, 512)]
    for (M, K, N) in shapes:
        x = torch.randn((M, K), device=DEV)
        w = torch.randn((N, K), device=DEV)
        packed = binary_linear_cuda.w_pack(w, 3, True)
        nbytes = M * K * 4 + K * N // 8 + M * N * 4
        reps = 20 if M < 10000 else 3
        us = ev_time(lambda i: binary_linear_cuda.forward(x, packed, 3, True), 1, reps)
        record(f"binary_linear_cuda.forward packed w, f32 {M}x{K}x{N}", us, nbytes, GOPs=round(2 * M * K * N / us / 1e3, 1))
        if ref is not None:
            try:
                pr = ref.w_pack(w, 3, True)
                us_r = ev_time(lambda i: ref.forward(x, pr, 3, True), 1, reps)
                record(f"REFERENCE binary_linear_cuda.forward {M}x{K}x{N}", us_r, nbytes, GOPs=round(2 * M * K * N / us_r / 1e3, 1))
            except Exception as e:      # noqa: BLE001
                record(f"REFERENCE binary forward {M}x{K}x{N} failed: {str(e)[:80]}", 0.0)
        if M <= 25088 or not quick:
            xc, wc = x.cpu(), w.cpu()
            pk = binary_linear_cpp.w_pack(wc, N, K)
            t0 = time.perf_counter()
            n_it = 3 if M > 1000 else 20
            for _ in range(n_it):
                binary_linear_cpp.forward(xc, pk, M, N, K)
            record(f"HOST binary_linear_cpp.forward (ours) {M}x{K}x{N}", (time.perf_counter() - t0) / n_it * 1e6, nbytes,
                   cores=torch.get_num_threads())
            if refc is not None and M <= 25088:
                pkr = refc.w_pack(wc, N, K)
                t0 = time.perf_counter()
                n_it = 1 if M > 1000 else 10
                for _ in range(n_it):
                    refc.forward(xc, pkr, M, N, K)
                record(f"REFERENCE HOST binary_linear_cpp.forward {M}x{K}x{N}", (time.perf_counter() - t0) / n_it * 1e6, nbytes,
                       cores=torch.get_num_threads())
        del x, w
        torch.cuda.empty_cache()


def bench_optim(quick):
    from bitorch_engine_b200.layers.qlinear.nbit import MPQWeightParameter
    from bitorch_engine_b200.layers.qlinear.binary import BinaryLinearParameter
    K, N = 4096, 11008
    for sdt, name in ((torch.float32, "f32 state"), (torch.float16, "f16 state")):
        g = torch.Generator(device=DEV).manual_seed(3)
        qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), dtype=torch.int32, device=DEV, generator=g)
        sc = (torch.rand((K // 128, N), device=DEV, generator=g) * 0.01 + 0.005).half()
        zr = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 128, N // 8), dtype=torch.int32, device=DEV, generator=g)
        gi = torch.arange(K, dtype=torch.int32, device=DEV) // 128
        qp = MPQWeightParameter(qw, requires_grad=False, scales=sc, zeros=zr, g_idx=gi, w_bit=4, asym=True, group_size=128, layer_type=1)
        m = torch.zeros((K, N), dtype=sdt, device=DEV)
        v = torch.zeros((K, N), dtype=sdt, device=DEV)
        grad = (torch.randn((K, N), device=DEV) * 0.05).half()
        step = torch.zeros(1)
        f = lambda i: MPQWeightParameter.update(qp, exp_avg_s=v, exp_avg_l=m, step=step, lr=1e-3, beta1=0.99, beta2=0.9999, eps=1e-6,
                                                dtype=sdt, correct_bias=True, grad=grad)
        us = ev_time(f, 1, 10)
        es = 4 if sdt == torch.float32 else 2
        record(f"diodemix_mpq_step (fused) w4g128 {K}x{N} {name}", us, K * N // 2 * 2 + K * N * 2 + 4 * K * N * es)
    w = torch.where(torch.rand((K, N), device=DEV) < 0.5, -1, 1).to(torch.int8)
    bp = BinaryLinearParameter(w, requires_grad=False)
    m = torch.zeros((K, N), device=DEV)
    v = torch.zeros((K, N), device=DEV)
    gq = torch.randint(-127, 127, (K, N), device=DEV, dtype=torch.int8)
    step = torch.zeros(1)
    f = lambda i: BinaryLinearParameter.update(bp, exp_avg_s=v, exp_avg_l=m, step=step, lr=1e-3, beta1=0.99, beta2=0.9999, dtype=torch.float32, grad=gq)
    record(f"diodemix_binary_step (fused) {K}x{N} f32 state", ev_time(f, 1, 10), K * N * (1 + 1 + 1 + 16))


def bench_functions(quick):
    n = 64 * 1024 * 1024
    codes = torch.randint(0, 16, (8192, n // 8192), device=DEV, dtype=torch.int32)
    f = lambda i: functions_cuda.q4_pack(codes)
    record("functions_cuda.q4_pack 64Mi codes", ev_time(f, 1, 5), n * 4 + n // 2)
    packed = functions_cuda.q4_pack(codes)
    record("functions_cuda.q4_unpack", ev_time(lambda i: functions_cuda.q4_unpack(packed), 1, 5), n * 4 + n // 2)
    record("functions_cuda.q4_unpack_and_scaling", ev_time(lambda i: functions_cuda.q4_unpack_and_scaling(packed, 0.5), 1, 5), n * 4 + n // 2)
    xf = torch.randn((n,), device=DEV)
    record("functions_cuda.tensor_pack_to_uint8 f32", ev_time(lambda i: functions_cuda.tensor_pack_to_uint8(xf.view(-1, 1024)), 1, 5), n * 4 + n // 8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "configs_r02.json"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    parts = {"nbit": bench_nbit, "exl2": bench_exl2, "binary": bench_binary, "optim": bench_optim, "functions": bench_functions}
    for name, fn in parts.items():
        if args.only and name not in args.only.split(","):
            continue
        try:
            fn(args.quick)
        except Exception as e:      # noqa: BLE001
            record(f"{name}: FAILED {type(e).__name__}: {str(e)[:200]}", 0.0)
    meta = {"hbm_peak_GBs": PEAK, "peak_source": PEAK_SRC, "gpu": torch.cuda.get_device_name(0), "host_threads": torch.get_num_threads(),
            "timing": "CUDA events, eager launches, weight pools > L2 for the n-bit kernels; host rows: time.perf_counter",
            "results": RESULTS}
    with open(args.out, "w") as fh:
        json.dump(meta, fh, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
