#!/usr/bin/env python
"""Per-shape sweep of the decode GEMV tuning knobs (lanes per row, warps per CTA, split-K, PDL) on one GPU.
Each configuration: a CUDA graph of `calls` GEMVs over a pool of distinct weights larger than L2, timed with CUDA
events; prints one JSON line per configuration (us per call, algorithmic GB/s, fraction of the measured HBM peak)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import algorithmic_bytes, hbm_peak  # noqa: E402
from bitorch_engine_b200 import _cabi  # noqa: E402
from bitorch_engine_b200.extensions import q_linear_cuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4096x4096,4096x11008,11008x4096")
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--group", type=int, default=128)
    ap.add_argument("--M", type=int, default=1)
    ap.add_argument("--configs", default="0:0:0,8:8:1,8:8:2,8:8:4,8:4:1,8:4:2,8:4:4,8:16:1,8:16:2,16:8:2,16:8:4,"
                                         "16:4:4,32:8:4,32:8:8,32:4:8,16:16:2,8:2:4,8:2:8")
    ap.add_argument("--pdl", default="0,1")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 CUDA-core gemv, 2 mma small-batch, 4 TMA stream")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    peak, _ = hbm_peak()
    _cabi.check(lib.b200bit_set_path(args.path, 1))
    for shp in args.shapes.split(","):
        K, N = (int(v) for v in shp.split("x"))
        per = K * N * args.bits // 8
        pool = max(4, int(400e6 // per) + 1)
        g = torch.Generator(device=dev).manual_seed(0)
        ws = []
        for i in range(pool):
            qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K * args.bits // 32, N), dtype=torch.int32, device=dev, generator=g)
            sc = (torch.rand((K // args.group, N), device=dev, generator=g) * 0.01 + 0.005).half()
            zr = (sc.float() * 8).half()
            ws.append((qw, sc, zr))
        gi = torch.arange(K, dtype=torch.int32, device=dev) // args.group
        x = torch.randn((args.M, K), device=dev, generator=g).half()
        calls = pool * 2
        nbytes = algorithmic_bytes(K, N, args.M, args.bits, args.group)
        for cfg in args.configs.split(","):
            L, wp, sk = (int(v) for v in cfg.split(":"))
            for pdl in (int(v) for v in args.pdl.split(",")):
                try:
                    _cabi.check(lib.b200bit_set_gemv_tuning(L, wp, sk))
                    stream = torch.cuda.Stream()
                    with torch.cuda.stream(stream):
                        def run():
                            for i in range(calls):
                                qw, sc, zr = ws[i % pool]
                                q_linear_cuda.mpq_forward(x, qw, sc, zr, gi, 16, args.bits, False, pdl=bool(pdl))
                        run()
                        stream.synchronize()
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph, stream=stream):
                            run()
                        for _ in range(3):
                            graph.replay()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        for _ in range(args.reps):
                            graph.replay()
                        e1.record(stream)
                        stream.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / (args.reps * calls)
                    gbs = nbytes / us / 1e3
                    print(json.dumps({"shape": shp, "bits": args.bits, "M": args.M, "path": args.path, "L": L, "warps": wp, "splitk": sk,
                                      "pdl": pdl, "us": round(us, 3), "GBs": round(gbs, 1),
                                      "frac": round(gbs / peak, 3)}), flush=True)
                except Exception as e:  # keep sweeping
                    print(json.dumps({"shape": shp, "cfg": cfg, "pdl": pdl, "error": str(e)[:200]}), flush=True)
        del ws
        torch.cuda.empty_cache()
    lib.b200bit_set_gemv_tuning(0, 0, 0)


if __name__ == "__main__":
    main()
