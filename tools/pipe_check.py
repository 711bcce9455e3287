#!/usr/bin/env python
"""Quick parity check of the pipelined decode GEMV (path 6) against the numpy oracle on a spread of configurations
(run this first on a new box: every case runs under the caller's timeout)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_mpq_inputs, to_np_f32, assert_close_to_oracles
from oracle import nbit
from bitorch_engine_b200 import _cabi
from bitorch_engine_b200.extensions import q_linear_cuda

lib = _cabi.lib()
PATH = int(sys.argv[1]) if len(sys.argv) > 1 else 7      # 7: integer-tensor-pipe kernel, 6: fp16-subnormal pipe kernels
lib.b200bit_set_path(PATH, 0)
cases = [(256, 64, 4, 128, "f16", False, (0, 0, 0)), (4096, 4096, 4, 128, "f16", False, (0, 0, 0)),
         (4096, 11008, 4, 128, "f16", False, (0, 0, 0)), (11008, 4096, 4, 128, "f16", False, (0, 0, 0)),
         (11008, 4096, 4, 128, "f16", False, (0, 0, 1)), (11008, 4096, 4, 128, "f16", False, (0, 2, 2)),
         (4096, 4096, 4, 128, "f16", False, (0, 2, 1)), (4096, 4096, 4, 128, "f16", False, (0, 1, 4)),
         (2048, 1024, 4, 32, "f16", False, (0, 0, 0)), (2048, 1024, 4, 64, "f16", True, (0, 0, 0)),
         (2048, 1024, 4, 128, "bf16", False, (0, 0, 0)), (2048, 1024, 4, 128, "bf16", True, (0, 0, 0)),
         (2048, 1024, 2, 128, "f16", False, (0, 0, 0)), (2048, 1024, 2, 64, "bf16", False, (0, 0, 0)),
         (2048, 1024, 8, 128, "f16", False, (0, 0, 0)), (2048, 1024, 8, 128, "f16", True, (0, 0, 0)),
         (2048, 1024, 4, 1024, "f16", False, (0, 0, 0)), (2048, 1024, 4, 256, "f16", True, (0, 0, 0)),
         (2176, 96, 4, 128, "f16", False, (0, 0, 0)),
         # CUDA-core FHFMA flavour forced (L = 32) where the default would pick the mma.sync flavour
         (256, 64, 4, 128, "f16", False, (32, 0, 0)), (4096, 4096, 4, 128, "f16", False, (32, 0, 0)),
         (11008, 4096, 4, 128, "f16", True, (32, 0, 0)), (2048, 1024, 4, 64, "f16", True, (32, 0, 0)),
         (4096, 4096, 4, 128, "f16", True, (0, 0, 0)), (4096, 4096, 4, 128, "f16", False, (0, 0, 2)),
         (2048, 1024, 4, 2048, "f16", False, (0, 0, 0)), (4096, 4096, 4, 32, "f16", False, (0, 0, 0)),
         (4096, 4096, 4, 64, "bf16", True, (0, 0, 0)), (11008, 4096, 4, 128, "bf16", False, (0, 0, 0)),
         (4096, 11008, 4, 128, "f16", True, (0, 0, 0)), (4096, 4096, 4, 4096, "f16", False, (0, 0, 0)),
         (4096, 11008, 4, 128, "f16", False, (0, 1, 0)), (11008, 4096, 4, 128, "f16", False, (0, 2, 0)),
         (8192, 1000, 4, 256, "f16", False, (0, 0, 0)), (14336, 4096, 4, 128, "f16", False, (0, 0, 0))]
bad = 0
for K, N, b, g, dt, asym, tune in cases:
    lib.b200bit_set_gemv_tuning(*tune)
    inp = make_mpq_inputs(K, N, b, g, dt, asym, M=1, seed=K + N + b + g, device="cuda")
    for pdl in (False, True):
        ys = [q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, b, asym, pdl=pdl)
              for _ in range(3)]
        torch.cuda.synchronize()
        zeros = inp["zeros"].cpu().numpy() if asym else to_np_f32(inp["zeros"])
        args = (to_np_f32(inp["x"]), inp["qweight"].cpu().numpy(), to_np_f32(inp["scales"]), zeros, None, b, asym)
        try:
            assert torch.equal(ys[0], ys[1]) and torch.equal(ys[0], ys[2]), "non-deterministic"
            nrm = assert_close_to_oracles(to_np_f32(ys[0]), nbit.mpq_forward(*args, dt), nbit.mpq_forward_exact(*args), dt, "pipe")
            print(f"ok   K={K} N={N} b={b} g={g} {dt} asym={asym} tune={tune} pdl={pdl} normwise={nrm:.2e}", flush=True)
        except AssertionError as e:
            bad += 1
            print(f"FAIL K={K} N={N} b={b} g={g} {dt} asym={asym} tune={tune} pdl={pdl}: {str(e)[:300]}", flush=True)
lib.b200bit_set_gemv_tuning(0, 0, 0); lib.b200bit_set_path(0, 0)
print("pipe_check:", "ALL OK" if not bad else f"{bad} FAILED")
sys.exit(1 if bad else 0)
