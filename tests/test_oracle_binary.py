"""CPU: pin the numpy binary oracle against the reference's own CPU extension compiled unmodified from /root/reference
(oracle/_ref/binary_linear_cpp_ref.so, built by oracle/build_ref.py).  Exact equality, as in the
reference's tests (tests/layers/test_binary_linear.py:25-66, 271-324)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import binary


def _ref_cpp():
    spec = importlib.util.spec_from_file_location("_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
    br = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(br)
    return br.load_ref("binary_linear_cpp")


REF = _ref_cpp()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref/binary_linear_cpp not built (python oracle/build_ref.py cpu)")


@needs_ref
@pytest.mark.parametrize("M,K,N", [(128, 512, 1000), (1, 8, 1), (7, 40, 13), (32, 576, 64)])
def test_forward_equals_reference_cpu_extension(M, K, N):
    g = torch.Generator().manual_seed(M * K + N)
    x = torch.randn((M, K), generator=g)
    w = torch.randn((N, K), generator=g)
    x[0, 0] = 0.0
    ref = REF.forward(x, w, M, N, K)
    assert np.array_equal(ref.numpy().astype(np.int64), binary.forward(x.numpy(), w.numpy()))


@needs_ref
@pytest.mark.parametrize("K,N", [(512, 1000), (64, 8), (128, 24)])
def test_cpu_weight_packing_layout(K, N):
    g = torch.Generator().manual_seed(K + N)
    w = torch.randn((N, K), generator=g)
    packed = REF.w_pack(w, N, K)
    assert np.array_equal(packed.numpy().reshape(-1), binary.pack_cpp(w.numpy()))
    x = torch.randn((16, K), generator=g)
    assert torch.equal(REF.forward(x, packed, 16, N, K), REF.forward(x, w, 16, N, K))   # packed == unpacked (:222-268)


def test_cuda_layout_restatements_are_permutations():
    rng = np.random.default_rng(0)
    for (N, K, layout) in [(8, 128, 2), (64, 256, 2), (32, 64, 1), (96, 160, 1)]:
        w = rng.standard_normal((N, K))
        packed = binary.pack_cuda(w, layout)
        assert packed.size == N * K // 8
        assert sorted(packed.tolist()) == sorted(binary.canonical_bits(w).reshape(-1).tolist())
