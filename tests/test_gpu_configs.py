"""GPU parity at the sizes BASELINE.json's configs name and round 1 never ran (VERDICT r1 "weak" 1):
  #3  2-bit g32 and 4-bit g128 at bs = 32 on the three Llama-7B shapes
  #5  Llama-3-8B shapes (4096->1024, 4096->14336, 14336->4096) at M = 1 and M = 512 (prefill)
      M in {512, 2048} forward and grad_input (sft training shape)
  #4  the binary Linear at the real conv-as-GEMM row counts (401408 x 576 x 64 ... 6272 x 4608 x 512) and the fc shape
The oracle's dequantised weight (oracle/nbit.py dequant_w16, pinned to the reference's unpack_qweight) is the reference
weight; the big matmuls of the CHECK run in float64 on the GPU (numpy would take minutes at M = 2048).  fp tolerance as
stated in tests/helpers.py; binary: integer-exact."""
import numpy as np
import pytest
import torch

from oracle import nbit
from helpers import make_mpq_inputs, to_np_f32, NORMWISE_TOL

pytestmark = pytest.mark.gpu


def _ref_weight(inp, w_bit, asym, dt="f16"):
    zeros = inp["zeros"].cpu().numpy() if asym else to_np_f32(inp["zeros"])
    W = nbit.dequant_w16(inp["qweight"].cpu().numpy(), to_np_f32(inp["scales"]), zeros, None, w_bit, asym, dt)
    return torch.from_numpy(np.asarray(W, dtype=np.float64)).cuda()


def _check(y, ref, what, tol=NORMWISE_TOL["f16"]):
    err = float((y.double() - ref).norm() / ref.norm())
    assert err <= tol, f"{what}: normwise rel err {err:.3e} > {tol}"
    # element-wise: fp16 output rounding + the fp16-accumulating dense GEMM of the large-M path
    bad = ((y.double() - ref).abs() > 2e-3 * ref.abs() + 2e-3 * ref.abs().mean()).float().mean().item()
    assert bad <= 1e-4, f"{what}: {bad:.2e} of the outputs outside 2e-3"


@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 11008), (11008, 4096)])
@pytest.mark.parametrize("w_bit,group", [(4, 128), (2, 32)])
def test_config3_bs32_llama7b(K, N, w_bit, group):
    from bitorch_engine_b200.extensions import q_linear_cuda
    inp = make_mpq_inputs(K, N, w_bit, group, "f16", False, M=32, seed=K + N + w_bit, device="cuda")
    y = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, w_bit, False)
    _check(y, inp["x"].double() @ _ref_weight(inp, w_bit, False), f"bs32 {K}x{N} w{w_bit}g{group}")


@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 1024), (4096, 14336), (14336, 4096)])
@pytest.mark.parametrize("M", [1, 512])
def test_config5_llama3_8b_shapes(K, N, M):
    from bitorch_engine_b200.extensions import q_linear_cuda
    inp = make_mpq_inputs(K, N, 4, 128, "f16", False, M=M, seed=K + N + M, device="cuda")
    y = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, False,
                                  pdl=(M == 1))
    _check(y, inp["x"].double() @ _ref_weight(inp, 4, False), f"llama3 {K}x{N} M={M}")


@pytest.mark.parametrize("K,N,M", [(4096, 4096, 512), (4096, 4096, 2048), (4096, 11008, 2048)])
@pytest.mark.parametrize("asym", [False, True])
def test_training_shapes_forward_and_grad_input(K, N, M, asym):
    from bitorch_engine_b200.extensions import q_linear_cuda
    inp = make_mpq_inputs(K, N, 4, 128, "f16", asym, M=M, seed=K + N + M + int(asym), device="cuda")
    W = _ref_weight(inp, 4, asym)
    y = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, asym)
    _check(y, inp["x"].double() @ W, f"forward {K}x{N} M={M} asym={asym}")
    dy = torch.randn((M, N), device="cuda").half()
    dx = q_linear_cuda.mpq_grad_input(inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], dy, 16, 4, asym)
    assert tuple(dx.shape) == (M, K) and dx.dtype == torch.float16
    _check(dx, dy.double() @ W.t(), f"grad_input {K}x{N} M={M} asym={asym}")


CONV_AS_GEMM = [(128, 512, 1000), (401408, 576, 64), (100352, 1152, 128), (25088, 2304, 256), (6272, 4608, 512)]


@pytest.mark.parametrize("M,K,N", CONV_AS_GEMM)
def test_config4_binary_at_real_row_counts(M, K, N):
    """integer-exact: y = sign(x) @ sign(w).T; +-1 sums up to 4608 are exact in fp32, so a float GEMM is the checker"""
    from bitorch_engine_b200.extensions import binary_linear_cuda
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn((M, K), device="cuda", generator=g)
    w = torch.randn((N, K), device="cuda", generator=g)
    y = binary_linear_cuda.forward(x, w, 3, True)
    ref = torch.where(x >= 0, 1.0, -1.0) @ torch.where(w >= 0, 1.0, -1.0).t()
    assert y.dtype == torch.float32 and torch.equal(y, ref)
    packed = binary_linear_cuda.w_pack(w, 3, True)
    assert torch.equal(binary_linear_cuda.forward(x, packed, 3, True), ref)
    assert torch.equal(binary_linear_cuda.forward(x, packed, 3, True), ref)        # second call: cached canonical layout
    packed.bitwise_not_()                                                           # in-place write invalidates the cache
    assert torch.equal(binary_linear_cuda.forward(x, packed, 3, True), -ref)
