"""GPU parity for the MBWQ layer family (the reference's "Q4" GPTQ-style path and the exl2 mixed-bit path):
  * the equivalences the reference's own test pins (tests/layers/test_nbit_linear.py:361-404): kernel == x @ q42fp(W),
    pack(q42fp(W)) -> unpack == W, MBWQ-q4 == MPQ(sym, default g_idx) on the same packed tensor
  * q42fp / exl2fp dequantisation BIT-EXACT against the numpy oracle (oracle/nbit.py style="kernel", oracle/exl2.py)
  * when oracle/_ref/q_linear_cuda (the reference CUDA extension compiled unmodified for sm_100a) is present:
    bit-exact dequantisation against it and forward outputs within the stated tolerance of it."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import nbit, exl2
from helpers import make_mpq_inputs, to_np_f32, assert_close_to_oracles, exl2_packed_info, exl2_q_groups, rel_fro

pytestmark = pytest.mark.gpu


def _ref_ext(name):
    spec = importlib.util.spec_from_file_location("_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
    br = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(br)
    return br.load_ref(name)


@pytest.mark.parametrize("bits,group", [(4, 128), (4, 32), (2, 32), (2, 64)])
@pytest.mark.parametrize("permute", [False, True])
def test_q4_path(bits, group, permute):
    from bitorch_engine_b200.extensions import q_linear_cuda
    from bitorch_engine_b200.layers.qlinear.nbit import MPQWeightParameter
    from bitorch_engine_b200.layers.qlinear.nbit.cuda.utils import pack_fp_weight, unpack_qweight
    K, N, M = 1024, 512, 2
    inp = make_mpq_inputs(K, N, bits, group, "f16", False, M=M, seed=bits * 100 + group, device="cuda")
    g = torch.Generator().manual_seed(1)
    q_perm = (torch.randperm(K, generator=g) if permute else torch.arange(K)).to(torch.short).cuda()
    W = q_linear_cuda.mbwq_q42fp_weight(inp["qweight"], inp["scales"], inp["zeros"], group, bits, q_perm)
    # (1) dequantisation bit-exact vs the oracle's kernel-style (fma) rounding, rows scattered through q_perm
    Wk = nbit.dequant_w16(inp["qweight"].cpu().numpy(), to_np_f32(inp["scales"]), to_np_f32(inp["zeros"]), None, bits,
                          False, "f16", style="kernel")
    Wo = np.zeros_like(Wk)
    Wo[q_perm.cpu().numpy().astype(np.int64) & 0xFFFF] = Wk
    assert np.array_equal(to_np_f32(W), Wo)
    # (2) kernel == x @ q42fp(W)  (reference: mean-abs < 2, test_nbit_linear.py:365; ours: the stated fp tolerance)
    y = q_linear_cuda.mbwq_q4_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], group, q_perm, bits)
    y_ref = to_np_f32(inp["x"]).astype(np.float64) @ Wo.astype(np.float64)
    assert rel_fro(to_np_f32(y), y_ref) <= 1e-3
    # (3) MBWQ-q4 == MPQ(sym) on the same packed tensor when there is no permutation (:397-404)
    if not permute:
        y_mpq = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, bits, False)
        assert torch.equal(y, y_mpq)
    # (4) pack(q42fp(W)) -> unpack reproduces W (rtol/atol 0.1 in the reference, :389-395)
    qp = MPQWeightParameter(inp["qweight"].clone(), requires_grad=False, scales=inp["scales"], zeros=inp["zeros"], g_idx=None,
                            w_bit=bits, asym=False, group_size=group, layer_type=2, q_perm=q_perm)
    qp.data = pack_fp_weight(W, qp)
    W2 = unpack_qweight(qp)
    assert torch.allclose(W2.float(), W.float(), rtol=0.1, atol=0.1)


STRATEGIES = [dict(group_size={"4": 32, "2": 32}, bits=[4, 2], bits_prop=[0.75, 0.25]),
              dict(group_size={"4": 32, "2": 32}, bits=[4, 2], bits_prop=[0.25, 0.75]),
              dict(group_size={"8": 32, "6": 32, "5": 32, "4": 32, "3": 32, "2": 32}, bits=[8, 6, 5, 4, 3, 2],
                   bits_prop=[0.125, 0.125, 0.125, 0.25, 0.125, 0.25])]


def _make_exl2(K, N, strat, seed, permute=False):
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MBWQLinearCuda
    groups, rows = exl2_packed_info(K, strat["bits"], strat["bits_prop"], strat["group_size"])
    g = torch.Generator().manual_seed(seed)
    layer = MBWQLinearCuda(in_channels=K, out_channels=N, w_bit=4, dtype=torch.half, group_size=32, dq_group_size=1,
                           use_gba_quant=True, asym=False, dq_mode=2, use_mbw=True, groups=groups, rows_packed=rows,
                           requires_grad=False)
    layer.set_qweight_data(torch.randint(-2 ** 31, 2 ** 31 - 1, (rows, N), dtype=torch.int32, generator=g))
    layer.set_scales((torch.rand((groups, N), generator=g) * 0.02 + 0.005).half())
    layer.set_zeros((torch.randn((groups, N), generator=g) * 0.05).half())
    layer.q_perm = (torch.randperm(K, generator=g) if permute else torch.arange(K)).to(torch.short)
    layer.q_groups = torch.tensor(exl2_q_groups(groups, strat["bits"], strat["group_size"], K, strat["bits_prop"]),
                                  dtype=torch.short)
    layer = layer.cuda()
    layer.qweight.data = layer.qweight.data.cuda()
    layer.set_scales(layer.scales.cuda()); layer.set_zeros(layer.zeros.cuda())
    layer.prepare_params()
    return layer


@pytest.mark.parametrize("K,N", [(128, 128), (1024, 256)])
@pytest.mark.parametrize("si", range(len(STRATEGIES)))
@pytest.mark.parametrize("permute", [False, True])
def test_exl2_dequant_and_forward(K, N, si, permute):
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MBWQLinearCuda
    if si == 2 and K < 1024:
        pytest.skip("six bit-widths need more rows")
    layer = _make_exl2(K, N, STRATEGIES[si], seed=K + N + si, permute=permute)
    W = MBWQLinearCuda.exl2fp_weight(layer.qweight, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map, layer.rows)
    assert tuple(W.shape) == (K, N) and W.dtype == torch.float16
    rows_o = exl2.rows_from_q_groups(layer.q_groups.cpu().numpy(), K)
    assert list(layer.rows) == rows_o
    gm = exl2.group_map(layer.q_groups.cpu().numpy(), layer.qweight.shape[0])
    assert np.array_equal(layer.q_group_map.cpu().numpy(), gm)
    Wo = exl2.dequant(layer.qweight.data.cpu().numpy(), to_np_f32(layer.scales), to_np_f32(layer.zeros),
                      layer.q_perm.cpu().numpy(), gm, rows_o)
    assert np.array_equal(to_np_f32(W), Wo), "exl2 dequantisation differs from the oracle"
    for M in (1, 2, 40):
        x = torch.randn((M, K), device="cuda").half()
        y = layer(x)
        ref = torch.matmul(x.mul(layer.channel_scale).view(M, K), W)
        assert torch.all(torch.isclose(y, ref, rtol=2, atol=2))        # the reference's own bound (mixbits test :108)
        assert rel_fro(to_np_f32(y), to_np_f32(x).astype(np.float64) @ Wo.astype(np.float64)) <= 2e-3


@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 11008), (11008, 4096)])
@pytest.mark.parametrize("M", [1, 5, 8, 32])
def test_exl2_fused_forward_llama_shapes(K, N, M):
    """BASELINE config #3 "(ii) MBWQ exl2 within-layer bits=[4,2], bits_prop=[0.75,0.25], group 32" at the Llama-7B
    shapes: the fused mixed-bit kernel (no cuBLAS, W never materialised) against x @ exl2fp(W), where exl2fp is pinned
    bit-exactly to the oracle / the reference extension above.  fp32 accumulation of the exact model vs the fp16-rounded
    W: 2e-3 normwise (the reference's own bound is |diff| < 2, test_nbit_linear_mixbits.py:108)."""
    from bitorch_engine_b200.extensions import q_linear_cuda
    layer = _make_exl2(K, N, STRATEGIES[0], seed=K + N, permute=True)
    x = torch.randn((M, K), device="cuda").half()
    y = q_linear_cuda.mbwq_exl2_forward(x, layer.qweight.data, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map,
                                        layer.rows)
    W = q_linear_cuda.mbwq_exl2fp_weight(layer.qweight.data, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map,
                                         layer.rows)
    ref = (x.double() @ W.double()).cpu().numpy()
    assert y.dtype == torch.float16 and tuple(y.shape) == (M, N)
    assert rel_fro(to_np_f32(y), ref) <= 2e-3
    yc = q_linear_cuda.mbwq_exl2_forward(x, layer.qweight.data, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map,
                                         layer.rows, use_cublas=True)
    assert rel_fro(to_np_f32(yc), ref) <= 2e-3


def test_against_reference_cuda_extension_when_available():
    ref = _ref_ext("q_linear_cuda")
    if ref is None:
        pytest.skip("oracle/_ref/q_linear_cuda not built")
    from bitorch_engine_b200.extensions import q_linear_cuda
    # MPQ forward / grad_input: the reference accumulates in fp16 with atomics, so the comparison is reported with the
    # reference's own error in mind (SURVEY.md section 7 "Numerics vs the reference kernels"): 3e-2 normwise.
    for (bits, group, asym) in [(4, 128, False), (2, 32, False), (4, 128, True), (8, 128, False)]:
        inp = make_mpq_inputs(1024, 512, bits, group, "f16", asym, M=2, seed=bits + group, device="cuda")
        args = (inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, bits, asym)
        y_ref, y = ref.mpq_forward(*args), q_linear_cuda.mpq_forward(*args)
        assert rel_fro(to_np_f32(y), to_np_f32(y_ref)) <= 3e-2
        dy = torch.randn((2, 512), device="cuda").half()
        gargs = (inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], dy, 16, bits, asym)
        assert rel_fro(to_np_f32(q_linear_cuda.mpq_grad_input(*gargs)), to_np_f32(ref.mpq_grad_input(*gargs))) <= 3e-2
    # q42fp: bit-exact
    for bits, group in [(4, 128), (2, 32)]:
        inp = make_mpq_inputs(1024, 512, bits, group, "f16", False, M=2, seed=9, device="cuda")
        q_perm = torch.randperm(1024).to(torch.short).cuda()
        a = ref.mbwq_q42fp_weight(inp["qweight"], inp["scales"], inp["zeros"], group, bits, q_perm)
        b = q_linear_cuda.mbwq_q42fp_weight(inp["qweight"], inp["scales"], inp["zeros"], group, bits, q_perm)
        assert torch.equal(a, b)
        ya = ref.mbwq_q4_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], group, q_perm, bits)
        yb = q_linear_cuda.mbwq_q4_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], group, q_perm, bits)
        assert rel_fro(to_np_f32(yb), to_np_f32(ya)) <= 3e-2
    # exl2: bit-exact dequantisation
    layer = _make_exl2(1024, 256, STRATEGIES[2], seed=3, permute=True)
    a = ref.mbwq_exl2fp_weight(layer.qweight.data, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map, layer.rows[:7])
    b = q_linear_cuda.mbwq_exl2fp_weight(layer.qweight.data, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map, layer.rows)
    assert torch.equal(a, b)
    # exl2 fused forward vs the reference's mixed-bit GEMV (fp16 accumulation + atomics on their side: 3e-2 normwise)
    for M in (1, 4):
        x = torch.randn((M, 1024), device="cuda").half()
        ya = ref.mbwq_exl2_forward(x, layer.qweight.data, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map,
                                   layer.rows[:7], False)
        yb = q_linear_cuda.mbwq_exl2_forward(x, layer.qweight.data, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map,
                                             layer.rows)
        assert rel_fro(to_np_f32(yb), to_np_f32(ya)) <= 3e-2
