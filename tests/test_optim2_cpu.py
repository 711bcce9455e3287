"""CPU parity of the optimizer-side host helpers against vectors produced by the reference's own Python
(oracle/gen_golden.py gen_optim2 -> tests/golden/optim2_cases.npz):
  gptq_style_unpacking (bitorch_engine/utils/quant_operators.py:310-345), update_zeros (utils/model_helper.py:330-360),
  GaLoreProjector.project / project_back (optim/galore_projector.py:26-107) incl. square gradients, where the side cannot
  be read off the factor's shape.  Bit-exact for the unpack / zero-point arithmetic, 1e-5 for the fp32 projections
  (same LAPACK SVD).  Also: the layer survives copy.deepcopy (peft / EMA flows) and the version-counter helper of the
  shim tolerates inference tensors."""
import copy
import os

import numpy as np
import pytest
import torch

from helpers import GOLD

Z = np.load(os.path.join(GOLD, "optim2_cases.npz"))


def t16(a, device="cpu"):
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.float16).to(device)
    return torch.from_numpy(a.copy()).to(device)


def make_param(name, kind, w_bit, group, device="cpu"):
    from bitorch_engine_b200.layers.qlinear.nbit import MPQWeightParameter
    kw = dict(scales=t16(Z[f"{name}_scales"], device), zeros=t16(Z[f"{name}_zeros0"], device), w_bit=w_bit,
              group_size=group)
    if kind.startswith("mpq"):
        kw.update(g_idx=t16(Z[f"{name}_g_idx"], device), asym=True, layer_type=1)
    else:
        kw.update(g_idx=None, asym=False, layer_type=2, q_perm=t16(Z[f"{name}_q_perm"], device))
    return MPQWeightParameter(t16(Z[f"{name}_qweight0"], device), requires_grad=False, **kw)


UNPACK = [str(c).split(",") for c in Z["unpack_cases"]]


@pytest.mark.parametrize("case", UNPACK, ids=lambda c: "-".join(c[:4]))
def test_gptq_style_unpacking_matches_reference(case):
    from bitorch_engine_b200.utils.quant_operators import gptq_style_unpacking
    name, kind, w_bit, group = case[0], case[1], int(case[2]), int(case[3])
    qp = make_param(name, kind, w_bit, group)
    w, z = gptq_style_unpacking(qp)
    assert torch.equal(w.view(torch.int16), t16(Z[f"{name}_w"]).view(torch.int16))
    zr = t16(Z[f"{name}_z"])
    assert z.dtype == zr.dtype and torch.equal(z.view(torch.int16) if z.dtype == torch.float16 else z,
                                               zr.view(torch.int16) if zr.dtype == torch.float16 else zr)


def test_gptq_style_unpacking_symmetric_with_g_idx_returns_no_zeros():
    """the reference raises UnboundLocalError here (quant_operators.py:330-345); documented deviation: zeros is None"""
    from bitorch_engine_b200.layers.qlinear.nbit import MPQWeightParameter
    from bitorch_engine_b200.utils.quant_operators import gptq_style_unpacking
    K, N = 64, 32
    g = torch.Generator().manual_seed(3)
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // 8, N), dtype=torch.int32, generator=g)
    sc = (torch.rand((2, N), generator=g) * 0.01 + 0.005).half()
    zr = (sc.float() * 8).half()
    gi = torch.arange(K, dtype=torch.int32) // 32
    qp = MPQWeightParameter(qw, requires_grad=False, scales=sc, zeros=zr, g_idx=gi, w_bit=4, asym=False, group_size=32,
                            layer_type=1)
    w, z = gptq_style_unpacking(qp)
    assert z is None
    q = ((qw.unsqueeze(1) >> torch.arange(0, 32, 4, dtype=torch.int32).view(1, -1, 1)) & 15).reshape(K, N)
    assert torch.equal(w, q.to(torch.int8) * sc[gi.long()] - zr[gi.long()])


@pytest.mark.parametrize("name,kind", [("z0", "mpq_act"), ("z1", "mbwq")])
def test_update_zeros_matches_reference(name, kind):
    from bitorch_engine_b200.utils.model_helper import update_zeros
    from bitorch_engine_b200.utils.quant_operators import gptq_style_unpacking
    qp = make_param(name, kind, 4, 64)
    w, z = gptq_style_unpacking(qp)
    update_zeros(qp, w.float(), t16(Z[f"{name}_norm_grad"]), 0.37, z.float())
    exp = t16(Z[f"{name}_zeros1"])
    assert qp.zeros.dtype == exp.dtype
    assert torch.equal(qp.zeros.view(torch.int16) if exp.dtype == torch.float16 else qp.zeros,
                       exp.view(torch.int16) if exp.dtype == torch.float16 else exp)


GALORE = [str(c).split(",") for c in Z["galore_cases"]]


@pytest.mark.parametrize("case", GALORE, ids=lambda c: "-".join(c[1:]))
def test_galore_projector_matches_reference(case):
    from bitorch_engine_b200.optim.galore_projector import GaLoreProjector
    name, pt = case[0], case[3]
    pr = GaLoreProjector(8, update_proj_gap=2, scale=0.25, proj_type=pt)
    for it in range(3):
        full = t16(Z[f"{name}_full{it}"])
        low = pr.project(full, it)
        back = pr.project_back(low)
        np.testing.assert_allclose(low.numpy(), Z[f"{name}_low{it}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(back.numpy(), Z[f"{name}_back{it}"], rtol=1e-5, atol=1e-5)
        assert back.shape == full.shape


def test_layer_deepcopy_and_gradient_carrier():
    """the [K,N] gradient placeholder used to be a non-leaf expand stored on the module -> deepcopy raised."""
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MPQLinearCuda
    layer = MPQLinearCuda(256, 128, w_bit=4, group_size=128, dq_group_size=128)
    twin = copy.deepcopy(layer)
    assert twin.privileged_grad.shape == (256, 128) and twin._grad_carrier is not layer._grad_carrier
    assert twin._grad_carrier.is_leaf and layer.privileged_grad.stride() == (0, 0)
    assert torch.equal(twin.g_idx, layer.g_idx) and set(twin.state_dict()) == set(layer.state_dict())
    frozen = MPQLinearCuda(256, 128, w_bit=4, group_size=128, dq_group_size=128, requires_grad=False)
    assert frozen.privileged_grad is None and copy.deepcopy(frozen).privileged_grad is None
    half = MPQLinearCuda(256, 128, w_bit=4, group_size=128, dq_group_size=128).to(torch.bfloat16)
    assert half.privileged_grad.dtype == torch.bfloat16 and half._grad_carrier.is_leaf


def test_version_helper_tolerates_inference_tensors():
    """reading ._version of an inference tensor raises; the shim must treat such tensors as 'unknown'."""
    from bitorch_engine_b200.extensions import q_linear_cuda
    with torch.inference_mode():
        t = torch.zeros(4)
    assert t.is_inference() and q_linear_cuda._version(t) is None
    assert q_linear_cuda._version(torch.zeros(4)) == 0
    assert q_linear_cuda._input_ready(t, 0) is False           # no proof of "unchanged" without a counter
    g = torch.arange(8, dtype=torch.int32) // 4
    with torch.inference_mode():
        gi = g.clone()
    assert q_linear_cuda._gidx_is_trivial(gi, 8, 2) and q_linear_cuda._gidx_is_trivial(gi, 8, 2)
