"""GPU parity for the fused DiodeMix updates against the reference's own qweight_update_fn run on CPU
(oracle/gen_golden.py gen_optim -> tests/golden/optim_cases.npz).

Tolerance: the reference tests pin nothing here ("parity unpinned"); the oracle is the reference
Python itself.  fp32 optimizer state: moments must agree to fp32 round-off (rtol 1e-6) and the re-packed integer codes
must be IDENTICAL except where a value sits within round-off of a .5 rounding boundary (bound: 1e-4 of the codes,
each off by one).  fp16 state: 1 half-ulp on the moments, 2e-3 of the codes."""
import os

import numpy as np
import pytest
import torch

from oracle import nbit
from helpers import GOLD

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(GOLD, "optim_cases.npz"))
CASES = [str(c).split(",") for c in Z["cases"]]
TDT = {"f32": torch.float32, "f16": torch.float16}


def _dev(a, dt=None):
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.float16).cuda()
    return torch.from_numpy(a.copy()).cuda()


@pytest.mark.parametrize("case", [c for c in CASES if c[1] == "mpq"], ids=lambda c: f"{c[0]}-b{c[2]}-g{c[3]}-{c[4]}")
def test_mpq_update_matches_reference(case):
    from bitorch_engine_b200.layers.qlinear.nbit import MPQWeightParameter
    name, _, w_bit, group, odt, K, N = case[0], case[1], int(case[2]), int(case[3]), case[4], int(case[5]), int(case[6])
    tdt = TDT[odt]
    qp = MPQWeightParameter(_dev(Z[f"{name}_qweight0"]), requires_grad=False, scales=_dev(Z[f"{name}_scales"]),
                            zeros=_dev(Z[f"{name}_zeros0"]), g_idx=(torch.arange(K, dtype=torch.int32) // group).cuda(),
                            w_bit=w_bit, asym=True, group_size=group, layer_type=1)
    m = torch.zeros((K, N), dtype=tdt, device="cuda")
    v = torch.zeros((K, N), dtype=tdt, device="cuda")
    step = torch.zeros(1)
    code_tol = 1e-4 if odt == "f32" else 2e-3
    for it in range(1, 7):
        grad = _dev(Z[f"{name}_grad{it}"])
        MPQWeightParameter.update(qp, exp_avg_s=v, exp_avg_l=m, step=step, lr=2e-3, weight_decay=0.0, beta1=0.99,
                                  beta2=0.9999, eps=1e-6, dtype=tdt, correct_bias=True, projector=None, grad=grad)
        torch.cuda.synchronize()
        m_ref, v_ref = _dev(Z[f"{name}_m{it}"]).float().cpu().numpy(), _dev(Z[f"{name}_v{it}"]).float().cpu().numpy()
        rt = 1e-6 if odt == "f32" else 1.5e-3
        np.testing.assert_allclose(m.float().cpu().numpy(), m_ref, rtol=rt, atol=1e-12 if odt == "f32" else 1e-7)
        np.testing.assert_allclose(v.float().cpu().numpy(), v_ref, rtol=rt, atol=1e-12 if odt == "f32" else 1e-7)
        got = nbit.unpack_int(qp.data.cpu().numpy(), w_bit).astype(np.int64)
        exp = nbit.unpack_int(Z[f"{name}_qweight{it}"], w_bit).astype(np.int64)
        diff = np.abs(got - exp)
        assert diff.max() <= 1 and (diff != 0).mean() <= code_tol, f"step {it}: {(diff != 0).mean():.2e} codes differ"
        zg = nbit.unpack_zeros_asym(qp.zeros.cpu().numpy(), w_bit)
        ze = nbit.unpack_zeros_asym(Z[f"{name}_zeros{it}"], w_bit)
        assert (zg != ze).mean() <= (0.0 if it < 5 else 5e-3), f"step {it}: zero points differ"
    assert int(step.item()) == 6


@pytest.mark.parametrize("case", [c for c in CASES if c[1] == "binary"], ids=lambda c: f"{c[0]}-{c[4]}")
def test_binary_update_matches_reference(case):
    from bitorch_engine_b200.layers.qlinear.binary import BinaryLinearParameter
    name, odt = case[0], case[4]
    tdt = TDT[odt]
    w = BinaryLinearParameter(_dev(Z[f"{name}_w0"]), requires_grad=False)
    m = torch.zeros(w.shape, dtype=tdt, device="cuda")
    v = _dev(Z[f"{name}_v0"]).to(tdt)
    step = torch.zeros(1)
    for it in range(1, 5):
        g = _dev(Z[f"{name}_grad{it}"])
        BinaryLinearParameter.update(w, exp_avg_s=v, exp_avg_l=m, step=step, lr=1e-3, beta1=0.99, beta2=0.9999, dtype=tdt,
                                     grad=g)
        torch.cuda.synchronize()
        rt = 1e-6 if odt == "f32" else 1.5e-3
        np.testing.assert_allclose(m.float().cpu().numpy(), _dev(Z[f"{name}_m{it}"]).float().cpu().numpy(), rtol=rt, atol=1e-7)
        np.testing.assert_allclose(v.float().cpu().numpy(), _dev(Z[f"{name}_v{it}"]).float().cpu().numpy(), rtol=rt, atol=1e-9)
        flips = (w.data.cpu().numpy() != Z[f"{name}_w{it}"]).mean()
        assert flips <= (0.0 if odt == "f32" else 2e-3), f"step {it}: {flips:.2e} of the signs differ"


def test_diodemix_end_to_end_reduces_loss():
    """A tiny regression: MPQLinearCuda + DiodeMix actually learns (loss goes down) with the fused kernels."""
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MPQLinearCuda
    from bitorch_engine_b200.optim import DiodeMix
    from helpers import make_mpq_inputs
    torch.manual_seed(0)
    layer = MPQLinearCuda(256, 128, w_bit=4, group_size=128, use_gba_quant=False, requires_grad=True)
    layer.prepare_params()
    layer = layer.cuda()
    inp = make_mpq_inputs(256, 128, 4, 128, "f16", True, M=32, seed=9, device="cuda")
    layer.qweight.data, layer.scales, layer.qzeros = inp["qweight"], inp["scales"], inp["zeros"]
    layer.zeros = layer.qzeros
    layer.train()
    opt = DiodeMix(layer.parameters(), lr=5e-4, dtype=torch.float)
    target = torch.randn((32, 128), device="cuda").half() * 0.1
    losses = []
    for _ in range(12):
        x = inp["x"].clone().requires_grad_(True)
        loss = ((layer(x) - target).float() ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses
