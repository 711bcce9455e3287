"""GPU parity of the optimizer paths round 1 left untested (VERDICT r1 "weak" 1.iv):
  * gptq_style_unpacking on CUDA tensors (one dequant kernel) -- bit-exact vs the reference's Python (optim2_cases.npz)
  * the UNFUSED weight update (optim/update.py _mpq_step_torch): act-order g_idx, MBWQ q_perm weights, GaLore projector,
    against the reference's qweight_update_fn run on CPU (gen_golden.py gen_optim2).  Codes may differ by one where a
    value sits on a .5 boundary (fp32: <= 2e-4 of them; fp16 state: <= 2e-2, see tests/test_gpu_optim.py)
  * the symmetric MPQ branch, which the reference cannot execute (UnboundLocalError): the documented rule
    zeros += step * mean_g(norm_grad) and re-packing against the updated zeros
  * DiodeMix end to end on stock torch through BinaryLinearCuda / MPQLinearCuda: integer weights update from the
    privileged gradient, a layer that did not run backward is skipped
  * M == 1 forward under torch.inference_mode()."""
import os

import numpy as np
import pytest
import torch

from oracle import nbit
from helpers import GOLD, make_mpq_inputs, to_np_f32, assert_close_to_oracles
from test_optim2_cpu import Z, t16, make_param, UNPACK

pytestmark = pytest.mark.gpu
TDT = {"f32": torch.float32, "f16": torch.float16}


@pytest.mark.parametrize("case", UNPACK, ids=lambda c: "-".join(c[:4]))
def test_gptq_style_unpacking_cuda_bit_exact(case):
    from bitorch_engine_b200.utils.quant_operators import gptq_style_unpacking
    name, kind, w_bit, group = case[0], case[1], int(case[2]), int(case[3])
    qp = make_param(name, kind, w_bit, group, device="cuda")
    w, z = gptq_style_unpacking(qp)
    assert w.is_cuda and torch.equal(w.cpu().view(torch.int16), t16(Z[f"{name}_w"]).view(torch.int16))
    zr = t16(Z[f"{name}_z"])
    zc = z.cpu()
    assert torch.equal(zc.view(torch.int16) if zc.dtype == torch.float16 else zc,
                       zr.view(torch.int16) if zr.dtype == torch.float16 else zr)


UPD = [str(c).split(",") for c in Z["update_cases"]]


@pytest.mark.parametrize("case", UPD, ids=lambda c: "-".join(c[:5]))
def test_unfused_update_matches_reference(case):
    from bitorch_engine_b200.layers.qlinear.nbit import MPQWeightParameter
    from bitorch_engine_b200.optim.galore_projector import GaLoreProjector
    name, kind, w_bit, group, odt, rank = case[0], case[1], int(case[2]), int(case[3]), case[4], int(case[5])
    tdt = TDT[odt]
    qp = make_param(name, "mpq" if kind.startswith("mpq") else "mbwq", w_bit, group, device="cuda")
    projector = GaLoreProjector(rank, update_proj_gap=1000, scale=0.5, proj_type="std") if rank else None
    step = torch.zeros(1)
    m = v = None
    code_tol = 2e-4 if odt == "f32" else 2e-2
    if rank:
        code_tol = 2e-3          # the SVD runs in cuSOLVER here and in LAPACK for the golden vectors
    for it in range(1, 7):
        grad = t16(Z[f"{name}_grad{it}"], "cuda")
        g = grad
        if projector is not None:
            g = projector.project(grad.to(tdt), step.item())
        if m is None:
            m = torch.zeros_like(g, dtype=tdt)
            v = torch.zeros_like(g, dtype=tdt)
        MPQWeightParameter.update(qp, exp_avg_s=v, exp_avg_l=m, step=step, lr=2e-3, weight_decay=0.0, beta1=0.99,
                                  beta2=0.9999, eps=1e-6, dtype=tdt, correct_bias=True, projector=projector, grad=g)
        torch.cuda.synchronize()
        got = nbit.unpack_int(qp.data.cpu().numpy(), w_bit).astype(np.int64)
        exp = nbit.unpack_int(Z[f"{name}_qweight{it}"], w_bit).astype(np.int64)
        diff = np.abs(got - exp)
        assert diff.max() <= 2 and (diff != 0).mean() <= code_tol, f"step {it}: {(diff != 0).mean():.2e} codes differ"
        ze = Z[f"{name}_zeros{it}"]
        if ze.dtype == np.uint16:
            np.testing.assert_allclose(qp.zeros.float().cpu().numpy(), t16(ze).float().numpy(), rtol=2e-3, atol=2e-5)
        else:
            zg = nbit.unpack_zeros_asym(qp.zeros.cpu().numpy(), w_bit)
            zx = nbit.unpack_zeros_asym(ze, w_bit)
            assert np.abs(zg - zx).max() <= 1 and (zg != zx).mean() <= (0.0 if it < 5 else 2e-2), f"step {it}: zero points"


def test_symmetric_mpq_update_rule():
    """fused kernel, sym branch: w' = w - step*m/(sqrt(v)+eps); every 5th step zeros += step*mean_g(norm_grad); codes =
    clamp(rint((w' + zeros')/s)).  Checked against the same arithmetic in torch fp32 (codes: <= 1e-3 off by one)."""
    from bitorch_engine_b200.layers.qlinear.nbit import MPQWeightParameter
    from bitorch_engine_b200.extensions import q_linear_cuda
    K, N, group, w_bit = 512, 256, 128, 4
    inp = make_mpq_inputs(K, N, w_bit, group, "f16", False, seed=41, device="cuda")
    qp = MPQWeightParameter(inp["qweight"].clone(), requires_grad=False, scales=inp["scales"], zeros=inp["zeros"].clone(),
                            g_idx=inp["g_idx"], w_bit=w_bit, asym=False, group_size=group, layer_type=1)
    m = torch.zeros((K, N), device="cuda"); v = torch.zeros((K, N), device="cuda")
    mr, vr = m.clone(), v.clone()
    zeros_ref = inp["zeros"].clone()
    qref = inp["qweight"].clone()
    step = torch.zeros(1)
    gen = torch.Generator(device="cuda").manual_seed(5)
    b1, b2, eps, lr = 0.99, 0.9999, 1e-6, 2e-3
    gi = inp["g_idx"].long()
    for it in range(1, 6):
        grad = (torch.randn((K, N), device="cuda", generator=gen) * 0.05).half()
        w = q_linear_cuda.mpq_dequant(qref, inp["scales"], zeros_ref, inp["g_idx"], w_bit, False).float()
        mr.mul_(b1).add_(grad.float(), alpha=1 - b1); vr.mul_(b2).addcmul_(grad.float(), grad.float(), value=1 - b2)
        ss = lr * (1 - b2 ** it) ** 0.5 / (1 - b1 ** it)
        ng = mr / (vr.sqrt() + eps)
        w = w - ss * ng
        if it % 5 == 0:
            zeros_ref = (zeros_ref.float() + ss * ng.view(K // group, group, N).mean(1)).half()
        codes = torch.round((w + zeros_ref.float()[gi]) / inp["scales"].float()[gi]).clamp(0, 15).to(torch.int32)
        qref = (codes.view(K // 8, 8, N) << (torch.arange(8, device="cuda", dtype=torch.int32) * 4).view(1, 8, 1)).sum(1).to(torch.int32)
        MPQWeightParameter.update(qp, exp_avg_s=v, exp_avg_l=m, step=step, lr=lr, beta1=b1, beta2=b2, eps=eps,
                                  dtype=torch.float32, correct_bias=True, grad=grad)
        got = nbit.unpack_int(qp.data.cpu().numpy(), w_bit).astype(np.int64)
        exp = nbit.unpack_int(qref.cpu().numpy(), w_bit).astype(np.int64)
        d = np.abs(got - exp)
        assert d.max() <= 1 and (d != 0).mean() <= 2e-3, f"step {it}: {(d != 0).mean():.2e}"
        qref = qp.data.clone()          # stay on the kernel's trajectory (off-by-one codes would otherwise accumulate)
        np.testing.assert_allclose(qp.zeros.float().cpu().numpy(), zeros_ref.float().cpu().numpy(), rtol=2e-3, atol=1e-5)
        zeros_ref = qp.zeros.clone()


def test_diodemix_end_to_end_on_stock_torch():
    """binary (int8) and MPQ (int32) weights cannot receive .grad on stock torch: backward leaves the weight gradient in
    privileged_grad; DiodeMix must use it, and must skip the layer whose backward did not run in this iteration."""
    from bitorch_engine_b200.layers.qlinear.binary.cuda import BinaryLinearCuda
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MPQLinearCuda
    from bitorch_engine_b200.optim import DiodeMix
    torch.manual_seed(0)
    used, unused = BinaryLinearCuda(256, 128, dtype=torch.float).cuda(), BinaryLinearCuda(256, 128, dtype=torch.float).cuda()
    for l in (used, unused):
        l.prepare_params()
        l.train()
    mpq = MPQLinearCuda(128, 256, w_bit=4, group_size=128, dq_group_size=128, use_gba_quant=False, dtype=torch.half).cuda()
    mpq.qweight.data = torch.randint(-2 ** 31, 2 ** 31 - 1, mpq.qweight.shape, dtype=torch.int32, device="cuda")
    mpq.scales.fill_(0.01)
    mpq.prepare_params()
    mpq.train()
    opt = DiodeMix([used.weight, unused.weight, mpq.qweight], lr=1e-2, dtype=torch.float)
    w_unused0, q0 = unused.weight.data.clone(), mpq.qweight.data.clone()
    flipped = 0
    for it in range(3):
        w0 = used.weight.data.clone()
        x = torch.randn(64, 256, device="cuda")
        y = used(x).half()
        loss = (mpq(y).float() ** 2).mean()
        opt.zero_grad()
        loss.backward()
        assert used.weight.grad is not None or getattr(used.weight, "privileged_grad", None) is not None
        opt.step()
        flipped += int((used.weight.data != w0).sum())
        assert torch.equal(unused.weight.data, w_unused0), "a layer without backward was stepped"
    assert used.weight.dtype == torch.int8 and flipped > 0, "the binary weight never updated"
    assert not torch.equal(mpq.qweight.data, q0), "the MPQ weight never updated"
    assert opt.state[used.weight]["step"].item() == 3 and "step" not in opt.state[unused.weight]


def test_decode_forward_under_inference_mode():
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MPQLinearCuda
    inp = make_mpq_inputs(4096, 4096, 4, 128, "f16", False, M=1, seed=9, device="cuda")
    layer = MPQLinearCuda(4096, 4096, w_bit=4, group_size=128, dq_group_size=128, dtype=torch.half, requires_grad=False).cuda()
    layer.prepare_params()
    layer.qweight.data, layer.scales, layer.zeros = inp["qweight"], inp["scales"], inp["zeros"]
    layer.eval()
    with torch.inference_mode():
        x = inp["x"].clone()
        ys = [layer(x) for _ in range(3)]          # eval: PDL on, the sibling detection must not touch ._version
    torch.cuda.synchronize()
    args = (to_np_f32(inp["x"]), inp["qweight"].cpu().numpy(), to_np_f32(inp["scales"]), to_np_f32(inp["zeros"]), None, 4, False)
    assert_close_to_oracles(to_np_f32(ys[0]), nbit.mpq_forward(*args, "f16"), nbit.mpq_forward_exact(*args), "f16", "inference_mode")
    assert all(torch.equal(ys[0], y) for y in ys)
