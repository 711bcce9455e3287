#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference Python
(/root/reference/bitorch_engine) on CPU in the build container.  /root/reference does not exist on the GPU box, so
the vectors are committed; re-run this script to regenerate them:

    python oracle/gen_golden.py            # writes tests/golden/nbit_cases.npz, ...

What runs from the reference (nothing is copied, only imported):
  * bitorch_engine/layers/qlinear/nbit/cuda/utils.py: unpack_qweight (5-69), pack_fp_weight (72-147)
  * bitorch_engine/utils/quant_operators.py: gptq_style_unpacking (310-345), gptq_style_zeros_packing (348-368)
  * bitorch_engine/utils/model_helper.py: qweight_update_fn (363-530)
  * bitorch_engine/optim/diode_beta.py: DiodeMix
The un-vendored `bitorch` dependency is satisfied by the import-only stub in oracle/_stubs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
os.environ["BIE_SKIP_TORCH_CHECK"] = "true"
sys.path.insert(0, os.path.join(ROOT, "oracle", "_stubs"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import torch  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TORCH_DT = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


def _bits(t):
    """torch half/bf16 -> uint16 numpy bits; float32 -> float32 numpy."""
    if t.dtype in (torch.float16, torch.bfloat16):
        return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()
    return t.contiguous().numpy().copy()


def make_inputs(K, N, w_bit, group, dt, asym, act_order, seed):
    """Synthetic inputs per SURVEY.md section 8(d)."""
    g = torch.Generator().manual_seed(seed)
    tdt = TORCH_DT[dt]
    nb = 32 // w_bit
    G = K // group
    qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // nb, N), dtype=torch.int32, generator=g)
    scales = (torch.rand((G, N), generator=g) * 0.01 + 0.005).to(tdt)
    if asym:
        zeros = torch.randint(-2 ** 31, 2 ** 31 - 1, (G, N // nb), dtype=torch.int32, generator=g)
    else:
        zeros = (scales.float() * (2 ** (w_bit - 1)) + torch.randn((G, N), generator=g) * 1e-3).to(tdt)
    g_idx = torch.arange(K, dtype=torch.int32) // group
    if act_order:
        g_idx = g_idx[torch.randperm(K, generator=g)].contiguous()
    x = torch.randn((3, K), generator=g).to(tdt)
    dy = torch.randn((3, N), generator=g).to(tdt)
    return qweight, scales, zeros, g_idx, x, dy


def gen_nbit():
    from bitorch_engine.layers.qlinear.nbit import MPQWeightParameter
    from bitorch_engine.layers.qlinear.nbit.cuda.utils import unpack_qweight, pack_fp_weight

    out = {}
    cases = []
    K, N = 256, 64
    cid = 0
    for w_bit in (4, 2, 8, 1):
        for dt in ("f16", "bf16", "f32"):
            if w_bit in (8, 1) and dt != "f16":
                continue
            for asym in (False, True):
                for act_order in (False, True):
                    if w_bit in (8, 1) and act_order:
                        continue
                    group = 32 if w_bit == 2 else 128
                    seed = 1000 + cid
                    qweight, scales, zeros, g_idx, x, dy = make_inputs(K, N, w_bit, group, dt, asym, act_order, seed)
                    qp = MPQWeightParameter(qweight.clone(), requires_grad=False, scales=scales, zeros=zeros,
                                            g_idx=g_idx, w_bit=w_bit, asym=asym, group_size=group, layer_type=1)
                    W = unpack_qweight(qp)                      # reference, dtype = scales dtype
                    assert W.dtype == TORCH_DT[dt], (W.dtype, dt)
                    y = x.float() @ W.float()                  # the tight oracle
                    dx = dy.float() @ W.float().t()
                    # pack: a perturbed weight so rounding / clamping paths are exercised
                    gen = torch.Generator().manual_seed(seed + 7)
                    Wp = (W.float() + torch.randn(W.shape, generator=gen) * 0.02).to(W.dtype)
                    packed = pack_fp_weight(Wp, qp)
                    roundtrip = pack_fp_weight(W, qp)
                    name = f"c{cid}"
                    cases.append((name, w_bit, dt, int(asym), int(act_order), group, K, N))
                    out[f"{name}_qweight"] = qweight.numpy()
                    out[f"{name}_scales"] = _bits(scales)
                    out[f"{name}_zeros"] = _bits(zeros) if not asym else zeros.numpy()
                    out[f"{name}_g_idx"] = g_idx.numpy()
                    out[f"{name}_x"] = _bits(x)
                    out[f"{name}_dy"] = _bits(dy)
                    out[f"{name}_W"] = _bits(W)
                    out[f"{name}_y"] = y.numpy()
                    out[f"{name}_dx"] = dx.numpy()
                    out[f"{name}_Wp"] = _bits(Wp)
                    out[f"{name}_packed"] = packed.numpy()
                    out[f"{name}_roundtrip_equal"] = np.array(int(torch.equal(roundtrip, qweight)))
                    cid += 1
    out["cases"] = np.array([",".join(map(str, c)) for c in cases])
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "nbit_cases.npz"), **out)
    print(f"nbit: {len(cases)} cases -> tests/golden/nbit_cases.npz")


def gen_optim():
    """DiodeMix weight updates by the reference's own qweight_update_fn (utils/model_helper.py:363-530), CPU."""
    from bitorch_engine.layers.qlinear.nbit import MPQWeightParameter
    from bitorch_engine.layers.qlinear.binary import BinaryLinearParameter
    from bitorch_engine.utils.model_helper import qweight_update_fn
    from bitorch_engine.utils.quant_operators import nv_tensor_quant

    out, cases = {}, []
    K, N = 256, 64
    cid = 0
    for w_bit, group in ((4, 128), (2, 32), (8, 64)):
        for odt in ("f32", "f16"):
            seed = 5000 + cid
            qweight, scales, zeros, g_idx, _, _ = make_inputs(K, N, w_bit, group, "f16", True, False, seed)
            qp = MPQWeightParameter(qweight.clone(), requires_grad=False, scales=scales, zeros=zeros.clone(), g_idx=g_idx,
                                    w_bit=w_bit, asym=True, group_size=group, layer_type=1)
            tdt = TORCH_DT[odt]
            m = torch.zeros((K, N), dtype=tdt)
            v = torch.zeros((K, N), dtype=tdt)
            step = torch.zeros(1)
            gen = torch.Generator().manual_seed(seed + 1)
            name = f"o{cid}"
            out[f"{name}_qweight0"] = qweight.numpy()
            out[f"{name}_scales"] = _bits(scales)
            out[f"{name}_zeros0"] = zeros.numpy()
            for it in range(1, 7):
                grad = (torch.randn((K, N), generator=gen) * 0.05).half()
                out[f"{name}_grad{it}"] = _bits(grad)
                qweight_update_fn(qweight=qp, exp_avg_s=v, exp_avg_l=m, step=step, lr=2e-3, weight_decay=0.0, beta1=0.99,
                                  beta2=0.9999, eps=1e-6, dtype=tdt, correct_bias=True, projector=None, grad=grad)
                out[f"{name}_qweight{it}"] = qp.data.numpy().copy()
                out[f"{name}_zeros{it}"] = qp.zeros.numpy().copy()
                out[f"{name}_m{it}"] = _bits(m) if odt != "f32" else m.numpy().copy()
                out[f"{name}_v{it}"] = _bits(v) if odt != "f32" else v.numpy().copy()
            cases.append((name, "mpq", w_bit, group, odt, K, N))
            cid += 1
    # binary branch
    for odt in ("f32", "f16"):
        gen = torch.Generator().manual_seed(7000 + cid)
        rows, cols = 64, 128
        w = torch.where(torch.rand((rows, cols), generator=gen) < 0.5, -1, 1).to(torch.int8)
        bp = BinaryLinearParameter(w.clone(), requires_grad=False)
        tdt = TORCH_DT[odt]
        m = torch.zeros((rows, cols), dtype=tdt)
        v = -(w.clone().sign().to(tdt) * (torch.rand((rows, cols), generator=gen).to(tdt) * 1e-3))
        step = torch.zeros(1)
        name = f"o{cid}"
        out[f"{name}_w0"] = w.numpy()
        out[f"{name}_v0"] = _bits(v) if odt != "f32" else v.numpy().copy()
        for it in range(1, 5):
            g = nv_tensor_quant(torch.randn((rows, cols), generator=gen))[0].to(torch.int8)
            out[f"{name}_grad{it}"] = g.numpy()
            bp.grad = None
            # the reference reads qweight.grad (integer grads need GreenBit's torch); hand it over the same way
            object.__setattr__(bp, "_grad_holder", g)
            type(bp).grad = property(lambda self: self._grad_holder)
            qweight_update_fn(qweight=bp, exp_avg_s=v, exp_avg_l=m, step=step, lr=1e-3, beta1=0.99, beta2=0.9999,
                              dtype=tdt)
            del type(bp).grad
            out[f"{name}_w{it}"] = bp.data.numpy().copy()
            out[f"{name}_m{it}"] = _bits(m) if odt != "f32" else m.numpy().copy()
            out[f"{name}_v{it}"] = _bits(v) if odt != "f32" else v.numpy().copy()
        cases.append((name, "binary", 1, 0, odt, rows, cols))
        cid += 1
    out["cases"] = np.array([",".join(map(str, c)) for c in cases])
    np.savez_compressed(os.path.join(GOLD, "optim_cases.npz"), **out)
    print(f"optim: {len(cases)} cases -> tests/golden/optim_cases.npz")


def gen_layers():
    """state_dict layout of the reference MPQLinearCuda before / after prepare_params, and the double-dequantised
    scales / zeros it produces (mpq_layer.py:163-204) for random statistics."""
    import json
    from bitorch_engine.layers.qlinear.nbit.cuda import MPQLinearCuda
    specs, tensors = [], {}
    configs = [dict(w_bit=4, group_size=128, dq_group_size=256), dict(w_bit=2, group_size=32, dq_group_size=32),
               dict(w_bit=4, group_size=128, use_gba_quant=False), dict(w_bit=4, group_size=64, dq_group_size=64, asym=True),
               dict(w_bit=2, group_size=32, dq_group_size=32, asym=True, dq_mode=1), dict(w_bit=8, group_size=256),
               dict(w_bit=1, group_size=128, dq_group_size=128)]
    for ci, kw in enumerate(configs):
        layer = MPQLinearCuda(256, 512, requires_grad=False, **kw)
        before = {k: [list(v.shape), str(v.dtype)] for k, v in layer.state_dict().items()}
        g = torch.Generator().manual_seed(9000 + ci)
        for name, buf in layer.named_buffers():
            if buf.dtype == torch.uint8:
                buf.copy_(torch.randint(0, 256, buf.shape, dtype=torch.uint8, generator=g))
            elif buf.is_floating_point() and name not in ("bias",):
                buf.copy_((torch.rand(buf.shape, generator=g) * 0.02 + 0.001).to(buf.dtype))
            if name in ("qstatistic", "qscales", "qscales_zeros", "qscales_scales", "qzeros_zeros", "qzeros_scales"):
                tensors[f"l{ci}_{name}"] = _bits(buf) if buf.dtype in (torch.float16, torch.bfloat16) else buf.numpy().copy()
        layer.prepare_params()
        after = {k: [list(v.shape), str(v.dtype)] for k, v in layer.state_dict().items()}
        tensors[f"l{ci}_scales_out"] = _bits(layer.scales)
        if layer.zeros.dtype != torch.int32:
            tensors[f"l{ci}_zeros_out"] = _bits(layer.zeros)
        specs.append(dict(kwargs=kw, before=before, after=after))
    with open(os.path.join(GOLD, "layer_specs.json"), "w") as fh:
        json.dump(specs, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(GOLD, "layer_prepare.npz"), **tensors)
    print(f"layers: {len(specs)} configs -> tests/golden/layer_specs.json, layer_prepare.npz")




def gen_helpers():
    """Host-side helpers either side of the path, by the reference's own Python (CPU): utils/model_helper.py
    pad_embedding_dim (54-82), pad_last_2_dims_to_multiple_of_128 (85-117), binary_matmul_forward_post_processing
    (120-155); utils/quant_operators.py get_binary_row / get_binary_col (118-231), q8_quantization / q4_quantization
    (234-307), gptq_style_zeros_packing (348-368).  -> tests/golden/helper_cases.npz"""
    from bitorch_engine.utils import model_helper as mh
    from bitorch_engine.utils import quant_operators as qo
    g = torch.Generator().manual_seed(77)
    out = {}
    w = torch.randn((5, 13), generator=g)
    out["pad_emb_in"], out["pad_emb_out"] = w.numpy(), mh.pad_embedding_dim(w).numpy()
    w8 = torch.randn((3, 16), generator=g)
    out["pad_emb8_in"], out["pad_emb8_out"] = w8.numpy(), mh.pad_embedding_dim(w8).numpy()
    t = torch.randn((2, 100, 130), generator=g)
    p, sec = mh.pad_last_2_dims_to_multiple_of_128(t)
    out["pad128_in"], out["pad128_out"], out["pad128_sec"] = t.numpy(), p.numpy(), np.array(sec)
    bm = torch.randint(0, 64, (2, 128, 128), generator=g).float()
    out["bmm_in"] = bm.numpy()
    out["bmm_out"] = mh.binary_matmul_forward_post_processing(bm, [2], 30, 28, 64).numpy()
    x = torch.randn((4, 64), generator=g)
    x[0, :3] = torch.tensor([0.0, -0.0, -1e-9])
    row = qo.get_binary_row(x.flatten().tolist(), [0] * (x.numel() // 32), x.numel(), 32)
    out["bin_in"], out["bin_row"] = x.numpy(), np.array(row, dtype=np.uint64)
    col = qo.get_binary_col(x.t().contiguous().flatten().tolist(), [0] * (64 // 32 * 4), 64, 4, 32)
    out["bin_col"] = np.array(col, dtype=np.uint64)
    a = torch.randn((7, 33), generator=g) * 3
    q8, s8 = qo.q8_quantization(a, eps=torch.tensor(1e-5))
    q4, s4 = qo.q4_quantization(a, eps=torch.tensor(1e-5))
    out["q_in"], out["q8"], out["q8_scale"], out["q4"], out["q4_scale"] = a.numpy(), q8.numpy(), s8.numpy(), q4.numpy(), s4.numpy()
    sc = torch.tensor(0.37)
    out["q8_given"] = qo.q8_quantization(a, sc, torch.tensor(1e-5)).numpy()
    out["q4_given"] = qo.q4_quantization(a, sc, torch.tensor(1e-5)).numpy()
    for b in (2, 4, 8):
        z = torch.randint(1, 2 ** b + 1, (4, 64), generator=g)
        out[f"zp{b}_in"] = z.numpy()
        out[f"zp{b}_out"] = qo.gptq_style_zeros_packing(z, b, 64, 32).numpy()
    np.savez_compressed(os.path.join(GOLD, "helper_cases.npz"), **out)
    print("helper_cases.npz:", {k: v.shape for k, v in out.items()})


def gen_optim2():
    """The unfused branches of the reference's MPQ weight update + its optimizer-side helpers, run on CPU:
      * gptq_style_unpacking (utils/quant_operators.py:310-345): asym act-order, asym contiguous, MBWQ (q_perm scatter)
      * update_zeros (utils/model_helper.py:330-360): both branches
      * qweight_update_fn (model_helper.py:485-523) with act-order g_idx, with an MBWQ (layer_type 2) weight, and with a
        GaLoreProjector (one SVD: update_proj_gap larger than the run, so CPU / GPU singular-vector signs cancel)
      * GaLoreProjector.project / project_back for every proj_type on tall, wide and square gradients."""
    from bitorch_engine.layers.qlinear.nbit import MPQWeightParameter
    from bitorch_engine.utils.model_helper import qweight_update_fn, update_zeros
    from bitorch_engine.utils.quant_operators import gptq_style_unpacking
    from bitorch_engine.optim.galore_projector import GaLoreProjector

    out = {}
    K, N = 256, 64

    def mpq(w_bit, group, act, seed):
        qweight, scales, zeros, g_idx, _, _ = make_inputs(K, N, w_bit, group, "f16", True, act, seed)
        return MPQWeightParameter(qweight.clone(), requires_grad=False, scales=scales, zeros=zeros.clone(), g_idx=g_idx,
                                  w_bit=w_bit, asym=True, group_size=group, layer_type=1)

    def mbwq(w_bit, group, seed):
        qweight, scales, zeros, _, _, _ = make_inputs(K, N, w_bit, group, "f16", False, False, seed)
        gen = torch.Generator().manual_seed(seed + 77)
        q_perm = torch.randperm(K, generator=gen).to(torch.short)
        return MPQWeightParameter(qweight.clone(), requires_grad=False, scales=scales, zeros=zeros.clone(), g_idx=None,
                                  w_bit=w_bit, asym=False, group_size=group, layer_type=2, q_perm=q_perm)

    def dump_param(name, qp):
        out[f"{name}_qweight0"] = qp.data.numpy().copy()
        out[f"{name}_scales"] = _bits(qp.scales)
        out[f"{name}_zeros0"] = qp.zeros.numpy().copy() if qp.zeros.dtype == torch.int32 else _bits(qp.zeros)
        if qp.g_idx is not None:
            out[f"{name}_g_idx"] = qp.g_idx.numpy().copy()
        if getattr(qp, "q_perm", None) is not None:
            out[f"{name}_q_perm"] = qp.q_perm.numpy().copy()

    # ---- gptq_style_unpacking ----
    unpack_cases = []
    for name, qp, meta in (("u0", mpq(4, 64, True, 8100), "mpq,4,64,1"), ("u1", mpq(2, 32, False, 8101), "mpq,2,32,0"),
                           ("u2", mpq(8, 128, True, 8102), "mpq,8,128,1"), ("u3", mbwq(4, 64, 8103), "mbwq,4,64,0"),
                           ("u4", mbwq(2, 32, 8104), "mbwq,2,32,0")):
        dump_param(name, qp)
        w, z = gptq_style_unpacking(qp)
        out[f"{name}_w"] = _bits(w)
        out[f"{name}_z"] = z.numpy().copy() if z.dtype in (torch.int8, torch.int16) else _bits(z)
        unpack_cases.append(f"{name},{meta}")
    out["unpack_cases"] = np.array(unpack_cases)

    # ---- update_zeros ----
    gen = torch.Generator().manual_seed(8200)
    for name, qp in (("z0", mpq(4, 64, True, 8201)), ("z1", mbwq(4, 64, 8202))):
        dump_param(name, qp)
        w, z = gptq_style_unpacking(qp)
        w, z = w.float(), z.float()
        ng = torch.randn((K, N), generator=gen)
        out[f"{name}_norm_grad"] = ng.numpy().copy()
        update_zeros(qp, w, ng, 0.37, z)
        out[f"{name}_zeros1"] = qp.zeros.numpy().copy() if qp.zeros.dtype == torch.int32 else _bits(qp.zeros)

    # ---- qweight_update_fn, unfused branches ----
    upd_cases = []
    def run_update(name, qp, odt, projector, iters=6):
        tdt = TORCH_DT[odt]
        gen = torch.Generator().manual_seed(8300 + len(upd_cases))
        dump_param(name, qp)
        step = torch.zeros(1)
        m = v = None
        for it in range(1, iters + 1):
            grad = (torch.randn((K, N), generator=gen) * 0.05).half()
            out[f"{name}_grad{it}"] = _bits(grad)
            g = grad
            if projector is not None:
                g = projector.project(grad.to(tdt), step.item())
            if m is None:
                m = torch.zeros_like(g, dtype=tdt)
                v = torch.zeros_like(g, dtype=tdt)
            qweight_update_fn(qweight=qp, exp_avg_s=v, exp_avg_l=m, step=step, lr=2e-3, weight_decay=0.0, beta1=0.99,
                              beta2=0.9999, eps=1e-6, dtype=tdt, correct_bias=True, projector=projector, grad=g)
            out[f"{name}_qweight{it}"] = qp.data.numpy().copy()
            out[f"{name}_zeros{it}"] = qp.zeros.numpy().copy() if qp.zeros.dtype == torch.int32 else _bits(qp.zeros)
    run_update("f0", mpq(4, 64, True, 8301), "f32", None); upd_cases.append("f0,mpq_act,4,64,f32,0")
    run_update("f1", mbwq(4, 64, 8302), "f32", None); upd_cases.append("f1,mbwq,4,64,f32,0")
    run_update("f2", mbwq(2, 32, 8303), "f16", None); upd_cases.append("f2,mbwq,2,32,f16,0")
    run_update("f3", mpq(4, 128, False, 8304), "f32", GaLoreProjector(16, update_proj_gap=1000, scale=0.5, proj_type="std"))
    upd_cases.append("f3,mpq_galore,4,128,f32,16")
    out["update_cases"] = np.array(upd_cases)

    # ---- GaLoreProjector ----
    gal = []
    gid = 0
    for shape in ((96, 48), (48, 96), (64, 64)):
        for pt in ("std", "reverse_std", "right", "left", "full"):
            gen = torch.Generator().manual_seed(8400 + gid)
            pr = GaLoreProjector(8, update_proj_gap=2, scale=0.25, proj_type=pt)
            name = f"g{gid}"
            for it in range(3):
                gfull = torch.randn(shape, generator=gen)
                low = pr.project(gfull, it)
                back = pr.project_back(low)
                out[f"{name}_full{it}"] = gfull.numpy().copy()
                out[f"{name}_low{it}"] = low.numpy().copy()
                out[f"{name}_back{it}"] = back.numpy().copy()
            gal.append(f"{name},{shape[0]},{shape[1]},{pt}")
            gid += 1
    out["galore_cases"] = np.array(gal)
    np.savez_compressed(os.path.join(GOLD, "optim2_cases.npz"), **out)
    print(f"optim2: {len(unpack_cases)} unpack, {len(upd_cases)} update, {len(gal)} galore cases -> tests/golden/optim2_cases.npz")


if __name__ == "__main__":
    what = sys.argv[1:] or ["nbit"]
    for w in what:
        globals()[f"gen_{w}"]()
