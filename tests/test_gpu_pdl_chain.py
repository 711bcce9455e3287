"""GPU parity of the programmatic-dependent-launch protocol of the decode GEMV (PipeParams::early, include/b200bit.h
B200BIT_FLAG_INPUT_READY): a decoder block's seven linears launched back to back with PDL -- siblings (k, v after q;
up after gate) read their activation before griddepcontrol.wait and run concurrently with the kernel in front, the
dependent ones (o, gate, down, next q) wait first -- must give, bit for bit, what the same kernels give one at a time
with full stream serialisation, eagerly and under CUDA-graph replay, and must match the numpy oracle.
Replaces nothing in the reference (its kernels run fully serialised on the legacy default stream,
mpq_linear_cuda_kernel.cu:482-577); the arithmetic contract is q_linear_cuda.mpq_forward's."""
import numpy as np
import pytest
import torch

from oracle import nbit
from helpers import make_mpq_inputs, to_np_f32, assert_close_to_oracles

pytestmark = pytest.mark.gpu

H, I = 4096, 11008
SHAPES = [("q", H, H), ("k", H, H), ("v", H, H), ("o", H, H), ("gate", H, I), ("up", H, I), ("down", I, H)]


def _block(w_bit, group, dt, asym, seed):
    layers = []
    for i, (name, K, N) in enumerate(SHAPES):
        inp = make_mpq_inputs(K, N, w_bit, group, dt, asym, M=1, seed=seed + i, device="cuda")
        # unit gain so that activations stay O(1) along the chain (codes uniform in [0, 2^b - 1])
        std = float(np.sqrt(((2 ** w_bit) ** 2 - 1) / 12.0))
        inp["scales"] = (inp["scales"].float() * (1.0 / (0.01 * np.sqrt(K) * std))).to(inp["scales"].dtype)
        if not asym:
            inp["zeros"] = (inp["scales"].float() * ((2 ** w_bit - 1) / 2.0)).to(inp["scales"].dtype)
        layers.append((name, inp))
    return layers


def _same(a, b):
    """bit-for-bit equality (NaN-safe: an asymmetric random chain may overflow fp16, identically on both sides)."""
    return torch.equal(a.view(torch.int16), b.view(torch.int16))


def _fwd(x, inp, w_bit, asym, pdl):
    from bitorch_engine_b200.extensions import q_linear_cuda
    return q_linear_cuda.mpq_forward(x, inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, w_bit, asym,
                                     pdl=pdl)


def _block_pass(layers, hid, w_bit, asym, pdl):
    """q,k,v <- hid; o <- v; gate,up <- o; down <- up (bench.py's dataflow).  Returns every output."""
    L = dict(layers)
    q = _fwd(hid, L["q"], w_bit, asym, pdl)
    k = _fwd(hid, L["k"], w_bit, asym, pdl)
    v = _fwd(hid, L["v"], w_bit, asym, pdl)
    o = _fwd(v, L["o"], w_bit, asym, pdl)
    g = _fwd(o, L["gate"], w_bit, asym, pdl)
    u = _fwd(o, L["up"], w_bit, asym, pdl)
    d = _fwd(u, L["down"], w_bit, asym, pdl)
    return [q, k, v, o, g, u, d]


@pytest.mark.parametrize("w_bit,group,dt,asym", [(4, 128, "f16", False), (4, 128, "f16", True), (4, 32, "f16", False),
                                                 (2, 32, "f16", False), (4, 128, "bf16", False), (8, 128, "f16", False)])
def test_block_pdl_matches_serial(w_bit, group, dt, asym):
    layers = _block(w_bit, group, dt, asym, seed=50)
    hid = layers[0][1]["x"]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ref = _block_pass(layers, hid, w_bit, asym, pdl=False)
        stream.synchronize()
        # eager, PDL: two blocks back to back (the second block's q follows the first block's down)
        for rep in range(3):
            out = _block_pass(layers, hid, w_bit, asym, pdl=True)
            out2 = _block_pass(layers, out[-1], w_bit, asym, pdl=True)
            stream.synchronize()
            for a, b, (name, _) in zip(out, ref, layers):
                assert _same(a, b), f"eager PDL rep {rep}: {name} differs from the serial result"
        ref2 = _block_pass(layers, ref[-1], w_bit, asym, pdl=False)
        stream.synchronize()
        for a, b, (name, _) in zip(out2, ref2, layers):
            assert _same(a, b), f"eager PDL second block: {name} differs from the serial result"
        # CUDA graph of two chained blocks, replayed
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            g1 = _block_pass(layers, hid, w_bit, asym, pdl=True)
            g2 = _block_pass(layers, g1[-1], w_bit, asym, pdl=True)
        for rep in range(5):
            for t in g1 + g2:
                t.fill_(float("nan"))
            graph.replay()
            stream.synchronize()
            for a, b, (name, _) in zip(g1 + g2, ref + ref2, layers + layers):
                assert _same(a, b), f"graph replay {rep}: {name} differs from the serial result"
    # and the serial result itself against the oracle (first three layers are enough: same kernels as the other tests)
    for (name, inp), y in list(zip(layers, ref))[:3]:
        zeros = inp["zeros"].cpu().numpy() if asym else to_np_f32(inp["zeros"])
        args = (to_np_f32(hid), inp["qweight"].cpu().numpy(), to_np_f32(inp["scales"]), zeros,
                inp["g_idx"].cpu().numpy(), w_bit, asym)
        assert_close_to_oracles(to_np_f32(y), nbit.mpq_forward(*args, dt), nbit.mpq_forward_exact(*args), dt, name)


def test_inplace_write_between_siblings_is_seen():
    """x.mul_(2) between two calls on the same buffer bumps the tensor version: the second call must not read early."""
    layers = _block(4, 128, "f16", False, seed=70)
    L = dict(layers)
    x = L["q"]["x"].clone()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        want_q = _fwd(x, L["q"], 4, False, pdl=False)
        want_k = _fwd(x * 2, L["k"], 4, False, pdl=False)
        stream.synchronize()
        for _ in range(5):
            xx = x.clone()
            q = _fwd(xx, L["q"], 4, False, pdl=True)
            xx.mul_(2)
            k = _fwd(xx, L["k"], 4, False, pdl=True)
            stream.synchronize()
            assert _same(q, want_q) and _same(k, want_k)


def test_output_fed_back_as_input_waits():
    """y of one call used as x of the next (same buffer address as an earlier x is irrelevant): never early."""
    layers = _block(4, 128, "f16", False, seed=90)
    L = dict(layers)
    x = L["q"]["x"]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        a = _fwd(x, L["q"], 4, False, pdl=False)
        b = _fwd(a, L["k"], 4, False, pdl=False)
        c = _fwd(b, L["v"], 4, False, pdl=False)
        stream.synchronize()
        for _ in range(5):
            a2 = _fwd(x, L["q"], 4, False, pdl=True)
            b2 = _fwd(a2, L["k"], 4, False, pdl=True)
            c2 = _fwd(b2, L["v"], 4, False, pdl=True)
            stream.synchronize()
            assert _same(a2, a) and _same(b2, b) and _same(c2, c)


def test_module_forward_uses_pdl_at_inference_and_matches_serial(monkeypatch):
    """MPQLinearCuda modules (requires_grad=False, the green-bit-llm inference configuration) launch with PDL by default
    (B200BIT_PDL): q/k/v-style calls on one hidden state and a dependent call behind them give the serial results."""
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MPQLinearCuda
    torch.manual_seed(3)
    mods = []
    for i in range(4):
        m = MPQLinearCuda(4096, 4096, w_bit=4, group_size=256, dq_group_size=256, use_gba_quant=True, requires_grad=False,
                          dtype=torch.float16).cuda()
        inp = make_mpq_inputs(4096, 4096, 4, 256, "f16", False, M=1, seed=200 + i, device="cuda")
        m.qweight.data.copy_(inp["qweight"])
        m.scales.data.copy_((inp["scales"].float() / (0.01 * 64 * 4.61)).half())
        m.zeros.data.copy_((m.scales.float() * 7.5).half())
        m.prepare_params()
        m.eval()
        mods.append(m)
    x = torch.randn(1, 1, 4096, device="cuda").half()

    def run():
        with torch.no_grad():
            q, k, v = mods[0](x), mods[1](x), mods[2](x)
            o = mods[3](v)
        torch.cuda.synchronize()
        return [q, k, v, o]

    monkeypatch.setenv("B200BIT_PDL", "0")
    ref = run()
    monkeypatch.setenv("B200BIT_PDL", "1")
    for _ in range(5):
        out = run()
        for a, b in zip(out, ref):
            assert a.shape == (1, 1, 4096) and _same(a, b)
