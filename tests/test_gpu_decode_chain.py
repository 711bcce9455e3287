"""GPU parity of the decode chain (include/b200bit.h b200bit_mpq_chain_*, csrc/mpq_chain.cuh): a list of batch-1 4-bit
Linear layers run by ONE persistent launch, dependencies resolved on the device, must give -- bit for bit -- what the
same layers give through q_linear_cuda.mpq_forward one launch at a time (which tests/test_gpu_mpq_forward.py pins against
the numpy oracle), on every replay, under CUDA-graph capture, and with reused buffers (write-after-read hazards).
Replaces n x q_linear_cuda.mpq_forward (q_linear_cuda.cpp:258-270 -> mpq_linear_cuda_kernel.cu:603-626)."""
import numpy as np
import pytest
import torch

from oracle import nbit
from helpers import make_mpq_inputs, to_np_f32, assert_close_to_oracles

pytestmark = pytest.mark.gpu

H, I = 4096, 11008
LLAMA7B = [("q", H, H), ("k", H, H), ("v", H, H), ("o", H, H), ("gate", H, I), ("up", H, I), ("down", I, H)]
LLAMA3_8B = [("q", 4096, 4096), ("k", 4096, 1024), ("v", 4096, 1024), ("o", 4096, 4096), ("gate", 4096, 14336),
             ("up", 4096, 14336), ("down", 14336, 4096)]
SMALL = [("q", 512, 512), ("k", 512, 256), ("v", 512, 512), ("o", 512, 512), ("gate", 512, 1280), ("up", 512, 1280),
         ("down", 1280, 512)]


def _block(shapes, group, dt, asym, seed):
    layers = {}
    for i, (name, K, N) in enumerate(shapes):
        inp = make_mpq_inputs(K, N, 4, group, dt, asym, M=1, seed=seed + i, device="cuda")
        inp["scales"] = (inp["scales"].float() * (1.0 / (0.01 * np.sqrt(K) * 4.61))).to(inp["scales"].dtype)
        if not asym:
            inp["zeros"] = (inp["scales"].float() * 7.5).to(inp["scales"].dtype)
        layers[name] = inp
    return layers


def _same(a, b):
    return torch.equal(a.view(torch.int16), b.view(torch.int16))


def _fwd(x, inp, asym):
    from bitorch_engine_b200.extensions import q_linear_cuda
    return q_linear_cuda.mpq_forward(x, inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, asym)


def _block_pass(L, hid, asym):
    """q,k,v <- hid; o <- v; gate,up <- o; down <- up (bench.py's dataflow)."""
    q, k, v = _fwd(hid, L["q"], asym), _fwd(hid, L["k"], asym), _fwd(hid, L["v"], asym)
    o = _fwd(v if v.shape == q.shape else q, L["o"], asym)
    g, u = _fwd(o, L["gate"], asym), _fwd(o, L["up"], asym)
    d = _fwd(u, L["down"], asym)
    return [q, k, v, o, g, u, d]


def _two_blocks(L, hid, asym):
    a = _block_pass(L, hid, asym)
    b = _block_pass(L, a[-1], asym)
    return a + b


@pytest.mark.parametrize("shapes,group,dt,asym", [
    (SMALL, 128, "f16", False), (SMALL, 32, "f16", False), (SMALL, 64, "bf16", True),
    (LLAMA7B, 128, "f16", False), (LLAMA7B, 128, "f16", True), (LLAMA7B, 128, "bf16", False), (LLAMA7B, 32, "f16", False),
    (LLAMA3_8B, 128, "f16", False)], ids=lambda v: str(len(v)) if isinstance(v, list) else str(v))
def test_chain_matches_per_layer_launches(shapes, group, dt, asym):
    from bitorch_engine_b200.decode_chain import DecodeChain
    L = _block(shapes, group, dt, asym, seed=70)
    hid = L["q"]["x"]
    ref = _two_blocks(L, hid, asym)
    torch.cuda.synchronize()
    chain = DecodeChain.capture(lambda: _two_blocks(L, hid, asym))
    assert len(chain.nodes) == 14
    for rep in range(3):                 # the plan's counters reset themselves
        for t in chain.outputs:
            t.fill_(float("nan"))
        out = chain.launch()
        chain.check()
        for i, (a, b) in enumerate(zip(out, ref)):
            assert _same(a, b), f"replay {rep}: node {i} differs from the per-layer launch"
    # under CUDA-graph capture
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            chain.launch()
        for t in chain.outputs:
            t.fill_(0)
        graph.replay()
        graph.replay()
        stream.synchronize()
    chain.check()
    for i, (a, b) in enumerate(zip(chain.outputs, ref)):
        assert _same(a, b), f"graph replay: node {i} differs"
    # and the first block against the oracle (the per-layer kernel's own contract)
    inp = L["q"]
    args = (to_np_f32(hid), inp["qweight"].cpu().numpy(), to_np_f32(inp["scales"]),
            inp["zeros"].cpu().numpy() if asym else to_np_f32(inp["zeros"]), None, 4, asym)
    assert_close_to_oracles(to_np_f32(chain.outputs[0]), nbit.mpq_forward(*args, dt), nbit.mpq_forward_exact(*args), dt,
                            "chain node 0")


def test_chain_reused_buffers_wait_for_their_readers():
    """y of a later node lands in the buffer an earlier node read / wrote (what torch's caching allocator does when
    intermediates are freed): the plan must order the write behind every reader."""
    from bitorch_engine_b200.decode_chain import DecodeChain
    L = _block(LLAMA7B, 128, "f16", False, seed=90)
    hid = L["q"]["x"]
    ref = _block_pass(L, hid, False)
    ref2 = _block_pass(L, ref[-1], False)
    torch.cuda.synchronize()
    chain = DecodeChain()
    w = lambda name: (L[name]["qweight"], L[name]["scales"], L[name]["zeros"], 4, False)
    q = chain.add(hid, *w("q")); k = chain.add(hid, *w("k")); v = chain.add(hid, *w("v"))
    o = chain.add(v, *w("o"))
    g = chain.add(o, *w("gate")); u = chain.add(o, *w("up"))
    d = chain.add(u, *w("down"))
    # second block: q, k, v are written into the first block's q, k, v buffers (v was read by o: write-after-read;
    # all three: write-after-write)
    chain.add(d, *w("q"), out=q); chain.add(d, *w("k"), out=k); chain.add(d, *w("v"), out=v)
    o2 = chain.add(v, *w("o"))
    chain.outputs = [q, k, v, o, g, u, d, o2]
    chain.build()
    for rep in range(3):
        chain.launch()
        chain.check()
        torch.cuda.synchronize()
        assert _same(o, ref[3]) and _same(g, ref[4]) and _same(u, ref[5]) and _same(d, ref[6]), f"rep {rep}: first block"
        assert _same(q, ref2[0]) and _same(k, ref2[1]) and _same(v, ref2[2]), f"rep {rep}: reused buffers"
        assert _same(o2, ref2[3])


def test_chain_of_a_single_segment_and_argument_checks():
    from bitorch_engine_b200.decode_chain import DecodeChain
    L = _block(LLAMA7B, 128, "f16", False, seed=95)
    hid = L["q"]["x"]
    ref = [_fwd(hid, L[n], False) for n in ("q", "k", "v")]
    chain = DecodeChain.capture(lambda: [_fwd(hid, L[n], False) for n in ("q", "k", "v")])
    chain.launch(); chain.check()
    for a, b in zip(chain.outputs, ref):
        assert _same(a, b)
    bad = DecodeChain()
    y = bad.add(hid, L["q"]["qweight"], L["q"]["scales"], L["q"]["zeros"], 4, False)
    bad.nodes[0] = (hid, hid, *bad.nodes[0][2:])          # a node that writes its own input
    with pytest.raises(ValueError):
        bad.build()
    with pytest.raises(ValueError):
        DecodeChain().add(torch.zeros((2, H), dtype=torch.float16, device="cuda"), L["q"]["qweight"], L["q"]["scales"],
                          L["q"]["zeros"], 4, False)


def test_chain_input_that_is_a_slice_of_an_earlier_output_uses_the_counter_path():
    """x is a proper sub-range of an earlier node's y (a view): not the shadow-word protocol but the counter wait."""
    from bitorch_engine_b200.decode_chain import DecodeChain
    a = make_mpq_inputs(512, 1024, 4, 128, "f16", False, M=1, seed=11, device="cuda")
    b = make_mpq_inputs(512, 768, 4, 128, "f16", False, M=1, seed=12, device="cuda")
    c = make_mpq_inputs(768, 512, 4, 128, "f16", False, M=1, seed=13, device="cuda")
    for inp, K in ((a, 512), (b, 512), (c, 768)):
        inp["scales"] = (inp["scales"].float() * (1.0 / (0.01 * np.sqrt(K) * 4.61))).half()
        inp["zeros"] = (inp["scales"].float() * 7.5).half()
    ya = _fwd(a["x"], a, False)
    yb = _fwd(ya[:, 256:768].contiguous(), b, False)
    yc = _fwd(yb, c, False)
    torch.cuda.synchronize()
    chain = DecodeChain()
    ca = chain.add(a["x"], a["qweight"], a["scales"], a["zeros"], 4, False)
    cb = chain.add(ca[:, 256:768], b["qweight"], b["scales"], b["zeros"], 4, False)
    cc = chain.add(cb, c["qweight"], c["scales"], c["zeros"], 4, False)
    chain.build()
    for rep in range(3):
        chain.launch(); chain.check()
        assert _same(ca, ya) and _same(cb, yb) and _same(cc, yc), f"rep {rep}"
