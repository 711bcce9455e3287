"""CPU: host-side logic of the decode path -- (1) the launch plan of the 4-bit decode kernel (strip geometry, grid, ring
depth; b200bit_mpq_decode_plan, no device needed), (2) the shim's B200BIT_FLAG_INPUT_READY inference (tensor identity +
version counter)."""
import ctypes

import torch


def _plan(K, N, G, w_bit=4, asym=0, dtype=1):
    from bitorch_engine_b200 import _cabi
    out = (ctypes.c_int * 8)()
    assert _cabi.lib().b200bit_mpq_decode_plan(K, N, G, w_bit, asym, dtype, out) == 0
    keys = ("ok", "grid", "strips", "n28", "tiles", "S", "F", "smem")
    return dict(zip(keys, list(out)))


def test_llama7b_plans():
    a = _plan(4096, 4096, 32)
    # 4096 columns: 136 strips of 28 + 12 of 24 = 148 strips, one per SM; the whole K range (2 tiles) fits the ring
    assert a["ok"] and a["grid"] == 148 and a["strips"] == 148 and a["n28"] == 136 and 136 * 28 + 12 * 24 == 4096
    assert a["tiles"] == 2 and a["S"] == 2 and a["F"] == 4
    b = _plan(4096, 11008, 32)
    # rounding 394 strips up to 444 would re-fetch 13 % of the columns: keep 28-wide strips, dealt out cyclically
    assert b["ok"] and b["grid"] == 148 and b["strips"] == 394 and b["n28"] == 394 and b["tiles"] == 2 and b["S"] == 3
    c = _plan(11008, 4096, 86)
    assert c["ok"] and c["grid"] == 148 and c["strips"] == 148 and c["tiles"] == 6 and c["S"] == 3
    for p in (a, b, c):
        assert p["smem"] <= 112 * 1024                      # two CTAs per SM


def test_plan_envelope():
    assert _plan(2048, 1024, 64)["F"] == 1 and _plan(2048, 1024, 32)["F"] == 2 and _plan(2048, 1024, 16)["F"] == 4   # groups 32, 64, 128
    small = _plan(2048, 1024, 16)
    assert small["grid"] == small["strips"] == 37 and small["tiles"] == 1 and small["S"] == 1
    assert _plan(2048, 1024, 1)["ok"]                      # one group over the whole K (2048 = power of two)
    assert not _plan(11008, 4096, 1)["ok"]                 # group of 11008 values: not a power of two -> other kernels
    assert not _plan(4096, 4096, 32, w_bit=2)["ok"] and not _plan(4096, 4096, 32, dtype=0)["ok"]   # 2-bit / fp32
    assert not _plan(4096, 4100, 32)["ok"]                 # N % 8 != 0: TMA row stride not a multiple of 16 bytes
    assert _plan(4096, 4096, 32, asym=1)["ok"] and not _plan(4096, 4104, 32, asym=1)["ok"]


def test_input_ready_inference():
    from bitorch_engine_b200.extensions import q_linear_cuda as q
    q._prev_x.clear()
    x = torch.zeros((1, 64), dtype=torch.float16)
    assert q._input_ready(x, 7) is False                   # first call on the stream
    assert q._input_ready(x, 7) is True                    # same buffer, untouched: a sibling call
    assert q._input_ready(x.view(1, 64), 7) is True        # a view of the same storage
    x.mul_(2)                                              # torch-visible in-place write
    assert q._input_ready(x, 7) is False
    assert q._input_ready(x, 7) is True
    assert q._input_ready(x, 8) is False                   # another stream has its own history
    y = torch.zeros((1, 64), dtype=torch.float16)
    assert q._input_ready(y, 7) is False                   # a different buffer
    assert q._input_ready(x, 7) is False                   # ... and back: not the immediately preceding input


def test_shims_reject_host_tensors_like_the_reference():
    """CHECK_CUDA of the reference (q_linear_cuda.cpp:255-256, functions_cuda.cpp) -> RuntimeError, before any library call."""
    import pytest
    from bitorch_engine_b200.extensions import q_linear_cuda, functions_cuda, binary_linear_cuda
    x = torch.zeros((1, 256), dtype=torch.float16)
    qw = torch.zeros((32, 64), dtype=torch.int32)
    sc = torch.ones((2, 64), dtype=torch.float16)
    with pytest.raises(RuntimeError):
        q_linear_cuda.mpq_forward(x, qw, sc, sc, None, 16, 4, False)
    with pytest.raises(RuntimeError):
        functions_cuda.q4_pack(torch.zeros((2, 2), dtype=torch.int32), False)
    with pytest.raises(RuntimeError):
        functions_cuda.tensor_pack_to_uint8(torch.zeros((2, 8)))
    with pytest.raises(NotImplementedError):
        functions_cuda.fp32toint4(torch.zeros(64))
    assert hasattr(binary_linear_cuda, "forward") and hasattr(binary_linear_cuda, "w_pack")


def test_default_pdl_switch(monkeypatch):
    from bitorch_engine_b200 import _cabi
    monkeypatch.delenv("B200BIT_PDL", raising=False)
    assert _cabi.default_pdl() is True
    monkeypatch.setenv("B200BIT_PDL", "0")
    assert _cabi.default_pdl() is False
