"""CPU parity of the host binary Linear (extensions/binary_linear_cpp -> csrc/binary_cpu.cpp) with the reference's CPU
extension: exact equality with sign(x) @ sign(w).T (the arithmetic of bitorch's QLinear(sign, sign)), with
the numpy oracle's packed layout (pinned against the reference extension in tests/test_oracle_binary.py) and -- when
oracle/_ref holds it -- with `binary_linear_cpp` compiled unmodified from the reference sources.  Mirrors
/root/reference/tests/layers/test_binary_linear.py:222-310 (packed vs unpacked equalities, python get_binary_col vs cpp)."""
import os

import numpy as np
import pytest
import torch

from oracle import binary as ob

SHAPES = [(1, 8, 8), (3, 64, 16), (5, 72, 10), (128, 512, 1000), (33, 576, 64), (7, 1152, 129)]


@pytest.mark.parametrize("m,k,n", SHAPES)
def test_forward_and_pack_match_the_oracle(m, k, n):
    from bitorch_engine_b200.extensions import binary_linear_cpp
    g = torch.Generator().manual_seed(m * 1000 + k + n)
    x = torch.randn((m, k), generator=g)
    w = torch.randn((n, k), generator=g)
    x[0, 0] = 0.0          # sign(0) = +1
    w[0, 1] = -0.0
    ref = ob.forward(x.numpy(), w.numpy()).astype(np.float32)
    y = binary_linear_cpp.forward(x, w, m, n, k)
    assert y.dtype == torch.float32 and tuple(y.shape) == (m, n) and np.array_equal(y.numpy(), ref)
    packed = binary_linear_cpp.w_pack(w, n, k)
    assert packed.dtype == torch.uint8 and np.array_equal(packed.numpy(), ob.pack_cpp(w.numpy()))
    assert np.array_equal(binary_linear_cpp.forward(x, packed, m, n, k).numpy(), ref)


def test_layer_surface_and_errors():
    from bitorch_engine_b200.layers.qlinear.binary.cpp import BinaryLinearCPP
    from bitorch_engine_b200.extensions import binary_linear_cpp
    layer = BinaryLinearCPP(64, 24)
    x = torch.randn(5, 64)
    layer.train()
    y_train = layer(x)
    layer.eval()
    y_eval = layer(x)                                  # packs the weight on first use (opt_weight)
    assert layer.qweight.dtype == torch.uint8 and layer.qweight.numel() == 64 * 24 // 8
    ref = ob.forward(x.numpy(), layer.weight.detach().numpy()).astype(np.float32)
    assert np.array_equal(y_train.detach().numpy(), ref) and np.array_equal(y_eval.numpy(), ref)
    layer.generate_quantized_weight(qweight_only=True)
    assert layer.weight is None and np.array_equal(layer(x).numpy(), ref)
    with pytest.raises(RuntimeError):
        binary_linear_cpp.forward(x.half(), torch.randn(24, 64), 5, 24, 64)        # fp32 only, as the reference
    with pytest.raises(ValueError):
        binary_linear_cpp.forward(x, torch.randn(24, 32), 5, 24, 64)


def test_against_the_compiled_reference_extension():
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "binary_linear_cpp",
                      "binary_linear_cpp_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/binary_linear_cpp not built")
    import importlib.util
    spec = importlib.util.spec_from_file_location("binary_linear_cpp_ref", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from bitorch_engine_b200.extensions import binary_linear_cpp
    g = torch.Generator().manual_seed(7)
    for m, k, n in ((128, 512, 1000), (16, 576, 64)):
        x, w = torch.randn((m, k), generator=g), torch.randn((n, k), generator=g)
        assert torch.equal(binary_linear_cpp.forward(x, w, m, n, k), ref.forward(x, w, m, n, k))
        pk = ref.w_pack(w, n, k)
        assert torch.equal(binary_linear_cpp.w_pack(w, n, k), pk)
        assert torch.equal(binary_linear_cpp.forward(x, pk, m, n, k), ref.forward(x, pk, m, n, k))
