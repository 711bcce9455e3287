"""CPU: pin the numpy oracle (oracle/nbit.py) against golden vectors produced by the reference's own Python
(oracle/gen_golden.py -> tests/golden/nbit_cases.npz).  Bit-exact for unpack / pack; exact fp values for W."""
import numpy as np
import pytest

from oracle import nbit
from helpers import load_nbit_cases

CASES = load_nbit_cases()


@pytest.mark.parametrize("c", CASES, ids=[c.id for c in CASES])
def test_dequant_matches_reference_unpack_qweight(c):
    zeros = c.zeros if c.asym else c.f("zeros")
    W = nbit.dequant_w16(c.qweight, c.f("scales"), zeros, c.g_idx, c.w_bit, c.asym, c.dt, style="python")
    assert np.array_equal(W, c.f("W")), "oracle W differs from reference unpack_qweight"


@pytest.mark.parametrize("c", CASES, ids=[c.id for c in CASES])
def test_forward_and_grad_input_match_reference(c):
    zeros = c.zeros if c.asym else c.f("zeros")
    y = nbit.mpq_forward(c.f("x"), c.qweight, c.f("scales"), zeros, c.g_idx, c.w_bit, c.asym, c.dt)
    dx = nbit.mpq_grad_input(c.f("dy"), c.qweight, c.f("scales"), zeros, c.g_idx, c.w_bit, c.asym, c.dt)
    # reference values were accumulated in fp32 by torch.matmul; the oracle accumulates in fp64
    np.testing.assert_allclose(y, c.y, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(dx, c.dx, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("c", CASES, ids=[c.id for c in CASES])
def test_pack_fp_weight_bit_exact(c):
    zeros = c.zeros if c.asym else c.f("zeros")
    packed = nbit.pack_fp_weight(c.f("Wp"), c.f("scales"), zeros, c.g_idx, c.w_bit, c.asym, c.dt)
    assert np.array_equal(packed, c.packed)
    # pack(unpack(q)) == q iff the reference said so
    rt = nbit.pack_fp_weight(c.f("W"), c.f("scales"), zeros, c.g_idx, c.w_bit, c.asym, c.dt)
    assert bool(c.roundtrip_equal) == bool(np.array_equal(rt, c.qweight))


@pytest.mark.parametrize("w_bit", [1, 2, 4, 8])
def test_pack_unpack_int_roundtrip(w_bit):
    rng = np.random.default_rng(w_bit)
    codes = rng.integers(0, 1 << w_bit, size=(256, 32))
    assert np.array_equal(nbit.unpack_int(nbit.pack_int(codes, w_bit), w_bit), codes)
    zq = rng.integers(1, (1 << w_bit) + 1, size=(4, 64))
    assert np.array_equal(nbit.unpack_zeros_asym(nbit.pack_zeros_asym(zq, w_bit), w_bit), zq)


def test_exact_vs_w16_distance_is_small():
    """The exact-math oracle and the reference-rounded one differ only by the fp16 rounding of W."""
    c = [c for c in CASES if c.w_bit == 4 and c.dt == "f16" and not c.asym and not c.act_order][0]
    y16 = nbit.mpq_forward(c.f("x"), c.qweight, c.f("scales"), c.f("zeros"), c.g_idx, 4, False, "f16")
    yex = nbit.mpq_forward_exact(c.f("x"), c.qweight, c.f("scales"), c.f("zeros"), c.g_idx, 4, False)
    rel = np.linalg.norm(y16 - yex) / np.linalg.norm(yex)
    assert rel < 2e-3
