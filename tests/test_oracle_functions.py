"""CPU: the numpy oracle of the `functions_cuda` conversions against the reference's own known-answer vectors
(reference tests/functions/test_quant_ops.py:124-144 and :199-221)."""
import numpy as np

from oracle import functions as F


def test_sign_unpack_known_answer():
    # test_quant_ops.py:124-144: bytes 0, 16, 35, 255 -> signs with the bit order reversed w.r.t. reading order (LSB first)
    emd = np.broadcast_to(np.array([0, 16, 35, 255], dtype=np.uint8), (2, 16, 4))
    scale = np.random.default_rng(0).random((2, 16, 1)).astype(np.float32)
    exp_last = np.array([-1] * 8 + [-1, -1, -1, +1, -1, -1, -1, -1][::-1] + [-1, -1, +1, -1, -1, -1, +1, +1][::-1] + [1] * 8,
                        dtype=np.float32)
    exp = np.broadcast_to(exp_last, (2, 16, 32)) * scale
    assert np.array_equal(F.uint8_to_unpacked_tensor(emd, scale), exp)


def test_sign_pack_is_the_inverse_of_unpack():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((7, 64)).astype(np.float32)
    x[0, :4] = [0.0, -0.0, np.nan, -1e-30]
    packed = F.tensor_pack_to_uint8(x)
    assert packed.dtype == np.uint8 and packed.shape == (7, 8)
    assert packed[0, 0] & 0b1111 == 0b0011                       # 0.0 and -0.0 are >= 0, NaN and negatives are not
    signs = F.uint8_to_unpacked_tensor(packed[None], np.ones((1, 7, 1), np.float32))[0]
    assert np.array_equal(signs > 0, x >= 0)


def test_q4_pack_unpack_round_trip():
    # test_quant_ops.py:199-221: random ints in [-8, 7]; the low four bits survive, the scaled unpack restores the value
    rng = np.random.default_rng(2)
    for _ in range(10):
        x = rng.integers(-8, 8, size=(10, 10)).astype(np.int32)
        packed = F.q4_pack(x)
        assert packed.dtype == np.int8 and packed.size * 2 == x.size
        assert np.array_equal(F.q4_unpack(packed) & 15, x & 15)
        scaled = F.q4_unpack_and_scaling(packed, 0.045)
        assert np.array_equal(np.rint(scaled / np.float32(0.045)).astype(np.int32), x)
    assert F.q4_pack(np.array([[0xA, 0x5]], dtype=np.int32)).view(np.uint8)[0, 0] == 0xA5     # first code -> high nibble
