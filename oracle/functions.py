"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's `functions_cuda` wire-format conversions
(bitorch_engine/functions/cuda/functions_cuda_kernel.cu).  The reference implements them in CUDA only, so this oracle is
pinned by (i) the reference's own known-answer test (tests/functions/test_quant_ops.py:124-144: bytes 0, 16, 35, 255 ->
signs, LSB first) and round-trip test (:199-221), restated in tests/test_oracle_functions.py, and (ii) on the GPU box by
the reference extension itself compiled unmodified for sm_100a (oracle/build_ref.py functions_cuda ->
oracle/_ref/functions_cuda, tests/test_gpu_functions.py)."""
import numpy as np


def q4_pack(codes):
    """q4_bit_packing_kernel (:136-159): out[i] = (in[2i] & 0xF) << 4 | (in[2i+1] & 0xF), as int8; last dim halves."""
    c = np.asarray(codes).astype(np.int64) & 15
    assert c.shape[-1] % 2 == 0
    return ((c[..., 0::2] << 4) | c[..., 1::2]).astype(np.uint8).view(np.int8)


def q4_unpack(packed):
    """q4_bit_unpacking_kernel (:162-182): high nibble first, unsigned 0..15, int32."""
    b = np.asarray(packed).view(np.uint8).astype(np.int32)
    out = np.empty(b.shape[:-1] + (b.shape[-1] * 2,), dtype=np.int32)
    out[..., 0::2] = b >> 4
    out[..., 1::2] = b & 15
    return out


def q4_unpack_and_scaling(packed, scale):
    """q4_bit_unpacking_scaling_kernel (:185-209): codes > 7 wrap to negative, times float32 scale."""
    u = q4_unpack(packed)
    s = np.where(u > 7, u - 16, u).astype(np.float32)
    return s * np.float32(scale)


def tensor_pack_to_uint8(x):
    """_to_uint8_array / bit_packing (:74-119): bit i of byte j = (x[.., 8j+i] >= 0), LSB first (NaN -> 0)."""
    x = np.asarray(x)
    bits = (x >= 0).astype(np.uint8)
    assert bits.shape[-1] % 8 == 0
    b = bits.reshape(bits.shape[:-1] + (bits.shape[-1] // 8, 8))
    return (b << np.arange(8, dtype=np.uint8)).sum(axis=-1).astype(np.uint8)


def uint8_to_unpacked_tensor(emd, scale):
    """unpack_uint8_to_float (:123-133): out[.., 8j+bit] = (+1 if bit set else -1) * scale[.., 0]."""
    e = np.asarray(emd, dtype=np.uint8)
    bits = (e[..., None] >> np.arange(8, dtype=np.uint8)) & 1
    sign = np.where(bits == 1, np.float32(1), np.float32(-1)).reshape(e.shape[:-1] + (e.shape[-1] * 8,))
    return sign * np.asarray(scale, dtype=np.float32)
