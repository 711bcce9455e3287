"""Torch-level quantisation helpers the binary Linear path uses (twin of the relevant part of
bitorch_engine/utils/quant_operators.py:7-90 and utils/model_helper.py:286-327).  Dense fp glue, left in torch on purpose
(SURVEY.md section 8a row a18)."""
from typing import Tuple, Type

import torch


def nv_tensor_quant(inputs, amax=None, num_bits=8, unsigned=False, narrow_range=True) -> Tuple[torch.Tensor, torch.Tensor]:
    """round(x * (2^(b-1)-1) / amax) clamped to the signed range; returns (quantised values, scale)
    (quant_operators.py:7-90; amax defaults to the tensor maximum, NOT the max magnitude, as in the reference)."""
    if isinstance(amax, torch.Tensor) and inputs.dim() != amax.dim():
        raise ValueError("amax %s has different shape than inputs %s. Make sure broadcast works as expected!",
                         amax.size(), inputs.size())
    if amax is None:
        amax = torch.amax(inputs, keepdim=True)
    if unsigned and inputs.min() < 0.0:
        raise TypeError("Negative values encountered in unsigned quantization.")
    in_dtype = inputs.dtype
    x = inputs.float() if in_dtype in (torch.bfloat16, torch.float16) else inputs
    if amax.dtype in (torch.bfloat16, torch.float16):
        amax = amax.float()
    lowest = amax.min()
    if lowest < 0:
        raise ValueError("Negative values in amax")
    top = torch.tensor((2.0 ** (num_bits - 1 + int(unsigned))) - 1.0, device=x.device)
    bottom = 0 if unsigned else (-top if narrow_range else -top - 1)
    scale = top / amax
    out = torch.clamp((x * scale).round_(), bottom, top)
    tiny = 1.0 / (1 << 24)
    if lowest <= tiny:
        scale[amax <= tiny] = 1.0
    if in_dtype in (torch.bfloat16, torch.float16):
        out = out.to(in_dtype)
    return out, scale


def init_weight(weight: torch.Tensor, cls: Type[torch.nn.Parameter] = torch.nn.Parameter):
    """fp weight -> (sign-preserving int8 parameter, mean-magnitude scale)  (model_helper.py:286-327)."""
    w = weight if weight.dtype == torch.float else weight.to(torch.float)
    scale_w = w.norm(p=1).div(w.nelement()).to(weight.device)
    centred = w - w.mean()
    q = nv_tensor_quant(centred)[0]
    q = torch.where(q == 0, centred.sign(), q)
    return cls(q.to(torch.int8)), scale_w


# ---------------------------------------------------------------------------------------------------------------
# Bit-packing specifications in plain Python and activation quantisers (quant_operators.py:93-307, 348-368).  Host-side
# helpers either side of the kernels: the packers are the written-down format of the CPU binary layer's words, the
# q8 / q4 quantisers are the activation side of the W4A4 / W8A8 layers, the zero-point packer is the inverse of what the
# fused optimizer kernel unpacks.
# ---------------------------------------------------------------------------------------------------------------
def bit_set(var: int, pos: int, val: int) -> int:
    """var | (val << pos)   (quant_operators.py:93-115)."""
    return var | (val << pos)


def get_binary_row(nd_row, binary_row, nd_size: int, bits_per_binary_word: int):
    """Sign-bit words of a flat row-major array: word w, bit j = (nd_row[w * bits + j] >= 0), LSB first
    (quant_operators.py:118-173)."""
    for w in range(0, nd_size, bits_per_binary_word):
        word = 0
        for j in range(bits_per_binary_word):
            word = bit_set(word, j, 1 if nd_row[w + j] >= 0 else 0)
        binary_row[w // bits_per_binary_word] = word
    return binary_row


def get_binary_col(nd_col, binary_col, dim_n: int, dim_k: int, bits_per_binary_word: int):
    """Sign-bit words down the columns of a flat [dim_n, dim_k] array: word (y, x), bit b = (nd_col[(y * bits + b) * dim_k
    + x] >= 0)   (quant_operators.py:176-231)."""
    for y in range(dim_n // bits_per_binary_word):
        for x in range(dim_k):
            word = 0
            for b in range(bits_per_binary_word):
                word = bit_set(word, b, 1 if nd_col[(y * bits_per_binary_word + b) * dim_k + x] >= 0 else 0)
            binary_col[y * dim_k + x] = word
    return binary_col


def _uniform_quant(input: torch.Tensor, scale_a, eps, default_divisor: float, lo: int, hi: int):
    scale_given = scale_a is not None
    x = input if input.dtype == torch.float else input.to(torch.float)
    if scale_a is None:
        scale_a = 2 * x.abs().mean() / default_divisor
    if eps is None:
        eps = torch.tensor(0.00001, dtype=x.dtype, device=x.device)
    scale_a = torch.where(scale_a > eps, scale_a, eps)
    q = (x / scale_a).round().clamp(lo, hi)
    return q if scale_given else (q, scale_a)


def q8_quantization(input: torch.Tensor, scale_a: torch.Tensor = None, eps: torch.Tensor = None):
    """round(x / max(scale, eps)) clamped to [-128, 127]; scale defaults to 2 * mean|x| / 11.269 and is then returned
    as well (quant_operators.py:234-269).  (The reference's own default for `eps` calls `.device(...)` on a tensor and
    raises; a float32 1e-5 on the input's device is used here.)"""
    return _uniform_quant(input, scale_a, eps, 11.269, -128, 127)


def q4_quantization(input: torch.Tensor, scale_a: torch.Tensor = None, eps: torch.Tensor = None):
    """4-bit twin: clamp to [-8, 7], default scale 2 * mean|x| / 5.6345 (quant_operators.py:272-307)."""
    return _uniform_quant(input, scale_a, eps, 5.6345, -8, 7)


def gptq_style_zeros_packing(zeros: torch.Tensor, w_bit: int, out_features: int, group_size: int) -> torch.Tensor:
    """Integer zero points [G, N] (value = stored field + 1) -> packed int32 [G, N * w_bit / 32], LSB first
    (quant_operators.py:348-368)."""
    per = 32 // w_bit
    z = (zeros.reshape(zeros.shape[0], out_features // 32 * w_bit, per).to(torch.int32) - 1) & ((1 << w_bit) - 1)
    shifts = torch.arange(0, 32, w_bit, device=zeros.device, dtype=torch.int32)
    return torch.bitwise_left_shift(z, shifts.view(1, 1, -1)).sum(dim=-1).to(torch.int32)


def gptq_style_unpacking(qweight):
    """(weights [K,N], zeros) of an MPQWeightParameter, the optimizer-side unpack (quant_operators.py:310-345).

      asym (packed qzeros)      : weights = scales[g] * (q - zq[g]),  zeros = integer zero points [G,N] (field + 1)
      sym, g_idx None (MBWQ)    : weights = q * scales - zeros, rows scattered through q_perm,
                                  zeros = the fp zeros repeated per row [K,N]
      sym with g_idx            : weights = q * scales[g] - zeros[g]; the reference returns an unbound name here
                                  (UnboundLocalError), this returns zeros = None
    CUDA tensors run the one-pass dequant kernel (bit-identical roundings, tests/test_gpu_mpq_aux.py); CPU tensors are
    unpacked with the reference's own torch arithmetic (host helper)."""
    w_bit, asym = qweight.w_bit, bool(qweight.asym)
    data = qweight.data
    g_idx = getattr(qweight, "g_idx", None)
    if asym:
        shifts = torch.arange(0, 32, w_bit, dtype=torch.int32, device=data.device)
        zq = torch.bitwise_right_shift(qweight.zeros.unsqueeze(2), shifts.view(1, 1, -1)) & ((1 << w_bit) - 1)
        zeros = (zq + 1).to(torch.int16 if w_bit == 8 else torch.int8).reshape(-1, data.size(-1))
    elif g_idx is None:
        rep = (data.size(0) * 32 // w_bit) // qweight.zeros.size(0)
        zeros = qweight.zeros.unsqueeze(1).repeat(1, rep, 1).view(-1, qweight.zeros.size(-1))
    else:
        zeros = None
    if data.is_cuda:
        from ..extensions import q_linear_cuda
        perm = None
        if not asym and g_idx is None:
            K = data.size(0) * 32 // w_bit
            perm = None if q_linear_cuda._perm_is_identity(qweight.q_perm, K) else qweight.q_perm
        weights = q_linear_cuda.mpq_dequant(data, qweight.scales, qweight.zeros, g_idx, w_bit, asym, perm=perm)
        return weights, zeros
    shifts = torch.arange(0, 32, w_bit, dtype=torch.int32)
    q = (torch.bitwise_right_shift(data.unsqueeze(1), shifts.view(1, -1, 1)) & ((1 << w_bit) - 1))
    q = q.to(torch.int16 if w_bit == 8 else torch.int8).view(-1, data.size(-1))
    if asym:
        gi = g_idx.long()
        weights = qweight.scales[gi] * (q - zeros[gi])
    elif g_idx is None:
        rep = q.size(0) // qweight.scales.size(0)
        scales = qweight.scales.unsqueeze(1).repeat(1, rep, 1).view(-1, qweight.scales.size(-1))
        weights = q.mul(scales) - zeros
        index = qweight.q_perm.unsqueeze(1).repeat(1, weights.size(1)).long()
        weights.scatter_(dim=0, index=index, src=weights.clone())
    else:
        gi = g_idx.long()
        weights = q * qweight.scales[gi] - qweight.zeros[gi]
    return weights, zeros
