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
