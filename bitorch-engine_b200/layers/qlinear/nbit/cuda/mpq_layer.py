"""MPQLinearCuda + its autograd Function: the reference surface (bitorch_engine/layers/qlinear/nbit/cuda/mpq_layer.py)
on top of the sm_100a kernels.  Same constructor, buffers, state_dict, attributes stamped on qweight, forward/backward
argument lists and return arity (10)."""
import math
import typing

import torch
from torch.autograd import Function

from .. import MPQLinearBase
from ..... import _cabi
from .....extensions import q_linear_cuda
from .....utils.model_helper import flatten_x, unflatten_x


def _stamp(qweight, scales, zeros, g_idx, w_bit, asym) -> None:
    qweight.scales, qweight.zeros, qweight.g_idx = scales, zeros, g_idx
    qweight.w_bit, qweight.asym, qweight.layer_type = w_bit, asym, 1


class MPQLinearCudaFunction(Function):
    """forward: y = x @ dequant(qweight); backward: grad_input through the same packed weight, dense weight gradient
    into qweight.privileged_grad (mpq_layer.py:14-120).  Unlike the reference there is no 32-row cliff: every M runs
    a fused kernel (the reference dequantises the whole matrix with torch ops and calls cuBLAS above 32 rows, :59-63)."""

    @staticmethod
    def forward(ctx, x, qweight, a_bit, w_bit, scales, zeros, g_idx, asym, is_training, privileged_grad=None):
        x2, lead = flatten_x(x)
        out = q_linear_cuda.mpq_forward(x2, qweight, scales, zeros, g_idx, a_bit, w_bit, asym,
                                        pdl=(not is_training) and _cabi.default_pdl())
        if is_training:
            qweight.privileged_grad = privileged_grad
            _stamp(qweight, scales, zeros, g_idx, w_bit, asym)
            ctx.a_bit = a_bit
            ctx.save_for_backward(x2, qweight)
        return unflatten_x(out, lead)

    @staticmethod
    @typing.no_type_check
    def backward(ctx, output_gradient):
        dy, lead = flatten_x(output_gradient)
        x2, qweight = ctx.saved_tensors
        wants_wgrad = qweight.requires_grad or getattr(qweight, "trainable", False)
        if wants_wgrad:
            assert qweight.privileged_grad is not None, \
                "The previledge gradient of qweight can not be None in backward pass."
        dy = dy.to(x2.dtype)
        grad_input = q_linear_cuda.mpq_grad_input(qweight.data, qweight.scales, qweight.zeros, qweight.g_idx, dy,
                                                  ctx.a_bit, qweight.w_bit, qweight.asym)
        if wants_wgrad:
            qweight.privileged_grad = x2.t().mm(dy)
            qweight._b200bit_grad_fresh = True      # DiodeMix.step consumes the mark (stock torch: no .grad on int32)                      # [K,M] @ [M,N], cuBLAS (mpq_layer.py:116)
        grad_q = qweight if qweight.requires_grad else None              # the reference returns the parameter itself
        return unflatten_x(grad_input, lead), grad_q, None, None, None, None, None, None, None, None


class MPQLinearCuda(MPQLinearBase):
    """W{1,2,4,8} x A16 Linear on CUDA (mpq_layer.py:123-224)."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.qweight.layer_type = 1
        self.check_parameters()

    def check_parameters(self) -> None:
        assert self.w_bit in [1, 2, 4, 8], f"The value of w_bit ({self.w_bit}) must be 1, 2, 4 or 8."
        assert self.a_bit == 16, f"The value of a_bit ({self.a_bit}) must be 16."

    def prepare_params(self) -> None:
        """Fold the double-quantised statistics into per-group scales / zeros and drop the load-only buffers
        (mpq_layer.py:163-204)."""
        try:
            if self.use_gba_quant:
                if self.group_size < 256:
                    shape = (math.ceil(self.in_channels / self.group_size), self.out_channels)
                    if self.asym:
                        qscales = self.qscales.unsqueeze(-1) if self.w_bit == 2 else self.qscales
                        self.zeros = self.qzeros
                    else:
                        stat = self.qstatistic.to(torch.uint8)
                        qscales, qzeros = stat >> 4, stat & 0x0F
                        self.zeros = ((qzeros.to(self.dtype) - self.qzeros_zeros) * self.qzeros_scales).view(shape)
                    self.scales = ((qscales.to(self.dtype) - self.qscales_zeros) * self.qscales_scales).view(shape)
                for name in ("qscales_zeros", "qscales_scales"):
                    delattr(self, name)
                for name in (("qscales",) if self.asym else ("qstatistic", "qzeros_zeros", "qzeros_scales")):
                    delattr(self, name)
            else:
                self.zeros = self.qzeros
            if self.disable_bias:
                del self.bias
            del self.wf
        except Exception as e:  # same wrapping as the reference
            raise RuntimeError(f"Error occurred during parameter preparation in MPQLinearCuda layer: {e}")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if any(t.device != x.device for t in (self.qweight, self.scales, self.zeros, self.g_idx)):
            raise RuntimeError("Some tensors are not on the correct device, please make sure to move the layer to "
                               "the correct device and call 'finalize_quantized_layers'.")
        out = MPQLinearCudaFunction.apply(x, self.qweight, self.a_bit, self.w_bit, self.scales, self.zeros, self.g_idx,
                                          self.asym, self.training, self.privileged_grad)
        if not self.disable_bias:
            out = out + self.bias
        return out
