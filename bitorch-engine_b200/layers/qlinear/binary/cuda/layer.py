"""BinaryLinearCuda + autograd Function (twin of bitorch_engine/layers/qlinear/binary/cuda/layer.py:25-284)."""
import math
import typing

import torch
from torch.autograd import Function

from .bmm import BMM
from ..layer import BinaryLinearBase, BinaryLinearParameter
from .....extensions import binary_linear_cuda
from .....utils.model_helper import flatten_x, unflatten_x
from .....utils.quant_operators import nv_tensor_quant, init_weight


class BinaryLinearForward(Function):
    @staticmethod
    def forward(ctx, input, weight, bmm_type, scale_a, scale_w, is_train):
        x2, lead = flatten_x(input)
        if is_train:
            ctx.save_for_backward(x2, weight, scale_w, scale_a)
        out = binary_linear_cuda.forward(x2, weight, bmm_type, True).to(x2.dtype)
        return unflatten_x(out, lead) * scale_a * scale_w

    @staticmethod
    @typing.no_type_check
    def backward(ctx, output_gradient):
        # straight-through estimator with clipping, dense fp matmuls (layer.py:65-123): torch / cuBLAS as in the reference
        dy, lead = flatten_x(output_gradient)
        x2, weight, scale_w, scale_a = ctx.saved_tensors
        wt = weight.type(dy.dtype)
        grad_input = dy.mm(wt.sign() * scale_w)
        x_sign = x2.sign()
        grad_weight = dy.t().mm(x_sign * scale_a)
        ratio = x2 / scale_a
        inside = 1.0 - (ratio < -1).float() - (ratio > 1).float()
        grad_input.mul_(inside)
        grad_scale_a = torch.sum(grad_input * x_sign * (1.0 / math.sqrt(x2.numel())))
        grad_weight = nv_tensor_quant(grad_weight)[0]
        if isinstance(weight, BinaryLinearParameter) and not weight.requires_grad:
            weight.privileged_grad = grad_weight          # stock torch: integer weights cannot receive .grad
            weight._b200bit_grad_fresh = True
            grad_weight = None
        return unflatten_x(grad_input, lead), grad_weight, None, grad_scale_a, None, None


class BinaryLinearCuda(BinaryLinearBase):
    def __init__(self, *args, bmm_type: BMM = BMM.ADAPTIVE, **kwargs):
        super().__init__(*args, **kwargs)
        self.bits_binary_word = 8
        self.bmm_type = bmm_type
        self.bias_a = torch.nn.Parameter(torch.zeros(self.input_features, dtype=self.dtype))
        self.scale_a = torch.nn.Parameter(torch.tensor(0, dtype=self.dtype))
        self.register_buffer("scale_w", torch.tensor(1, dtype=self.dtype))

    def prepare_params(self) -> None:
        self.weight, self.scale_w.data = init_weight(self.weight, cls=BinaryLinearParameter)

    def generate_quantized_weight(self, qweight_only: bool = False) -> None:
        self.qweight = torch.nn.Parameter(binary_linear_cuda.w_pack(self.weight, self.bmm_type.value, True),
                                          requires_grad=False)
        if qweight_only:
            self.weight = None

    @staticmethod
    def w_pack(weights: torch.Tensor, bmm_type: BMM) -> torch.Tensor:
        return binary_linear_cuda.w_pack(weights, bmm_type.value, True)

    def set_activation(self, x: torch.Tensor) -> torch.Tensor:
        if not self.scale_a.is_nonzero():
            scale = (2 * x.abs().mean()) if self.symmetric else (4 * x.abs().mean())
            self.scale_a.data = scale.to(self.dtype)
        return x + self.bias_a.expand_as(x)

    def set_weight_data(self, x: torch.Tensor) -> None:
        super().set_weight_data(x)
        self.prepare_params()

    def forward(self, x: torch.Tensor, bmm_type: BMM = BMM.ADAPTIVE) -> torch.Tensor:
        self._check_forward(x)
        if self.bmm_type is not bmm_type:
            self.bmm_type = bmm_type
        if self.bmm_type is BMM.BTC32:
            m, k, n = x.size(dim=0), x.size(dim=1), self.output_features
            if m % 8 != 0 or k % 128 != 0 or n % 8 != 0:
                raise Exception("Invalid matrix dimensions for bit-tensorcore (BTC) kernel m:{}, n:{}, k:{}. "
                                "Guidelines: m and n must be multiplies of 8, and k must be multiplies of 128.".format(m, n, k))
        x = self.set_activation(x)
        return BinaryLinearForward.apply(x, self.opt_weight, self.bmm_type.value, self.scale_a, self.scale_w,
                                         self.training)
