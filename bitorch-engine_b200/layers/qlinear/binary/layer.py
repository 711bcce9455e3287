"""Binary Linear base + parameter type (twin of bitorch_engine/layers/qlinear/binary/layer.py)."""
import math

import torch
from torch import nn


class BinaryLinearParameter(nn.Parameter):
    """int8 (+-1 valued after prepare_params) weight whose update rule is DiodeMix's sign descent
    (binary/layer.py:8-60)."""

    def __new__(cls, data=None, requires_grad: bool = True):
        from ....utils import TORCH_INT_GRADIENTS
        want = bool(requires_grad)
        integer = data is not None and not torch.is_floating_point(data)
        obj = super().__new__(cls, data, requires_grad=want and (TORCH_INT_GRADIENTS or not integer))
        obj.trainable = want
        return obj

    @staticmethod
    def update(qweight, exp_avg_s=None, exp_avg_l=None, step=None, lr=1e-4, weight_decay=0.0, beta1=0.99, beta2=0.9999,
               eps=1e-6, dtype=torch.half, correct_bias=None, projector=None, grad=None) -> None:
        if not isinstance(qweight, BinaryLinearParameter):
            raise TypeError("qweight must be a BinaryLinearParameter")
        from ....optim.update import qweight_update_fn
        qweight_update_fn(qweight=qweight, exp_avg_s=exp_avg_s, exp_avg_l=exp_avg_l, step=step, lr=lr,
                          weight_decay=weight_decay, beta1=beta1, beta2=beta2, correct_bias=correct_bias, eps=eps,
                          dtype=dtype, projector=projector, grad=grad)


class BinaryLinearBase(nn.Module):
    """weight [out, in] (fp at construction, int8 after prepare_params), optional packed qweight for inference
    (binary/layer.py:63-231)."""

    def __init__(self, input_features: int, out_features: int, device: torch.device = None,
                 dtype: torch.dtype = torch.float, symmetric: bool = True) -> None:
        super().__init__()
        self.bits_binary_word = 8
        self.input_features, self.output_features = input_features, out_features
        self.qweight = None
        self.device, self.dtype, self.symmetric = device, dtype, symmetric
        self.reset_parameters()

    def reset_parameters(self) -> None:
        self.weight = nn.Parameter(torch.empty(self.output_features, self.input_features, dtype=self.dtype))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def set_weight_data(self, x: torch.Tensor) -> None:
        assert self.dtype == x.dtype, "dtype mismatch. Expected: '{}', but '{}' found".format(torch.float, x.dtype)
        self.weight = nn.Parameter(x)

    def prepare_params(self) -> None:
        raise NotImplementedError("Subclasses should implement this method.")

    def set_quantized_weight_data(self, x: torch.Tensor) -> None:
        self.qweight = nn.Parameter(x, requires_grad=False)

    def generate_quantized_weight(self, qweight_only: bool = False) -> None:
        raise NotImplementedError("Subclasses should implement this method.")

    def _check_forward(self, x: torch.Tensor) -> None:
        packed_in = x.dtype is torch.uint8
        last = x.size(dim=-1)
        if not packed_in:
            assert last % self.bits_binary_word == 0, \
                "Input tensor dimension ({}) must be divisible by {}.".format(last, self.bits_binary_word)
        if self.qweight is not None:
            per = 1 if packed_in else self.bits_binary_word
            expect = last * self.output_features / per
            assert self.qweight.nelement() == expect, \
                "Weight and input tensor mismatch. {}:{}".format(self.qweight.nelement(), expect)
        elif packed_in:
            assert self.weight.size(dim=1) / self.bits_binary_word == last, "Weight and input tensor mismatch."
        else:
            assert self.weight.size(dim=1) == last, "Weight and input tensor mismatch."

    @property
    def opt_weight(self) -> nn.Parameter:
        if not self.training and self.qweight is None:
            self.generate_quantized_weight()
        return self.weight if self.training else self.qweight

    def set_bits_binary_word(self, num_bit: int) -> None:
        self.bits_binary_word = num_bit
