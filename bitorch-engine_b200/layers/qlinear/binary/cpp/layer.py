"""BinaryLinearCPP: the CPU binary Linear of the reference (bitorch_engine/layers/qlinear/binary/cpp/layer.py:18-125) on
the host kernels behind extensions/binary_linear_cpp.  Inference only, fp32 activations, as the reference."""
import torch
from torch.autograd import Function

from ..layer import BinaryLinearBase
from .....extensions import binary_linear_cpp
from .....utils.model_helper import flatten_x, unflatten_x


class BinaryLinearForward(Function):
    @staticmethod
    def forward(ctx, input: torch.Tensor, weights: torch.Tensor, m: int, n: int, k: int) -> torch.Tensor:
        x2, lead = flatten_x(input)
        return unflatten_x(binary_linear_cpp.forward(x2, weights, m, n, k), lead)


class BinaryLinearCPP(BinaryLinearBase):
    def __init__(self, input_features: int, out_features: int, device: torch.device = None) -> None:
        super().__init__(input_features, out_features, device)

    def prepare_params(self) -> None:
        pass

    def generate_quantized_weight(self, qweight_only: bool = False) -> None:
        self.qweight = binary_linear_cpp.w_pack(self.weight, self.output_features, self.input_features)
        if qweight_only:
            self.weight = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._check_forward(x)
        m, k, n = x.size(dim=0), x.size(dim=1), self.output_features
        return BinaryLinearForward.apply(x, self.opt_weight, m, n, k)
