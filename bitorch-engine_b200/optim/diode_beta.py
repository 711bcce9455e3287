"""DiodeMix optimizer (twin of bitorch_engine/optim/diode_beta.py:37-196): sign descent for binary weights, Adam on the
de-quantised weight followed by re-quantisation for n-bit MPQ weights, plain AdamW-style update for everything else.
State layout and hyper-parameters are the reference's; the quantised branches run as fused kernels (optim/update.py)."""
import math
from typing import Callable, Iterable, Tuple

import torch
from torch import nn
from torch.optim import Optimizer

from .galore_projector import GaLoreProjector


class DiodeMix(Optimizer):
    def __init__(self, params: Iterable[nn.parameter.Parameter], lr: float = 1e-4,
                 betas: Tuple[float, float] = (0.99, 0.9999), eps: float = 1e-6, weight_decay: float = 0.0,
                 correct_bias: bool = True, dtype: torch.dtype = torch.float):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr} - should be >= 0.0")
        for i, b in enumerate(betas):
            if not 0.0 <= b < 1.0:
                raise ValueError(f"Invalid beta parameter: {b} - should be in [0.0, 1.0)")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps} - should be >= 0.0")
        self.dtype = dtype
        super().__init__(params, {"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay,
                                  "correct_bias": correct_bias})

    @torch.no_grad()
    def step(self, closure: Callable = None):
        from ..layers.qlinear.nbit import MPQWeightParameter
        from ..layers.qlinear.binary import BinaryLinearParameter
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                is_mpq = isinstance(p, MPQWeightParameter)
                is_bin = isinstance(p, BinaryLinearParameter)
                # The reference skips a parameter whose .grad is None (diode_beta.py:118-119).  On stock torch an integer
                # weight never receives .grad: backward() then leaves the weight gradient in p.privileged_grad and marks
                # it fresh, and step() consumes the mark -- a layer that did not run backward in this iteration is
                # skipped exactly as in the reference (neither the forward pass's placeholder nor a gradient of an
                # earlier iteration is ever applied).
                if is_mpq:
                    fresh = getattr(p, "_b200bit_grad_fresh", False) or p.grad is not None
                    grad = p.privileged_grad if fresh else None
                elif is_bin and p.grad is None:
                    grad = getattr(p, "privileged_grad", None) if getattr(p, "_b200bit_grad_fresh", False) else None
                else:
                    grad = p.grad
                if grad is None:
                    continue
                if is_mpq or is_bin:
                    p._b200bit_grad_fresh = False
                if grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                state = self.state[p]
                if "step" not in state:
                    state["step"] = torch.zeros(1)
                projector = None
                if "rank" in group:
                    if "projector" not in state:
                        state["projector"] = GaLoreProjector(group["rank"], update_proj_gap=group["update_proj_gap"],
                                                             scale=group["scale"], proj_type=group["proj_type"])
                    projector = state["projector"]
                    grad = projector.project(grad.to(self.dtype), state["step"].item())
                if "exp_avg_s" not in state:
                    if isinstance(p, BinaryLinearParameter):
                        delta = torch.rand_like(p, dtype=self.dtype).mul_(1e-3)
                        state["exp_avg_l"] = torch.zeros_like(p, dtype=self.dtype)
                        state["exp_avg_s"] = -(p.data.clone().sign_().to(self.dtype).mul_(delta))
                    else:
                        state["exp_avg_l"] = torch.zeros_like(grad, dtype=self.dtype)
                        state["exp_avg_s"] = torch.zeros_like(grad, dtype=self.dtype)
                if is_mpq or is_bin:
                    type(p).update(qweight=p, exp_avg_s=state["exp_avg_s"], exp_avg_l=state["exp_avg_l"],
                                   step=state["step"], lr=group["lr"], weight_decay=group["weight_decay"], beta1=beta1,
                                   beta2=beta2, correct_bias=group["correct_bias"], eps=group["eps"], dtype=self.dtype,
                                   projector=projector, grad=grad)
                    continue
                m, v, step = state["exp_avg_l"], state["exp_avg_s"], state["step"]
                step.add_(1)
                m.mul_(beta1).add_(grad, alpha=(1.0 - beta1))
                v.mul_(beta2).addcmul_(grad, grad, value=1.0 - beta2)
                denom = v.sqrt().add_(group["eps"])
                step_size = group["lr"]
                if group["correct_bias"]:
                    n = step.item()
                    step_size = step_size * math.sqrt(1.0 - beta2 ** n) / (1.0 - beta1 ** n)
                norm_grad = m / denom
                if projector is not None:
                    norm_grad = projector.project_back(norm_grad)
                p.add_(norm_grad, alpha=-step_size)
                if group["weight_decay"] > 0.0:
                    p.add_(p, alpha=(-group["lr"] * group["weight_decay"]))
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        """Also drops the privileged gradients of quantised parameters (stock torch keeps them outside .grad)."""
        super().zero_grad(set_to_none=set_to_none)
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "_b200bit_grad_fresh", False):
                    p._b200bit_grad_fresh = False
                    p.privileged_grad = None
