"""Persistent format of the n-bit Linear: parameter type + base module
(twin of bitorch_engine/layers/qlinear/nbit/layer.py:8-119 `MPQWeightParameter`, :270-511 `MPQLinearBase`).
state_dict keys, shapes and dtypes are byte-compatible with the reference."""
import math

import torch
from torch import nn

from ....utils import TORCH_INT_GRADIENTS

_META = ("privileged_grad", "scales", "zeros", "g_idx", "w_bit", "asym", "group_size", "layer_type", "q_perm",
         "qscales_zeros", "qscales_scales", "qzeros_zeros", "qzeros_scales", "q_group_map", "rows")
_DEFAULTS = dict(w_bit=-1, asym=False, group_size=-1, layer_type=-1)


class MPQWeightParameter(nn.Parameter):
    """int32 packed weight + the quantisation metadata the kernels / optimizer need, carried as plain attributes
    (nbit/layer.py:8-83).  layer_type: 1 = MPQLinear, 2 = MBWQLinear.  On stock torch an integer tensor cannot
    require grad; `trainable` then records the caller's wish and the weight gradient travels via privileged_grad."""

    def __new__(cls, data=None, requires_grad: bool = True, **meta):
        want = bool(requires_grad)
        obj = super().__new__(cls, data, requires_grad=want and TORCH_INT_GRADIENTS)
        obj.trainable = want
        return obj

    def __init__(self, data=None, requires_grad: bool = True, **meta):
        unknown = set(meta) - set(_META)
        if unknown:
            raise TypeError(f"unexpected MPQWeightParameter attributes: {sorted(unknown)}")
        for name in _META:
            setattr(self, name, meta.get(name, _DEFAULTS.get(name)))

    @staticmethod
    def update(qweight, exp_avg_s=None, exp_avg_l=None, step=None, lr=1e-4, weight_decay=0.0, beta1=0.99,
               beta2=0.9999, eps=1e-6, dtype=torch.half, correct_bias=None, projector=None, grad=None) -> None:
        """Optimizer hook, same signature as nbit/layer.py:85-119; forwards to the fused update."""
        if not isinstance(qweight, MPQWeightParameter):
            raise TypeError("qweight must be an MPQWeightParameter")
        from ....optim.update import qweight_update_fn
        qweight_update_fn(qweight=qweight, exp_avg_s=exp_avg_s, exp_avg_l=exp_avg_l, step=step, lr=lr,
                          weight_decay=weight_decay, beta1=beta1, beta2=beta2, correct_bias=correct_bias, eps=eps,
                          dtype=dtype, projector=projector, grad=grad)


class MPQLinearBase(nn.Module):
    """Buffers and packed layouts of the mixed-precision (W n-bit, A 16-bit) Linear (nbit/layer.py:270-511).

    qweight int32 [K*w_bit/32, N]; g_idx int32 [K]; GPTQ style: qzeros int32 [G, N*w_bit/32], scales [G,N];
    GBA style: scales/zeros [G,N] + double-quantisation statistics that prepare_params() folds away."""

    def __init__(self, in_channels: int, out_channels: int, a_bit: int = 16, w_bit: int = 4, dtype=torch.half,
                 group_size=-1, use_gba_quant=True, dq_group_size=-1, dq_mode=2, disable_bias=True, asym=False,
                 requires_grad=True) -> None:
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.a_bit, self.w_bit, self.dtype = a_bit, w_bit, dtype
        self.maxq = 2 ** w_bit - 1
        self.group_size = group_size if group_size > -1 else in_channels
        self.asym, self.disable_bias = asym, disable_bias
        self.use_gba_quant, self.dq_group_size, self.dq_mode = use_gba_quant, dq_group_size, dq_mode
        self.requires_grad = requires_grad
        self.initialize()

    # -- allocation ------------------------------------------------------------------------------------------
    def initialize(self) -> None:
        K, N, b = self.in_channels, self.out_channels, self.w_bit
        self.qweight = MPQWeightParameter(torch.empty((K // 32 * b, N), dtype=torch.int32),
                                          requires_grad=self.requires_grad, w_bit=b, asym=self.asym,
                                          group_size=self.group_size)
        # weight-gradient slot handed to the autograd Function (nbit/layer.py:382).  The reference allocates a dense
        # [K,N] tensor per layer; here only a one-element LEAF is stored (deepcopy / peft / EMA copy the module like any
        # other) and the `privileged_grad` property hands out a stride-0 [K,N] view of it.  When integer tensors cannot
        # require grad the leaf doubles as the autograd hook that makes backward() run.
        self._grad_carrier = (torch.zeros((1,), dtype=self.dtype, requires_grad=not TORCH_INT_GRADIENTS)
                              if self.requires_grad else None)
        self.register_buffer("g_idx", (torch.arange(K, dtype=torch.int32) // self.group_size).contiguous())
        self.register_buffer("bias", torch.zeros((N,), dtype=self.dtype))
        self.register_buffer("wf", torch.arange(0, 32, b, dtype=torch.int32).unsqueeze(0))
        if self.use_gba_quant:
            self.init_gba()
        else:
            self.init_gptq()

    @property
    def privileged_grad(self):
        """[K, N] view of the gradient carrier (None for frozen layers); built on access, never stored."""
        c = self.__dict__.get("_grad_carrier")
        if c is None:
            return None
        return c.expand(self.in_channels, self.out_channels)

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .half(): move the carrier with the module (it is neither a Parameter nor a buffer)
        out = super()._apply(fn, *args, **kwargs)
        c = self.__dict__.get("_grad_carrier")
        if c is not None:
            moved = fn(c.detach())
            self._grad_carrier = moved.clone().requires_grad_(c.requires_grad) if moved.is_floating_point() else c
        return out

    def _groups(self) -> int:
        return math.ceil(self.in_channels / self.group_size)

    def init_gptq(self) -> None:
        G, N = self._groups(), self.out_channels
        self.register_buffer("qzeros", torch.zeros((G, N // 32 * self.w_bit), dtype=torch.int32))
        self.register_buffer("scales", torch.ones((G, N), dtype=self.dtype))
        self.asym = True

    def init_gba(self) -> None:
        G, N = self._groups(), self.out_channels
        if self.dq_group_size == -1:
            self.dq_group_size = N
        dq_groups = math.ceil(N / self.dq_group_size)
        stat_shape, meta_shape = (G, dq_groups, self.dq_group_size), (G, dq_groups, 1)
        if self.asym:
            self.register_buffer("qzeros", torch.zeros((G, N // 32 * self.w_bit), dtype=torch.int32))
            qs_shape = stat_shape if self.w_bit == 4 else (G, N)
            self.register_buffer("qscales", torch.ones(qs_shape, dtype=torch.uint8))
        else:
            self.register_buffer("qstatistic", torch.ones(stat_shape, dtype=torch.uint8))
            self.register_buffer("qzeros_zeros", torch.zeros(meta_shape, dtype=self.dtype))
            self.register_buffer("qzeros_scales", torch.ones(meta_shape, dtype=self.dtype))
        qs_meta = (1, N, 1) if self.dq_mode == 1 else meta_shape
        self.register_buffer("qscales_zeros", torch.zeros(qs_meta, dtype=self.dtype))
        self.register_buffer("qscales_scales", torch.ones(qs_meta, dtype=self.dtype))
        self.register_buffer("scales", torch.ones((G, N), dtype=self.dtype))
        self.register_buffer("zeros", torch.zeros((G, N), dtype=self.dtype))

    # -- reference API -----------------------------------------------------------------------------------------
    def set_qweight_data(self, data: torch.Tensor) -> None:
        self.qweight.data = data

    def generate_quantized_weight(self, qweight_only: bool = False) -> None:
        raise NotImplementedError("this method has not been implemented.")

    def check_parameters(self) -> None:
        raise NotImplementedError("Subclasses should implement this method.")

    def prepare_params(self) -> None:
        raise NotImplementedError("Subclasses should implement this method.")
