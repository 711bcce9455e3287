"""qweight_update_fn: the per-parameter-type weight update DiodeMix dispatches to
(twin of bitorch_engine/utils/model_helper.py:363-530), with the MPQ and binary branches running as single fused
kernels (csrc/optim.cu)."""
import math

import torch

from .. import _cabi


def _bias_corrected_step(lr, beta1, beta2, step, correct_bias):
    """model_helper.py:501-506."""
    if not correct_bias:
        return lr
    n = step.item()
    return lr * math.sqrt(1.0 - beta2 ** n) / (1.0 - beta1 ** n)


def _mpq_step_torch(qweight, exp_avg_s, exp_avg_l, step, step_size, beta1, beta2, eps, dtype, projector, grad):
    """Unfused path for what the kernel does not cover (act-order g_idx, GaLore projector, MBWQ q_perm): the reference's
    own sequence (model_helper.py:485-523) on top of the dequant / pack kernels -- gptq_style_unpacking, Adam moments,
    projection back, update_zeros every 5th step, pack_fp_weight."""
    from ..layers.qlinear.nbit.cuda.utils import pack_fp_weight
    from ..utils.quant_operators import gptq_style_unpacking
    from ..utils.model_helper import update_zeros
    w, z_unpacked = gptq_style_unpacking(qweight)
    w = w.to(dtype)
    if z_unpacked is not None:
        z_unpacked = z_unpacked.to(dtype)
    exp_avg_l.mul_(beta1).add_(grad, alpha=(1.0 - beta1))
    exp_avg_s.mul_(beta2).addcmul_(grad, grad, value=1.0 - beta2)
    denom = exp_avg_s.sqrt().add_(eps)
    norm_grad = exp_avg_l / denom
    if projector is not None:
        norm_grad = projector.project_back(norm_grad.to(dtype))
    w.add_(norm_grad, alpha=-step_size)
    if int(step.item()) % 5 == 0:
        if qweight.layer_type == 1 and not qweight.asym:
            # symmetric MPQ: the reference cannot reach this point (its unpack raises); the MBWQ rule without the gather
            order = torch.argsort(qweight.g_idx.long(), dim=0)
            G = qweight.scales.shape[0]
            zg = norm_grad[order].view(G, w.shape[0] // G, -1).mean(1)
            qweight.zeros.add_((step_size * zg).to(qweight.zeros.dtype))
        else:
            update_zeros(qweight, w, norm_grad, step_size, z_unpacked)
    qweight.data = pack_fp_weight(w, qweight, z_unpacked if qweight.asym else None)


def qweight_update_fn(qweight, exp_avg_s=None, exp_avg_l=None, step=None, lr=1e-4, weight_decay=0.0, beta1=0.99,
                      beta2=0.9999, eps=1e-6, dtype=torch.half, correct_bias=None, projector=None, grad=None) -> None:
    """Same contract as the reference: updates `qweight` (and the optimizer state tensors) in place."""
    from ..layers.qlinear.nbit import MPQWeightParameter
    from ..layers.qlinear.binary import BinaryLinearParameter
    from ..extensions.q_linear_cuda import _gidx_is_trivial

    step.add_(1)
    lib = _cabi.lib()

    if isinstance(qweight, BinaryLinearParameter):
        # model_helper.py:437-445
        g = qweight.grad if grad is None else grad
        if not qweight.is_cuda:
            raise RuntimeError("b200bit: the fused binary update needs CUDA tensors")
        g_i8 = g.contiguous() if g.dtype == torch.int8 else None
        g_c = None if g_i8 is not None else g.to(dtype).contiguous()
        with torch.cuda.device(qweight.device):
            rc = lib.b200bit_diodemix_binary_step(
                qweight.data.data_ptr(), None if g_i8 is None else g_i8.data_ptr(),
                None if g_c is None else g_c.data_ptr(), exp_avg_l.data_ptr(), exp_avg_s.data_ptr(),
                qweight.numel(), _cabi.dtype_code(dtype), beta1, beta2, lr, torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc)
        return

    if isinstance(qweight, MPQWeightParameter):
        if grad is None:
            grad = qweight.privileged_grad
        step_size = _bias_corrected_step(lr, beta1, beta2, step, correct_bias)
        K = qweight.shape[0] * 32 // qweight.w_bit if qweight.layer_type == 1 else None
        sdt = qweight.scales.dtype
        fused = (qweight.layer_type == 1 and projector is None and qweight.is_cuda
                 and (dtype == torch.float32 or dtype == sdt) and grad.dtype in (sdt, dtype)
                 and _gidx_is_trivial(qweight.g_idx, K, qweight.scales.shape[0])
                 and exp_avg_l.dtype == dtype and exp_avg_s.dtype == dtype
                 and exp_avg_l.is_contiguous() and exp_avg_s.is_contiguous())
        if not fused:
            _mpq_step_torch(qweight, exp_avg_s, exp_avg_l, step, step_size, beta1, beta2, eps, dtype, projector, grad)
            return
        N = qweight.shape[1]
        G = qweight.scales.shape[0]
        zeros = qweight.zeros
        upd = int(step.item()) % 5 == 0
        if upd and qweight.asym:
            zeros = zeros.clone()            # the layer's buffer object is replaced, as in the reference (:357)
        with torch.cuda.device(qweight.device):
            rc = lib.b200bit_diodemix_mpq_step(
                qweight.data.data_ptr(), qweight.scales.contiguous().data_ptr(), zeros.data_ptr(),
                grad.contiguous().data_ptr(), exp_avg_l.data_ptr(), exp_avg_s.data_ptr(), K, N, G, qweight.w_bit,
                int(bool(qweight.asym)), _cabi.dtype_code(sdt), _cabi.dtype_code(dtype), _cabi.dtype_code(grad.dtype),
                beta1, beta2, eps, step_size, int(upd), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc)
        if upd and qweight.asym:
            qweight.zeros = zeros
        return

    raise NotImplementedError("qweight.dtype '{}' has not been supported yet.".format(str(qweight.data.dtype)))
