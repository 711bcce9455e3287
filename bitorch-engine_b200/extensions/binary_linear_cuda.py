"""Twin of the reference pybind module `binary_linear_cuda`
(bitorch_engine/layers/qlinear/binary/cuda/binary_linear_cuda.cpp:92-122): forward / w_pack / mm.

bmm_type: 1 = BSTC32, 2 = BTC32, 3 = ADAPTIVE (binary/cuda/bmm.py).  On sm_100a there is one compute kernel; the
bmm_type only selects (exactly as in the reference) the byte order of PACKED weights:
  packed weight <-> BTC order when bmm_type == 2 or (3 and the shape admits it), else BSTC order."""
import torch

from .. import _cabi

BSTC, BTC, ADAPTIVE = 1, 2, 3


def _check_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _stride_bytes(K):
    return ((K + 31) // 32) * 4


def _pack(t, rows, K, transposed):
    """float/half/bf16/int8 matrix -> canonical bit matrix [rows, stride] uint8."""
    stride = _stride_bytes(K)
    out = torch.empty((rows, stride), dtype=torch.uint8, device=t.device)
    t = t.contiguous()
    with torch.cuda.device(t.device):
        rc = _cabi.lib().b200bit_binary_pack(t.data_ptr(), _cabi.dtype_code(t.dtype), rows, K, int(transposed),
                                             out.data_ptr(), stride, torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc)
    return out


def _weight_layout(bmm_type, m, n, k, for_pack):
    """The reference's rule for which packed byte order a weight has / gets:
    w_pack: BTC iff type 2, or type 3 with k%128==0 and n%8==0 (_get_binary_weight_cuda, kernel.cu:852);
    forward: BTC iff type 2, or type 3 with m%8==0, k%128==0, n%8==0 (binary_linear_forward_combined, :621)."""
    if bmm_type == BTC:
        return BTC
    if bmm_type == ADAPTIVE and k % 128 == 0 and n % 8 == 0 and (for_pack or m % 8 == 0):
        return BTC
    return BSTC


def _relayout(src, n, k, layout, to_reference):
    stride = _stride_bytes(k)
    dst = torch.empty((n * k // 8,) if to_reference else (n, stride), dtype=torch.uint8, device=src.device)
    if not to_reference and stride * 8 != k:
        dst.zero_()
    with torch.cuda.device(src.device):
        rc = _cabi.lib().b200bit_binary_relayout(src.data_ptr(), dst.data_ptr(), n, k, layout, int(to_reference), stride,
                                                 torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc)
    return dst


def _gemm(xb, wb, m, n, k, dtype):
    out = torch.empty((m, n), dtype=dtype, device=xb.device)
    with torch.cuda.device(xb.device):
        rc = _cabi.lib().b200bit_binary_gemm(xb.data_ptr(), wb.data_ptr(), out.data_ptr(), m, n, k, _stride_bytes(k),
                                             _cabi.dtype_code(dtype), torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc)
    return out


def forward(input, weight, bmm_type, transpose):
    """out[m,n] = sum_k sign(input[m,k]) * sign(weight[n,k]) in input.dtype (kernel.cu:629-659, 709-752).
    weight: [n,k] int8/float (unpacked) or the flat uint8 stream produced by w_pack (numel == k*n/8)."""
    _check_cuda(input, "input")
    _check_cuda(weight, "weights")
    if input.dim() != 2:
        raise ValueError("binary_linear_cuda.forward expects a 2-D input")
    m, k = input.shape
    xb = _pack(input, m, k, transposed=False)
    if weight.dtype == torch.uint8:
        n = weight.numel() * 8 // k
        layout = _weight_layout(bmm_type, m, n, k, for_pack=False)
        # the canonical bit matrix of a packed (inference) weight is derived once and kept ON the weight tensor, keyed by
        # torch's version counter: the packed stream is re-laid out again only after an in-place write (the reference
        # re-reads its own byte order directly; here the relayout launch + allocation left the per-call path)
        ver = None if weight.is_inference() else weight._version
        tag = getattr(weight, "_b200bit_canon", None)
        if tag is not None and tag[0] == (ver, layout, weight.data_ptr()) and ver is not None:
            wb = tag[1]
        else:
            wb = _relayout(weight.contiguous().view(-1), n, k, layout, to_reference=False)
            if ver is not None:
                weight._b200bit_canon = ((ver, layout, weight.data_ptr()), wb)
    else:
        n = weight.shape[0]
        # `transpose` only tells the reference to make a [k,n] copy first; packing [n,k] rows directly is the same bits
        wb = _pack(weight, n, k, transposed=False) if transpose else _pack(weight, weight.shape[1], k, transposed=True)
        if not transpose:
            n = weight.shape[1]
    return _gemm(xb, wb, m, n, k, input.dtype)


def w_pack(weight, bmm_type, transpose):
    """[n,k] weight -> flat uint8 [k*n/8] in the reference's byte order (_get_binary_weight_cuda, kernel.cu:830-889)."""
    _check_cuda(weight, "weights")
    n, k = weight.shape
    layout = _weight_layout(bmm_type, 0, n, k, for_pack=True)
    canon = _pack(weight, n, k, transposed=False)
    return _relayout(canon, n, k, layout, to_reference=True)


def mm(x, y, bmm_type):
    """out[m,n] = sum_k sign(x[m,k]) * sign(y[k,n]) (binary_mm_cuda, kernel.cu:754-795)."""
    _check_cuda(x, "x")
    _check_cuda(y, "y")
    if x.dtype != y.dtype:
        raise ValueError(f"The input tensors must have the same dtype. x_dtype: {x.dtype} and y_dtype: {y.dtype}")
    m, k = x.shape
    n = y.shape[1]
    return _gemm(_pack(x, m, k, transposed=False), _pack(y, n, k, transposed=True), m, n, k, x.dtype)
