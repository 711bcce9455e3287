"""Twin of the reference pybind module `functions_cuda`
(bitorch_engine/functions/cuda/functions_cuda.cpp:160-200): same function names, argument order and return types."""
import torch

from .. import _cabi


def _check_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")      # reference: CHECK_CUDA -> AT_ASSERTM


def _dense(t):
    """contiguous and 16-byte aligned (a contiguous view at an odd storage offset is cloned)."""
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t


def _stream():
    return torch.cuda.current_stream().cuda_stream


def fp32toint4(input):
    """Not provided: the reference's kernel pair has no defined result (functions_cuda_kernel.cu:23-69 reduces 1024
    shared-memory slots with 256 threads and quantises 7 of every 4 inputs; include/b200bit.h)."""
    raise NotImplementedError("fp32toint4: the reference implementation reads uninitialised shared memory; "
                              "no defined behaviour to reproduce")


def tensor_pack_to_uint8(data):
    """[m, k] int8 / float32 / bfloat16 / half -> uint8 [m, k/8], bit i of byte j = (data[.., 8j+i] >= 0)
    (functions_cuda_kernel.cu:283-336)."""
    _check_cuda(data, "data")
    if data.dim() != 2 or data.shape[1] % 8 != 0:
        raise ValueError(f"tensor_pack_to_uint8 expects [m, k] with k % 8 == 0; got {tuple(data.shape)}")
    data = _dense(data)
    m, k = data.shape
    out = torch.empty((m, k // 8), dtype=torch.uint8, device=data.device)
    with torch.cuda.device(data.device):
        rc = _cabi.lib().b200bit_sign_pack_u8(data.data_ptr(), _cabi.dtype_code(data.dtype), out.data_ptr(),
                                              out.numel(), _stream())
    _cabi.check(rc)
    return out


def uint8_to_unpacked_tensor(emd, scl):
    """uint8 [bs, seq, pd] and float32 scale [bs, seq, 1] -> float32 [bs, seq, pd*8] of +-scale
    (functions_cuda_kernel.cu:365-402)."""
    _check_cuda(emd, "emd")
    _check_cuda(scl, "scl")
    if emd.dim() != 3 or emd.dtype != torch.uint8:
        raise ValueError("uint8_to_unpacked_tensor expects a uint8 tensor [batch, seq, packed_dim]")
    if scl.dtype != torch.float32 or scl.numel() != emd.shape[0] * emd.shape[1]:
        raise ValueError("scale must be float32 [batch, seq, 1]")
    emd, scl = emd.contiguous(), scl.contiguous()
    bs, seq, pd = emd.shape
    out = torch.empty((bs, seq, pd * 8), dtype=torch.float32, device=scl.device)
    with torch.cuda.device(emd.device):
        rc = _cabi.lib().b200bit_sign_unpack_u8(emd.data_ptr(), scl.data_ptr(), out.data_ptr(), emd.numel(), pd, _stream())
    _cabi.check(rc)
    return out


def q4_pack(data, is_transpose=False):
    """int32 [n, k] or [b, n, k] -> int8 [.., k/2], first code in the high nibble (functions_cuda_kernel.cu:431-466)."""
    _check_cuda(data, "data")
    if data.dtype != torch.int32:
        raise ValueError("q4_pack expects an int32 tensor")
    if data.dim() not in (2, 3):
        raise ValueError(f"tensor sizes not supported: {data.dim()}")     # the reference exit()s here
    k = data.shape[-1]
    if k % 2 != 0:
        raise ValueError("q4_pack: the last dimension must be even (the reference writes out of bounds otherwise)")
    data = _dense(data)
    out = torch.empty(tuple(data.shape[:-1]) + (k // 2,), dtype=torch.int8, device=data.device)
    with torch.cuda.device(data.device):
        rc = _cabi.lib().b200bit_q4_pack(data.data_ptr(), out.data_ptr(), out.numel(), _stream())
    _cabi.check(rc)
    return out.transpose(-1, -2).contiguous() if is_transpose else out


def _q4_unpack(packed_data, scale, is_transpose):
    _check_cuda(packed_data, "packed_data")
    if packed_data.dtype != torch.int8:
        raise ValueError("q4_unpack expects an int8 tensor")
    packed_data = _dense(packed_data)
    shape = tuple(packed_data.shape[:-1]) + (packed_data.shape[-1] * 2,)
    out = torch.empty(shape, dtype=torch.int32 if scale is None else torch.float32, device=packed_data.device)
    with torch.cuda.device(packed_data.device):
        if scale is None:
            rc = _cabi.lib().b200bit_q4_unpack(packed_data.data_ptr(), out.data_ptr(), packed_data.numel(), _stream())
        else:
            rc = _cabi.lib().b200bit_q4_unpack_scale(packed_data.data_ptr(), float(scale), out.data_ptr(),
                                                     packed_data.numel(), _stream())
    _cabi.check(rc)
    return out.transpose(-1, -2).contiguous() if is_transpose else out


def q4_unpack(packed_data, is_transpose=False):
    """int8 [.., k/2] -> int32 codes 0..15 [.., k] (functions_cuda_kernel.cu:477-505)."""
    return _q4_unpack(packed_data, None, is_transpose)


def q4_unpack_and_scaling(packed_data, scale, is_transpose=False):
    """int8 [.., k/2] -> float32 [.., k]: codes read as signed 4-bit times `scale`; 2-, 3- and 4-D (NHWC) inputs
    (functions_cuda_kernel.cu:518-549)."""
    return _q4_unpack(packed_data, scale, is_transpose)
