"""Twin of the reference pybind module `q_linear_cuda`
(bitorch_engine/layers/qlinear/nbit/cuda/q_linear_cuda.cpp:357-369)."""
import torch

from .. import _cabi

def _check_cuda(t, name):
    # reference: CHECK_CUDA -> AT_ASSERTM -> RuntimeError (q_linear_cuda.cpp:255-256)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _gidx_is_trivial(g_idx, K, G):
    """True when g_idx == arange(K) // (K // G) (nbit/layer.py:385-386), i.e. groups are contiguous: the fast
    kernels then derive the group from k and never read g_idx.  Checked once per tensor object and version (one
    D2H sync) and remembered ON the tensor, so steady-state calls stay asynchronous / graph-capturable."""
    if g_idx is None:
        return True
    tag = getattr(g_idx, "_b200bit_trivial", None)
    key = (g_idx._version, K, G)
    if tag is not None and tag[0] == key:
        return tag[1]
    if K % G != 0 or g_idx.numel() != K:
        verdict = False
    else:
        ref = torch.arange(K, device=g_idx.device, dtype=g_idx.dtype) // (K // G)
        verdict = bool(torch.equal(g_idx, ref))
    g_idx._b200bit_trivial = (key, verdict)
    return verdict


def mpq_forward(x, qweight, scales, zeros, g_idx, a_bit, w_bit, asym, pdl=False):
    """y[M,N] = x[M,K] @ dequant(qweight)  (q_linear_cuda.cpp:258-270 -> mpq_linear_cuda_kernel.cu:603-626).

    x: [M,K] f16/bf16/f32; qweight int32 [K*w_bit/32, N]; scales [G,N] (x.dtype); zeros [G,N] (sym) or packed int32
    [G, N*w_bit/32] (asym); g_idx int32 [K].  Differences to the reference, all relaxations: any M (the reference's
    caller switches to dequant+matmul above 32 rows), no K%256/N%256 requirement, unsupported configurations raise
    instead of exit()."""
    _check_cuda(x, "x")
    _check_cuda(qweight, "qweight")
    if a_bit != 16:
        raise NotImplementedError(f"a_bit:{a_bit} has not been supported yet!")
    if x.dim() != 2:
        raise ValueError("mpq_forward expects a 2-D input (use flatten_x)")
    M, K = x.shape
    N = qweight.shape[1]
    G = scales.shape[0]
    if qweight.dtype != torch.int32 or qweight.shape[0] * 32 != K * w_bit:
        raise ValueError(f"qweight must be int32 [K*w_bit/32, N]; got {tuple(qweight.shape)} for K={K}, w_bit={w_bit}")
    if scales.dtype != x.dtype:
        raise ValueError(f"scales dtype {scales.dtype} must match x dtype {x.dtype}")
    if M == 0:
        return torch.empty((0, N), dtype=x.dtype, device=x.device)
    x = x.contiguous()
    qweight = qweight.contiguous()
    scales = scales.contiguous()
    zeros = zeros.contiguous()
    lib = _cabi.lib()
    trivial = _gidx_is_trivial(g_idx, K, G)
    y = torch.empty((M, N), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream().cuda_stream
        need = lib.b200bit_mpq_forward_workspace_bytes(M, K, N, w_bit)
        ws = _cabi.workspace(x.device, stream, need)
        rc = lib.b200bit_mpq_forward(
            x.data_ptr(), qweight.data_ptr(), scales.data_ptr(), zeros.data_ptr(),
            None if trivial else g_idx.contiguous().data_ptr(), y.data_ptr(),
            M, K, N, G, w_bit, int(bool(asym)), _cabi.dtype_code(x.dtype),
            ws.data_ptr(), ws.numel(), _cabi.FLAG_PDL if pdl else 0, stream)
    _cabi.check(rc)
    return y
