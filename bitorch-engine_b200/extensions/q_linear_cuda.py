"""Twin of the reference pybind module `q_linear_cuda`
(bitorch_engine/layers/qlinear/nbit/cuda/q_linear_cuda.cpp:357-369)."""
import contextlib

import torch

from .. import _cabi
from .. import decode_chain

_NULL_CTX = contextlib.nullcontext()
_ws_bytes = {}


def _on_device(device):
    """device guard only when the tensor does not live on the current device (the guard costs ~4 us per call, as much as
    a whole decode kernel; the reference pays an OptionalCUDAGuard in C++, mpq_linear_cuda_kernel.cu:612)"""
    return _NULL_CTX if device.index == torch.cuda.current_device() else torch.cuda.device(device)


def _raw_stream(device):
    return torch._C._cuda_getCurrentRawStream(device.index)


def _check_cuda(t, name):
    # reference: CHECK_CUDA -> AT_ASSERTM -> RuntimeError (q_linear_cuda.cpp:255-256)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _version(t):
    """torch's in-place version counter, or None for inference tensors (torch.inference_mode(): they do not track one;
    reading ._version raises there -- the reference works in that mode, e.g. under vLLM-style serving)."""
    if t.is_inference():
        return None
    return t._version


def _gidx_is_trivial(g_idx, K, G):
    """True when g_idx == arange(K) // (K // G) (nbit/layer.py:385-386), i.e. groups are contiguous: the fast
    kernels then derive the group from k and never read g_idx.  Checked once per tensor object and version (one
    D2H sync) and remembered ON the tensor, so steady-state calls stay asynchronous / graph-capturable."""
    if g_idx is None:
        return True
    tag = getattr(g_idx, "_b200bit_trivial", None)
    key = (_version(g_idx), g_idx.data_ptr(), K, G)
    if tag is not None and tag[0] == key:
        return tag[1]
    if K % G != 0 or g_idx.numel() != K:
        verdict = False
    else:
        ref = torch.arange(K, device=g_idx.device, dtype=g_idx.dtype) // (K // G)
        verdict = bool(torch.equal(g_idx, ref))
    g_idx._b200bit_trivial = (key, verdict)
    return verdict


# (device index, stream) -> (x tensor of the previous decode call, its _version at that call).  The strong reference
# keeps the storage alive, so an equal data_ptr() really is the same buffer and not a re-allocation.
_prev_x = {}


def _input_ready(x, stream):
    """B200BIT_FLAG_INPUT_READY (include/b200bit.h): True when this call's activation is the very buffer the previous
    mpq_forward call on this stream read (k_proj / v_proj after q_proj, up_proj after gate_proj) and torch has recorded
    no in-place write to it since -- it was complete before that call, so this kernel may read it without waiting for
    the kernels in front.  Writes torch does not version (a foreign kernel writing through data_ptr()) are the usual
    caveat; B200BIT_EARLY=0 switches the overlap off."""
    key = (x.device.index, stream)
    ver = _version(x)
    if ver is None:                  # inference tensor: no version counter, so no proof that x is unchanged
        _prev_x.pop(key, None)
        return False
    prev = _prev_x.get(key)
    _prev_x[key] = (x, ver)
    if prev is None:
        return False
    px, pv = prev
    return (px.data_ptr() == x.data_ptr() and px.shape == x.shape and px.dtype == x.dtype and _version(px) == pv
            and ver == pv)


MPQ_FUSED_MAX_ROWS = 32
# 2-bit: the small-batch kernel needs 85 - 205 us at 32 rows on the Llama-7B shapes, dequantise + dense GEMM 30 - 46 us
# (profiles/configs_r02.json)
MPQ_FUSED_MAX_ROWS_2BIT = 8
TC_MIN_ROWS = 16           # from 17 rows on the tcgen05 kernel takes every shape it covers
GRAD_INPUT_FUSED_MAX_ROWS = 4


_tc_verdicts = {}
_sm_counts = {}
TC_SPLIT_WS_CAP = 64 << 20


def _tc_workspace_bytes(M, N, device):
    """Split-K scratch for the batched kernel: fp32 partial tiles for up to 8 k-slices.  Always for small batches; for
    larger ones only when the tiles alone would leave more than half of the SMs idle (narrow layers: the 1024-column k / v
    projections of a GQA model at a 512-token prefill are 32 tiles) and the scratch stays under 64 MB."""
    if M <= 256:
        return _cabi.WS_TICKET_BYTES + 8 * M * N * 4
    idx = device.index if device.index is not None else torch.cuda.current_device()
    sms = _sm_counts.get(idx)
    if sms is None:
        sms = _sm_counts[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    tiles128 = -(-N // 128) * -(-M // 128)
    need = _cabi.WS_TICKET_BYTES + 8 * M * N * 4
    return need if (tiles128 * 2 <= sms and need <= TC_SPLIT_WS_CAP) else 0


def _tc_supported(M, K, N, G, w_bit, asym, ws_bytes):
    """b200bit_mpq_forward_tc_supported, remembered per problem (shape rules + does the group table of a k-slice fit)."""
    key = (M, K, N, G, w_bit, asym, ws_bytes)
    v = _tc_verdicts.get(key)
    if v is None:
        v = _tc_verdicts[key] = bool(_cabi.lib().b200bit_mpq_forward_tc_supported(M, K, N, G, w_bit, int(asym), _cabi.F16, ws_bytes))
    return v


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
    rec = decode_chain.recording()
    if rec is not None:
        # DecodeChain.capture: the call becomes a node of the chain (nothing is launched here)
        if not _gidx_is_trivial(g_idx, K, G):
            raise NotImplementedError("DecodeChain: act-order g_idx is not supported inside a chain")
        return rec.add(x, qweight, scales, zeros, w_bit, asym)
    if M > (TC_MIN_ROWS if w_bit != 2 else MPQ_FUSED_MAX_ROWS_2BIT) and x.dtype != torch.float32:
        # batches (bs = 32 serving, prefill, training).  The tcgen05 kernel (csrc/mpq_tc.cu) -- weights dequantised straight
        # into tensor memory, one pass over the packed matrix for any M, split-K for small M, no fp16 copy of W in HBM --
        # takes every shape it covers from 17 rows on (measured against the mma.sync small-batch kernel at 32 rows:
        # 26.7 vs 39.1 us on 4096x11008, 26.2 vs 35.4 us on 11008x4096, 20.9 vs 18.8 us on 4096x4096; against dequantise +
        # cuBLAS it is ahead at every M, profiles/r2_25_tc_kernel_split_k.jsonl).  Shapes it does not cover: the
        # small-batch kernels up to 32 rows (2-bit: 8), then dequantise ONCE (one kernel, bit-identical to unpack_qweight) +
        # dense GEMM -- the switch the reference makes at 32 rows (mpq_layer.py:59-63).
        ws_bytes = _tc_workspace_bytes(M, N, x.device)
        tc_ok = (w_bit in (2, 4) and x.dtype == torch.float16 and _gidx_is_trivial(g_idx, K, G) and
                 _tc_supported(M, K, N, G, w_bit, bool(asym), ws_bytes))
        if tc_ok:
            x = x.contiguous()
            if x.data_ptr() % 16 == 0:
                y = torch.empty((M, N), dtype=x.dtype, device=x.device)
                with _on_device(x.device):
                    stream = _raw_stream(x.device)
                    ws = _cabi.workspace(x.device, stream, ws_bytes)
                    rc = _cabi.lib().b200bit_mpq_forward_tc(x.data_ptr(), qweight.contiguous().data_ptr(),
                                                            scales.contiguous().data_ptr(), zeros.contiguous().data_ptr(),
                                                            y.data_ptr(), M, K, N, G, w_bit, int(bool(asym)), _cabi.F16,
                                                            ws.data_ptr(), ws.numel(), stream)
                if rc:
                    _cabi.check(rc)
                return y
        if M > (MPQ_FUSED_MAX_ROWS if w_bit != 2 else MPQ_FUSED_MAX_ROWS_2BIT):
            return torch.matmul(x, mpq_dequant(qweight, scales, zeros, g_idx, w_bit, asym))
    x = x.contiguous()
    qweight = qweight.contiguous()
    scales = scales.contiguous()
    zeros = zeros.contiguous()
    lib = _cabi.lib()
    trivial = _gidx_is_trivial(g_idx, K, G)
    y = torch.empty((M, N), dtype=x.dtype, device=x.device)
    device = x.device
    with _on_device(device):
        stream = _raw_stream(device)
        key = (M, K, N, w_bit)
        need = _ws_bytes.get(key)
        if need is None:
            need = _ws_bytes[key] = lib.b200bit_mpq_forward_workspace_bytes(M, K, N, w_bit)
        ws = _cabi.workspace(device, stream, need)
        flags = 0
        if pdl:
            flags = _cabi.FLAG_PDL
            if M == 1 and _input_ready(x, stream):
                flags |= _cabi.FLAG_INPUT_READY
        rc = lib.b200bit_mpq_forward(
            x.data_ptr(), qweight.data_ptr(), scales.data_ptr(), zeros.data_ptr(),
            None if trivial else g_idx.contiguous().data_ptr(), y.data_ptr(),
            M, K, N, G, w_bit, int(bool(asym)), _cabi.dtype_code(x.dtype),
            ws.data_ptr(), ws.numel(), flags, stream)
    if rc:
        _cabi.check(rc)
    return y


def _ptr(t):
    return None if t is None else t.data_ptr()


def mpq_grad_input(qweight, scales, zeros, g_idx, output_gradient, a_bit, w_bit, asym):
    """dx[M,K] = dy[M,N] @ dequant(qweight)^T  (q_linear_cuda.cpp:272-284 -> mpq_linear_cuda_kernel.cu:1198-1223)."""
    _check_cuda(output_gradient, "output_gradient")
    _check_cuda(qweight, "qweight")
    if a_bit != 16:
        raise NotImplementedError(f"a_bit:{a_bit} has not been supported yet!")
    dy = output_gradient.contiguous()
    M, N = dy.shape
    K = qweight.shape[0] * 32 // w_bit
    G = scales.shape[0]
    if scales.dtype != dy.dtype:
        raise ValueError(f"scales dtype {scales.dtype} must match output_gradient dtype {dy.dtype}")
    if M > GRAD_INPUT_FUSED_MAX_ROWS and dy.dtype != torch.float32:
        # measured on B200 (profiles/configs_r02.json): the warp-per-packed-row kernel re-reads W once per 4 rows of dy and
        # needs 190 - 530 us at 32 rows; dequantise once (one kernel) + the dense GEMM needs ~55 us there and runs at
        # tensor-core speed for the training shapes (M = 2048).  The reference's own kernel (back_quant_mm_kernel,
        # mpq_linear_cuda_kernel.cu:920-983) is slower than either.
        return torch.matmul(dy, mpq_dequant(qweight, scales, zeros, g_idx, w_bit, asym).t())
    trivial = _gidx_is_trivial(g_idx, K, G)
    dx = torch.empty((M, K), dtype=dy.dtype, device=dy.device)
    if M == 0:
        return dx
    with torch.cuda.device(dy.device):
        rc = _cabi.lib().b200bit_mpq_grad_input(
            dy.data_ptr(), qweight.contiguous().data_ptr(), scales.contiguous().data_ptr(), zeros.contiguous().data_ptr(),
            None if trivial else g_idx.contiguous().data_ptr(), dx.data_ptr(), M, K, N, G, w_bit, int(bool(asym)),
            _cabi.dtype_code(dy.dtype), torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc)
    return dx


def mpq_dequant(qweight, scales, zeros, g_idx, w_bit, asym, fused=False, perm=None):
    """fp weight [K,N] in scales.dtype, bit-identical to the reference's Python unpack_qweight (layer_type 1).
    Not part of the reference extension (it does this with ~6 torch kernels, utils.py:31-51); exposed here because the
    host-side unpack_qweight / M>32 forward / MBWQ backward call it."""
    _check_cuda(qweight, "qweight")
    K = qweight.shape[0] * 32 // w_bit
    N = qweight.shape[1]
    G = scales.shape[0]
    trivial = _gidx_is_trivial(g_idx, K, G)
    out = torch.empty((K, N), dtype=scales.dtype, device=qweight.device)
    with torch.cuda.device(qweight.device):
        rc = _cabi.lib().b200bit_mpq_dequant(
            qweight.contiguous().data_ptr(), scales.contiguous().data_ptr(), zeros.contiguous().data_ptr(),
            None if trivial else g_idx.contiguous().data_ptr(), out.data_ptr(), K, N, G, w_bit, int(bool(asym)),
            _cabi.dtype_code(scales.dtype), int(bool(fused)), _ptr(None if perm is None else perm.contiguous()),
            torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc)
    return out


def mpq_pack_weight(weight, scales, zeros, g_idx, w_bit, asym, zeros_unpacked=False, perm=None):
    """int32 [K*w_bit/32, N] from fp weight [K,N], bit-identical to the reference's pack_fp_weight (utils.py:72-147)."""
    _check_cuda(weight, "weight")
    K, N = weight.shape
    G = scales.shape[0]
    # an fp32 weight against half parameters keeps its precision: torch promotes the expression to fp32 (utils.py:118-131)
    if weight.dtype != torch.float32:
        weight = weight.to(scales.dtype)
    weight = weight.contiguous()
    zeros = zeros.contiguous()
    if asym and zeros_unpacked:
        zeros = zeros.to(scales.dtype).contiguous()
    trivial = _gidx_is_trivial(g_idx, K, G)
    out = torch.empty((K * w_bit // 32, N), dtype=torch.int32, device=weight.device)
    with torch.cuda.device(weight.device):
        rc = _cabi.lib().b200bit_mpq_pack_weight(
            weight.data_ptr(), scales.contiguous().data_ptr(), zeros.data_ptr(),
            None if trivial else g_idx.contiguous().data_ptr(), _ptr(None if perm is None else perm.contiguous()),
            out.data_ptr(), K, N, G, w_bit, int(bool(asym)), int(bool(zeros_unpacked)), _cabi.dtype_code(scales.dtype),
            _cabi.dtype_code(weight.dtype), torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc)
    return out


def unpack_zeros(qzeros, w_bit):
    """packed int32 [G, N*w_bit/32] -> integer zero points [G, N] including the +1 (utils.py:36-41).  [G,N] is
    K/group times smaller than the weight: plain torch ops."""
    nb = 32 // w_bit
    shifts = torch.arange(0, 32, w_bit, dtype=torch.int32, device=qzeros.device).view(1, 1, nb)
    z = (qzeros.unsqueeze(2) >> shifts) & ((1 << w_bit) - 1)
    return (z + 1).reshape(qzeros.shape[0], -1)


def pack_zeros(zeros, w_bit):
    """integer(-valued) zero points [G, N] -> packed int32 [G, N*w_bit/32]: trunc, (z-1) & mask, LSB first along N
    (gptq_style_zeros_packing, quant_operators.py:348-368)."""
    nb = 32 // w_bit
    G, N = zeros.shape
    z = (zeros.reshape(G, N // nb, nb).to(torch.int32) - 1) & ((1 << w_bit) - 1)
    shifts = torch.arange(0, 32, w_bit, dtype=torch.int32, device=zeros.device).view(1, 1, nb)
    return (z << shifts).sum(dim=-1).to(torch.int32)


# ---------------------------------------------------------------------------------------------------------------
# MBWQ ("Q4" GPTQ-style and exl2 mixed-bit) entry points (q_linear_cuda.cpp:286-354)
# ---------------------------------------------------------------------------------------------------------------
def _perm_is_identity(q_perm, K):
    """q_perm == 0 everywhere is the reference's "no permutation" marker (mbwq_linear_cuda_kernel.cu:777 evaluates
    torch::all(q_perm == 0).item() on EVERY call -- a device sync); arange(K) is the same map.  Verdict cached on the
    tensor object / version."""
    if q_perm is None:
        return True
    tag = getattr(q_perm, "_b200bit_identity", None)
    key = (_version(q_perm), q_perm.data_ptr())
    if tag is not None and tag[0] == key:
        return tag[1]
    ident = bool(torch.all(q_perm == 0).item()) or bool(
        torch.equal(q_perm.to(torch.int64) & 0xFFFF, torch.arange(K, device=q_perm.device)))
    q_perm._b200bit_identity = (key, ident)
    return ident


def mbwq_trans_qweight(qweight, q_groups, use_mbw, height, groups, bits):
    """(qweight, rows) -- q_linear_cuda.cpp:286-296 -> mbwq_linear_trans_qweight_cuda (:536-626).  The in-place
    "shuffle" of the reference is compiled to a no-op (exl2/config.h:16-21, all QMODE_* = 0), so qweight is returned
    untouched; rows = cumulative weight-row counts per bit-width (8,6,5,4,3,2) + the bit-width mask, [] for use_mbw=False."""
    _check_cuda(qweight, "qweight")
    if not use_mbw:
        if bits not in (2, 4):
            raise NotImplementedError(f"Error: weight bit width:{bits} has not been supported yet!")
        return qweight, []
    qg = q_groups.detach().to("cpu").to(torch.int64).view(-1)[: 2 * groups] & 0xFFFF     # one D2H copy, as :562
    counts = {8: 0, 6: 0, 5: 0, 4: 0, 3: 0, 2: 0}
    mask, row = 0, 0
    for i in range(groups):
        b = int(qg[2 * i])
        if b not in counts:
            raise NotImplementedError(f"Error: weight bit width:{b} has not been supported yet!")
        mask |= 1 << (b - 1)
        rows = (int(qg[2 * i + 3]) - int(qg[2 * i + 1])) * 32 // b if i < groups - 1 else height - row
        counts[b] += rows
        row += rows
    out, acc = [], 0
    for b in (8, 6, 5, 4, 3, 2):
        acc += counts[b]
        out.append(acc)
    return qweight, out + [mask]


def mbwq_q42fp_weight(qweight, scales, zeros, group_size, bits, q_perm):
    """fp16 weight [K,N] of a GPTQ-style 4/2-bit matrix, rows scattered through q_perm (q_linear_cuda.cpp:298-308)."""
    _check_cuda(qweight, "qweight")
    if scales.dtype != torch.float16:
        raise TypeError("mbwq_q42fp_weight: scales must be torch.half")
    K = qweight.shape[0] * 32 // bits
    perm = None if _perm_is_identity(q_perm, K) else q_perm
    return mpq_dequant(qweight, scales, zeros, None, bits, False, fused=True, perm=perm)


def mbwq_q4_forward(x, qweight, scales, zeros, group_size, q_perm, bits):
    """y = x[:, q_perm] @ W  (q_linear_cuda.cpp:310-321 -> mbwq_linear_q4_forward_cuda :742-825).  Same packed layout as
    MPQ-sym with contiguous groups, so it runs on the same kernels; the activation gather is one index_select."""
    _check_cuda(x, "x")
    if x.dtype != torch.float16:
        raise TypeError("mbwq_q4_forward: x must be torch.half")          # TORCH_CHECK at :753-754
    K = x.shape[1]
    if not _perm_is_identity(q_perm, K):
        x = x.index_select(1, q_perm.to(torch.int64) & 0xFFFF)
    return mpq_forward(x, qweight, scales, zeros, None, 16, bits, False)


def mbwq_exl2fp_weight(qweight, scales, zeros, q_perm, q_group_map, rows):
    """fp16 weight [K,N] of an exl2 mixed-bit matrix (q_linear_cuda.cpp:323-336 -> :849-897)."""
    _check_cuda(qweight, "qweight")
    import ctypes
    K, N = q_group_map.numel() // 2, qweight.shape[1]
    out = torch.empty((K, N), dtype=torch.float16, device=qweight.device)
    rows6 = (ctypes.c_int * 6)(*[int(r) for r in rows[:6]])
    perm = None if (q_perm is None or _perm_is_identity(q_perm, K)) else q_perm.contiguous()
    with torch.cuda.device(qweight.device):
        rc = _cabi.lib().b200bit_exl2_dequant(qweight.contiguous().data_ptr(), scales.contiguous().data_ptr(),
                                              zeros.contiguous().data_ptr(), _ptr(perm),
                                              q_group_map.contiguous().data_ptr(), out.data_ptr(), K, N, rows6,
                                              torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc)
    return out


# above this many rows the dequantise-once + dense GEMM pair wins over re-streaming the packed matrix per 8-row group
# (the reference switches to its reconstruct + cuBLAS path at 50 rows, mbwq_linear_cuda_kernel.cu:947-957)
EXL2_FUSED_MAX_ROWS = 32


def mbwq_exl2_forward(x, qweight, scales, zeros, q_perm, q_group_map, rows, use_cublas=False):
    """y = x[:, q_perm] @ W_exl2 (q_linear_cuda.cpp:338-354 -> :926-1007).  Up to 32 rows: ONE fused mixed-bit kernel
    (csrc/exl2_gemv.cu: the packed matrix is read once, no fp16 copy of W, no cuBLAS); more rows, or use_cublas=True:
    dequantise (one kernel) + dense matmul, the path the reference itself takes for large batches."""
    _check_cuda(x, "x")
    _check_cuda(qweight, "qweight")
    if x.dtype != torch.float16:
        raise TypeError("mbwq_exl2_forward: x must be torch.half")
    if x.dim() != 2:
        raise ValueError("mbwq_exl2_forward expects a 2-D input (use flatten_x)")
    M, K = x.shape
    N = qweight.shape[1]
    if q_group_map.numel() != 2 * K:
        raise ValueError(f"q_group_map has {q_group_map.numel()} entries, expected 2*K = {2 * K}")
    if use_cublas or M > EXL2_FUSED_MAX_ROWS:
        W = mbwq_exl2fp_weight(qweight, scales, zeros, q_perm, q_group_map, rows)
        return torch.matmul(x, W)
    import ctypes
    y = torch.empty((M, N), dtype=torch.float16, device=x.device)
    if M == 0:
        return y
    rows6 = (ctypes.c_int * 6)(*[int(r) for r in rows[:6]])
    perm = None if (q_perm is None or _perm_is_identity(q_perm, K)) else q_perm.contiguous()
    with _on_device(x.device):
        rc = _cabi.lib().b200bit_exl2_forward(x.contiguous().data_ptr(), qweight.contiguous().data_ptr(),
                                              scales.contiguous().data_ptr(), zeros.contiguous().data_ptr(), _ptr(perm),
                                              q_group_map.contiguous().data_ptr(), y.data_ptr(), M, K, N, int(scales.shape[0]), rows6,
                                              _raw_stream(x.device))
    if rc:
        _cabi.check(rc)
    return y
