"""ctypes binding of libb200bit.so (C ABI declared in include/b200bit.h).

The library is the product: if it is missing or a symbol is absent this module raises -- it never falls back to
PyTorch or to the oracle.  ctypes releases the GIL around every call (SURVEY.md section 8b "Threading")."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200bit.so")

F32, F16, BF16, I8, I32 = 0, 1, 2, 3, 4
FLAG_PDL = 1
FLAG_INPUT_READY = 2
WS_TICKET_BYTES = 16384

_c_int, _c_size_t, _c_void_p, _c_uint = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_uint

# name -> (restype, argtypes); mirrors include/b200bit.h one to one (tests/test_cabi_symbols.py checks both ways)
PROTOTYPES = {
    "b200bit_version": (_c_int, []),
    "b200bit_last_error": (ctypes.c_char_p, []),
    "b200bit_device_info": (_c_int, [ctypes.POINTER(_c_int)] * 3),
    "b200bit_set_gemv_tuning": (_c_int, [_c_int, _c_int, _c_int]),
    "b200bit_set_path": (_c_int, [_c_int, _c_int]),
    "b200bit_set_trace_buffer": (_c_int, [_c_void_p]),
    "b200bit_mpq_decode_plan": (_c_int, [_c_int] * 6 + [ctypes.POINTER(_c_int)]),
    "b200bit_mpq_forward_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "b200bit_mpq_forward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                     _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                     _c_void_p, _c_size_t, _c_uint, _c_void_p]),
    "b200bit_mpq_chain_plan_bytes": (_c_size_t, [_c_void_p, _c_int]),
    "b200bit_mpq_chain_build": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_size_t,
                                         ctypes.POINTER(_c_int)]),
    "b200bit_mpq_chain_launch": (_c_int, [_c_void_p, ctypes.POINTER(_c_int), _c_uint, _c_void_p]),
    "b200bit_mpq_chain_plan_host": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, ctypes.POINTER(_c_int)]),
    "b200bit_mpq_chain_status": (_c_int, [_c_void_p, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int), _c_void_p]),
    "b200bit_mpq_forward_tc_supported": (_c_int, [_c_int] * 7 + [_c_size_t]),
    "b200bit_mpq_forward_tc": (_c_int, [_c_void_p] * 5 + [_c_int] * 7 + [_c_void_p, _c_size_t, _c_void_p]),
    "b200bit_mpq_grad_input": (_c_int, [_c_void_p] * 6 + [_c_int] * 7 + [_c_void_p]),
    "b200bit_mpq_dequant": (_c_int, [_c_void_p] * 5 + [_c_int] * 7 + [_c_void_p, _c_void_p]),
    "b200bit_exl2_dequant": (_c_int, [_c_void_p] * 6 + [_c_int, _c_int, ctypes.POINTER(_c_int), _c_void_p]),
    "b200bit_exl2_forward": (_c_int, [_c_void_p] * 7 + [_c_int, _c_int, _c_int, _c_int, ctypes.POINTER(_c_int), _c_void_p]),
    "b200bit_mpq_pack_weight": (_c_int, [_c_void_p] * 6 + [_c_int] * 8 + [_c_void_p]),
    "b200bit_diodemix_mpq_step": (_c_int, [_c_void_p] * 6 + [_c_int] * 8 + [ctypes.c_double] * 4 + [_c_int, _c_void_p]),
    "b200bit_diodemix_binary_step": (_c_int, [_c_void_p] * 5 + [_c_size_t, _c_int] + [ctypes.c_double] * 3 + [_c_void_p]),
    "b200bit_binary_pack": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_void_p]),
    "b200bit_binary_relayout": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "b200bit_binary_gemm": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "b200bit_cpu_binary_pack": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int]),
    "b200bit_cpu_binary_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int]),
    "b200bit_q4_pack": (_c_int, [_c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "b200bit_q4_unpack": (_c_int, [_c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "b200bit_q4_unpack_scale": (_c_int, [_c_void_p, ctypes.c_float, _c_void_p, _c_size_t, _c_void_p]),
    "b200bit_sign_pack_u8": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "b200bit_sign_unpack_u8": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_size_t, _c_void_p]),
}



class ChainNode(ctypes.Structure):
    """b200bit_chain_node (include/b200bit.h)."""
    _fields_ = [("x", _c_void_p), ("y", _c_void_p), ("qweight", _c_void_p), ("scales", _c_void_p), ("zeros", _c_void_p),
                ("K", _c_int), ("N", _c_int), ("G", _c_int), ("reserved", _c_int)]


class ChainPlanNode(ctypes.Structure):
    """One 64-byte record of the plan's node table (csrc/mpq_chain.cuh ChainNode), as b200bit_mpq_chain_plan_host writes it."""
    _fields_ = [("x", _c_void_p), ("y", _c_void_p), ("xll", _c_void_p), ("yll", _c_void_p), ("R", _c_int), ("N", _c_int),
                ("strips", _c_int), ("n28", _c_int), ("tiles", _c_int), ("off_sig", _c_int), ("wx_node", _c_int),
                ("wy_node", _c_int)]


_lib = None
_lock = threading.Lock()


def lib():
    """Load (once) and return the ctypes handle; raises RuntimeError if the CUDA library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"b200bit: {LIB_PATH} is missing -- build it with `python bitorch-engine_b200/build.py` "
                    "(there is no CPU / PyTorch fallback for this path)")
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(handle, name)      # AttributeError if the symbol is not exported
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


class B200BitError(RuntimeError):
    pass


def check(rc):
    """Translate a negative return code into a Python exception (the reference raised RuntimeError through
    AT_ASSERTM / TORCH_CHECK, or killed the process with exit(); SURVEY.md section 8b "Error convention")."""
    if rc == 0:
        return
    msg = lib().b200bit_last_error().decode("utf-8", "replace")
    if rc in (-1, -2):
        raise ValueError(f"b200bit: {msg}")
    if rc == -3:
        raise NotImplementedError(f"b200bit: {msg}")
    raise B200BitError(f"b200bit (code {rc}): {msg}")


def dtype_code(dtype):
    import torch
    if dtype == torch.float16:
        return F16
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float32:
        return F32
    if dtype == torch.int8:
        return I8
    if dtype == torch.int32:
        return I32
    raise NotImplementedError(f"b200bit: tensor type not supported: {dtype}")


def default_pdl():
    """Inference-time default of the `pdl` argument of the forward shims (B200BIT_PDL=0 switches it off): launch with the
    programmatic-dependent-launch attribute, so a layer's weight prefetch overlaps the kernel in front of it.  Training
    keeps full stream serialisation (the optimizer rewrites the packed weights right in front of the next forward)."""
    return os.environ.get("B200BIT_PDL", "1") != "0"


_workspaces = {}


def workspace(device, stream_ptr, nbytes):
    """Per (device, stream) scratch: split-K partials + self-resetting tickets.  Zero-initialised once; grown on
    demand (never inside a CUDA-graph capture: warm up first)."""
    import torch
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws
