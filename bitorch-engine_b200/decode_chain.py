"""Decode chain: consecutive batch-1 MPQ (4-bit) Linear layers executed by ONE persistent launch
(include/b200bit.h `b200bit_mpq_chain_*`, csrc/mpq_chain.cuh).

The reference runs every layer as its own kernel (`MPQLinearCudaFunction.forward` -> `q_linear_cuda.mpq_forward`,
bitorch_engine/layers/qlinear/nbit/cuda/mpq_layer.py:28-75); at batch 1 a layer is 1 - 4 us of HBM time and the
boundaries between kernels cost as much.  A chain keeps the reference's per-layer API for BUILDING (the same modules /
the same `mpq_forward` arguments) and replaces only the launching:

    chain = DecodeChain.capture(lambda: block(hidden))   # runs the nn.Modules once in recording mode: nothing is
    out = chain.outputs                                  # launched, every MPQ layer call becomes a node
    chain.launch()                                       # one kernel for all recorded layers; graph-capturable

Recording is only valid for code whose CUDA work consists of M == 1 MPQ 4-bit layer calls (views / reshapes between them
are fine): fused q/k/v and gate/up segments of a decoder block, or the bare chain of linear layers that the
"linear-layer tokens/s" metric measures.  Anything else must stay outside the recorded function.
"""
import ctypes

import torch

from . import _cabi

_recorder = None     # the DecodeChain that q_linear_cuda.mpq_forward feeds while a capture is active


def recording():
    return _recorder


class DecodeChain:
    def __init__(self):
        self.nodes = []          # (x, y, qweight, scales, zeros, K, N, G)
        self.w_bit = None
        self.asym = None
        self.dtype = None
        self.device = None
        self.plan = None
        self.info = (ctypes.c_int * 16)()
        self.outputs = None

    # ---- building -------------------------------------------------------------------------------------------
    def add(self, x, qweight, scales, zeros, w_bit, asym, out=None):
        """Append y = x @ dequant(qweight) (M == 1); returns the output tensor the launch will fill (`out`, a contiguous
        [1, N] tensor of x's dtype, or a fresh one)."""
        if self.plan is not None:
            raise RuntimeError("DecodeChain: already built")
        if x.dim() != 2 or x.shape[0] != 1:
            raise ValueError("DecodeChain: nodes are batch-1 layers (x must be [1, K])")
        if not x.is_contiguous():
            raise ValueError("DecodeChain: x must be contiguous (a chain cannot launch a copy kernel)")
        K = x.shape[1]
        N = qweight.shape[1]
        G = scales.shape[0]
        if self.w_bit is None:
            self.w_bit, self.asym, self.dtype, self.device = int(w_bit), bool(asym), x.dtype, x.device
        if (int(w_bit), bool(asym), x.dtype, x.device) != (self.w_bit, self.asym, self.dtype, self.device):
            raise ValueError("DecodeChain: every node must share w_bit / asym / dtype / device")
        if scales.dtype != x.dtype:
            raise ValueError(f"scales dtype {scales.dtype} must match x dtype {x.dtype}")
        if out is None:
            y = torch.empty((1, N), dtype=x.dtype, device=x.device)
        else:
            if tuple(out.shape) != (1, N) or out.dtype != x.dtype or out.device != x.device or not out.is_contiguous():
                raise ValueError("DecodeChain: out must be a contiguous [1, N] tensor of x's dtype on x's device")
            y = out
        self.nodes.append((x, y, qweight.contiguous(), scales.contiguous(), zeros.contiguous(), K, N, G))
        return y

    def build(self):
        """Hazard analysis + plan image (host) -> device buffer.  Once, outside CUDA-graph capture."""
        n = len(self.nodes)
        if n == 0:
            raise ValueError("DecodeChain: no nodes")
        lib = _cabi.lib()
        arr = (_cabi.ChainNode * n)()
        for i, (x, y, qw, sc, zr, K, N, G) in enumerate(self.nodes):
            arr[i].x, arr[i].y = x.data_ptr(), y.data_ptr()
            arr[i].qweight, arr[i].scales, arr[i].zeros = qw.data_ptr(), sc.data_ptr(), zr.data_ptr()
            arr[i].K, arr[i].N, arr[i].G = K, N, G
        nbytes = lib.b200bit_mpq_chain_plan_bytes(ctypes.cast(arr, ctypes.c_void_p), n)
        with torch.cuda.device(self.device):
            self.plan = torch.zeros(nbytes + 128, dtype=torch.uint8, device=self.device)
            base = (self.plan.data_ptr() + 127) & ~127
            torch.cuda.current_stream().synchronize()        # the zero fill is complete before the synchronous copy
            _cabi.check(lib.b200bit_mpq_chain_build(ctypes.cast(arr, ctypes.c_void_p), n, self.w_bit, int(self.asym),
                                                    _cabi.dtype_code(self.dtype), base, nbytes, self.info))
        self._base = base
        return self

    # ---- running --------------------------------------------------------------------------------------------
    def launch(self):
        """Enqueue the whole chain on the current stream (one cooperative kernel launch)."""
        if self.plan is None:
            self.build()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().b200bit_mpq_chain_launch(self._base, self.info, 0,
                                                             torch.cuda.current_stream().cuda_stream))
        return self.outputs

    def check(self):
        """Synchronise and raise if a dependency wait inside the kernel gave up (a broken plan; never expected)."""
        flag = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().b200bit_mpq_chain_status(self._base, self.info, ctypes.byref(flag),
                                                             torch.cuda.current_stream().cuda_stream))
        if flag.value:
            raise _cabi.B200BitError("b200bit: decode chain: a dependency wait timed out inside the kernel")

    @property
    def grid(self):
        return int(self.info[2])

    @property
    def ring_slots(self):
        return int(self.info[3])

    # ---- recording through the reference-facing API ----------------------------------------------------------
    @classmethod
    def capture(cls, fn):
        """Run `fn()` with every M == 1 `q_linear_cuda.mpq_forward` call (i.e. every MPQLinearCuda.forward) recorded as a
        node instead of launched; returns the built chain, `chain.outputs` = what fn returned."""
        global _recorder
        if _recorder is not None:
            raise RuntimeError("DecodeChain.capture: already recording")
        chain = cls()
        _recorder = chain
        try:
            chain.outputs = fn()
        finally:
            _recorder = None
        return chain.build()
