"""bitorch-engine_b200 -- B200 (sm_100a) implementation of bitorch-engine's low-bit Linear hot path.

Layout:
    csrc/          hand-written CUDA kernels + the C ABI (include/b200bit.h) -> lib/libb200bit.so
    _cabi.py       ctypes loader (fails loudly when the library is missing; there is NO CPU / eager fallback)
    extensions/    python-visible twins of the reference's pybind modules (q_linear_cuda, binary_linear_cuda)
    layers/        host-side mirror of the reference nn.Module / autograd.Function surface
    optim/         DiodeMix with the fused update kernels
"""
__version__ = "0.1.0"
