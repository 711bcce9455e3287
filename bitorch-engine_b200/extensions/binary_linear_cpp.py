"""Twin of the reference CPU extension `binary_linear_cpp`
(bitorch_engine/layers/qlinear/binary/cpp/binary_linear.cpp:494-518): forward(input, weights, m, n, k) and
w_pack(weights, n, k) on HOST tensors, backed by the from-scratch host kernels in csrc/binary_cpu.cpp."""
import torch

from .. import _cabi


def _check_cpu_float(t, name):
    if not isinstance(t, torch.Tensor) or t.is_cuda:
        raise RuntimeError(f"{name} must be a CPU tensor")
    if t.dtype != torch.float32:
        # the reference reads data_ptr<float>() (binary_linear.cpp:425): any other dtype raises there too
        raise RuntimeError(f"expected scalar type Float but found {t.dtype} ({name})")


def forward(input: torch.Tensor, weights: torch.Tensor, m: int, n: int, k: int) -> torch.Tensor:
    """[m, n] float32 = sign(input) @ sign(weights).T; `weights` is float [n, k] or the packed uint8 [k*n/8] stream --
    told apart by numel, as the reference does (:506-511)."""
    _check_cpu_float(input, "input")
    x = input.contiguous()
    if x.numel() != m * k:
        raise ValueError(f"input has {x.numel()} elements, expected m*k = {m * k}")
    packed = weights.numel() == k * n // 8
    if packed:
        if weights.dtype != torch.uint8:
            raise RuntimeError("packed weights must be uint8")
    else:
        _check_cpu_float(weights, "weights")
        if weights.numel() != n * k:
            raise ValueError(f"weights has {weights.numel()} elements, expected n*k = {n * k}")
    w = weights.contiguous()
    out = torch.empty((m, n), dtype=torch.float32)
    _cabi.check(_cabi.lib().b200bit_cpu_binary_forward(x.data_ptr(), w.data_ptr(), int(packed), out.data_ptr(), m, n, k,
                                                       torch.get_num_threads()))
    return out


def w_pack(weights: torch.Tensor, n: int, k: int) -> torch.Tensor:
    """uint8 [k*n/8]: byte (k/8)*n + j holds the signs of weights[j, 8*(k/8) .. +7], LSB first (:447-466)."""
    _check_cpu_float(weights, "weights")
    w = weights.contiguous()
    if w.numel() != n * k:
        raise ValueError(f"weights has {w.numel()} elements, expected n*k = {n * k}")
    out = torch.empty((k * n // 8,), dtype=torch.uint8)
    _cabi.check(_cabi.lib().b200bit_cpu_binary_pack(w.data_ptr(), out.data_ptr(), n, k))
    return out
