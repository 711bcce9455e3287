"""Mirror of bitorch_engine/functions/cuda/functions.py: the user-facing wrappers around the `functions_cuda` extension
(same names, arguments and assertions; :8-177)."""
import torch

from ...extensions import functions_cuda


def fp32toint4(input: torch.Tensor) -> torch.Tensor:
    return functions_cuda.fp32toint4(input)                                     # functions.py:8-32


def tensor_to_packed_uint8(input: torch.Tensor) -> torch.Tensor:
    return functions_cuda.tensor_pack_to_uint8(input)                           # functions.py:35-56


def unpack_uint8_tensor(input: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    return functions_cuda.uint8_to_unpacked_tensor(input, scale)                # functions.py:59-88


def q4_pack_tensor(input: torch.Tensor, is_transpose: bool = False) -> torch.Tensor:
    assert input.dtype == torch.int32, "Error: input tensor dtype should be int32"
    return functions_cuda.q4_pack(input, is_transpose)                          # functions.py:91-122


def q4_unpack_tensor(input: torch.Tensor, is_transpose: bool = False) -> torch.Tensor:
    assert input.dtype == torch.int8, "Error: input tensor dtype should be int8."
    return functions_cuda.q4_unpack(input, is_transpose)                        # functions.py:125-150


def q4_unpack_and_scaling_tensor(input: torch.Tensor, scale: float, is_transpose: bool = False) -> torch.Tensor:
    assert input.dtype == torch.int8, "Error: input tensor dtype should be int8."
    return functions_cuda.q4_unpack_and_scaling(input, scale, is_transpose)     # functions.py:153-177
