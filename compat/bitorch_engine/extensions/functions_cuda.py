from bitorch_engine_b200.extensions.functions_cuda import (fp32toint4, tensor_pack_to_uint8, uint8_to_unpacked_tensor,  # noqa: F401
                                                           q4_pack, q4_unpack, q4_unpack_and_scaling)
