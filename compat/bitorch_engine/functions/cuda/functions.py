from bitorch_engine_b200.functions.cuda.functions import *  # noqa: F401,F403
from bitorch_engine_b200.functions.cuda.functions import (fp32toint4, tensor_to_packed_uint8, unpack_uint8_tensor,  # noqa: F401
                                                          q4_pack_tensor, q4_unpack_tensor, q4_unpack_and_scaling_tensor)
