from bitorch_engine_b200.utils.quant_operators import (nv_tensor_quant, bit_set, get_binary_row, get_binary_col,  # noqa: F401
                                                       q8_quantization, q4_quantization, gptq_style_zeros_packing,
                                                       gptq_style_unpacking)
