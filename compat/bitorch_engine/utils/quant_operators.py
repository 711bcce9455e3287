from bitorch_engine_b200.utils.quant_operators import nv_tensor_quant  # noqa: F401
