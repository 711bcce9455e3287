from bitorch_engine_b200.extensions.binary_linear_cpp import forward, w_pack  # noqa: F401
