from bitorch_engine_b200.extensions.binary_linear_cuda import forward, w_pack, mm  # noqa: F401
