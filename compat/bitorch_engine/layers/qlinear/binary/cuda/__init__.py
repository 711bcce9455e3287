from bitorch_engine_b200.layers.qlinear.binary.cuda import BinaryLinearCuda, BinaryLinearForward, BMM  # noqa: F401
