from bitorch_engine_b200.layers.qlinear.binary.cpp import BinaryLinearCPP, BinaryLinearForward  # noqa: F401
