from .layer import BinaryLinearCPP, BinaryLinearForward  # noqa: F401
