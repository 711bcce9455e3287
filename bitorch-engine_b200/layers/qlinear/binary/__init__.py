from .layer import BinaryLinearBase, BinaryLinearParameter  # noqa: F401
