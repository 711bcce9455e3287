from bitorch_engine_b200.layers.qlinear.binary import BinaryLinearBase, BinaryLinearParameter  # noqa: F401
from bitorch_engine_b200.layers.qlinear.binary.cuda import BinaryLinearCuda as BinaryLinear  # noqa: F401  (best impl, binary/__init__.py:5-14)
