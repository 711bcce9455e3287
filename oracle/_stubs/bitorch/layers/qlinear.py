from . import QLinearBase  # noqa: F401
from .register import QLinearImplementation  # noqa: F401
