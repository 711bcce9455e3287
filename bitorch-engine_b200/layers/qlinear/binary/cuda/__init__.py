from .bmm import BMM  # noqa: F401
from .layer import BinaryLinearCuda, BinaryLinearForward  # noqa: F401
