from .mpq_layer import MPQLinearCuda, MPQLinearCudaFunction  # noqa: F401
from .mbwq_layer import MBWQLinearCuda, MBWQLinearCudaFunction  # noqa: F401
