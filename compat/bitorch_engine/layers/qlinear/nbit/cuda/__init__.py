from bitorch_engine_b200.layers.qlinear.nbit.cuda import (MPQLinearCuda, MPQLinearCudaFunction,  # noqa: F401
                                                          MBWQLinearCuda, MBWQLinearCudaFunction)
