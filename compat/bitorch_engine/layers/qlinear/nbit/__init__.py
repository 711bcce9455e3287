from bitorch_engine_b200.layers.qlinear.nbit import MPQLinearBase, MPQWeightParameter  # noqa: F401
