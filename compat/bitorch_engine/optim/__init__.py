from bitorch_engine_b200.optim import DiodeMix, GaLoreProjector  # noqa: F401
