from .layer import MPQLinearBase, MPQWeightParameter  # noqa: F401
