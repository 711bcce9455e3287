from .diode_beta import DiodeMix  # noqa: F401
from .galore_projector import GaLoreProjector  # noqa: F401
