from bitorch_engine_b200.utils.model_helper import flatten_x, unflatten_x, prepare_bie_layers  # noqa: F401
from bitorch_engine_b200.utils.quant_operators import init_weight  # noqa: F401
from bitorch_engine_b200.optim.update import qweight_update_fn  # noqa: F401
