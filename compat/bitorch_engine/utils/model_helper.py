from bitorch_engine_b200.utils.model_helper import (flatten_x, unflatten_x, prepare_bie_layers, pad_embedding_dim,  # noqa: F401
                                                    pad_last_2_dims_to_multiple_of_128,
                                                    binary_matmul_forward_post_processing, pack_bie_layers,
                                                    save_checkpoint, load_checkpoint, update_zeros)
from bitorch_engine_b200.utils.quant_operators import init_weight  # noqa: F401
from bitorch_engine_b200.optim.update import qweight_update_fn  # noqa: F401
