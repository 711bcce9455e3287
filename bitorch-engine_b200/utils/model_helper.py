"""Twins of the helpers in bitorch_engine/utils/model_helper.py that sit on the hot path."""
from typing import List, Tuple

import torch


def flatten_x(x: torch.Tensor) -> Tuple[torch.Tensor, List[int]]:
    """[..., K] -> ([prod(...), K] view, leading shape)   (model_helper.py:10-29)."""
    lead = list(x.shape[:-1])
    return x.reshape(-1, x.shape[-1]) if not x.is_contiguous() else x.view(-1, x.shape[-1]), lead


def unflatten_x(x: torch.Tensor, shape: List[int]) -> torch.Tensor:
    """inverse of flatten_x on the output (model_helper.py:32-50)."""
    return x.view(list(shape) + [x.shape[-1]])


def prepare_bie_layers(model: torch.nn.Module, layers=None) -> None:
    """Call prepare_params() on every quantised layer of `model` (model_helper.py: prepare_bie_layers)."""
    for module in model.modules():
        if layers is not None and not isinstance(module, tuple(layers)):
            continue
        fn = getattr(module, "prepare_params", None)
        if callable(fn) and module is not model:
            fn()
