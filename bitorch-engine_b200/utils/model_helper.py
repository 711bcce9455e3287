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


def pad_embedding_dim(weight: torch.Tensor) -> torch.Tensor:
    """Pad the embedding dimension (dim 1) with -1 columns up to the next multiple of 8, the unit of the sign-bit
    packing (model_helper.py:54-82).  -1 packs to bit 0, so the padding never flips a popcount."""
    extra = -weight.shape[1] % 8
    if extra == 0:
        return weight
    fill = torch.full((weight.shape[0], extra), -1.0, dtype=torch.float, device=weight.device)
    return torch.cat([weight, fill], dim=1)


def pad_last_2_dims_to_multiple_of_128(tensor: torch.Tensor):
    """Zero-pad the last two dimensions up to multiples of 128 (the BTC tile of the binary matmul); returns the padded
    tensor and the number of rows added to the second-to-last dimension (model_helper.py:85-117)."""
    pad_last = -tensor.shape[-1] % 128
    pad_sec = -tensor.shape[-2] % 128
    if pad_last or pad_sec:
        tensor = torch.nn.functional.pad(tensor, (0, pad_last, 0, pad_sec), mode="constant", value=0)
    return tensor, pad_sec


def binary_matmul_forward_post_processing(tensor: torch.Tensor, shape_pre: list, x_pad_sec_last: int,
                                          y_pad_sec_last: int, k: int) -> torch.Tensor:
    """Undo the padding of a batched binary matmul result [b, m, n], restore the leading shape and map popcounts back
    to the +-1 domain: k - 2 * popc (model_helper.py:120-155)."""
    if x_pad_sec_last > 0:
        tensor = tensor[:, :-x_pad_sec_last, :]
    if y_pad_sec_last > 0:
        tensor = tensor[:, :, :-y_pad_sec_last]
    tensor = tensor.reshape(list(shape_pre) + [tensor.size(-2), tensor.size(-1)])
    return k - 2 * tensor


def _quantised_layer_bases():
    from ..layers.qlinear.binary import BinaryLinearBase
    from ..layers.qlinear.nbit import MPQLinearBase
    return [BinaryLinearBase, MPQLinearBase]


def pack_bie_layers(model: torch.nn.Module, qweight_only: bool = True, layers=None) -> None:
    """Call generate_quantized_weight(qweight_only) on every quantised sub-module (model_helper.py:199-235): the step in
    front of torch.save.  The default layer list holds the bases this package provides (binary and MPQ Linear)."""
    layers = tuple(layers) if layers else tuple(_quantised_layer_bases())
    for idx, module in enumerate(model.modules()):
        if idx > 0 and isinstance(module, layers):
            module.generate_quantized_weight(qweight_only=qweight_only)


def save_checkpoint(model: torch.nn.Module, name: str, qweight_only: bool = True) -> None:
    """Pack, then torch.save({'state_dict': ...}) (model_helper.py:238-263)."""
    pack_bie_layers(model, qweight_only)
    torch.save({"state_dict": model.state_dict()}, name)


def load_checkpoint(model: torch.nn.Module, checkpoint_path: str, qweight_only: bool = True) -> None:
    """Pack (so that the packed buffers exist), then load the state dict non-strictly (model_helper.py:266-283)."""
    pack_bie_layers(model, qweight_only)
    checkpoint = torch.load(checkpoint_path)
    model.load_state_dict(checkpoint["state_dict"], strict=False)


def update_zeros(qweight, w, norm_grad, step_size, z_unpacked=None):
    """The every-5th-step zero-point update of DiodeMix (model_helper.py:330-360).

    MBWQ (layer_type 2): zeros += step_size * mean over each group of the q_perm-gathered normalised gradient.
    MPQ  (layer_type 1, g_idx given): the integer zero points follow their rows' gradients,
         zeros = pack(mean_g(z_unpacked[g_idx] + step_size * norm_grad)), stored packed (the buffer object is replaced)."""
    from .quant_operators import gptq_style_zeros_packing
    groups, cols = qweight.scales.size(0), qweight.scales.size(-1)
    if qweight.layer_type == 2:
        index = qweight.q_perm.unsqueeze(1).repeat(1, w.size(1)).long()
        zeros_grad = torch.gather(norm_grad, dim=0, index=index)
        qweight.zeros.add_(step_size * zeros_grad.view(-1, w.size(0) // groups, cols).mean(1))
    elif qweight.layer_type == 1 and qweight.g_idx is not None:
        g_idx = qweight.g_idx.long()
        zeros_unpack = z_unpacked[g_idx]
        zeros_unpack.add_(step_size * norm_grad)
        order = torch.argsort(g_idx, dim=0)
        zeros = zeros_unpack[order, :].view(-1, w.size(0) // groups, cols).mean(1)
        qweight.zeros = gptq_style_zeros_packing(zeros, qweight.w_bit, zeros.size(-1), qweight.group_size)
    else:
        raise NotImplementedError(
            "qweight.layer_type: '{}' has not been supported yet.".format(str(qweight.layer_type)))
