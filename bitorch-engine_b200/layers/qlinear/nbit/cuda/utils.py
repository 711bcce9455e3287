"""unpack_qweight / pack_fp_weight / make_group_map with the reference's signatures
(bitorch_engine/layers/qlinear/nbit/cuda/utils.py:5-186), each a single kernel launch instead of a chain of torch
elementwise ops; results are bit-identical (tests/test_gpu_mpq_aux.py)."""
import torch

from .....extensions import q_linear_cuda


def _meta(qweight, name, default=None):
    return getattr(qweight, name, default)


def unpack_qweight(qweight) -> torch.Tensor:
    """fp weight [K,N] from an MPQWeightParameter (utils.py:5-69)."""
    layer_type = _meta(qweight, "layer_type")
    if layer_type is None or layer_type == -1:
        raise ValueError("Error: invalid attribute of qweight in 'unpack_qweight'.")
    if layer_type == 1:
        return q_linear_cuda.mpq_dequant(qweight.data, qweight.scales, qweight.zeros, qweight.g_idx, qweight.w_bit,
                                         qweight.asym)
    if layer_type == 2:
        if _meta(qweight, "q_group_map") is None:
            return q_linear_cuda.mbwq_q42fp_weight(qweight.data, qweight.scales, qweight.zeros, qweight.group_size,
                                                   qweight.w_bit, qweight.q_perm)
        return q_linear_cuda.mbwq_exl2fp_weight(qweight.data, qweight.scales, qweight.zeros, qweight.q_perm,
                                                qweight.q_group_map, qweight.rows)
    raise NotImplementedError("Error: 'layer_type' not yet supported!")


def pack_fp_weight(weight: torch.Tensor, qweight, unpacked_zeros: torch.Tensor = None) -> torch.Tensor:
    """int32 packed weight from fp weight + the quantisation attributes on `qweight` (utils.py:72-147)."""
    layer_type = _meta(qweight, "layer_type")
    if layer_type is None or layer_type == -1:
        raise ValueError("Error: invalid 'layer_type' attribute in 'unpack_qweight' method.")
    if not (layer_type == 1 or (layer_type == 2 and _meta(qweight, "q_group_map") is None)):
        raise NotImplementedError("Error: pack_fp_weight for MBWQLinear using channel-mix quantization not supported yet.")
    asym = bool(qweight.asym)
    zeros = qweight.zeros
    zeros_unpacked = False
    if asym:
        if unpacked_zeros is not None:
            zeros, zeros_unpacked = unpacked_zeros, True
        elif zeros.dtype != torch.int32:
            raise ValueError("Error: Got invalid dtype of qweight.zeros while packing fp weight.")
    perm = None
    if not asym and qweight.g_idx is None:
        perm = _meta(qweight, "q_perm")
    return q_linear_cuda.mpq_pack_weight(weight, qweight.scales, zeros, qweight.g_idx, qweight.w_bit, asym,
                                         zeros_unpacked=zeros_unpacked, perm=perm)


def make_group_map(q_groups: torch.Tensor, num_qrows: int) -> torch.Tensor:
    """(group, rows-left-in-group) pairs for every weight row of an exl2-style matrix (utils.py:150-186).
    Vectorised: the reference walks the rows in a Python loop."""
    qg = q_groups.to(torch.int64).cpu().view(-1, 2)
    bits, first = qg[:, 0], qg[:, 1]
    last = torch.cat([first[1:], torch.tensor([num_qrows], dtype=torch.int64)])
    rows = (last - first) * 32 // bits
    group = torch.repeat_interleave(torch.arange(len(rows), dtype=torch.int64), rows)
    start = torch.cumsum(rows, 0) - rows
    left = rows[group] - (torch.arange(int(rows.sum()), dtype=torch.int64) - start[group])
    return torch.stack([group, left], dim=1).reshape(-1).to(torch.short).to(q_groups.device)
