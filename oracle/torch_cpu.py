"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- multi-threaded torch-CPU restatement of the reference's only
CPU-capable n-bit path: dequantise the whole weight with elementwise tensor ops, then matmul
(bitorch_engine/layers/qlinear/nbit/cuda/utils.py:31-51 `unpack_qweight` + mpq_layer.py:59-63 `torch.matmul`;
BASELINE.md section 3 / SURVEY.md section 8d config #1 "Q4LinearCPU").

bench.py times this on the GPU box's host cores (`cpu_baseline`, `--impl reference`, kind = "port" because
/root/reference does not exist there).  tests/test_oracle_torch_cpu.py pins it against the golden vectors."""
import torch


def unpack_codes(qweight: torch.Tensor, w_bit: int) -> torch.Tensor:
    """int32 [K*b/32, N] -> integer codes [K, N] (int8, int16 for 8-bit); same op sequence as utils.py:31-34:
    broadcast right-shift over the nb sub-positions, narrow, mask."""
    nb = 32 // w_bit
    shifts = torch.arange(0, 32, w_bit, dtype=torch.int32, device=qweight.device).view(1, nb, 1)
    codes = (qweight.unsqueeze(1) >> shifts).to(torch.int16 if w_bit == 8 else torch.int8)
    codes = codes.reshape(-1, qweight.shape[-1])
    return codes.bitwise_and_((1 << w_bit) - 1)


def unpack_zeros(qzeros: torch.Tensor, w_bit: int) -> torch.Tensor:
    """packed int32 [G, N*b/32] -> [G, N] integer zero points incl. the +1 (utils.py:36-41)."""
    nb = 32 // w_bit
    shifts = torch.arange(0, 32, w_bit, dtype=torch.int32, device=qzeros.device).view(1, 1, nb)
    z = (qzeros.unsqueeze(2) >> shifts).to(torch.int16 if w_bit == 8 else torch.int8)
    z = z.bitwise_and_((1 << w_bit) - 1) + 1
    return z.reshape(qzeros.shape[0], -1)


def dequant(qweight, scales, zeros, g_idx, w_bit, asym) -> torch.Tensor:
    """fp weight [K, N] in scales.dtype with the reference's per-op rounding (utils.py:36-51)."""
    codes = unpack_codes(qweight, w_bit)
    gi = g_idx.long()
    if asym:
        return scales[gi] * (codes - unpack_zeros(zeros, w_bit)[gi])
    return codes * scales[gi] - zeros[gi]


def mpq_forward(x, qweight, scales, zeros, g_idx, w_bit, asym, cached_weight=None):
    """The reference's M>32 branch, used here as its CPU path: y = x @ unpack_qweight(q) (mpq_layer.py:59-63)."""
    w = cached_weight if cached_weight is not None else dequant(qweight, scales, zeros, g_idx, w_bit, asym)
    return torch.matmul(x, w)
