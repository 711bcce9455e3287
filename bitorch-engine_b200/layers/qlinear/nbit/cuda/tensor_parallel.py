"""Optional tensor parallelism for an MPQ layer (no counterpart in the reference, which is
single-GPU): the two classic splits expressed on the packed format, so that a shard is again a valid
(qweight, scales, zeros) triple for `q_linear_cuda.mpq_forward`.

  column parallel  split N: every rank computes its slice of y from the whole x; nothing is exchanged (outputs can be
                   gathered by the caller if the next layer is not row parallel).
  row parallel     split K at group boundaries: every rank computes a partial y from its slice of x; ONE all-reduce
                   (sum) on the [M, N] output -- the only collective of the whole path (`torch.distributed`, NCCL on GPUs).

At decode batch sizes the all-reduce of an 8 KB vector is latency-bound and slower than running independent replicas,
which is why bench.py shards by replica; this module is for models that do not fit one GPU.
Status: shard arithmetic and the reduction are covered on CPU (gloo, world size 2) with the oracle standing in for the
kernel (tests/test_tensor_parallel_cpu.py); the NCCL path has a 2-GPU test (tests/test_gpu_tensor_parallel.py) that
needs `gpurun --gpus 2`."""
import torch
import torch.distributed as dist


def _bounds(total, unit, rank, world):
    """[lo, hi) of `total` split into `world` contiguous parts that are multiples of `unit`."""
    if total % unit != 0:
        raise ValueError(f"{total} is not a multiple of {unit}")
    units = total // unit
    if units < world:
        raise ValueError(f"cannot split {units} units of {unit} over {world} ranks")
    lo = (units * rank // world) * unit
    hi = (units * (rank + 1) // world) * unit
    return lo, hi


def shard_column_parallel(qweight, scales, zeros, w_bit, asym, rank, world):
    """Slice of the output columns owned by `rank`: returns (qweight, scales, zeros, (n_lo, n_hi)).  Column boundaries
    are multiples of 32 so that packed asymmetric zero points (32 / w_bit per word) split cleanly."""
    N = qweight.shape[1]
    lo, hi = _bounds(N, 32, rank, world)
    z = zeros[:, lo * w_bit // 32: hi * w_bit // 32] if asym else zeros[:, lo:hi]
    return qweight[:, lo:hi].contiguous(), scales[:, lo:hi].contiguous(), z.contiguous(), (lo, hi)


def shard_row_parallel(qweight, scales, zeros, w_bit, group_size, rank, world):
    """Slice of the input channels owned by `rank` (whole groups, whole packed words): returns
    (qweight, scales, zeros, (k_lo, k_hi)).  g_idx of a shard is arange(k_hi - k_lo) // group_size again."""
    nb = 32 // w_bit
    K = qweight.shape[0] * nb
    unit = group_size if group_size % nb == 0 else group_size * nb
    lo, hi = _bounds(K, unit, rank, world)
    return (qweight[lo // nb: hi // nb].contiguous(), scales[lo // group_size: hi // group_size].contiguous(),
            zeros[lo // group_size: hi // group_size].contiguous(), (lo, hi))


def row_parallel_forward(x, shard, w_bit, asym, group=None, forward_fn=None):
    """y = all_reduce_sum(x[:, k_lo:k_hi] @ W_shard).  `shard` is the tuple returned by shard_row_parallel;
    `forward_fn(x_local, qweight, scales, zeros)` defaults to the CUDA plugin function."""
    qweight, scales, zeros, (lo, hi) = shard
    if forward_fn is None:
        from .....extensions import q_linear_cuda

        def forward_fn(xl, q, s, z):
            return q_linear_cuda.mpq_forward(xl, q, s, z, None, 16, w_bit, asym)
    y = forward_fn(x[:, lo:hi].contiguous(), qweight, scales, zeros)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y


def column_parallel_forward(x, shard, w_bit, asym, gather=False, group=None, forward_fn=None):
    """y[:, n_lo:n_hi] = x @ W[:, n_lo:n_hi]; with gather=True every rank returns the full [M, N] (all-gather)."""
    qweight, scales, zeros, (lo, hi) = shard
    if forward_fn is None:
        from .....extensions import q_linear_cuda

        def forward_fn(xl, q, s, z):
            return q_linear_cuda.mpq_forward(xl, q, s, z, None, 16, w_bit, asym)
    y = forward_fn(x.contiguous(), qweight, scales, zeros)
    if not gather or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return y
    parts = [torch.empty_like(y) for _ in range(dist.get_world_size(group))]      # equal widths: N / 32 divisible by world
    dist.all_gather(parts, y, group=group)
    return torch.cat(parts, dim=1)
