"""CPU (gloo, world size 2): the shard arithmetic of layers/qlinear/nbit/cuda/tensor_parallel.py and the one collective of
the path (all-reduce of the row-parallel partial outputs), with the numpy oracle standing in for the CUDA kernel."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from helpers import make_mpq_inputs, to_np_f32


def _oracle_forward(w_bit, asym, dt):
    from oracle import nbit

    def f(xl, q, s, z):
        zeros = z.numpy() if asym else to_np_f32(z)
        y = nbit.mpq_forward_exact(to_np_f32(xl), q.numpy(), to_np_f32(s), zeros, None, w_bit, asym)
        return torch.from_numpy(np.asarray(y, dtype=np.float32))
    return f


@pytest.mark.parametrize("w_bit,group,asym", [(4, 128, False), (4, 32, True), (2, 64, False), (8, 128, True)])
def test_shards_recompose_the_layer(w_bit, group, asym):
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import tensor_parallel as tp
    K, N, world = 1024, 256, 4
    inp = make_mpq_inputs(K, N, w_bit, group, "f16", asym, M=3, seed=5)
    full = _oracle_forward(w_bit, asym, "f16")(inp["x"], inp["qweight"], inp["scales"], inp["zeros"])
    rows = [tp.shard_row_parallel(inp["qweight"], inp["scales"], inp["zeros"], w_bit, group, r, world) for r in range(world)]
    assert [s[3] for s in rows] == [(r * K // world, (r + 1) * K // world) for r in range(world)]
    part = sum(tp.row_parallel_forward(inp["x"], s, w_bit, asym, forward_fn=_oracle_forward(w_bit, asym, "f16")) for s in rows)
    np.testing.assert_allclose(part.numpy(), full.numpy(), rtol=1e-5, atol=1e-5)
    cols = [tp.shard_column_parallel(inp["qweight"], inp["scales"], inp["zeros"], w_bit, asym, r, world) for r in range(world)]
    ys = [tp.column_parallel_forward(inp["x"], s, w_bit, asym, forward_fn=_oracle_forward(w_bit, asym, "f16")) for s in cols]
    np.testing.assert_allclose(torch.cat(ys, dim=1).numpy(), full.numpy(), rtol=1e-6, atol=1e-6)


def test_bad_splits_raise():
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import tensor_parallel as tp
    inp = make_mpq_inputs(256, 64, 4, 128, "f16", False)
    with pytest.raises(ValueError):
        tp.shard_row_parallel(inp["qweight"], inp["scales"], inp["zeros"], 4, 128, 0, 4)     # 2 groups over 4 ranks
    with pytest.raises(ValueError):
        tp.shard_column_parallel(inp["qweight"], inp["scales"], inp["zeros"], 4, False, 0, 4)  # 2 x 32 columns over 4 ranks


def _worker(rank, world, port, q):
    import sys
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import tensor_parallel as tp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = make_mpq_inputs(512, 128, 4, 128, "f16", False, M=2, seed=9)          # same seed on every rank: same layer
    f = _oracle_forward(4, False, "f16")
    full = f(inp["x"], inp["qweight"], inp["scales"], inp["zeros"])
    shard = tp.shard_row_parallel(inp["qweight"], inp["scales"], inp["zeros"], 4, 128, rank, world)
    y = tp.row_parallel_forward(inp["x"], shard, 4, False, forward_fn=f)         # all-reduce inside
    cshard = tp.shard_column_parallel(inp["qweight"], inp["scales"], inp["zeros"], 4, False, rank, world)
    yc = tp.column_parallel_forward(inp["x"], cshard, 4, False, gather=True, forward_fn=f)
    q.put((rank, float((y - full).abs().max()), float((yc - full).abs().max()), tuple(yc.shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_all_reduce_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[0] for r in res] == [0, 1]
    for _, err_row, err_col, shape in res:
        assert err_row < 1e-4 and err_col < 1e-6 and shape == (2, 128)
