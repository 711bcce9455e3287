"""2-GPU check of the optional tensor-parallel wrappers (NCCL all-reduce / all-gather around the CUDA kernels).
Needs two devices (`gpurun --gpus 2`); passed on 2 x B200 in round 2 (tools/r2_gpu24.sh).  On a one-GPU box it skips."""
import os

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import sys
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_mpq_inputs
    from bitorch_engine_b200.extensions import q_linear_cuda
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import tensor_parallel as tp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    inp = make_mpq_inputs(4096, 4096, 4, 128, "f16", False, M=1, seed=9, device=f"cuda:{rank}")
    full = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, False)
    shard = tp.shard_row_parallel(inp["qweight"], inp["scales"], inp["zeros"], 4, 128, rank, world)
    y = tp.row_parallel_forward(inp["x"], shard, 4, False)
    cshard = tp.shard_column_parallel(inp["qweight"], inp["scales"], inp["zeros"], 4, False, rank, world)
    yc = tp.column_parallel_forward(inp["x"], cshard, 4, False, gather=True)
    torch.cuda.synchronize()
    rel = float((y.float() - full.float()).norm() / full.float().norm())
    relc = float((yc.float() - full.float()).norm() / full.float().norm())
    q.put((rank, rel, relc))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2 or os.environ.get("B200BIT_TEST_TP") == "0",
                    reason="needs two GPUs (B200BIT_TEST_TP=0 switches it off)")
def test_two_gpu_row_and_column_parallel():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for _, rel, relc in res:
        assert rel < 2e-3          # two fp16 roundings (one per partial) instead of one
        assert relc < 1e-3         # column slices: same arithmetic per column (expected bit-identical)
