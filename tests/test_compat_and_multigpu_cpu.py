"""CPU: (1) the `compat/bitorch_engine` alias exposes the reference's module paths; (2) the N>1 host logic of bench.py
(rank-sharded replicas, barrier, max-over-ranks timing, whole-job aggregation) under a world_size-2 gloo group."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_module_paths_resolve():
    code = ("import importlib;"
            "from bitorch_engine.layers.qlinear.nbit.cuda import MPQLinearCuda, MBWQLinearCuda;"
            "from bitorch_engine.layers.qlinear.nbit import MPQWeightParameter;"
            "from bitorch_engine.layers.qlinear.binary.cuda import BinaryLinearCuda, BMM;"
            "from bitorch_engine.optim import DiodeMix;"
            "from bitorch_engine.utils.model_helper import flatten_x, prepare_bie_layers, save_checkpoint, load_checkpoint, pad_embedding_dim;"
            "from bitorch_engine.utils.quant_operators import q4_quantization, q8_quantization, get_binary_row, gptq_style_zeros_packing;"
            "m = importlib.import_module('bitorch_engine.extensions.q_linear_cuda');"
            "assert all(hasattr(m, n) for n in ['mpq_forward','mpq_grad_input','mbwq_trans_qweight','mbwq_q42fp_weight',"
            "'mbwq_q4_forward','mbwq_exl2fp_weight','mbwq_exl2_forward']);"
            "b = importlib.import_module('bitorch_engine.extensions.binary_linear_cuda');"
            "assert all(hasattr(b, n) for n in ['forward','w_pack','mm']);"
            "f = importlib.import_module('bitorch_engine.extensions.functions_cuda');"
            "assert all(hasattr(f, n) for n in ['fp32toint4','tensor_pack_to_uint8','uint8_to_unpacked_tensor','q4_pack',"
            "'q4_unpack','q4_unpack_and_scaling']);"
            "from bitorch_engine.functions.cuda import q4_pack_tensor, q4_unpack_tensor, q4_unpack_and_scaling_tensor;"
            "from bitorch_engine.functions.cuda.functions import tensor_to_packed_uint8, unpack_uint8_tensor;"
            "c = importlib.import_module('bitorch_engine.extensions.binary_linear_cpp');"
            "assert all(hasattr(c, n) for n in ['forward','w_pack']);"
            "from bitorch_engine.layers.qlinear.binary.cpp import BinaryLinearCPP;"
            "from bitorch_engine.utils.quant_operators import gptq_style_unpacking, gptq_style_zeros_packing;"
            "from bitorch_engine.utils.model_helper import update_zeros, qweight_update_fn;"
            "print('ok')")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "compat"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms_local, ms_e2e_local = (10.0, 12.0) if rank == 0 else (14.0, 11.0)
    ms, ms_e2e = bench.max_over_ranks([ms_local, ms_e2e_local], world, device="cpu")
    seeds = bench.rank_seed(rank)
    q.put((rank, ms, ms_e2e, seeds, bench.whole_job_rate(world, steps=7, ms=ms)))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_aggregation():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ms0, e0, s0, rate0), (r1, ms1, e1, s1, rate1) = res
    assert ms0 == ms1 == 14.0 and e0 == e1 == 12.0           # max over ranks, identical on every rank
    assert s0 != s1                                           # independent request streams
    assert rate0 == pytest.approx(2 * 7 / 14.0e-3)            # whole-job tokens/s = all ranks' units / max time
