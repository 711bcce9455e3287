"""one exl2 forward (for ncu): python tools/exl2_one.py M K N"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bitorch_engine_b200  # noqa
from bitorch_engine_b200.extensions import q_linear_cuda
from test_gpu_mbwq import _make_exl2, STRATEGIES
M, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
layer = _make_exl2(K, N, STRATEGIES[0], seed=1, permute=True)
x = torch.randn((M, K), device="cuda").half()
for _ in range(4):
    y = q_linear_cuda.mbwq_exl2_forward(x, layer.qweight.data, layer.scales, layer.zeros, layer.q_perm, layer.q_group_map, layer.rows, False)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
