import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bitorch_engine_b200  # noqa
from bitorch_engine_b200 import _cabi
from helpers import make_mpq_inputs
lib = _cabi.lib()
M, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
inp = make_mpq_inputs(K, N, 4, 128, "f16", False, M=M, seed=1, device="cuda")
y = torch.empty((M, N), dtype=torch.float16, device="cuda")
for _ in range(3):
    _cabi.check(lib.b200bit_mpq_forward_tc(inp["x"].data_ptr(), inp["qweight"].data_ptr(), inp["scales"].data_ptr(), inp["zeros"].data_ptr(),
                                           y.data_ptr(), M, K, N, K // 128, 4, 0, _cabi.F16, None, 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok")
# the product route (workspace for split-K at small M comes from the shim)
from bitorch_engine_b200.extensions import q_linear_cuda
for _ in range(3):
    y2 = q_linear_cuda.mpq_forward(inp["x"], inp["qweight"], inp["scales"], inp["zeros"], inp["g_idx"], 16, 4, False)
torch.cuda.synchronize()
print("shim ok", float((y2.float() - y.float()).abs().max()))
