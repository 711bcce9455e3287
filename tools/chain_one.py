"""a 4-layer decode chain at small sizes (for compute-sanitizer): python tools/chain_one.py [layers]
`layers`: launch the same layers one by one instead (mpq_imma_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bitorch_engine_b200  # noqa
from bitorch_engine_b200.extensions import q_linear_cuda
from bitorch_engine_b200.decode_chain import DecodeChain
from helpers import make_mpq_inputs
shapes = [(1024, 1024), (1024, 2048), (2048, 1024), (1024, 1024)]
ls = [make_mpq_inputs(K, N, 4, 128, "f16", False, M=1, seed=i, device="cuda") for i, (K, N) in enumerate(shapes)]


def run():
    h = ls[0]["x"]
    for d in ls:
        h = q_linear_cuda.mpq_forward(h, d["qweight"], d["scales"], d["zeros"], d["g_idx"], 16, 4, False)
    return h


want = run()
torch.cuda.synchronize()
if len(sys.argv) > 1:
    print("layers ok")
    sys.exit(0)
chain = DecodeChain.capture(run)
for _ in range(2):
    chain.launch()
chain.check()
print("chain ok", bool(torch.equal(chain.outputs, want)))
