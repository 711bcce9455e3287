"""CPU: the torch-CPU port used as the timed CPU baseline reproduces the reference's unpack_qweight bit for bit."""
import numpy as np
import pytest
import torch

from oracle import torch_cpu
from helpers import load_nbit_cases

CASES = load_nbit_cases()
TDT = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


def _t(c, key):
    a = getattr(c, key)
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16).copy()).view(TDT[c.dt])
    return torch.from_numpy(a.copy())


@pytest.mark.parametrize("c", CASES, ids=[c.id for c in CASES])
def test_torch_cpu_dequant_equals_reference(c):
    W = torch_cpu.dequant(_t(c, "qweight"), _t(c, "scales"), _t(c, "zeros"), _t(c, "g_idx"), c.w_bit, c.asym)
    assert W.dtype == TDT[c.dt]
    assert torch.equal(W, _t(c, "W"))
    y = torch_cpu.mpq_forward(_t(c, "x").float(), None, None, None, None, c.w_bit, c.asym, cached_weight=W.float())
    np.testing.assert_allclose(y.numpy(), c.y, rtol=1e-5, atol=1e-5)
