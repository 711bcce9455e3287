"""CPU: the host-side mirror keeps the reference's persistent format: state_dict keys / shapes / dtypes before and after
prepare_params(), and the double-dequantised scales / zeros, against fixtures produced by the reference itself
(oracle/gen_golden.py gen_layers -> tests/golden/layer_specs.json, layer_prepare.npz)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLD

SPECS = json.load(open(os.path.join(GOLD, "layer_specs.json")))
PREP = np.load(os.path.join(GOLD, "layer_prepare.npz"))


def _spec(layer):
    return {k: [list(v.shape), str(v.dtype)] for k, v in layer.state_dict().items()}


@pytest.mark.parametrize("ci", range(len(SPECS)))
def test_state_dict_layout_and_prepare_params(ci):
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MPQLinearCuda
    spec = SPECS[ci]
    layer = MPQLinearCuda(256, 512, requires_grad=False, **spec["kwargs"])
    assert _spec(layer) == spec["before"]
    for name, buf in layer.named_buffers():
        key = f"l{ci}_{name}"
        if key in PREP.files:
            a = PREP[key]
            t = torch.from_numpy(a.view(np.int16).copy()).view(buf.dtype) if a.dtype == np.uint16 else torch.from_numpy(a.copy())
            buf.copy_(t)
    layer.prepare_params()
    assert _spec(layer) == spec["after"]
    derived = spec["kwargs"].get("use_gba_quant", True) and spec["kwargs"]["group_size"] < 256
    if not derived:      # scales / zeros are plain buffers there: nothing is computed by prepare_params
        return
    got = layer.scales.contiguous().view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(got, PREP[f"l{ci}_scales_out"]), "double-dequantised scales differ from the reference"
    if f"l{ci}_zeros_out" in PREP.files:
        gz = layer.zeros.contiguous().view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(gz, PREP[f"l{ci}_zeros_out"])


def test_parameter_metadata_and_update_signature():
    import inspect
    from bitorch_engine_b200.layers.qlinear.nbit import MPQWeightParameter
    p = MPQWeightParameter(torch.zeros((4, 8), dtype=torch.int32), requires_grad=False, w_bit=4, group_size=32, layer_type=1)
    for name in ("privileged_grad", "scales", "zeros", "g_idx", "w_bit", "asym", "group_size", "layer_type", "q_perm",
                 "q_group_map", "rows"):
        assert hasattr(p, name)
    sig = list(inspect.signature(MPQWeightParameter.update).parameters)
    assert sig == ["qweight", "exp_avg_s", "exp_avg_l", "step", "lr", "weight_decay", "beta1", "beta2", "eps", "dtype",
                   "correct_bias", "projector", "grad"]            # nbit/layer.py:86-89


def test_make_group_map_matches_reference_loop():
    from bitorch_engine_b200.layers.qlinear.nbit.cuda.utils import make_group_map
    q_groups = torch.tensor([4, 0, 4, 4, 2, 8, 2, 10], dtype=torch.short)   # (bits, first packed row) pairs
    got = make_group_map(q_groups, 12).tolist()
    exp = []
    for i in range(4):                                     # restatement of utils.py:171-184
        bits = int(q_groups[2 * i])
        qrows = (int(q_groups[2 * i + 3]) if i < 3 else 12) - int(q_groups[2 * i + 1])
        rows = qrows * 32 // bits
        for j in range(rows):
            exp += [i, rows - j]
    assert got == exp


def test_flatten_unflatten():
    from bitorch_engine_b200.utils.model_helper import flatten_x, unflatten_x
    x = torch.randn(2, 3, 8)
    f, lead = flatten_x(x)
    assert f.shape == (6, 8) and lead == [2, 3]
    assert unflatten_x(f[:, :4], lead).shape == (2, 3, 4)
