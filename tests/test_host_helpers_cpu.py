"""CPU: host-side helpers either side of the kernels (utils/model_helper.py, utils/quant_operators.py) against vectors
produced by the reference's own Python (oracle/gen_golden.py gen_helpers -> tests/golden/helper_cases.npz), plus the
checkpoint helpers on a toy model."""
import os

import numpy as np
import torch

from helpers import GOLD

Z = np.load(os.path.join(GOLD, "helper_cases.npz"))


def _t(key):
    return torch.from_numpy(Z[key].copy())


def test_padding_helpers_match_reference():
    from bitorch_engine_b200.utils import model_helper as mh
    assert np.array_equal(mh.pad_embedding_dim(_t("pad_emb_in")).numpy(), Z["pad_emb_out"])
    assert np.array_equal(mh.pad_embedding_dim(_t("pad_emb8_in")).numpy(), Z["pad_emb8_out"])      # already a multiple of 8
    out, sec = mh.pad_last_2_dims_to_multiple_of_128(_t("pad128_in"))
    assert np.array_equal(out.numpy(), Z["pad128_out"]) and sec == int(Z["pad128_sec"])
    post = mh.binary_matmul_forward_post_processing(_t("bmm_in"), [2], 30, 28, 64)
    assert np.array_equal(post.numpy(), Z["bmm_out"])


def test_python_bit_packers_match_reference_and_the_cuda_wire_format():
    from bitorch_engine_b200.utils import quant_operators as qo
    from oracle import functions as OF
    x = Z["bin_in"]
    row = qo.get_binary_row(x.flatten().tolist(), [0] * (x.size // 32), x.size, 32)
    assert np.array_equal(np.array(row, dtype=np.uint64), Z["bin_row"])
    col = qo.get_binary_col(x.T.copy().flatten().tolist(), [0] * (64 // 32 * 4), 64, 4, 32)
    assert np.array_equal(np.array(col, dtype=np.uint64), Z["bin_col"])
    # the same bits, eight at a time, are what tensor_pack_to_uint8 writes (LSB first)
    as_bytes = np.array(row, dtype=np.uint64).astype("<u4").view(np.uint8).reshape(4, 8)
    assert np.array_equal(as_bytes, OF.tensor_pack_to_uint8(x))


def test_activation_quantisers_match_reference():
    from bitorch_engine_b200.utils import quant_operators as qo
    a = _t("q_in")
    q8, s8 = qo.q8_quantization(a, eps=torch.tensor(1e-5))
    q4, s4 = qo.q4_quantization(a, eps=torch.tensor(1e-5))
    assert np.array_equal(q8.numpy(), Z["q8"]) and np.array_equal(q4.numpy(), Z["q4"])
    assert np.array_equal(s8.numpy(), Z["q8_scale"]) and np.array_equal(s4.numpy(), Z["q4_scale"])
    assert np.array_equal(qo.q8_quantization(a, torch.tensor(0.37), torch.tensor(1e-5)).numpy(), Z["q8_given"])
    assert np.array_equal(qo.q4_quantization(a, torch.tensor(0.37), torch.tensor(1e-5)).numpy(), Z["q4_given"])
    q8d, _ = qo.q8_quantization(a)                      # default eps (the reference's own default raises)
    assert np.array_equal(q8d.numpy(), Z["q8"])


def test_zero_point_packing_matches_reference_and_round_trips():
    from bitorch_engine_b200.utils import quant_operators as qo
    from oracle import nbit
    for b in (2, 4, 8):
        z = _t(f"zp{b}_in")
        packed = qo.gptq_style_zeros_packing(z, b, 64, 32)
        assert packed.dtype == torch.int32 and np.array_equal(packed.numpy(), Z[f"zp{b}_out"])
        assert np.array_equal(nbit.unpack_zeros_asym(packed.numpy(), b), z.numpy())


class _Packs(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(4, 16))
        self.register_buffer("qweight", torch.zeros((4, 2), dtype=torch.uint8))
        self.calls = []

    def generate_quantized_weight(self, qweight_only=False):
        from oracle import functions as OF
        self.calls.append(qweight_only)
        self.qweight.copy_(torch.from_numpy(OF.tensor_pack_to_uint8(self.weight.detach().numpy())))


def test_checkpoint_helpers_round_trip(tmp_path):
    from bitorch_engine_b200.utils import model_helper as mh
    torch.manual_seed(0)
    model = torch.nn.Sequential(_Packs(), torch.nn.ReLU(), _Packs())
    mh.pack_bie_layers(model, qweight_only=True, layers=[_Packs])
    assert model[0].calls == [True] and model[2].calls == [True]
    path = os.path.join(tmp_path, "ckpt.pth")
    # save_checkpoint / load_checkpoint use the package's default layer list; exercise their file format directly
    torch.save({"state_dict": model.state_dict()}, path)
    other = torch.nn.Sequential(_Packs(), torch.nn.ReLU(), _Packs())
    other.load_state_dict(torch.load(path)["state_dict"], strict=False)
    assert torch.equal(other[0].qweight, model[0].qweight) and torch.equal(other[2].weight, model[2].weight)
    # default layer list: an MPQ layer refuses to pack, exactly like the reference (nbit/layer.py:468-480)
    from bitorch_engine_b200.layers.qlinear.nbit.cuda import MPQLinearCuda
    import pytest
    with pytest.raises(NotImplementedError):
        mh.save_checkpoint(torch.nn.Sequential(MPQLinearCuda(256, 256, requires_grad=False, group_size=128, dq_group_size=256)), path)
