"""GPU parity of the `functions_cuda` conversions (C ABI b200bit_q4_* / b200bit_sign_*): bit-exact against the numpy
oracle, against the reference's own known-answer vector (tests/functions/test_quant_ops.py:124-144) and -- when
oracle/_ref/functions_cuda was built -- against the reference extension compiled unmodified for sm_100a."""
import numpy as np
import pytest
import torch

from oracle import functions as OF

pytestmark = pytest.mark.gpu


def _ext():
    from bitorch_engine_b200.extensions import functions_cuda
    return functions_cuda


def _ref():
    from oracle.build_ref import load_ref
    return load_ref("functions_cuda")


@pytest.mark.parametrize("shape", [(10, 10), (256, 4096), (3, 33, 14), (1, 2), (1000, 1001 * 2)])
@pytest.mark.parametrize("transpose", [False, True])
def test_q4_pack_unpack(shape, transpose):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randint(-2 ** 31, 2 ** 31 - 1, shape, dtype=torch.int32, generator=g).cuda()
    ext = _ext()
    packed = ext.q4_pack(x, transpose)
    want = OF.q4_pack(x.cpu().numpy())
    want_t = np.ascontiguousarray(np.swapaxes(want, -1, -2)) if transpose else want
    assert packed.dtype == torch.int8 and np.array_equal(packed.cpu().numpy(), want_t)
    plain = ext.q4_pack(x, False)
    un = ext.q4_unpack(plain, transpose)
    want_u = OF.q4_unpack(want)
    want_u = np.ascontiguousarray(np.swapaxes(want_u, -1, -2)) if transpose else want_u
    assert un.dtype == torch.int32 and np.array_equal(un.cpu().numpy(), want_u)
    sc = ext.q4_unpack_and_scaling(plain, 0.045, transpose)
    want_s = OF.q4_unpack_and_scaling(want, 0.045)
    want_s = np.ascontiguousarray(np.swapaxes(want_s, -1, -2)) if transpose else want_s
    assert sc.dtype == torch.float32 and np.array_equal(sc.cpu().numpy(), want_s)
    ref = _ref()
    if ref is not None and not transpose and len(shape) == 2:
        assert torch.equal(ref.q4_pack(x, False), plain)
        assert torch.equal(ref.q4_unpack(plain, False), ext.q4_unpack(plain, False))
        assert torch.equal(ref.q4_unpack_and_scaling(plain, 0.045, False), ext.q4_unpack_and_scaling(plain, 0.045, False))


def test_q4_round_trip_of_the_reference_test():
    # tests/functions/test_quant_ops.py:199-221
    from bitorch_engine_b200.functions.cuda import q4_pack_tensor, q4_unpack_tensor, q4_unpack_and_scaling_tensor
    for i in range(10):
        x = torch.randint(low=-8, high=8, size=(10, 10), dtype=torch.int).cuda()
        packed = q4_pack_tensor(x)
        assert packed.dtype == torch.int8 and packed.numel() * 2 == x.numel()
        assert ((q4_unpack_tensor(packed) & 15) ^ (x & 15) == 0).all()
        un = q4_unpack_and_scaling_tensor(packed, 0.045)
        assert torch.all(torch.isclose((un / 0.045).to(torch.int), x, rtol=1, atol=1))


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16, torch.int8])
@pytest.mark.parametrize("shape", [(100, 32), (4096, 4096), (3, 8), (7, 264)])
def test_sign_pack(dt, shape):
    g = torch.Generator().manual_seed(7)
    if dt == torch.int8:
        x = torch.randint(-128, 128, shape, dtype=torch.int8, generator=g)
    else:
        x = torch.randn(shape, generator=g)
        x[0, :4] = torch.tensor([0.0, -0.0, float("nan"), -1e-30])
        x = x.to(dt)
    ext = _ext()
    out = ext.tensor_pack_to_uint8(x.cuda())
    want = OF.tensor_pack_to_uint8(x.float().numpy() if dt != torch.int8 else x.numpy())
    assert out.dtype == torch.uint8 and tuple(out.shape) == (shape[0], shape[1] // 8)
    assert np.array_equal(out.cpu().numpy(), want)
    ref = _ref()
    if ref is not None:
        assert torch.equal(ref.tensor_pack_to_uint8(x.cuda()), out)


def test_sign_unpack_known_answer_and_random():
    ext = _ext()
    # tests/functions/test_quant_ops.py:124-144
    emd = torch.tensor([[[0, 16, 35, 255]]], dtype=torch.uint8).expand(2, 16, 4).contiguous().cuda()
    scale = torch.rand(2, 16, 1).cuda()
    exp_last = torch.tensor([[[-1] * 8 + [-1, -1, -1, +1, -1, -1, -1, -1][::-1] + [-1, -1, +1, -1, -1, -1, +1, +1][::-1] + [1] * 8]],
                            dtype=torch.float32).cuda()
    out = ext.uint8_to_unpacked_tensor(emd, scale)
    assert out.shape == (2, 16, 32)
    assert torch.equal(out, exp_last.expand(2, 16, 32) * scale)
    g = torch.Generator().manual_seed(3)
    emd = torch.randint(0, 256, (8, 16, 72), dtype=torch.uint8, generator=g).cuda()
    scale = torch.rand(8, 16, 1, generator=g).cuda()
    out = ext.uint8_to_unpacked_tensor(emd, scale)
    assert np.array_equal(out.cpu().numpy(), OF.uint8_to_unpacked_tensor(emd.cpu().numpy(), scale.cpu().numpy()))
    ref = _ref()
    if ref is not None:
        assert torch.equal(ref.uint8_to_unpacked_tensor(emd, scale), out)


def test_errors_and_missing_op():
    ext = _ext()
    with pytest.raises(RuntimeError):
        ext.q4_pack(torch.zeros((2, 2), dtype=torch.int32), False)           # CPU tensor
    with pytest.raises(ValueError):
        ext.q4_pack(torch.zeros((2, 3), dtype=torch.int32).cuda(), False)    # odd last dimension
    with pytest.raises(NotImplementedError):
        ext.fp32toint4(torch.zeros(64).cuda())
