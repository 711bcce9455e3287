"""GPU parity for the binary (1-bit) Linear: integer-exact against sign(x) @ sign(w).T (the arithmetic of bitorch's
QLinear(sign, sign) that the reference tests compare with, tests/layers/test_binary_linear.py:69-220), packed == unpacked
(:222-268), shapes incl. non-multiples, and -- when oracle/_ref holds the reference extensions compiled from
/root/reference -- bit-exact agreement with the reference's own binary_linear_cuda / binary_linear_cpp."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sign(t):
    return torch.where(t >= 0, 1.0, -1.0)


def _ref(x, w):
    return (_sign(x.double()) @ _sign(w.double()).t())


SHAPES = [(128, 512, 1000), (8, 128, 8), (32, 64, 32), (7, 40, 13), (100, 576, 64), (64, 1152, 128), (1, 4608, 512),
          (256, 2304, 256)]


@pytest.mark.parametrize("M,K,N", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("bmm", [1, 2, 3])
def test_forward_exact(M, K, N, dtype, bmm):
    from bitorch_engine_b200.extensions import binary_linear_cuda
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn((M, K), device="cuda", generator=g).to(dtype)
    w = torch.randn((N, K), device="cuda", generator=g).to(dtype)
    x[0, 0] = 0.0        # sign(0) = +1 everywhere in the reference (binary_linear.cpp:49, kernel.cu:70)
    out = binary_linear_cuda.forward(x, w, bmm, True)
    assert out.dtype == dtype and tuple(out.shape) == (M, N)
    exp = _ref(x, w)
    if dtype == torch.float32 or K <= 2048:
        assert torch.equal(out.double(), exp.to(out.device))
    else:   # half cannot hold every integer above 2048; the reference converts the int32 result the same way
        assert torch.equal(out, exp.to(dtype))
    wi = torch.where(w >= 0, 1, -1).to(torch.int8)
    assert torch.equal(binary_linear_cuda.forward(x, wi, bmm, True), out)


@pytest.mark.parametrize("N,K,bmm", [(1000, 512, 3), (8, 128, 2), (64, 128, 3), (32, 64, 1), (128, 96, 3), (512, 4608, 3)])
def test_packed_weights_equal_unpacked(N, K, bmm):
    from bitorch_engine_b200.extensions import binary_linear_cuda
    g = torch.Generator(device="cuda").manual_seed(N * 3 + K)
    w = torch.randn((N, K), device="cuda", generator=g)
    packed = binary_linear_cuda.w_pack(w, bmm, True)
    assert packed.dtype == torch.uint8 and packed.numel() == K * N // 8
    from oracle import binary as ob
    layout = 2 if (bmm == 2 or (bmm == 3 and K % 128 == 0 and N % 8 == 0)) else 1
    assert np.array_equal(packed.cpu().numpy(), ob.pack_cuda(w.cpu().numpy(), layout))
    for M in (8, 16):
        x = torch.randn((M, K), device="cuda", generator=g)
        assert torch.equal(binary_linear_cuda.forward(x, packed, bmm, True), binary_linear_cuda.forward(x, w, bmm, True))


def test_mm_and_layer():
    from bitorch_engine_b200.extensions import binary_linear_cuda
    from bitorch_engine_b200.layers.qlinear.binary.cuda import BinaryLinearCuda
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((24, 256), device="cuda", generator=g)
    y = torch.randn((256, 40), device="cuda", generator=g)
    assert torch.equal(binary_linear_cuda.mm(x, y, 3).double(), _sign(x.double()) @ _sign(y.double()))
    layer = BinaryLinearCuda(256, 64).cuda()
    layer.prepare_params()
    layer.train()
    xin = torch.randn((16, 256), device="cuda", generator=g, requires_grad=True)
    out = layer(xin)
    exp = _ref((xin + layer.bias_a).detach(), layer.weight.data.float()) * layer.scale_a.detach().double() * layer.scale_w.double()
    assert torch.allclose(out.detach().double(), exp.double(), rtol=1e-6, atol=1e-6)
    out.sum().backward()
    assert xin.grad is not None and layer.scale_a.grad is not None
    layer.eval()
    out_eval = layer(xin.detach())          # packed inference path
    assert torch.allclose(out_eval, out.detach(), rtol=1e-6, atol=1e-6)


def test_against_reference_extensions_when_available():
    """Bit-exact against the reference's own compiled extensions (oracle/_ref, built by oracle/build_ref.py)."""
    import importlib.util, os, sys
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
    br = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(br)
    ref_cuda = br.load_ref("binary_linear_cuda")
    ref_cpp = br.load_ref("binary_linear_cpp")
    if ref_cuda is None and ref_cpp is None:
        pytest.skip("oracle/_ref not built (needs /root/reference in the build container)")
    from bitorch_engine_b200.extensions import binary_linear_cuda
    g = torch.Generator(device="cuda").manual_seed(11)
    for (M, K, N) in [(128, 512, 1000), (8, 128, 8), (32, 64, 32), (64, 1152, 128)]:
        x = torch.randn((M, K), device="cuda", generator=g)
        w = torch.randn((N, K), device="cuda", generator=g)
        if ref_cpp is not None:
            exp = ref_cpp.forward(x.cpu(), w.cpu(), M, N, K)
            assert torch.equal(binary_linear_cuda.forward(x, w, 3, True).cpu(), exp)
        if ref_cuda is not None:
            for bmm in (1, 2, 3):
                if bmm == 2 and (M % 8 or K % 128 or N % 8):
                    continue
                assert torch.equal(binary_linear_cuda.forward(x, w, bmm, True), ref_cuda.forward(x, w, bmm, True)), (M, K, N, bmm)
                btc = bmm == 2 or (bmm == 3 and K % 128 == 0 and N % 8 == 0)
                if btc or (K % 32 == 0 and N % 32 == 0):   # the BSTC byte stream is only well defined for multiples of 32
                    assert torch.equal(binary_linear_cuda.w_pack(w, bmm, True), ref_cuda.w_pack(w, bmm, True)), (K, N, bmm)
