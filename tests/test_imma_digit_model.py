"""CPU: the arithmetic model of the integer-tensor-pipe decode kernel (bitorch-engine_b200/csrc/mpq_imma.cuh), restated in
numpy: fixed-point expansion of a 128-value activation unit into balanced base-256 digits, the unmasked-byte operand
trick (raw byte x Xo digits + masked nibble x (Xe - Xo) digits), the m16n8k32 fragment mapping and the integer bounds the
kernel relies on.  The exact quantised model  s * sum(q x) - z * sum(x)  must come out to fp32 round-off.  (The kernel
itself is checked against the oracle on the GPU; this test pins the algebra a reviewer would otherwise have to trust.)"""
import numpy as np
import pytest


def _digits(x_unit, bf16=False):
    """stage_x of mpq_imma.cuh: per packed row (8 values) -> (Xo digit words, Xe - Xo digit words), unit weight."""
    v = x_unit.astype(np.float32).reshape(16, 8)
    amax = np.float32(np.abs(v).max())
    e = max(int(amax.view(np.uint32) >> 23), 67)
    scale = np.uint32((283 - e) << 23).view(np.float32)
    wt = np.uint32((e - 29) << 23).view(np.float32)
    Xo = np.rint(v[:, 1::2] * np.float32(scale * np.float32(0.0625))).astype(np.int64)
    Xe = np.rint(v[:, 0::2] * scale).astype(np.int64) - Xo
    assert np.abs(Xe).max() < 2 ** 31 and np.abs(Xo).max() < 2 ** 27

    def balanced(X):                       # P = (X + 0x00808080) ^ 0x00808080: four signed bytes, least significant first
        P = ((X + 0x00808080) & 0xFFFFFFFF) ^ 0x00808080
        d = np.stack([(P >> (8 * i)) & 0xFF for i in range(4)], axis=-1).astype(np.int64)
        d = np.where(d >= 128, d - 256, d)
        assert np.array_equal((d * (256 ** np.arange(4))).sum(-1), X)
        return d
    return balanced(Xo), balanced(Xe), wt              # [16 rows, 4 pairs, 4 digits]


@pytest.mark.parametrize("kind", ["normal", "one_huge", "tiny", "zeros", "mixed_sign_max"])
def test_unit_dot_product_is_exact_to_fp32_roundoff(kind):
    rng = np.random.default_rng(hash(kind) & 0xFFFF)
    x = rng.standard_normal(128).astype(np.float16)
    if kind == "one_huge":
        x[5] = np.float16(60000.0)
    elif kind == "tiny":
        x = (x.astype(np.float32) * 1e-6).astype(np.float16)          # fp16 subnormals
    elif kind == "zeros":
        x[:] = 0
    elif kind == "mixed_sign_max":
        x[:] = np.float16(65504.0) * np.where(rng.random(128) < 0.5, -1, 1)
    words = rng.integers(0, 2 ** 32, size=(16, 32), dtype=np.uint64)   # 16 packed rows x 32 columns of the strip window
    dXo, dXe, wt = _digits(x)
    # per column: integer sums per digit over the unit: raw byte (16*code_odd + code_even) x Xo + masked nibble x (Xe - Xo)
    raw = np.stack([(words >> (8 * i)) & 0xFF for i in range(4)], axis=-1).astype(np.int64)      # [16, 32, 4 byte slots]
    low = raw & 15
    acc = np.einsum("rcb,rbd->cd", raw, dXo) + np.einsum("rcb,rbd->cd", low, dXe)               # [32 columns, 4 digits]
    assert np.abs(acc).max() < 2 ** 23                                     # bound quoted in the kernel header
    lane0 = acc[:, 1] * 256 + acc[:, 0]                                    # digit pair of lanes c = 0 ...
    lane1 = acc[:, 3] * 256 + acc[:, 2]                                    # ... and c = 1 (weight 65536)
    assert max(np.abs(lane0).max(), np.abs(lane1).max()) < 2 ** 31
    got = (np.float32(lane0) * wt + np.float32(lane1) * np.float32(wt * np.float32(65536.0))).astype(np.float64)
    codes = np.stack([(words >> (4 * i)) & 15 for i in range(8)], axis=-1).astype(np.float64)    # [16, 32, 8]
    exact = np.einsum("rck,rk->c", codes, x.astype(np.float64).reshape(16, 8))
    scale_ref = max(np.abs(x.astype(np.float64)).max(), 1e-30) * 15 * 128
    assert np.abs(got - exact).max() <= 2.0 ** -22 * scale_ref, (kind, np.abs(got - exact).max(), scale_ref)


def test_fragment_mapping_of_the_kernel():
    """m16n8k32 (PTX ISA fragment layouts) with the kernel's lane -> (packed row, column, digit) assignment: the lanes with
    c = 0 / 1 end up with the digit pairs of columns 2g, 2g+1 (alpha) and 16+2g, 17+2g (beta)."""
    rng = np.random.default_rng(3)
    W = rng.integers(0, 2 ** 32, size=(16, 32), dtype=np.uint64)
    x = rng.standard_normal(128).astype(np.float16)
    dXo, dXe, wt = _digits(x)

    def imma(a_regs, b_regs, acc):
        A = np.zeros((16, 32), np.int64); B = np.zeros((32, 8), np.int64)
        for lane in range(32):
            g, c = lane >> 2, lane & 3
            for r in range(4):
                row, kb = g + (8 if r in (1, 3) else 0), 4 * c + (16 if r >= 2 else 0)
                A[row, kb:kb + 4] = [(a_regs[lane][r] >> (8 * i)) & 0xFF for i in range(4)]
            for r in range(2):
                kb = 4 * c + (16 if r else 0)
                byt = np.array([(b_regs[lane][r] >> (8 * i)) & 0xFF for i in range(4)])
                B[kb:kb + 4, g] = np.where(byt >= 128, byt - 256, byt)
        D = A @ B
        for lane in range(32):
            g, c = lane >> 2, lane & 3
            acc[lane] += [D[g, 2 * c], D[g, 2 * c + 1], D[g + 8, 2 * c], D[g + 8, 2 * c + 1]]
    accA = np.zeros((32, 4), np.int64); accB = np.zeros((32, 4), np.int64)
    pack = lambda d: int(sum((int(v) & 0xFF) << (8 * i) for i, v in enumerate(d)))
    for ks in range(4):
        b, j = ks >> 1, ks & 1
        aA, aB, bb = [], [], []
        for lane in range(32):
            g, c = lane >> 2, lane & 3
            row = 8 * b + 2 * c + j
            w = [int(W[row, 2 * g]), int(W[row, 2 * g + 1]), int(W[row, 16 + 2 * g]), int(W[row, 17 + 2 * g])]
            aA.append([w[0], w[1], w[0] & 0x0F0F0F0F, w[1] & 0x0F0F0F0F])
            aB.append([w[2], w[3], w[2] & 0x0F0F0F0F, w[3] & 0x0F0F0F0F])
            dgt = g & 3                                                   # lanes g >= 4: don't-care columns of B
            bb.append([pack(dXo[row, :, dgt]), pack(dXe[row, :, dgt])])
        imma(aA, bb, accA); imma(aB, bb, accB)
    codes = np.stack([(W >> (4 * i)) & 15 for i in range(8)], axis=-1).astype(np.float64)
    exact = np.einsum("rck,rk->c", codes, x.astype(np.float64).reshape(16, 8))
    for g in range(8):
        for q, col in enumerate((2 * g, 2 * g + 1, 16 + 2 * g, 17 + 2 * g)):
            acc = accA if q < 2 else accB
            tot = 0.0
            for c, lane_w in ((0, 1.0), (1, 65536.0)):
                d = acc[4 * g + c]
                tot += float(np.float32(d[(q & 1) * 2 + 1] * 256 + d[(q & 1) * 2]) * np.float32(wt * np.float32(lane_w)))
            assert abs(tot - exact[col]) <= 2.0 ** -20 * max(np.abs(exact).max(), 1.0)
