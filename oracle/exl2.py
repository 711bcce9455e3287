"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's exl2 mixed bit-width dequantisation
(reconstruct_exl2_kernel, bitorch_engine/layers/qlinear/nbit/cuda/mbwq_linear_cuda_kernel.cu:92-308; bit-stream
primitives exl2/quant/qdq_{2,3,4,5,6,8}.cuh #else branches, qdq_util.cuh:56-64) and of the q4 GPTQ-style dequantisation
(reconstruct_q4_gptq_kernel :314-411).  The reference's implementation of this path is CUDA-only, so it is pinned on the
GPU box against oracle/_ref/q_linear_cuda (tests/test_gpu_mbwq.py); without that build the header status is
"parity unpinned" for exl2."""
import numpy as np

WIDTHS = (8, 6, 5, 4, 3, 2)


def rows_from_q_groups(q_groups, height):
    """cumulative section ends + bit mask (mbwq_linear_trans_qweight_cuda :565-600)."""
    qg = np.asarray(q_groups).astype(np.int64).reshape(-1, 2)
    counts = {b: 0 for b in WIDTHS}
    mask, row = 0, 0
    for i in range(len(qg)):
        b = int(qg[i, 0])
        mask |= 1 << (b - 1)
        rows = (int(qg[i + 1, 1]) - int(qg[i, 1])) * 32 // b if i < len(qg) - 1 else height - row
        counts[b] += rows
        row += rows
    out, acc = [], 0
    for b in WIDTHS:
        acc += counts[b]
        out.append(acc)
    return out + [mask]


def group_map(q_groups, num_qrows):
    """(group, rows left) per weight row (nbit/cuda/utils.py:150-186)."""
    qg = np.asarray(q_groups).astype(np.int64).reshape(-1, 2)
    gm = []
    for i in range(len(qg)):
        b = int(qg[i, 0])
        qrows = (int(qg[i + 1, 1]) if i < len(qg) - 1 else num_qrows) - int(qg[i, 1])
        rows = qrows * 32 // b
        for j in range(rows):
            gm += [i, rows - j]
    return np.array(gm, dtype=np.int16)


def dequant(qweight, scales, zeros, q_perm, q_group_map, rows):
    """float32 values of the fp16 weight [K,N]: w[perm[k], n] = half(fma(q, s, -z))."""
    qw = np.asarray(qweight).astype(np.int64) & 0xFFFFFFFF
    N = qw.shape[1]
    K = len(q_group_map) // 2
    out = np.zeros((K, N), dtype=np.float32)
    prev, prow = 0, 0
    for b, end in zip(WIDTHS, rows[:6]):
        for k in range(prev, end):
            bitpos = (k - prev) * b
            w0, sh = prow + bitpos // 32, bitpos % 32
            v = qw[w0] >> sh
            if sh + b > 32:
                v = v | (qw[w0 + 1] << (32 - sh))
            v = (v & ((1 << b) - 1)).astype(np.float64)
            g = int(q_group_map[2 * k])
            s, z = np.asarray(scales[g], dtype=np.float64), np.asarray(zeros[g], dtype=np.float64)
            row = int(q_perm[k]) & 0xFFFF if q_perm is not None else k
            out[row] = (v * s - z).astype(np.float16).astype(np.float32)
        prow += (end - prev) * b // 32
        prev = end
    return out
