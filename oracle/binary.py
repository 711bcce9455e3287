"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's binary (1-bit) Linear.

  forward : y[m,n] = sum_k sgn(x[m,k]) * sgn(w[n,k]),  sgn(v) = +1 for v >= 0 else -1
            (binary_linear.cpp:43-54 bit = (v >= 0), :256-288 xnor/popcount, epilogue -(2C - 8K');
             binary_linear_cuda_kernel.cu:70, :174-176  K - 2*popc)
  packing : CPU extension  b_col[(k/8)*N + n], bit j = sign(w[n, 8*(k/8) + j])   (binary_linear.cpp:80-145;
            python twin utils/quant_operators.py:118-231 get_binary_col)
            CUDA BTC / BSTC byte streams (binary_linear_cuda_kernel.cu:118, :201, :22-41), see btc_index / bstc_index.
Pinned against oracle/_ref/binary_linear_cpp (compiled unmodified from /root/reference) in tests/test_oracle_binary.py
and, on the GPU box, against oracle/_ref/binary_linear_cuda in tests/test_gpu_binary.py."""
import numpy as np


def sgn(a):
    return np.where(np.asarray(a) >= 0, 1, -1).astype(np.int64)


def forward(x, w):
    return sgn(x) @ sgn(w).T


def canonical_bits(mat):
    """[rows, K] -> uint8 [rows, ceil(K/8)], bit (7 - k%8) of byte k//8 = (v >= 0)."""
    bits = (np.asarray(mat) >= 0).astype(np.uint8)
    rows, K = bits.shape
    pad = (-K) % 8
    if pad:
        bits = np.concatenate([bits, np.zeros((rows, pad), np.uint8)], axis=1)
    return np.packbits(bits, axis=1, bitorder="big")


def pack_cpp(w):
    """CPU extension layout: byte (k//8)*N + n, bit j (LSB first) = sign(w[n, 8*(k//8)+j])."""
    bits = (np.asarray(w) >= 0).astype(np.uint8)            # [N, K]
    N, K = bits.shape
    by = np.packbits(bits.reshape(N, K // 8, 8), axis=2, bitorder="little")[:, :, 0]   # [N, K/8]
    return by.T.reshape(-1).copy()


def btc_index(n, kb, N, K):
    return ((n // 8) * (K // 128) + kb // 16) * 128 + (n % 8) * 16 + kb % 16


def bstc_index(n, kb, N, K):
    return 4 * ((kb // 4) * N + n) + kb % 4


def pack_cuda(w, layout):
    """flat uint8 [K*N/8] as binary_linear_cuda.w_pack emits it; layout 2 = BTC, 1 = BSTC."""
    w = np.asarray(w)
    N, K = w.shape
    canon = canonical_bits(w)
    n_idx, kb_idx = np.meshgrid(np.arange(N), np.arange(K // 8), indexing="ij")
    idx = btc_index(n_idx, kb_idx, N, K) if layout == 2 else bstc_index(n_idx, kb_idx, N, K)
    out = np.zeros(N * K // 8, dtype=np.uint8)
    out[idx.reshape(-1)] = canon[:, : K // 8].reshape(-1)
    return out
