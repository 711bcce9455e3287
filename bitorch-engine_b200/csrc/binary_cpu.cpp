// binary_cpu.cpp -- host (CPU) twin of the reference's `binary_linear_cpp` extension: the 1-bit Linear for
// BinaryLinearCPP (bitorch_engine/layers/qlinear/binary/cpp/layer.py:18-125, binary_linear.cpp:43-518).  This is the
// one CPU layer SURVEY.md section 8 puts on the path (row a17); it shares nothing with the CUDA kernels but the C ABI.
//
//   y[m, n] = K - 2 * popcount(bits(x[m, :]) xor bits(w[n, :])),   bit = (v >= 0)
//   packed weight: byte (k / 8) * N + n, bit j (LSB first) = sign of w[n, 8 * (k / 8) + j]   (binary_linear.cpp:80-145)
//
// Written from scratch: the reference walks B byte-wise with one OpenMP loop over M (:249-295); here both operands
// become rows of 64-bit words (weights transposed once per call: K * N / 8 bytes), the (m, n) grid is blocked 4 x 4 so
// each loaded word is used four times, POPCNT / AVX-512 VPOPCNTDQ clones are picked at load time, and the rows of x are
// split over std::thread workers (no OpenMP runtime needed).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/b200bit.h"

namespace {

inline int words_of(int k) { return (k + 63) / 64; }

// floats -> sign bits, LSB first inside each byte (= inside each little-endian 64-bit word), zero padded
void pack_row(const float* v, int k, uint64_t* out) {
    const int W = words_of(k);
    for (int w = 0; w < W; ++w) {
        uint64_t bits = 0;
        const int lim = std::min(64, k - 64 * w);
        for (int j = 0; j < lim; ++j) bits |= uint64_t(v[64 * w + j] >= 0.0f) << j;
        out[w] = bits;
    }
}

__attribute__((target_clones("arch=sapphirerapids", "arch=icelake-server", "arch=znver4", "popcnt", "default")))
void gemm_rows(const uint64_t* xw, const uint64_t* wt, float* out, int m0, int m1, int n, int k, int W) {
    int m = m0;
    for (; m + 4 <= m1; m += 4) {
        const uint64_t* x0 = xw + size_t(m) * W;
        int j = 0;
        for (; j + 4 <= n; j += 4) {
            int acc[4][4] = {};
            const uint64_t* w0 = wt + size_t(j) * W;
            for (int w = 0; w < W; ++w) {
                const uint64_t a[4] = {x0[w], x0[W + w], x0[2 * W + w], x0[3 * W + w]};
                const uint64_t b[4] = {w0[w], w0[W + w], w0[2 * W + w], w0[3 * W + w]};
                for (int r = 0; r < 4; ++r)
                    for (int c = 0; c < 4; ++c) acc[r][c] += __builtin_popcountll(a[r] ^ b[c]);
            }
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) out[size_t(m + r) * n + j + c] = float(k - 2 * acc[r][c]);
        }
        for (; j < n; ++j)
            for (int r = 0; r < 4; ++r) {
                int acc = 0;
                for (int w = 0; w < W; ++w) acc += __builtin_popcountll(x0[size_t(r) * W + w] ^ wt[size_t(j) * W + w]);
                out[size_t(m + r) * n + j] = float(k - 2 * acc);
            }
    }
    for (; m < m1; ++m)
        for (int j = 0; j < n; ++j) {
            int acc = 0;
            for (int w = 0; w < W; ++w) acc += __builtin_popcountll(xw[size_t(m) * W + w] ^ wt[size_t(j) * W + w]);
            out[size_t(m) * n + j] = float(k - 2 * acc);
        }
}

template <typename F>
void parallel_rows(int rows, int threads, F&& fn) {
    int T = threads > 0 ? threads : int(std::thread::hardware_concurrency());
    T = std::max(1, std::min(T, (rows + 3) / 4));
    if (T == 1) { fn(0, rows); return; }
    std::vector<std::thread> pool;
    const int chunk = ((rows + T - 1) / T + 3) & ~3;
    for (int t = 0; t < T; ++t) {
        const int a = t * chunk, b = std::min(rows, a + chunk);
        if (a >= b) break;
        pool.emplace_back([=, &fn] { fn(a, b); });
    }
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

// weights float [n, k] -> packed uint8 [k * n / 8] in the reference's layout (binary_linear_cpp.w_pack)
int b200bit_cpu_binary_pack(const float* weights, uint8_t* out, int n, int k) {
    if (!weights || !out) return B200BIT_ERR_ARG;
    if (n <= 0 || k <= 0 || k % 8 != 0) return B200BIT_ERR_SHAPE;
    for (int kb = 0; kb < k / 8; ++kb)
        for (int j = 0; j < n; ++j) {
            const float* v = weights + size_t(j) * k + 8 * kb;
            uint8_t b = 0;
            for (int i = 0; i < 8; ++i) b |= uint8_t(v[i] >= 0.0f) << i;
            out[size_t(kb) * n + j] = b;
        }
    return B200BIT_OK;
}

// x float [m, k]; weights: packed uint8 [k * n / 8] (weights_packed = 1) or float [n, k]; out float [m, n]
// (binary_linear_cpp.forward).  threads <= 0: one worker per hardware thread.
int b200bit_cpu_binary_forward(const float* x, const void* weights, int weights_packed, float* out, int m, int n, int k,
                               int threads) {
    if (!x || !weights || !out) return B200BIT_ERR_ARG;
    if (m < 0 || n <= 0 || k <= 0 || k % 8 != 0) return B200BIT_ERR_SHAPE;
    if (m == 0) return B200BIT_OK;
    const int W = words_of(k);
    std::vector<uint64_t> wt(size_t(n) * W, 0), xw(size_t(m) * W);
    if (weights_packed) {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(weights);
        uint8_t* dst = reinterpret_cast<uint8_t*>(wt.data());          // little-endian: byte kb of row j
        for (int kb = 0; kb < k / 8; ++kb)
            for (int j = 0; j < n; ++j) dst[size_t(j) * W * 8 + kb] = p[size_t(kb) * n + j];
    } else {
        const float* w = reinterpret_cast<const float*>(weights);
        parallel_rows(n, threads, [&](int a, int b) { for (int j = a; j < b; ++j) pack_row(w + size_t(j) * k, k, wt.data() + size_t(j) * W); });
    }
    parallel_rows(m, threads, [&](int a, int b) {
        for (int i = a; i < b; ++i) pack_row(x + size_t(i) * k, k, xw.data() + size_t(i) * W);
        gemm_rows(xw.data(), wt.data(), out, a, b, n, k, W);
    });
    return B200BIT_OK;
}

}  // extern "C"
