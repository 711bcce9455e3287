// mpq_aux.cu -- the data-format kernels either side of the n-bit Linear: dequantise (unpack), quantise + bit-pack,
// and grad_input.  All three are HBM-bound byte/word streams: coalesced along N (the contiguous dimension of every
// tensor involved), one pass over the data, no temporaries.
//
// Reference functions replaced:
//   b200bit_mpq_dequant      : unpack_qweight, layer_type 1   (bitorch_engine/layers/qlinear/nbit/cuda/utils.py:5-69;
//                              ~6 torch elementwise kernels, int32 [K/nb, nb, N] + int8 [K,N] temporaries)
//   b200bit_mpq_pack_weight  : pack_fp_weight                 (utils.py:72-147; ~8 torch kernels, int32 [K,N] temporaries)
//   b200bit_mpq_grad_input   : mpq_grad_input -> back_quant_mm_kernel{,_asym}
//                              (q_linear_cuda.cpp:272-284, mpq_linear_cuda_kernel.cu:635-1049, 1079-1223)
// Rounding contract (bit-exactness with the reference's Python): torch evaluates every half / bfloat16 elementwise op
// in fp32 and rounds the result to the tensor dtype, per op.  The helpers below do exactly that.
#include "common.cuh"

namespace b200bit {

template <int DT> struct El;
template <> struct El<B200BIT_F32> {
    using T = float;
    __device__ static float ld(const void* p, size_t i) { return reinterpret_cast<const float*>(p)[i]; }
    __device__ static void st(void* p, size_t i, float v) { reinterpret_cast<float*>(p)[i] = v; }
    __device__ static float rnd(float v) { return v; }
};
template <> struct El<B200BIT_F16> {
    using T = __half;
    __device__ static float ld(const void* p, size_t i) { return __half2float(reinterpret_cast<const __half*>(p)[i]); }
    __device__ static void st(void* p, size_t i, float v) { reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); }
    __device__ static float rnd(float v) { return __half2float(__float2half_rn(v)); }
};
template <> struct El<B200BIT_BF16> {
    using T = __nv_bfloat16;
    __device__ static float ld(const void* p, size_t i) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]); }
    __device__ static void st(void* p, size_t i, float v) { reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); }
    __device__ static float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

// eight consecutive elements as floats: 128-bit accesses (one for the 16-bit types, two for fp32)
template <int DT>
__device__ __forceinline__ void load8(const void* p, size_t i, float (&v)[8]) {
    if constexpr (DT == B200BIT_F32) {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + i);
        const uint32_t r[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[2 * q] = cvt16_lo<DT == B200BIT_BF16>(r[q]);
            v[2 * q + 1] = cvt16_hi<DT == B200BIT_BF16>(r[q]);
        }
    }
}
template <int DT>
__device__ __forceinline__ void store8(void* p, size_t i, const float (&v)[8]) {
    if constexpr (DT == B200BIT_F32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            r[q] = uint32_t(f32_to_16<DT == B200BIT_BF16>(v[2 * q])) | (uint32_t(f32_to_16<DT == B200BIT_BF16>(v[2 * q + 1])) << 16);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p) + i) = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

__device__ __forceinline__ int group_of(const int32_t* g_idx, int k, int gs) { return g_idx ? g_idx[k] : k / gs; }

// ---------------------------------------------------------------------------------------------------------------
// dequantise: out[k, n] in dtype DT.   sym: rnd(rnd(q*s) - z)   asym: rnd(s * (q - (qz+1)))    (utils.py:43, :51)
// one thread per (packed row, column); consecutive threads -> consecutive columns
// ---------------------------------------------------------------------------------------------------------------
// fused = 1: sym value is rnd(fma(s, q, -z)) -- the rounding of the reference's CUDA dequant kernels
// (reconstruct_q4_gptq_kernel, mbwq_linear_cuda_kernel.cu:395: __hfma2(scales, dq, -zeros)); perm (int16 [K],
// nullable): row k of the packed matrix is written to row perm[k] of the output (:398-399, MBWQ q_perm).
template <int DT>
__global__ void __launch_bounds__(256) mpq_dequant_kernel(const uint32_t* __restrict__ qw, const void* __restrict__ scales,
                                                          const void* __restrict__ zeros, const int32_t* __restrict__ g_idx,
                                                          void* __restrict__ out, int K, int N, int G, int w_bit, int asym,
                                                          int fused, const uint16_t* __restrict__ perm) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (n >= N) return;
    const int nb = 32 / w_bit, gs = K / G;
    const uint32_t mask = (1u << w_bit) - 1u;
    const uint32_t w = qw[size_t(r) * N + n];
    int g_prev = -1;
    float s = 0.f, z = 0.f;
    for (int j = 0; j < nb; ++j) {
        const int k = r * nb + j;
        const int g = group_of(g_idx, k, gs);
        if (g != g_prev) {
            s = El<DT>::ld(scales, size_t(g) * N + n);
            if (asym) {
                const uint32_t zw = reinterpret_cast<const uint32_t*>(zeros)[size_t(g) * (N / nb) + n / nb];
                z = float(((zw >> ((n % nb) * w_bit)) & mask) + 1u);
            } else {
                z = El<DT>::ld(zeros, size_t(g) * N + n);
            }
            g_prev = g;
        }
        const float q = float((w >> (j * w_bit)) & mask);
        float v;
        if (asym) v = El<DT>::rnd(__fmul_rn(s, __fsub_rn(q, z)));
        else if (fused) v = El<DT>::rnd(fmaf(s, q, -z));
        else v = El<DT>::rnd(__fsub_rn(El<DT>::rnd(__fmul_rn(q, s)), z));
        const int ko = perm ? int(perm[k]) : k;
        El<DT>::st(out, size_t(ko) * N + n, v);
    }
}

// Vector flavour (N % 8 == 0): a thread owns eight consecutive columns of one packed row -- two 128-bit loads of packed
// words, 128-bit loads of the group's scales / zeros, one 128-bit store per output row (two for fp32).  Same arithmetic,
// same roundings as the scalar kernel above (which stays for odd N).
template <int DT>
__global__ void __launch_bounds__(128) mpq_dequant_vec_kernel(const uint32_t* __restrict__ qw, const void* __restrict__ scales,
                                                              const void* __restrict__ zeros, const int32_t* __restrict__ g_idx,
                                                              void* __restrict__ out, int K, int N, int G, int w_bit, int asym,
                                                              int fused, const uint16_t* __restrict__ perm) {
    const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int r = blockIdx.y;
    if (n0 >= N) return;
    const int nb = 32 / w_bit, gs = K / G;
    const uint32_t mask = (1u << w_bit) - 1u;
    const uint4 wa = ldg_stream_v4(qw + size_t(r) * N + n0), wb = ldg_stream_v4(qw + size_t(r) * N + n0 + 4);
    const uint32_t w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    int g_prev = -1;
    float s[8], z[8];
    for (int j = 0; j < nb; ++j) {
        const int k = r * nb + j;
        const int g = group_of(g_idx, k, gs);
        if (g != g_prev) {
            load8<DT>(scales, size_t(g) * N + n0, s);
            if (asym) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int n = n0 + c;
                    const uint32_t zw = reinterpret_cast<const uint32_t*>(zeros)[size_t(g) * (N / nb) + n / nb];
                    z[c] = float(((zw >> ((n % nb) * w_bit)) & mask) + 1u);
                }
            } else {
                load8<DT>(zeros, size_t(g) * N + n0, z);
            }
            g_prev = g;
        }
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float q = float((w[c] >> (j * w_bit)) & mask);
            if (asym) v[c] = __fmul_rn(s[c], __fsub_rn(q, z[c]));
            else if (fused) v[c] = fmaf(s[c], q, -z[c]);
            else v[c] = __fsub_rn(El<DT>::rnd(__fmul_rn(q, s[c])), z[c]);
        }
        const int ko = perm ? int(perm[k]) : k;
        store8<DT>(out, size_t(ko) * N + n0, v);           // the store rounds to DT (the last rnd of every formula)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// quantise + pack: codes = clamp(rint(t), 0, 2^b-1);  sym: t = rnd(rnd(w + z) / s);  asym: t = rnd(rnd(w / s) + zq)
// (utils.py:118, :128, :131; torch.round == rint, half-to-even).  zeros: sym dtype [G,N]; asym either packed int32
// [G, N/nb] (zeros_unpacked == 0) or already-unpacked integer zero points stored as DT [G, N] (zeros_unpacked == 1,
// the `unpacked_zeros` argument of pack_fp_weight).  perm (int16 [K], nullable) gathers rows: w[perm[k], n]
// (utils.py:124-126, MBWQ q_perm).
// ---------------------------------------------------------------------------------------------------------------
// WT = dtype of `weight` (the optimizer hands over an fp32 weight next to half scales: torch then promotes, i.e. every op
// of the expression is evaluated -- and rounded -- in fp32)
template <int DT, int WT>
__global__ void __launch_bounds__(256) mpq_pack_kernel(const void* __restrict__ weight, const void* __restrict__ scales,
                                                       const void* __restrict__ zeros, const int32_t* __restrict__ g_idx,
                                                       const int16_t* __restrict__ perm, uint32_t* __restrict__ out, int K,
                                                       int N, int G, int w_bit, int asym, int zeros_unpacked) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (n >= N) return;
    const int nb = 32 / w_bit, gs = K / G;
    const uint32_t mask = (1u << w_bit) - 1u;
    const float maxq = float(mask);
    uint32_t word = 0;
    int g_prev = -1;
    float s = 1.f, z = 0.f;
    for (int j = 0; j < nb; ++j) {
        const int k = r * nb + j;
        const int g = group_of(g_idx, k, gs);
        if (g != g_prev) {
            s = El<DT>::ld(scales, size_t(g) * N + n);
            if (asym && !zeros_unpacked) {
                const uint32_t zw = reinterpret_cast<const uint32_t*>(zeros)[size_t(g) * (N / nb) + n / nb];
                z = float(((zw >> ((n % nb) * w_bit)) & mask) + 1u);
            } else {
                z = El<DT>::ld(zeros, size_t(g) * N + n);
            }
            g_prev = g;
        }
        const int ks = perm ? int(uint16_t(perm[k])) : k;
        constexpr int CT = (WT == B200BIT_F32) ? B200BIT_F32 : DT;       // torch's type promotion
        const float wv = El<WT>::ld(weight, size_t(ks) * N + n);
        float t;
        if (asym) t = El<CT>::rnd(__fadd_rn(El<CT>::rnd(__fdiv_rn(wv, s)), z));
        else t = El<CT>::rnd(__fdiv_rn(El<CT>::rnd(__fadd_rn(wv, z)), s));
        float c = rintf(t);
        c = fminf(fmaxf(c, 0.f), maxq);
        if (!(c == c)) c = 0.f;   // NaN -> 0 (torch: undefined conversion; keep the word well-formed)
        word |= (uint32_t(c) & mask) << (j * w_bit);
    }
    out[size_t(r) * N + n] = word;
}

// Vector flavour of the pack kernel (N % 8 == 0): eight columns per thread, 128-bit loads of the weight rows and of
// the group parameters, two 128-bit stores of packed words.
template <int DT, int WT>
__global__ void __launch_bounds__(128) mpq_pack_vec_kernel(const void* __restrict__ weight, const void* __restrict__ scales,
                                                           const void* __restrict__ zeros, const int32_t* __restrict__ g_idx,
                                                           const int16_t* __restrict__ perm, uint32_t* __restrict__ out,
                                                           int K, int N, int G, int w_bit, int asym, int zeros_unpacked) {
    const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int r = blockIdx.y;
    if (n0 >= N) return;
    constexpr int CT = (WT == B200BIT_F32) ? B200BIT_F32 : DT;
    const int nb = 32 / w_bit, gs = K / G;
    const uint32_t mask = (1u << w_bit) - 1u;
    const float maxq = float(mask);
    uint32_t word[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    int g_prev = -1;
    float s[8], z[8];
    for (int j = 0; j < nb; ++j) {
        const int k = r * nb + j;
        const int g = group_of(g_idx, k, gs);
        if (g != g_prev) {
            load8<DT>(scales, size_t(g) * N + n0, s);
            if (asym && !zeros_unpacked) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int n = n0 + c;
                    const uint32_t zw = reinterpret_cast<const uint32_t*>(zeros)[size_t(g) * (N / nb) + n / nb];
                    z[c] = float(((zw >> ((n % nb) * w_bit)) & mask) + 1u);
                }
            } else {
                load8<DT>(zeros, size_t(g) * N + n0, z);
            }
            g_prev = g;
        }
        const int ks = perm ? int(uint16_t(perm[k])) : k;
        float wv[8];
        load8<WT>(weight, size_t(ks) * N + n0, wv);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float t;
            if (asym) t = El<CT>::rnd(__fadd_rn(El<CT>::rnd(__fdiv_rn(wv[c], s[c])), z[c]));
            else t = El<CT>::rnd(__fdiv_rn(El<CT>::rnd(__fadd_rn(wv[c], z[c])), s[c]));
            float cq = rintf(t);
            cq = fminf(fmaxf(cq, 0.f), maxq);
            if (!(cq == cq)) cq = 0.f;
            word[c] |= (uint32_t(cq) & mask) << (j * w_bit);
        }
    }
    *reinterpret_cast<uint4*>(out + size_t(r) * N + n0) = make_uint4(word[0], word[1], word[2], word[3]);
    *reinterpret_cast<uint4*>(out + size_t(r) * N + n0 + 4) = make_uint4(word[4], word[5], word[6], word[7]);
}

// ---------------------------------------------------------------------------------------------------------------
// grad_input: dx[m, k] = sum_n dy[m, n] * W[k, n],  W = s*q - z (sym) / s*(q - zq) (asym), fp32 accumulation.
// One warp per packed row (nb consecutive k); lanes stride over n (coalesced word / scale / dy loads); MB batch rows
// per pass; nb*MB partial sums per lane, reduced with shuffles; deterministic.
// ---------------------------------------------------------------------------------------------------------------
template <int DT, int BITS, int MB>
__global__ void __launch_bounds__(128) mpq_grad_input_kernel(const void* __restrict__ dy, const uint32_t* __restrict__ qw,
                                                             const void* __restrict__ scales,
                                                             const void* __restrict__ zeros,
                                                             const int32_t* __restrict__ g_idx, void* __restrict__ dx,
                                                             int M, int K, int N, int G, int asym) {
    constexpr int NB = 32 / BITS;
    constexpr uint32_t mask = (1u << BITS) - 1u;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int m0 = blockIdx.y * MB;
    if (r >= K / NB) return;
    const int gs = K / G;
    float acc[MB][NB];
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[m][j] = 0.f;
    // groups of the NB k-values of this packed row (uniform across the warp)
    const bool one_group = (g_idx == nullptr) && ((r * NB) / gs == (r * NB + NB - 1) / gs);
    for (int n = lane; n < N; n += 32) {
        const uint32_t w = qw[size_t(r) * N + n];
        float dyv[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m) dyv[m] = (m0 + m < M) ? El<DT>::ld(dy, size_t(m0 + m) * N + n) : 0.f;
        float s = 0.f, z = 0.f;
        int g_prev = -1;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int k = r * NB + j;
            const int g = one_group ? (r * NB) / gs : group_of(g_idx, k, gs);
            if (g != g_prev) {
                s = El<DT>::ld(scales, size_t(g) * N + n);
                if (asym) {
                    const uint32_t zw = reinterpret_cast<const uint32_t*>(zeros)[size_t(g) * (N / NB) + n / NB];
                    z = s * float(((zw >> ((n % NB) * BITS)) & mask) + 1u);
                } else {
                    z = El<DT>::ld(zeros, size_t(g) * N + n);
                }
                g_prev = g;
            }
            const float wv = fmaf(s, float((w >> (j * BITS)) & mask), -z);
#pragma unroll
            for (int m = 0; m < MB; ++m) acc[m][j] = fmaf(dyv[m], wv, acc[m][j]);
        }
    }
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            float v = acc[m][j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0 && m0 + m < M) El<DT>::st(dx, size_t(m0 + m) * K + r * NB + j, v);
        }
}

template <int DT, int BITS>
static int launch_grad_input(const void* dy, const uint32_t* qw, const void* scales, const void* zeros,
                             const int32_t* g_idx, void* dx, int M, int K, int N, int G, int asym, cudaStream_t st) {
    constexpr int NB = 32 / BITS;
    constexpr int MB = (NB >= 16) ? 2 : 4;         // keep nb*MB accumulators <= 32
    const int rows = K / NB;
    dim3 grid((rows + 3) / 4, (M + MB - 1) / MB);
    mpq_grad_input_kernel<DT, BITS, MB><<<grid, 128, 0, st>>>(dy, qw, scales, zeros, g_idx, dx, M, K, N, G, asym);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}
template <int DT>
static int launch_grad_input_bits(int w_bit, const void* dy, const uint32_t* qw, const void* scales, const void* zeros,
                                  const int32_t* g_idx, void* dx, int M, int K, int N, int G, int asym, cudaStream_t st) {
    switch (w_bit) {
        case 1: return launch_grad_input<DT, 1>(dy, qw, scales, zeros, g_idx, dx, M, K, N, G, asym, st);
        case 2: return launch_grad_input<DT, 2>(dy, qw, scales, zeros, g_idx, dx, M, K, N, G, asym, st);
        case 4: return launch_grad_input<DT, 4>(dy, qw, scales, zeros, g_idx, dx, M, K, N, G, asym, st);
        case 8: return launch_grad_input<DT, 8>(dy, qw, scales, zeros, g_idx, dx, M, K, N, G, asym, st);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "grad_input: w_bit=%d", w_bit);
}

// ---------------------------------------------------------------------------------------------------------------
// exl2 (mixed bit-width) dequantise: rows of W are sorted by bit-width 8,6,5,4,3,2; inside a section of width b the
// codes of one column form an LSB-first bit stream over consecutive packed rows (32 codes per b words).
// w[perm[k], n] = rnd(fma(q, s[g,n], -z[g,n])), g = q_group_map[2k]     (reconstruct_exl2_kernel,
// mbwq_linear_cuda_kernel.cu:92-308; dequant primitives exl2/quant/qdq_*.cuh #else branches).  fp16 only.
// rows[6] = cumulative row counts (rows_8, rows_6, ..., rows_2) from mbwq_trans_qweight.
// ---------------------------------------------------------------------------------------------------------------
struct Exl2Rows { int end[6]; int prow[6]; };   // section end (weight rows) and first packed row, order 8,6,5,4,3,2

__global__ void __launch_bounds__(256) exl2_dequant_kernel(const uint32_t* __restrict__ qw, const __half* __restrict__ scales,
                                                           const __half* __restrict__ zeros,
                                                           const uint16_t* __restrict__ perm,
                                                           const uint16_t* __restrict__ group_map, __half* __restrict__ out,
                                                           int K, int N, const Exl2Rows rows) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (n >= N) return;
    const int widths[6] = {8, 6, 5, 4, 3, 2};
    int sec = 0;
    while (sec < 5 && k >= rows.end[sec]) ++sec;
    const int b = widths[sec];
    const int k_sec = sec == 0 ? 0 : rows.end[sec - 1];
    const int bitpos = (k - k_sec) * b;
    const size_t w0 = size_t(rows.prow[sec] + bitpos / 32) * N + n;
    const int sh = bitpos % 32;
    uint32_t v = qw[w0] >> sh;
    if (sh + b > 32) v |= qw[w0 + N] << (32 - sh);
    v &= (1u << b) - 1u;
    const int g = group_map[2 * k];
    const float s = __half2float(scales[size_t(g) * N + n]), z = __half2float(zeros[size_t(g) * N + n]);
    out[size_t(perm ? perm[k] : k) * N + n] = __float2half_rn(fmaf(float(v), s, -z));
}

static int check_common(const char* who, const void* a, const void* b, const void* c, const void* d, int K, int N, int G,
                        int w_bit, int asym, int dtype) {
    B200_REQUIRE(a && b && c && d, B200BIT_ERR_ARG, "%s: null pointer argument", who);
    B200_REQUIRE(dtype == B200BIT_F32 || dtype == B200BIT_F16 || dtype == B200BIT_BF16, B200BIT_ERR_ARG,
                 "%s: bad dtype code %d", who, dtype);
    B200_REQUIRE(w_bit == 1 || w_bit == 2 || w_bit == 4 || w_bit == 8, B200BIT_ERR_UNSUPPORTED,
                 "%s: w_bit=%d not supported (1, 2, 4, 8)", who, w_bit);
    B200_REQUIRE(K > 0 && N > 0 && G > 0, B200BIT_ERR_SHAPE, "%s: bad sizes K=%d N=%d G=%d", who, K, N, G);
    const int nb = 32 / w_bit;
    B200_REQUIRE(K % nb == 0, B200BIT_ERR_SHAPE, "%s: K=%d must be a multiple of %d", who, K, nb);
    B200_REQUIRE(!asym || N % nb == 0, B200BIT_ERR_SHAPE, "%s: asym needs N %% %d == 0 (N=%d)", who, nb, N);
    return B200BIT_OK;
}

}  // namespace b200bit

using namespace b200bit;

extern "C" {

int b200bit_mpq_dequant(const int32_t* qweight, const void* scales, const void* zeros, const int32_t* g_idx, void* out,
                        int K, int N, int G, int w_bit, int asym, int dtype, int fused, const int16_t* perm_,
                        void* stream_) {
    const uint16_t* perm = reinterpret_cast<const uint16_t*>(perm_);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_common("mpq_dequant", qweight, scales, zeros, out, K, N, G, w_bit, asym, dtype);
    if (rc != B200BIT_OK) return rc;
    B200_REQUIRE(g_idx || K % G == 0, B200BIT_ERR_SHAPE, "mpq_dequant: K=%d not divisible by G=%d", K, G);
    const int nb = 32 / w_bit;
    dim3 grid((N + 255) / 256, K / nb);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(qweight);
    const bool aligned = (reinterpret_cast<uintptr_t>(qweight) | reinterpret_cast<uintptr_t>(scales) | reinterpret_cast<uintptr_t>(zeros) |
                          reinterpret_cast<uintptr_t>(out)) % 16 == 0;
    if (N % 8 == 0 && aligned) {
        dim3 vgrid((N / 8 + 127) / 128, K / nb);
        if (dtype == B200BIT_F32) mpq_dequant_vec_kernel<B200BIT_F32><<<vgrid, 128, 0, st>>>(q, scales, zeros, g_idx, out, K, N, G, w_bit, asym, fused, perm);
        else if (dtype == B200BIT_F16) mpq_dequant_vec_kernel<B200BIT_F16><<<vgrid, 128, 0, st>>>(q, scales, zeros, g_idx, out, K, N, G, w_bit, asym, fused, perm);
        else mpq_dequant_vec_kernel<B200BIT_BF16><<<vgrid, 128, 0, st>>>(q, scales, zeros, g_idx, out, K, N, G, w_bit, asym, fused, perm);
        B200_CUDA_OK(cudaGetLastError());
        return B200BIT_OK;
    }
    if (dtype == B200BIT_F32) mpq_dequant_kernel<B200BIT_F32><<<grid, 256, 0, st>>>(q, scales, zeros, g_idx, out, K, N, G, w_bit, asym, fused, perm);
    else if (dtype == B200BIT_F16) mpq_dequant_kernel<B200BIT_F16><<<grid, 256, 0, st>>>(q, scales, zeros, g_idx, out, K, N, G, w_bit, asym, fused, perm);
    else mpq_dequant_kernel<B200BIT_BF16><<<grid, 256, 0, st>>>(q, scales, zeros, g_idx, out, K, N, G, w_bit, asym, fused, perm);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_exl2_dequant(const int32_t* qweight, const void* scales, const void* zeros, const int16_t* perm,
                         const int16_t* q_group_map, void* out, int K, int N, const int* rows6, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(qweight && scales && zeros && q_group_map && out && rows6, B200BIT_ERR_ARG,
                 "exl2_dequant: null pointer argument");
    B200_REQUIRE(K > 0 && N > 0 && K <= 65535, B200BIT_ERR_SHAPE, "exl2_dequant: bad sizes K=%d N=%d", K, N);
    const int widths[6] = {8, 6, 5, 4, 3, 2};
    Exl2Rows r{};
    int prev = 0, prow = 0;
    for (int i = 0; i < 6; ++i) {
        B200_REQUIRE(rows6[i] >= prev && rows6[i] <= K && (rows6[i] - prev) % 32 == 0, B200BIT_ERR_SHAPE,
                     "exl2_dequant: rows[%d]=%d is not a cumulative multiple of 32 within K=%d", i, rows6[i], K);
        r.end[i] = rows6[i];
        r.prow[i] = prow;
        prow += (rows6[i] - prev) * widths[i] / 32;
        prev = rows6[i];
    }
    B200_REQUIRE(prev == K, B200BIT_ERR_SHAPE, "exl2_dequant: rows cover %d of K=%d weight rows", prev, K);
    dim3 grid((N + 255) / 256, K);
    exl2_dequant_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(qweight),
                                              reinterpret_cast<const __half*>(scales), reinterpret_cast<const __half*>(zeros),
                                              reinterpret_cast<const uint16_t*>(perm),
                                              reinterpret_cast<const uint16_t*>(q_group_map),
                                              reinterpret_cast<__half*>(out), K, N, r);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_mpq_pack_weight(const void* weight, const void* scales, const void* zeros, const int32_t* g_idx,
                            const int16_t* perm, int32_t* qweight_out, int K, int N, int G, int w_bit, int asym,
                            int zeros_unpacked, int dtype, int weight_dtype, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(weight_dtype == dtype || weight_dtype == B200BIT_F32, B200BIT_ERR_UNSUPPORTED,
                 "mpq_pack_weight: weight dtype code %d with parameter dtype code %d (same dtype, or an fp32 weight)", weight_dtype, dtype);
    int rc = check_common("mpq_pack_weight", weight, scales, zeros, qweight_out, K, N, G, w_bit, asym, dtype);
    if (rc != B200BIT_OK) return rc;
    B200_REQUIRE(g_idx || K % G == 0, B200BIT_ERR_SHAPE, "mpq_pack_weight: K=%d not divisible by G=%d", K, G);
    const int nb = 32 / w_bit;
    dim3 grid((N + 255) / 256, K / nb);
    uint32_t* o = reinterpret_cast<uint32_t*>(qweight_out);
    const bool wf32 = weight_dtype == B200BIT_F32;
    const bool aligned = (reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(scales) | reinterpret_cast<uintptr_t>(zeros) |
                          reinterpret_cast<uintptr_t>(qweight_out)) % 16 == 0;
    if (N % 8 == 0 && aligned) {
        dim3 vgrid((N / 8 + 127) / 128, K / nb);
#define B200_PACK_VEC(DT_, WT_) mpq_pack_vec_kernel<DT_, WT_><<<vgrid, 128, 0, st>>>(weight, scales, zeros, g_idx, perm, o, K, N, G, w_bit, asym, zeros_unpacked)
        if (dtype == B200BIT_F32) B200_PACK_VEC(B200BIT_F32, B200BIT_F32);
        else if (dtype == B200BIT_F16 && wf32) B200_PACK_VEC(B200BIT_F16, B200BIT_F32);
        else if (dtype == B200BIT_F16) B200_PACK_VEC(B200BIT_F16, B200BIT_F16);
        else if (wf32) B200_PACK_VEC(B200BIT_BF16, B200BIT_F32);
        else B200_PACK_VEC(B200BIT_BF16, B200BIT_BF16);
#undef B200_PACK_VEC
        B200_CUDA_OK(cudaGetLastError());
        return B200BIT_OK;
    }
    if (dtype == B200BIT_F32) mpq_pack_kernel<B200BIT_F32, B200BIT_F32><<<grid, 256, 0, st>>>(weight, scales, zeros, g_idx, perm, o, K, N, G, w_bit, asym, zeros_unpacked);
    else if (dtype == B200BIT_F16 && wf32) mpq_pack_kernel<B200BIT_F16, B200BIT_F32><<<grid, 256, 0, st>>>(weight, scales, zeros, g_idx, perm, o, K, N, G, w_bit, asym, zeros_unpacked);
    else if (dtype == B200BIT_F16) mpq_pack_kernel<B200BIT_F16, B200BIT_F16><<<grid, 256, 0, st>>>(weight, scales, zeros, g_idx, perm, o, K, N, G, w_bit, asym, zeros_unpacked);
    else if (wf32) mpq_pack_kernel<B200BIT_BF16, B200BIT_F32><<<grid, 256, 0, st>>>(weight, scales, zeros, g_idx, perm, o, K, N, G, w_bit, asym, zeros_unpacked);
    else mpq_pack_kernel<B200BIT_BF16, B200BIT_BF16><<<grid, 256, 0, st>>>(weight, scales, zeros, g_idx, perm, o, K, N, G, w_bit, asym, zeros_unpacked);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_mpq_grad_input(const void* dy, const int32_t* qweight, const void* scales, const void* zeros,
                           const int32_t* g_idx, void* dx, int M, int K, int N, int G, int w_bit, int asym, int dtype,
                           void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_common("mpq_grad_input", dy, qweight, scales, zeros, K, N, G, w_bit, asym, dtype);
    if (rc != B200BIT_OK) return rc;
    B200_REQUIRE(dx != nullptr, B200BIT_ERR_ARG, "mpq_grad_input: null output");
    B200_REQUIRE(M >= 0, B200BIT_ERR_SHAPE, "mpq_grad_input: M=%d", M);
    B200_REQUIRE(g_idx || K % G == 0, B200BIT_ERR_SHAPE, "mpq_grad_input: K=%d not divisible by G=%d", K, G);
    if (M == 0) return B200BIT_OK;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(qweight);
    if (dtype == B200BIT_F32) return launch_grad_input_bits<B200BIT_F32>(w_bit, dy, q, scales, zeros, g_idx, dx, M, K, N, G, asym, st);
    if (dtype == B200BIT_F16) return launch_grad_input_bits<B200BIT_F16>(w_bit, dy, q, scales, zeros, g_idx, dx, M, K, N, G, asym, st);
    return launch_grad_input_bits<B200BIT_BF16>(w_bit, dy, q, scales, zeros, g_idx, dx, M, K, N, G, asym, st);
}

}  // extern "C"
