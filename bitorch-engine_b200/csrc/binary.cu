// binary.cu -- 1-bit (xnor-popcount) Linear for sm_100a:  y[m,n] = K - 2 * popc(bits(x[m,:]) xor bits(w[n,:])),
// sign bit = (v >= 0), integer-exact.
//
// Reference replaced (bitorch_engine/layers/qlinear/binary/cuda/binary_linear_cuda_kernel.cu):
//   packers    BMMA_toBit32Row_new / BMMA_toBit32Col_new (:59-150), ToBit32RowUd / ToBit32ColUd (:186-301),
//              uint8_to_uint32 / uint32_to_uint8 (:22-41)
//   GEMMs      BMMAS_new (wmma b1, :155-179) and BMM32_Arow_Brow_UD (:308-394)
//   host flows binary_linear_forward_BTC / _BSTC / _combined (:481-626), _get_binary_weight_cuda (:830-889)
// sm_100 has no 1-bit MMA any more (nvcc lowers wmma b1 to a bit-expansion + legacy int8 IMMA emulation, SURVEY.md
// section 2.3), so the product path is a shared-memory-tiled XOR/POPC kernel on a canonical bit layout:
//   canonical: one byte = 8 consecutive k of one row, bit (7 - k%8) = sign bit, rows padded with zero bits to whole
//   32-bit words.  This is exactly the byte content of both reference layouts (their 32-bit words are MSB-first and
//   stored big-endian), only the byte ORDER differs:
//     BTC  (k%128==0, n%8==0): byte(n, kb) at ((n/8)*(K/128) + kb/16)*128 + (n%8)*16 + kb%16     (:118, :22-41)
//     BSTC (k%32==0,  n%32==0): byte(n, kb) at 4*((kb/4)*N + n) + kb%4                            (:201, :22-41)
//   so packed reference weights are re-ordered, never re-interpreted, and w_pack emits them bit-exactly.
#include "common.cuh"


namespace b200bit {

template <int DT>
__device__ __forceinline__ bool ge0(const void* p, size_t i) {
    if constexpr (DT == B200BIT_F32) return reinterpret_cast<const float*>(p)[i] >= 0.f;
    else if constexpr (DT == B200BIT_F16) return __hge(reinterpret_cast<const __half*>(p)[i], __float2half(0.f));
    else if constexpr (DT == B200BIT_BF16) return __hge(reinterpret_cast<const __nv_bfloat16*>(p)[i], __float2bfloat16(0.f));
    else return reinterpret_cast<const int8_t*>(p)[i] >= 0;
}

// in [R, K] row-major -> out[r][kb]; thread <-> (r, kb)
template <int DT>
__global__ void __launch_bounds__(256) binary_pack_rows_kernel(const void* __restrict__ in, uint8_t* __restrict__ out,
                                                               int R, int K, int stride_bytes) {
    const size_t t = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (t >= size_t(R) * stride_bytes) return;
    const int r = int(t / stride_bytes), kb = int(t % stride_bytes);
    unsigned b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = kb * 8 + j;
        if (k < K && ge0<DT>(in, size_t(r) * K + k)) b |= 0x80u >> j;
    }
    out[t] = uint8_t(b);
}

// in [K, R] row-major (the "transposed" weight / mm operand) -> out[r][kb]; thread <-> (r, kb), coalesced over r
template <int DT>
__global__ void __launch_bounds__(256) binary_pack_cols_kernel(const void* __restrict__ in, uint8_t* __restrict__ out,
                                                               int R, int K, int stride_bytes) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int kb = blockIdx.y;
    if (r >= R) return;
    unsigned b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = kb * 8 + j;
        if (k < K && ge0<DT>(in, size_t(k) * R + r)) b |= 0x80u >> j;
    }
    out[size_t(r) * stride_bytes + kb] = uint8_t(b);
}

__device__ __forceinline__ size_t ref_byte_index(int n, int kb, int N, int K, int layout) {
    if (layout == 2) return (size_t(n / 8) * (K / 128) + kb / 16) * 128 + (n % 8) * 16 + kb % 16;   // BTC
    return 4 * (size_t(kb / 4) * N + n) + kb % 4;                                                   // BSTC
}

__global__ void __launch_bounds__(256) binary_relayout_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                              int N, int K, int layout, int to_reference,
                                                              int canon_stride) {
    const size_t t = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (t >= size_t(N) * (K / 8)) return;
    const int n = int(t / (K / 8)), kb = int(t % (K / 8));
    const size_t c = size_t(n) * canon_stride + kb, r = ref_byte_index(n, kb, N, K, layout);
    if (to_reference) out[r] = in[c];
    else out[c] = in[r];
}

template <int DT>
__device__ __forceinline__ void store_out(void* p, size_t i, int v) {
    if constexpr (DT == B200BIT_F32) reinterpret_cast<float*>(p)[i] = float(v);
    else if constexpr (DT == B200BIT_F16) reinterpret_cast<__half*>(p)[i] = __int2half_rn(v);
    else if constexpr (DT == B200BIT_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __int2bfloat16_rn(v);
    else reinterpret_cast<int32_t*>(p)[i] = v;
}

// 64 x 64 output tile per CTA, 16 x 16 threads x (4 x 4) outputs, K consumed in chunks of 32 words
constexpr int BG_T = 64, BG_KC = 32;
template <int DT>
__global__ void __launch_bounds__(256) binary_gemm_kernel(const uint32_t* __restrict__ xb, const uint32_t* __restrict__ wb,
                                                          void* __restrict__ out, int M, int N, int K, int kwords) {
    __shared__ uint32_t xs[BG_T][BG_KC + 1];
    __shared__ uint32_t ws[BG_T][BG_KC + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * BG_T, n0 = blockIdx.x * BG_T;
    int acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    for (int k0 = 0; k0 < kwords; k0 += BG_KC) {
        for (int e = threadIdx.x; e < BG_T * BG_KC; e += 256) {
            const int row = e / BG_KC, kw = e % BG_KC;
            const bool kin = k0 + kw < kwords;
            xs[row][kw] = (kin && m0 + row < M) ? xb[size_t(m0 + row) * kwords + k0 + kw] : 0u;
            ws[row][kw] = (kin && n0 + row < N) ? wb[size_t(n0 + row) * kwords + k0 + kw] : 0u;
        }
        __syncthreads();
#pragma unroll 8
        for (int kw = 0; kw < BG_KC; ++kw) {
            uint32_t xv[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = xs[ty * 4 + i][kw];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = ws[tx * 4 + j][kw];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += __popc(xv[i] ^ wv[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) store_out<DT>(out, size_t(m) * N + n, K - 2 * acc[i][j]);
        }
}

}  // namespace b200bit

using namespace b200bit;

extern "C" {

int b200bit_binary_pack(const void* in, int in_dtype, int rows, int K, int transposed_input, uint8_t* out,
                        int stride_bytes, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(in && out, B200BIT_ERR_ARG, "binary_pack: null pointer argument");
    B200_REQUIRE(rows > 0 && K > 0 && stride_bytes * 8 >= K && stride_bytes % 4 == 0, B200BIT_ERR_SHAPE,
                 "binary_pack: bad sizes rows=%d K=%d stride=%d", rows, K, stride_bytes);
    B200_REQUIRE(in_dtype >= 0 && in_dtype <= B200BIT_I8, B200BIT_ERR_ARG, "binary_pack: bad dtype code %d", in_dtype);
    if (!transposed_input) {
        dim3 grid(unsigned((size_t(rows) * stride_bytes + 255) / 256));
        switch (in_dtype) {
            case B200BIT_F32: binary_pack_rows_kernel<B200BIT_F32><<<grid, 256, 0, st>>>(in, out, rows, K, stride_bytes); break;
            case B200BIT_F16: binary_pack_rows_kernel<B200BIT_F16><<<grid, 256, 0, st>>>(in, out, rows, K, stride_bytes); break;
            case B200BIT_BF16: binary_pack_rows_kernel<B200BIT_BF16><<<grid, 256, 0, st>>>(in, out, rows, K, stride_bytes); break;
            default: binary_pack_rows_kernel<B200BIT_I8><<<grid, 256, 0, st>>>(in, out, rows, K, stride_bytes);
        }
    } else {
        dim3 grid((rows + 255) / 256, stride_bytes);
        switch (in_dtype) {
            case B200BIT_F32: binary_pack_cols_kernel<B200BIT_F32><<<grid, 256, 0, st>>>(in, out, rows, K, stride_bytes); break;
            case B200BIT_F16: binary_pack_cols_kernel<B200BIT_F16><<<grid, 256, 0, st>>>(in, out, rows, K, stride_bytes); break;
            case B200BIT_BF16: binary_pack_cols_kernel<B200BIT_BF16><<<grid, 256, 0, st>>>(in, out, rows, K, stride_bytes); break;
            default: binary_pack_cols_kernel<B200BIT_I8><<<grid, 256, 0, st>>>(in, out, rows, K, stride_bytes);
        }
    }
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_binary_relayout(const uint8_t* in, uint8_t* out, int N, int K, int layout, int to_reference,
                            int canon_stride_bytes, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(in && out, B200BIT_ERR_ARG, "binary_relayout: null pointer argument");
    B200_REQUIRE(layout == 1 || layout == 2, B200BIT_ERR_ARG, "binary_relayout: layout must be 1 (BSTC) or 2 (BTC)");
    if (layout == 2)
        B200_REQUIRE(K % 128 == 0 && N % 8 == 0, B200BIT_ERR_SHAPE,
                     "binary_relayout: BTC layout needs k %% 128 == 0 and n %% 8 == 0 (n=%d, k=%d)", N, K);
    else
        B200_REQUIRE(K % 32 == 0 && N % 32 == 0, B200BIT_ERR_SHAPE,
                     "binary_relayout: BSTC layout is only well defined for k %% 32 == 0 and n %% 32 == 0 (n=%d, k=%d)", N, K);
    dim3 grid(unsigned((size_t(N) * (K / 8) + 255) / 256));
    binary_relayout_kernel<<<grid, 256, 0, st>>>(in, out, N, K, layout, to_reference, canon_stride_bytes);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_binary_gemm(const uint8_t* x_bits, const uint8_t* w_bits, void* out, int M, int N, int K, int stride_bytes,
                        int out_dtype, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(x_bits && w_bits && out, B200BIT_ERR_ARG, "binary_gemm: null pointer argument");
    B200_REQUIRE(M >= 0 && N > 0 && K > 0 && stride_bytes % 4 == 0 && stride_bytes * 8 >= K, B200BIT_ERR_SHAPE,
                 "binary_gemm: bad sizes M=%d N=%d K=%d stride=%d", M, N, K, stride_bytes);
    if (M == 0) return B200BIT_OK;
    const uint32_t* xb = reinterpret_cast<const uint32_t*>(x_bits);
    const uint32_t* wb = reinterpret_cast<const uint32_t*>(w_bits);
    const int kwords = stride_bytes / 4;
    dim3 grid((N + BG_T - 1) / BG_T, (M + BG_T - 1) / BG_T);
    switch (out_dtype) {
        case B200BIT_F32: binary_gemm_kernel<B200BIT_F32><<<grid, 256, 0, st>>>(xb, wb, out, M, N, K, kwords); break;
        case B200BIT_F16: binary_gemm_kernel<B200BIT_F16><<<grid, 256, 0, st>>>(xb, wb, out, M, N, K, kwords); break;
        case B200BIT_BF16: binary_gemm_kernel<B200BIT_BF16><<<grid, 256, 0, st>>>(xb, wb, out, M, N, K, kwords); break;
        case B200BIT_I32: binary_gemm_kernel<B200BIT_I32><<<grid, 256, 0, st>>>(xb, wb, out, M, N, K, kwords); break;
        default: return set_error(B200BIT_ERR_ARG, "binary_gemm: bad dtype code %d", out_dtype);
    }
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

}  // extern "C"
