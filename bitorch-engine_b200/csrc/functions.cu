// functions.cu -- the wire-format conversions next to the low-bit Linear path: the ops of the
// reference's `functions_cuda` extension.  All of them are one-pass HBM streams: 128-bit loads and stores, a grid of a few
// CTAs per SM walking the array with a grid stride, no temporaries, nothing allocated.
//
// Reference functions replaced (bitorch_engine/functions/cuda/functions_cuda_kernel.cu):
//   b200bit_q4_pack          q4_pack  -> q4_bit_packing_kernel            (:136-159, host flow :431-466)
//   b200bit_q4_unpack        q4_unpack -> q4_bit_unpacking_kernel         (:162-182, :477-505)
//   b200bit_q4_unpack_scale  q4_unpack_and_scaling -> q4_bit_unpacking_scaling_kernel (:185-209, :518-549)
//   b200bit_sign_pack_u8     tensor_pack_to_uint8 -> _to_uint8_array<T>   (:74-119, :283-336)
//   b200bit_sign_unpack_u8   uint8_to_unpacked_tensor -> unpack_uint8_to_float (:123-133, :365-402)
// Formats (fixed by the reference): a q4 byte holds two codes, the FIRST in the HIGH nibble (:146-153); a sign byte holds
// eight signs LSB first, bit = (v >= 0) (:74-83; NaN packs as 0, -0.0 as 1, as the reference's comparisons do).
// fp32toint4 is not provided: the reference kernel pair reads shared memory it never wrote (256 threads launched over a
// 1024-wide reduction, :23-52, :252-253) and writes half of its output buffer, so it has no defined result to match.
#include "common.cuh"

namespace b200bit {

constexpr int FN_THREADS = 256;

static int fn_grid(size_t work_items) {
    const size_t want = (work_items + FN_THREADS - 1) / FN_THREADS;
    const size_t cap = size_t(sm_count()) * 8;            // a few resident CTAs per SM, grid-stride beyond that
    return int(want < cap ? (want ? want : 1) : cap);
}

__device__ __forceinline__ uint4 ldg_v4(const void* p) { return *reinterpret_cast<const uint4*>(p); }

// ---- q4 pack: 16 int32 codes (64 B) -> 8 bytes per item ----
__global__ void __launch_bounds__(FN_THREADS) q4_pack_kernel(const int32_t* __restrict__ in, uint8_t* __restrict__ out,
                                                             size_t n_bytes) {
    const size_t items = n_bytes / 8;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t it = size_t(blockIdx.x) * blockDim.x + threadIdx.x; it < items; it += stride) {
        uint32_t w[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint4 a = ldg_v4(in + it * 16 + h * 8), b = ldg_v4(in + it * 16 + h * 8 + 4);
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint32_t r = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) r |= (((v[2 * j] & 15u) << 4) | (v[2 * j + 1] & 15u)) << (8 * j);
            w[h] = r;
        }
        *reinterpret_cast<uint2*>(out + it * 8) = make_uint2(w[0], w[1]);
    }
    // tail (n_bytes % 8 bytes), one byte per thread of the first CTA
    if (blockIdx.x == 0) {
        const size_t done = items * 8;
        for (size_t b = done + threadIdx.x; b < n_bytes; b += blockDim.x)
            out[b] = uint8_t(((uint32_t(in[2 * b]) & 15u) << 4) | (uint32_t(in[2 * b + 1]) & 15u));
    }
}

// ---- q4 unpack: 8 bytes -> 16 codes.  SCALE: signed (-8..7) x scale -> f32 (:198-206), else unsigned -> int32 (:170-171) ----
template <bool SCALE>
__global__ void __launch_bounds__(FN_THREADS) q4_unpack_kernel(const uint8_t* __restrict__ in, void* __restrict__ out_,
                                                               float scale, size_t n_bytes) {
    const size_t items = n_bytes / 8;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    auto emit = [&](size_t byte_index, uint32_t byte) {
        const int hi = int(byte >> 4), lo = int(byte & 15u);
        if constexpr (SCALE) {
            float* o = reinterpret_cast<float*>(out_) + 2 * byte_index;
            o[0] = float(hi > 7 ? hi - 16 : hi) * scale;
            o[1] = float(lo > 7 ? lo - 16 : lo) * scale;
        } else {
            int32_t* o = reinterpret_cast<int32_t*>(out_) + 2 * byte_index;
            o[0] = hi; o[1] = lo;
        }
    };
    // one (thread, step) = two packed bytes -> four codes = ONE 16-byte store, so that a warp's store instruction covers 512
    // contiguous bytes (a thread that unpacks a whole 8-byte item writes 64 bytes of its own and every store instruction of
    // the warp touches 32 half-filled sectors); four steps in flight per thread
    const size_t pairs = items * 4;
    const uint16_t* in2 = reinterpret_cast<const uint16_t*>(in);
    for (size_t base = (size_t(blockIdx.x) * blockDim.x) * 4; base < pairs; base += stride * 4) {
        uint32_t h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t it = base + size_t(u) * blockDim.x + threadIdx.x;
            h[u] = it < pairs ? uint32_t(in2[it]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t it = base + size_t(u) * blockDim.x + threadIdx.x;
            if (it >= pairs) continue;
            const uint32_t b0 = h[u] & 0xffu, b1 = h[u] >> 8;
            const uint32_t v[4] = {b0 >> 4, b0 & 15u, b1 >> 4, b1 & 15u};
            if constexpr (SCALE) {
                float f[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) f[j] = float(int(v[j]) > 7 ? int(v[j]) - 16 : int(v[j])) * scale;
                reinterpret_cast<float4*>(out_)[it] = make_float4(f[0], f[1], f[2], f[3]);
            } else {
                reinterpret_cast<uint4*>(out_)[it] = make_uint4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    if (blockIdx.x == 0)
        for (size_t b = items * 8 + threadIdx.x; b < n_bytes; b += blockDim.x) emit(b, in[b]);
}

// ---- sign pack: eight values -> one byte, LSB first ----
template <int DT> __device__ __forceinline__ bool fn_nonneg(const void* p, size_t i);
template <> __device__ __forceinline__ bool fn_nonneg<B200BIT_F32>(const void* p, size_t i) { return reinterpret_cast<const float*>(p)[i] >= 0.f; }
template <> __device__ __forceinline__ bool fn_nonneg<B200BIT_F16>(const void* p, size_t i) { return __hge(reinterpret_cast<const __half*>(p)[i], __float2half(0.f)); }
template <> __device__ __forceinline__ bool fn_nonneg<B200BIT_BF16>(const void* p, size_t i) { return reinterpret_cast<const __nv_bfloat16*>(p)[i] >= __float2bfloat16(0.f); }
template <> __device__ __forceinline__ bool fn_nonneg<B200BIT_I8>(const void* p, size_t i) { return reinterpret_cast<const int8_t*>(p)[i] >= 0; }

// one thread packs 16 values (two output bytes) from 128-bit loads where the element size allows
template <int DT>
__global__ void __launch_bounds__(FN_THREADS) sign_pack_kernel(const void* __restrict__ in, uint8_t* __restrict__ out,
                                                               size_t n_bytes) {
    const size_t items = n_bytes / 2;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t it = size_t(blockIdx.x) * blockDim.x + threadIdx.x; it < items; it += stride) {
        uint32_t bits = 0;
        if constexpr (DT == B200BIT_F32) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = reinterpret_cast<const float4*>(in)[it * 4 + q];
                bits |= (uint32_t(v.x >= 0.f) | (uint32_t(v.y >= 0.f) << 1) | (uint32_t(v.z >= 0.f) << 2) | (uint32_t(v.w >= 0.f) << 3)) << (4 * q);
            }
        } else if constexpr (DT == B200BIT_I8) {
            const uint4 v = reinterpret_cast<const uint4*>(in)[it];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) bits |= (((w[j >> 2] >> (8 * (j & 3) + 7)) & 1u) ^ 1u) << j;
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) bits |= uint32_t(fn_nonneg<DT>(in, it * 16 + j)) << j;
        }
        *reinterpret_cast<uint16_t*>(out + it * 2) = uint16_t(bits);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (n_bytes & 1)) {
        const size_t b = n_bytes - 1;
        uint32_t bits = 0;
        for (int j = 0; j < 8; ++j) bits |= uint32_t(fn_nonneg<DT>(in, b * 8 + j)) << j;
        out[b] = uint8_t(bits);
    }
}

// ---- sign unpack: byte -> eight +-scale floats; scale index = byte index / packed_dim (:128) ----
__global__ void __launch_bounds__(FN_THREADS) sign_unpack_kernel(const uint8_t* __restrict__ in, const float* __restrict__ scale,
                                                                 float* __restrict__ out, size_t n_bytes, size_t packed_dim) {
    // one (thread, step) = one nibble -> four floats = one 16-byte store (a warp's store instruction covers 512 contiguous bytes)
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    const size_t halves = n_bytes * 2;
    for (size_t it = size_t(blockIdx.x) * blockDim.x + threadIdx.x; it < halves; it += stride) {
        const size_t b = it >> 1;
        const uint32_t w = uint32_t(in[b]) >> ((it & 1) * 4);
        const float sc = scale[b / packed_dim];
        reinterpret_cast<float4*>(out)[it] = make_float4((w & 1u) ? sc : -sc, (w & 2u) ? sc : -sc, (w & 4u) ? sc : -sc, (w & 8u) ? sc : -sc);
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace b200bit

using namespace b200bit;

extern "C" {

int b200bit_q4_pack(const int32_t* in, int8_t* out, size_t n_bytes, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(in && out, B200BIT_ERR_ARG, "q4_pack: null pointer argument");
    B200_REQUIRE(aligned16(in) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0, B200BIT_ERR_ARG, "q4_pack: in must be 16-byte, out 8-byte aligned");
    if (n_bytes == 0) return B200BIT_OK;
    q4_pack_kernel<<<fn_grid(n_bytes / 8 + 1), FN_THREADS, 0, st>>>(in, reinterpret_cast<uint8_t*>(out), n_bytes);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_q4_unpack(const int8_t* in, int32_t* out, size_t n_bytes, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(in && out, B200BIT_ERR_ARG, "q4_unpack: null pointer argument");
    B200_REQUIRE(aligned16(out) && (reinterpret_cast<uintptr_t>(in) & 7u) == 0, B200BIT_ERR_ARG, "q4_unpack: in must be 8-byte, out 16-byte aligned");
    if (n_bytes == 0) return B200BIT_OK;
    q4_unpack_kernel<false><<<fn_grid(n_bytes / 8 + 1), FN_THREADS, 0, st>>>(reinterpret_cast<const uint8_t*>(in), out, 0.f, n_bytes);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_q4_unpack_scale(const int8_t* in, float scale, float* out, size_t n_bytes, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(in && out, B200BIT_ERR_ARG, "q4_unpack_scale: null pointer argument");
    B200_REQUIRE(aligned16(out) && (reinterpret_cast<uintptr_t>(in) & 7u) == 0, B200BIT_ERR_ARG, "q4_unpack_scale: in must be 8-byte, out 16-byte aligned");
    if (n_bytes == 0) return B200BIT_OK;
    q4_unpack_kernel<true><<<fn_grid(n_bytes / 8 + 1), FN_THREADS, 0, st>>>(reinterpret_cast<const uint8_t*>(in), out, scale, n_bytes);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_sign_pack_u8(const void* in, int in_dtype, uint8_t* out, size_t n_bytes, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(in && out, B200BIT_ERR_ARG, "sign_pack_u8: null pointer argument");
    B200_REQUIRE(aligned16(in) && (reinterpret_cast<uintptr_t>(out) & 1u) == 0, B200BIT_ERR_ARG, "sign_pack_u8: in must be 16-byte, out 2-byte aligned");
    if (n_bytes == 0) return B200BIT_OK;
    const int grid = fn_grid(n_bytes / 2 + 1);
    switch (in_dtype) {
        case B200BIT_F32: sign_pack_kernel<B200BIT_F32><<<grid, FN_THREADS, 0, st>>>(in, out, n_bytes); break;
        case B200BIT_F16: sign_pack_kernel<B200BIT_F16><<<grid, FN_THREADS, 0, st>>>(in, out, n_bytes); break;
        case B200BIT_BF16: sign_pack_kernel<B200BIT_BF16><<<grid, FN_THREADS, 0, st>>>(in, out, n_bytes); break;
        case B200BIT_I8: sign_pack_kernel<B200BIT_I8><<<grid, FN_THREADS, 0, st>>>(in, out, n_bytes); break;
        default: return set_error(B200BIT_ERR_UNSUPPORTED, "sign_pack_u8: tensor type not supported (dtype code %d)", in_dtype);
    }
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_sign_unpack_u8(const uint8_t* in, const float* scale, float* out, size_t n_bytes, size_t packed_dim, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(in && scale && out, B200BIT_ERR_ARG, "sign_unpack_u8: null pointer argument");
    B200_REQUIRE(packed_dim > 0, B200BIT_ERR_SHAPE, "sign_unpack_u8: packed_dim must be positive");
    B200_REQUIRE(aligned16(out), B200BIT_ERR_ARG, "sign_unpack_u8: out must be 16-byte aligned");
    if (n_bytes == 0) return B200BIT_OK;
    sign_unpack_kernel<<<fn_grid(n_bytes * 2), FN_THREADS, 0, st>>>(in, scale, out, n_bytes, packed_dim);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

}  // extern "C"
