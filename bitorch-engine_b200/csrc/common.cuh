// common.cuh -- shared host/device helpers for libb200bit (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/b200bit.h"

namespace b200bit {

// ---------------------------------------------------------------------------------------------------------------
// error reporting (thread-local message, negative return codes; never exit())
// ---------------------------------------------------------------------------------------------------------------
char* err_buf();
int set_error(int code, const char* fmt, ...);

#define B200_CUDA_OK(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::b200bit::set_error(B200BIT_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define B200_REQUIRE(cond, code, ...)                              \
    do {                                                           \
        if (!(cond)) return ::b200bit::set_error(code, __VA_ARGS__); \
    } while (0)

int sm_count();  // cached per current device

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
// 128-bit streaming load: weights are read exactly once per call -> do not allocate in L1.
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_nc_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// Blackwell mixed-precision FMA (SASS FHFMA / FHFMA.BF16): d = a(16-bit) * b(16-bit) + c(f32), product exact.
// HI selects the upper 16 bits of the packed register.
template <bool BF16, bool A_HI, bool B_HI>
__device__ __forceinline__ float fhfma(uint32_t a2, uint32_t b2, float c) {
    const unsigned short a = (unsigned short)(A_HI ? (a2 >> 16) : (a2 & 0xffffu));
    const unsigned short b = (unsigned short)(B_HI ? (b2 >> 16) : (b2 & 0xffffu));
    float d;
    if constexpr (BF16)
        asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    else
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    return d;
}

template <bool BF16>
__device__ __forceinline__ float cvt16_lo(uint32_t v) {
    if constexpr (BF16) return __uint_as_float(v << 16);
    else return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu)));
}
template <bool BF16>
__device__ __forceinline__ float cvt16_hi(uint32_t v) {
    if constexpr (BF16) return __uint_as_float(v & 0xffff0000u);
    else return __half2float(__ushort_as_half((unsigned short)(v >> 16)));
}
template <bool BF16>
__device__ __forceinline__ unsigned short f32_to_16(float f) {
    if constexpr (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(f));
    else return __half_as_ushort(__float2half_rn(f));
}

// Programmatic dependent launch (no-ops when the kernel was not launched with the PDL attribute).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace b200bit
