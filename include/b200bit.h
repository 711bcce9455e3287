/*
 * b200bit.h -- C ABI of libb200bit.so: the B200 (sm_100a) implementation of bitorch-engine's low-bit Linear hot path.
 *
 * Drop-in boundary.  The reference exposes this path as pybind11 extension modules that take
 * torch::Tensor arguments (bitorch_engine/layers/qlinear/nbit/cuda/q_linear_cuda.cpp:357-369,
 * bitorch_engine/layers/qlinear/binary/cuda/binary_linear_cuda.cpp:120-122).  This header is the torch-free
 * equivalent: plain device pointers, sizes, an explicit stream.  The Python shims in
 * bitorch-engine_b200/extensions/ re-create the reference's python-visible functions on top of it.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; inputs are borrowed for the duration of the
 *     call only; nothing is allocated or freed by the library; outputs are fully overwritten (never accumulated into,
 *     so no memset is needed -- the reference's torch::zeros at mpq_linear_cuda_kernel.cu:618 disappears).
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; all work is enqueued on it, nothing synchronises,
 *     so every entry point is CUDA-graph capturable.
 *   - return value: 0 on success, negative B200BIT_ERR_* otherwise; b200bit_last_error() returns a thread-local
 *     message.  The library never calls exit() (the reference does: mpq_linear_cuda_kernel.cu:507-508,574-575).
 *   - dtype codes: B200BIT_F32 / F16 / BF16 name the activation/scale/output element type (the reference switches on
 *     x.dtype(), mpq_linear_cuda_kernel.cu:517-575).
 *   - workspace: caller-owned scratch, >= the matching *_workspace_bytes() and 16-byte aligned.  Its first
 *     B200BIT_WS_TICKET_BYTES bytes must be zero before the FIRST use; kernels leave them zero again
 *     (split-K tickets in [0, B200BIT_WS_ZERO_OFFSET), a read-only zero page after it).
 */
#ifndef B200BIT_H_
#define B200BIT_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200BIT_API __attribute__((visibility("default")))
#else
#define B200BIT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define B200BIT_VERSION 100

#define B200BIT_F32 0
#define B200BIT_F16 1
#define B200BIT_BF16 2

#define B200BIT_OK 0
#define B200BIT_ERR_ARG (-1)          /* null pointer / bad enum                                   */
#define B200BIT_ERR_SHAPE (-2)        /* divisibility / size constraint violated                   */
#define B200BIT_ERR_UNSUPPORTED (-3)  /* bit-width / dtype combination not implemented             */
#define B200BIT_ERR_WORKSPACE (-4)    /* workspace missing or too small                            */
#define B200BIT_ERR_CUDA (-5)         /* a CUDA runtime call failed (message has the CUDA string)  */

#define B200BIT_WS_TICKET_BYTES 16384
#define B200BIT_WS_ZERO_OFFSET 12288   /* [12288, 16384) of the head is a page the kernels only ever read as zeros */

/* flags for the `flags` argument of the forward entry points */
#define B200BIT_FLAG_PDL 1u           /* launch with programmatic dependent launch (graph/stream overlap) */

B200BIT_API int b200bit_version(void);
B200BIT_API const char* b200bit_last_error(void);
/* number of SMs / compute capability of the current device (host helper for the shims; returns <0 on error) */
B200BIT_API int b200bit_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------------------
 * MPQ n-bit Linear forward:  y[M,N] = x[M,K] @ W[K,N],  W dequantised on the fly from
 *   qweight int32 [K*w_bit/32, N]  (value k of column n: word k/nb, bits (k%nb)*w_bit, LSB first; nb = 32/w_bit)
 *   sym  (asym=0): W = scales[g,n]*q - zeros[g,n]          zeros: same dtype as scales, [G,N]
 *   asym (asym=1): W = scales[g,n]*(q - (qz[g,n]+1))       zeros: packed int32 [G, N*w_bit/32]
 *   g = g_idx[k] (int32 [K]) or, when g_idx == NULL, k / (K/G).
 * Replaces q_linear_cuda.mpq_forward (q_linear_cuda.cpp:258-270 -> mpq_linear_cuda_kernel.cu:603-626, 482-577,
 * kernels :67-451).  w_bit in {1,2,4,8}; dtype F16/BF16 (fast paths) or F32.
 * Unlike the reference there is no M<=32 restriction and no K%256 / N%256 requirement; K % (32/w_bit) == 0 and
 * K % G == 0 are required.
 * ------------------------------------------------------------------------------------------------------------ */
B200BIT_API size_t b200bit_mpq_forward_workspace_bytes(int M, int K, int N, int w_bit);
B200BIT_API int b200bit_mpq_forward(const void* x, const int32_t* qweight, const void* scales, const void* zeros,
                        const int32_t* g_idx, void* y, int M, int K, int N, int G, int w_bit, int asym, int dtype,
                        void* workspace, size_t workspace_bytes, unsigned flags, void* stream);

/* Sweep hook for bench.py / tests (process-wide; 0 = built-in heuristic): lanes per packed-row segment (8, 16, 32),
 * warps per CTA (1..16), split-K factor of the decode GEMV.  No reference counterpart. */
B200BIT_API int b200bit_set_gemv_tuning(int lanes_per_row, int warps, int splitk);
/* Kernel-path override (process-wide): 0 auto, 1 CUDA-core FHFMA GEMV, 2 small-batch mma.sync kernel, 3 general
 * fallback; mma_for_m1 selects, in auto mode, whether M == 1 uses the tensor kernel (1) or the CUDA-core GEMV (0). */
B200BIT_API int b200bit_set_path(int path, int mma_for_m1);
/* Diagnostics: device buffer of [grid][16][8] uint64 that receives %globaltimer stamps from the TMA-streamed kernel
 * (per CTA, per warp: start, init done, dependency wait done, first tile landed, first run done, all runs done,
 * CTA barrier passed, end); NULL switches tracing off. */
B200BIT_API int b200bit_set_trace_buffer(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* B200BIT_H_ */
