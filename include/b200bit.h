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
#define B200BIT_I8 3    /* binary path: int8 +-1 weights */
#define B200BIT_I32 4   /* binary path: raw integer output */

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
/* Caller's promise about x (decode GEMV, M == 1, with B200BIT_FLAG_PDL): x was completely written BEFORE the previous
 * b200bit_mpq_forward launch on this stream was enqueued and has not been written since -- the "sibling" case of a
 * decoder block (k_proj / v_proj after q_proj, up_proj after gate_proj: same hidden state).  The kernel then reads x
 * and computes BEFORE griddepcontrol.wait, i.e. concurrently with the kernels in front of it, and only its output
 * write waits for them.  The library honours the flag only when it knows that the launch in front on `stream` was one
 * of its own late-trigger decode kernels (or a chain of such siblings) whose outputs do not overlap x; otherwise the
 * call behaves exactly as without the flag.  No reference counterpart (the reference kernels run on the legacy
 * default stream, fully serialised). */
#define B200BIT_FLAG_INPUT_READY 2u

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

/* Batched forward on the tcgen05 tensor cores (csrc/mpq_tc.cu): y[M,N] = x[M,K] @ fp16(dequant(qweight)), ONE pass over
 * the packed weight for any M -- the weights are dequantised into tensor memory (TMEM A operand), the activations arrive
 * by TMA, fp32 accumulation.  The numerics are those of the reference's large-batch path (fp16 weight, dense GEMM:
 * mpq_layer.py:59-63), which this replaces without materialising the fp16 matrix.  2- / 4-bit, f16, contiguous groups of 32 * 2^i,
 * K % 64 == 0, N % 8 == 0, x 16-byte aligned.  b200bit_mpq_forward routes here for M > 16 (2-bit: M > 8) when the shape qualifies.
 * workspace (optional, as for b200bit_mpq_forward): enables split-K when the tiles alone would leave most SMs idle. */
B200BIT_API int b200bit_mpq_forward_tc_supported(int M, int K, int N, int G, int w_bit, int asym, int dtype,
                                                 size_t workspace_bytes);      /* 1: the call below runs this problem */
B200BIT_API int b200bit_mpq_forward_tc(const void* x, const int32_t* qweight, const void* scales, const void* zeros, void* y,
                                       int M, int K, int N, int G, int w_bit, int asym, int dtype, void* workspace,
                                       size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Decode chain: ONE persistent launch runs a list of batch-1 (M == 1) 4-bit Linear layers whose dataflow is given by
 * their pointers (node j reads what node i < j wrote when the ranges overlap).  Same arithmetic and bit-identical
 * results as b200bit_mpq_forward at M == 1; the weight stream is prefetched across layer boundaries and dependent layers
 * synchronise on the device -- through the data itself ({values, launch epoch} words) or through counters -- instead of
 * kernel boundaries (csrc/mpq_chain.cuh).  Use it for consecutive
 * Linear layers with nothing else between them: fused q/k/v and gate/up segments of a decoder block, or the whole
 * linear-layer chain of a token.  Replaces n x mpq_forward (q_linear_cuda.cpp:258-270 -> mpq_linear_cuda_kernel.cu:603-626).
 *   b200bit_mpq_chain_plan_bytes : size of the caller-owned device buffer that holds the plan (128-byte aligned)
 *   b200bit_mpq_chain_build      : host array of nodes -> plan image in plan_device (synchronous copy; call it once, outside
 *                                  graph capture) + info16 (16 host ints the caller keeps and hands to launch)
 *   b200bit_mpq_chain_launch     : enqueue the chain on `stream` (graph-capturable; launches of ONE plan must be
 *                                  stream-ordered; cooperative launch: the device must be able to hold grid = #SMs CTAs)
 *   b200bit_mpq_chain_status     : 0 / 1 = a dependency wait inside the kernel gave up (synchronises the stream)
 * Constraints: w_bit 4, f16 / bf16, contiguous groups of 32 * 2^i values (one group size per chain), K % 128 == 0,
 * N % 8 == 0 (asym: N % 32 == 0), no node may write its own input.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const void* x;            /* [K] activations (dtype of the chain) */
    void* y;                  /* [N] output */
    const int32_t* qweight;   /* int32 [K/8, N] */
    const void* scales;       /* [G, N] */
    const void* zeros;        /* sym: [G, N]; asym: packed int32 [G, N/8] */
    int K, N, G;
    int reserved;
} b200bit_chain_node;
B200BIT_API size_t b200bit_mpq_chain_plan_bytes(const b200bit_chain_node* nodes_host, int n_nodes);
B200BIT_API int b200bit_mpq_chain_build(const b200bit_chain_node* nodes_host, int n_nodes, int w_bit, int asym, int dtype,
                                        void* plan_device, size_t plan_bytes, int* info16_host);
/* Planning only, on the host, no device needed (diagnostics / tests): writes the n 64-byte node records of the plan --
 * {x, y, xll, yll (pointers; shadows as offsets from a fake base), R, N, strips, n28, tiles, off_sig, wx_node, wy_node},
 * see csrc/mpq_chain.cuh -- and info16.  Assumes 148 SMs when no device is present. */
B200BIT_API int b200bit_mpq_chain_plan_host(const b200bit_chain_node* nodes_host, int n_nodes, int w_bit, int asym, int dtype,
                                            void* node_table_host, int* info16_host);
B200BIT_API int b200bit_mpq_chain_launch(void* plan_device, const int* info16_host, unsigned flags, void* stream);
B200BIT_API int b200bit_mpq_chain_status(const void* plan_device, const int* info16_host, int* error_flag_host, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * MPQ grad_input:  dx[M,K] = dy[M,N] @ W^T  (same dequantisation as the forward; fp32 accumulation; deterministic).
 * Replaces q_linear_cuda.mpq_grad_input (q_linear_cuda.cpp:272-284 -> mpq_linear_cuda_kernel.cu:1198-1223,
 * back_quant_mm_kernel{,_asym} :635-1049).  Argument meaning as b200bit_mpq_forward.
 * ------------------------------------------------------------------------------------------------------------ */
B200BIT_API int b200bit_mpq_grad_input(const void* dy, const int32_t* qweight, const void* scales, const void* zeros,
                                       const int32_t* g_idx, void* dx, int M, int K, int N, int G, int w_bit, int asym,
                                       int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Dequantise the whole weight to out[K,N] in `dtype`, bit-identical to the reference's Python unpack_qweight
 * (layer_type 1; bitorch_engine/layers/qlinear/nbit/cuda/utils.py:5-69): sym rnd(rnd(q*s) - z), asym rnd(s*(q - zq)),
 * every op rounded to dtype.  One pass, no temporaries (the reference materialises int32 [K/nb,nb,N] and int8 [K,N]).
 * ------------------------------------------------------------------------------------------------------------ */
B200BIT_API int b200bit_mpq_dequant(const int32_t* qweight, const void* scales, const void* zeros, const int32_t* g_idx,
                                    void* out, int K, int N, int G, int w_bit, int asym, int dtype, int fused,
                                    const int16_t* perm, void* stream);
/* `fused` = 1 selects the rounding of the reference's CUDA dequant kernels, rnd(fma(s, q, -z)) (mbwq_q42fp_weight ->
 * reconstruct_q{4,2}_gptq_kernel, mbwq_linear_cuda_kernel.cu:314-501); `perm` (int16 [K], nullable) scatters packed row
 * k to output row perm[k] (MBWQ q_perm, :398-399).
 *
 * exl2 mixed bit-width dequantise (mbwq_exl2fp_weight -> reconstruct_exl2_kernel, :92-308): rows6 = the six
 * cumulative section ends (8,6,5,4,3,2 bit) returned by mbwq_trans_qweight, a HOST array; fp16 only. */
B200BIT_API int b200bit_exl2_dequant(const int32_t* qweight, const void* scales, const void* zeros, const int16_t* perm,
                                     const int16_t* q_group_map, void* out, int K, int N, const int* rows6_host,
                                     void* stream);

/* Fused exl2 mixed bit-width forward: y[M,N] = x[:, perm] @ W without materialising W (fp16; fp32 accumulation; M rows
 * through the grid, no cuBLAS).  Replaces mbwq_exl2_forward -> gemm_half_q_half_kernel (q_linear_cuda.cpp:338-354,
 * mbwq_linear_cuda_kernel.cu:926-1007, exl2/q_gemm_kernel.cuh:90-549).  rows6_host as for b200bit_exl2_dequant; groups
 * must be runs of 32*i weight rows (the exl2 packing guarantees it); G = rows of the scales / zeros tables.  Up to 4 rows
 * the packed rows stream through a TMA ring (one tensor map per bit width), 5 - 32 rows use register-prefetched loads. */
B200BIT_API int b200bit_exl2_forward(const void* x, const int32_t* qweight, const void* scales, const void* zeros,
                                     const int16_t* perm, const int16_t* q_group_map, void* y, int M, int K, int N, int G,
                                     const int* rows6_host, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Quantise + bit-pack weight[K,N] (dtype) into qweight_out int32 [K*w_bit/32, N], bit-identical to the reference's
 * pack_fp_weight (utils.py:72-147): sym clamp(rint(rnd(rnd(w+z)/s))), asym clamp(rint(rnd(rnd(w/s)+zq))).
 * zeros: sym dtype [G,N]; asym packed int32 [G,N*w_bit/32] (zeros_unpacked=0) or integer zero points stored as dtype
 * [G,N] (zeros_unpacked=1: the `unpacked_zeros` argument).  perm: optional int16 [K] row gather (MBWQ q_perm).
 * weight_dtype: dtype of `weight` -- `dtype` (that of scales / zeros) or F32: the optimizer packs its fp32 master weight
 * against half parameters, torch then promotes and evaluates (and rounds) the whole expression in fp32.
 * ------------------------------------------------------------------------------------------------------------ */
B200BIT_API int b200bit_mpq_pack_weight(const void* weight, const void* scales, const void* zeros, const int32_t* g_idx,
                                        const int16_t* perm, int32_t* qweight_out, int K, int N, int G, int w_bit,
                                        int asym, int zeros_unpacked, int dtype, int weight_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused DiodeMix update of an MPQ weight (one kernel): unpack -> Adam moments -> normalised gradient -> step ->
 * (update_zeros: the every-5th-step zero-point update) -> re-quantise + re-pack.  Replaces the MPQWeightParameter
 * branch of qweight_update_fn (bitorch_engine/utils/model_helper.py:485-530, update_zeros :330-360,
 * gptq_style_unpacking / gptq_style_zeros_packing quant_operators.py:310-368, pack_fp_weight utils.py:72-147),
 * ~25 torch kernels.  Contiguous groups only (g_idx == arange(K) // group).  storage_dtype: dtype of scales (and of
 * sym zeros); compute_dtype: optimizer dtype of exp_avg_l / exp_avg_s (f32 or == storage); grad_dtype: storage or
 * compute dtype.  step_size already carries the bias correction (model_helper.py:501-506).
 * ------------------------------------------------------------------------------------------------------------ */
B200BIT_API int b200bit_diodemix_mpq_step(int32_t* qweight, const void* scales, void* zeros, const void* grad,
                                          void* exp_avg_l, void* exp_avg_s, int K, int N, int G, int w_bit, int asym,
                                          int storage_dtype, int compute_dtype, int grad_dtype, double beta1, double beta2,
                                          double eps, double step_size, int update_zeros, void* stream);

/* Fused DiodeMix update of a binary (int8 +-1) weight: two lerps, sign descent, conditional flip.  Replaces the
 * BinaryLinearParameter branch of qweight_update_fn (model_helper.py:437-445).  Exactly one of grad_i8 / grad_c
 * (compute dtype) is non-NULL. */
B200BIT_API int b200bit_diodemix_binary_step(int8_t* weight, const int8_t* grad_i8, const void* grad_c, void* exp_avg_l,
                                             void* exp_avg_s, size_t numel, int compute_dtype, double beta1, double beta2,
                                             double lr, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Binary (1-bit) Linear:  y[m,n] = K - 2*popc(bits(x[m,:]) xor bits(w[n,:])),  sign bit = (v >= 0); integer-exact.
 * Replaces binary_linear_cuda.{forward, w_pack, mm} (binary_linear_cuda.cpp:92-122; kernels
 * binary_linear_cuda_kernel.cu:59-394, host flows :481-626, :830-889) without the per-call cudaMalloc/cudaFree.
 *   canonical bit matrix: [rows][stride_bytes] bytes, byte kb holds k = 8*kb..8*kb+7, bit (7 - k%8), zero padded,
 *                         stride_bytes a multiple of 4 with stride_bytes*8 >= K.
 *   b200bit_binary_pack     : float/half/bf16/int8 matrix -> canonical bits.  transposed_input = 0: in is [rows, K];
 *                             1: in is [K, rows] (the reference's weight.t().contiguous() / mm operand y).
 *   b200bit_binary_relayout : canonical <-> the reference's packed uint8 weight streams; layout 2 = BTC
 *                             (k%128==0, n%8==0; :118 + :22-41), 1 = BSTC (k%32==0, n%32==0; :201 + :22-41).
 *   b200bit_binary_gemm     : canonical x bits [M] and w bits [N] -> out [M,N] in out_dtype (F32/F16/BF16/I32).
 * ------------------------------------------------------------------------------------------------------------ */
B200BIT_API int b200bit_binary_pack(const void* in, int in_dtype, int rows, int K, int transposed_input, uint8_t* out,
                                    int stride_bytes, void* stream);
B200BIT_API int b200bit_binary_relayout(const uint8_t* in, uint8_t* out, int N, int K, int layout, int to_reference,
                                        int canon_stride_bytes, void* stream);
B200BIT_API int b200bit_binary_gemm(const uint8_t* x_bits, const uint8_t* w_bits, void* out, int M, int N, int K,
                                    int stride_bytes, int out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * HOST (CPU) binary Linear -- the reference's `binary_linear_cpp` extension for BinaryLinearCPP
 * (binary/cpp/binary_linear.cpp:494-518: forward(input, weights, m, n, k), w_pack(weights, n, k)).  All pointers are
 * HOST pointers; fp32 input only, as the reference (data_ptr<float>, :425).  Same arithmetic as the CUDA path; packed
 * weight layout of the CPU extension: byte (k/8)*N + n, bit j (LSB first) = sign(w[n, 8*(k/8)+j]) (:80-145).
 * ------------------------------------------------------------------------------------------------------------ */
B200BIT_API int b200bit_cpu_binary_pack(const float* weights_host, uint8_t* out_host, int n, int k);
B200BIT_API int b200bit_cpu_binary_forward(const float* x_host, const void* weights_host, int weights_packed,
                                           float* out_host, int m, int n, int k, int threads);

/* ------------------------------------------------------------------------------------------------------------
 * Wire-format conversions of the reference's `functions_cuda` extension (bitorch_engine/functions/cuda/
 * functions_cuda.cpp:160-200; kernels functions_cuda_kernel.cu:74-209).  One-pass HBM streams.
 *   b200bit_q4_pack         : int32 codes -> bytes, out[i] = (in[2i] & 15) << 4 | (in[2i+1] & 15)   (:136-159);
 *                             n_bytes = number of OUTPUT bytes (= codes / 2)
 *   b200bit_q4_unpack       : bytes -> int32 codes 0..15, high nibble first                       (:162-182)
 *   b200bit_q4_unpack_scale : bytes -> f32, codes read as signed 4-bit (-8..7) times `scale`       (:185-209)
 *   b200bit_sign_pack_u8    : eight values -> one byte, bit i = (v[8j+i] >= 0), LSB first; in_dtype F32 / F16 / BF16 /
 *                             I8; n_bytes = number of OUTPUT bytes                                  (:74-119)
 *   b200bit_sign_unpack_u8  : byte -> eight floats +-scale[byte_index / packed_dim], LSB first      (:123-133)
 * `in` / `out` of the vectorised side must be 16-byte aligned (the shims clone unaligned views).
 * fp32toint4 (:23-69, :239-262) is deliberately absent: the reference reads uninitialised shared memory there.
 * ------------------------------------------------------------------------------------------------------------ */
B200BIT_API int b200bit_q4_pack(const int32_t* in, int8_t* out, size_t n_bytes, void* stream);
B200BIT_API int b200bit_q4_unpack(const int8_t* in, int32_t* out, size_t n_bytes, void* stream);
B200BIT_API int b200bit_q4_unpack_scale(const int8_t* in, float scale, float* out, size_t n_bytes, void* stream);
B200BIT_API int b200bit_sign_pack_u8(const void* in, int in_dtype, uint8_t* out, size_t n_bytes, void* stream);
B200BIT_API int b200bit_sign_unpack_u8(const uint8_t* in, const float* scale, float* out, size_t n_bytes, size_t packed_dim,
                                       void* stream);

/* Host-side plan of the 4-bit decode kernel (mpq_imma.cuh) for a layer shape, without launching anything (tests of the
 * host logic; needs no device: the SM count falls back to 148).  out8 receives {ok, grid, strips, strips_28_wide,
 * tiles_per_strip, ring_slots, k_steps_per_flush, shared_memory_bytes}; returns 0, or B200BIT_ERR_ARG for a null pointer. */
B200BIT_API int b200bit_mpq_decode_plan(int K, int N, int G, int w_bit, int asym, int dtype, int* out8);

/* Sweep hook for bench.py / tests (process-wide; 0 = built-in heuristic): lanes per packed-row segment (8, 16, 32),
 * warps per CTA (1..16), split-K factor of the decode GEMV.  No reference counterpart. */
B200BIT_API int b200bit_set_gemv_tuning(int lanes_per_row, int warps, int splitk);
/* Kernel-path override (process-wide): 0 auto, 1 CUDA-core FHFMA GEMV, 2 small-batch mma.sync kernel, 3 general
 * fallback, 4 TMA-streamed small-batch kernel, 5 tcgen05 batched kernel, 6 pipelined decode GEMV (fp16-subnormal
 * FHFMA / HMMA), 7 persistent integer-tensor-pipe decode GEMV (4-bit); mma_for_m1 selects, in auto mode, whether M == 1 uses the tensor kernel (1) or the CUDA-core GEMV (0). */
B200BIT_API int b200bit_set_path(int path, int mma_for_m1);
/* Diagnostics: device buffer of [grid][16][8] uint64 that receives %globaltimer stamps from the TMA-streamed kernel
 * (per CTA, per warp: start, init done, dependency wait done, first tile landed, first run done, all runs done,
 * CTA barrier passed, end); NULL switches tracing off. */
B200BIT_API int b200bit_set_trace_buffer(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* B200BIT_H_ */
