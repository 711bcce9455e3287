// optim.cu -- fused DiodeMix weight updates (one pass over [K,N], no temporaries).
//
// Reference replaced: DiodeMix.step -> Parameter.update -> qweight_update_fn
//   (bitorch_engine/optim/diode_beta.py:93-196, bitorch_engine/utils/model_helper.py:363-530):
//   * MPQWeightParameter branch (:485-530): gptq_style_unpacking (quant_operators.py:310-345) -> Adam moments ->
//     normalised gradient -> w -= step*ng -> every 5th step update_zeros (:330-360) + gptq_style_zeros_packing
//     (quant_operators.py:348-368) -> pack_fp_weight (nbit/cuda/utils.py:72-147): ~25 torch elementwise kernels over
//     [K,N] plus empty_cache() every step.  Here: ONE kernel, thread <-> (group, column), reads code / grad / moments
//     once, writes moments / code once, and reduces the group's zero-point statistic in registers.
//   * BinaryLinearParameter branch (:437-445): 8 torch kernels -> one elementwise kernel.
// Arithmetic follows torch's op sequence and rounding (every op evaluated in fp32 and rounded to the compute dtype C;
// add_(alpha=) and addcmul_ are fused multiply-adds in fp32, as in torch's CPU and CUDA kernels).
#include "common.cuh"

namespace b200bit {

template <int DT> struct OEl;
template <> struct OEl<B200BIT_F32> {
    __device__ static float ld(const void* p, size_t i) { return reinterpret_cast<const float*>(p)[i]; }
    __device__ static void st(void* p, size_t i, float v) { reinterpret_cast<float*>(p)[i] = v; }
    __device__ static float rnd(float v) { return v; }
};
template <> struct OEl<B200BIT_F16> {
    __device__ static float ld(const void* p, size_t i) { return __half2float(reinterpret_cast<const __half*>(p)[i]); }
    __device__ static void st(void* p, size_t i, float v) { reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); }
    __device__ static float rnd(float v) { return __half2float(__float2half_rn(v)); }
};
template <> struct OEl<B200BIT_BF16> {
    __device__ static float ld(const void* p, size_t i) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]); }
    __device__ static void st(void* p, size_t i, float v) { reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); }
    __device__ static float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

struct MpqStepParams {
    uint32_t* qweight;        // [K/nb, N] in/out
    const void* scales;       // [G, N]   storage dtype S
    void* zeros;              // asym: packed u32 [G, N/nb] in/out (rewritten when update_zeros); sym: S [G, N] in/out
    const void* grad;         // [K, N]   storage dtype S (privileged_grad) or C when it went through a projector
    void* exp_avg_l;          // [K, N]   compute dtype C
    void* exp_avg_s;          // [K, N]   compute dtype C
    int K, N, G, w_bit, asym;
    int grad_is_c;            // grad stored in C (1) or S (0)
    int update_zeros;
    float beta1, beta2, one_m_beta1, one_m_beta2, eps, neg_step, step_size;
};

// thread <-> (group g, column n); loops over the group's packed rows
template <int SDT, int CDT>
__global__ void __launch_bounds__(128) diodemix_mpq_kernel(const MpqStepParams p) {
    using S = OEl<SDT>;
    using C = OEl<CDT>;
    const int n_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    const bool ok = n_raw < p.N;               // out-of-range lanes shadow the last column (no stores) so that the
    const int n = ok ? n_raw : p.N - 1;        // whole warp reaches the shuffle of the zero-point packing
    const int nb = 32 / p.w_bit, gs = p.K / p.G, rpg = gs / nb;
    const uint32_t mask = (1u << p.w_bit) - 1u;
    const float maxq = float(mask);
    const float s = S::ld(p.scales, size_t(g) * p.N + n);
    float z;        // asym: integer zero point (+1 form); sym: fp zero
    uint32_t zword = 0;
    if (p.asym) {
        zword = reinterpret_cast<const uint32_t*>(p.zeros)[size_t(g) * (p.N / nb) + n / nb];
        z = float(((zword >> ((n % nb) * p.w_bit)) & mask) + 1u);
    } else {
        z = S::ld(p.zeros, size_t(g) * p.N + n);
    }
    float zsum = 0.f;   // sum over the group of the per-element zero statistic
    for (int rr = 0; rr < rpg; ++rr) {
        const int r = g * rpg + rr;
        const uint32_t w_in = p.qweight[size_t(r) * p.N + n];
        uint32_t w_out = 0;
        for (int j = 0; j < nb; ++j) {
            const size_t idx = size_t(r * nb + j) * p.N + n;
            const float q = float((w_in >> (j * p.w_bit)) & mask);
            // gptq_style_unpacking: dequantise in the storage dtype, then .to(C)
            float w = p.asym ? S::rnd(__fmul_rn(s, __fsub_rn(q, z))) : S::rnd(__fsub_rn(S::rnd(__fmul_rn(q, s)), z));
            w = C::rnd(w);
            const float gr = p.grad_is_c ? C::ld(p.grad, idx) : S::ld(p.grad, idx);
            float m = C::ld(p.exp_avg_l, idx), v = C::ld(p.exp_avg_s, idx);
            m = C::rnd(__fmul_rn(m, p.beta1));
            m = C::rnd(fmaf(p.one_m_beta1, gr, m));                       // add_(grad, alpha=1-beta1)
            v = C::rnd(__fmul_rn(v, p.beta2));
            v = C::rnd(fmaf(__fmul_rn(p.one_m_beta2, gr), gr, v));        // addcmul_(grad, grad, value=1-beta2)
            const float denom = C::rnd(__fadd_rn(C::rnd(__fsqrt_rn(v)), p.eps));
            const float ng = C::rnd(__fdiv_rn(m, denom));
            w = C::rnd(fmaf(p.neg_step, ng, w));                          // w.add_(norm_grad, alpha=-step_size)
            if (ok) {
                C::st(p.exp_avg_l, idx, m);
                C::st(p.exp_avg_s, idx, v);
            }
            if (p.update_zeros) {
                // asym: zeros_unpack.add_(step_size * norm_grad) per element; sym: mean(norm_grad) of the group
                zsum += p.asym ? C::rnd(__fadd_rn(C::rnd(z), C::rnd(__fmul_rn(p.step_size, ng)))) : ng;
            }
            // pack_fp_weight with the PRE-update zero points (model_helper.py:525 passes z_unpacked)
            float t;
            if (p.asym) t = C::rnd(__fadd_rn(C::rnd(__fdiv_rn(w, s)), z));
            else t = C::rnd(__fdiv_rn(C::rnd(__fadd_rn(w, z)), s));
            float c = rintf(t);
            c = fminf(fmaxf(c, 0.f), maxq);
            if (!(c == c)) c = 0.f;
            w_out |= (uint32_t(c) & mask) << (j * p.w_bit);
        }
        if (ok) p.qweight[size_t(r) * p.N + n] = w_out;
    }
    if (p.update_zeros) {
        const float mean = C::rnd(zsum / float(gs));
        if (p.asym) {
            // gptq_style_zeros_packing: trunc to int32, (z - 1) & mask, LSB-first along N.  nb adjacent columns share
            // a word: combine them with shuffles (nb <= 32 columns sit in one warp because blockDim.x % 32 == 0
            // and N % nb == 0) and let the first column of the word write it.
            const int zi = int(mean);                                    // .to(torch.int32) truncates toward zero
            uint32_t field = (uint32_t(zi - 1) & mask) << ((n % nb) * p.w_bit);
            for (int off = 1; off < nb; off <<= 1) field |= __shfl_xor_sync(0xffffffffu, field, off);
            if (ok && n % nb == 0) reinterpret_cast<uint32_t*>(p.zeros)[size_t(g) * (p.N / nb) + n / nb] = field;
        } else {
            // sym: zeros.add_(step_size * mean(norm_grad))  (MBWQ rule, model_helper.py:345-347)
            if (ok) S::st(p.zeros, size_t(g) * p.N + n, S::rnd(__fadd_rn(z, S::rnd(C::rnd(__fmul_rn(p.step_size, mean))))));
        }
    }
}

// binary branch: exp_avg_l.lerp_(grad, 1-b1); v = sign(exp_avg_l)*lr; exp_avg_s.lerp_(v, 1-b2); u = -sign(exp_avg_s),
// u[u==0] = 1; flip w where u != sign(w)   (model_helper.py:437-445).  w: int8 [numel]; grad: int8 (nv_tensor_quant
// output, binary/cuda/layer.py:120) or any float dtype converted by the caller to C.
__device__ __forceinline__ float lerp_torch(float a, float b, float w) {
    // at::native lerp: weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w), evaluated with fused multiply-adds
    const float d = __fsub_rn(b, a);
    return (w < 0.5f) ? fmaf(w, d, a) : fmaf(-d, __fsub_rn(1.0f, w), b);
}
__device__ __forceinline__ float sgn(float x) { return float((x > 0.f) - (x < 0.f)); }

template <int CDT>
__global__ void __launch_bounds__(256) diodemix_binary_kernel(int8_t* __restrict__ w, const int8_t* __restrict__ grad_i8,
                                                              const void* __restrict__ grad_c, void* __restrict__ m_,
                                                              void* __restrict__ s_, size_t numel, float w1, float w2,
                                                              float lr) {
    using C = OEl<CDT>;
    // four elements per thread, one block-width apart (every load instruction of a warp stays contiguous): all twelve
    // loads are in flight before the first result is needed
    const size_t base = blockIdx.x * size_t(blockDim.x) * 4 + threadIdx.x;
    float g4[4], m4[4], s4[4];
    int8_t w4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const size_t i = base + size_t(u) * blockDim.x;
        g4[u] = m4[u] = s4[u] = 0.f;
        w4[u] = 0;
        if (i < numel) {
            g4[u] = grad_i8 ? float(grad_i8[i]) : C::ld(grad_c, i);
            m4[u] = C::ld(m_, i);
            s4[u] = C::ld(s_, i);
            w4[u] = w[i];
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const size_t i = base + size_t(u) * blockDim.x;
        if (i >= numel) continue;
        const float m = C::rnd(lerp_torch(m4[u], C::rnd(g4[u]), w1));
        const float v = C::rnd(__fmul_rn(sgn(m), lr));
        const float s = C::rnd(lerp_torch(s4[u], v, w2));
        C::st(m_, i, m);
        C::st(s_, i, s);
        float uu = -sgn(s);
        if (uu == 0.f) uu = 1.f;
        const int8_t wv = w4[u];
        const float sw = float((wv > 0) - (wv < 0));
        if (uu != sw) w[i] = int8_t(-wv);
    }
}

}  // namespace b200bit

using namespace b200bit;

extern "C" {

int b200bit_diodemix_mpq_step(int32_t* qweight, const void* scales, void* zeros, const void* grad, void* exp_avg_l,
                              void* exp_avg_s, int K, int N, int G, int w_bit, int asym, int storage_dtype,
                              int compute_dtype, int grad_dtype, double beta1, double beta2, double eps, double step_size,
                              int update_zeros, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(qweight && scales && zeros && grad && exp_avg_l && exp_avg_s, B200BIT_ERR_ARG,
                 "diodemix_mpq_step: null pointer argument");
    B200_REQUIRE(w_bit == 1 || w_bit == 2 || w_bit == 4 || w_bit == 8, B200BIT_ERR_UNSUPPORTED,
                 "diodemix_mpq_step: w_bit=%d", w_bit);
    const int nb = 32 / w_bit;
    B200_REQUIRE(K > 0 && N > 0 && G > 0 && K % G == 0 && (K / G) % nb == 0, B200BIT_ERR_SHAPE,
                 "diodemix_mpq_step: need K %% G == 0 and group size %% %d == 0 (K=%d G=%d)", nb, K, G);
    B200_REQUIRE(!asym || N % nb == 0, B200BIT_ERR_SHAPE, "diodemix_mpq_step: asym needs N %% %d == 0", nb);
    B200_REQUIRE(compute_dtype == B200BIT_F32 || compute_dtype == storage_dtype, B200BIT_ERR_UNSUPPORTED,
                 "diodemix_mpq_step: compute dtype must be f32 or equal the storage dtype");
    B200_REQUIRE(grad_dtype == storage_dtype || grad_dtype == compute_dtype, B200BIT_ERR_UNSUPPORTED,
                 "diodemix_mpq_step: grad dtype must equal the storage or the compute dtype");
    MpqStepParams p{};
    p.qweight = reinterpret_cast<uint32_t*>(qweight);
    p.scales = scales; p.zeros = zeros; p.grad = grad; p.exp_avg_l = exp_avg_l; p.exp_avg_s = exp_avg_s;
    p.K = K; p.N = N; p.G = G; p.w_bit = w_bit; p.asym = asym;
    p.grad_is_c = (grad_dtype == compute_dtype && compute_dtype != storage_dtype) ? 1 : 0;
    p.update_zeros = update_zeros;
    p.beta1 = float(beta1); p.beta2 = float(beta2);
    p.one_m_beta1 = float(1.0 - beta1); p.one_m_beta2 = float(1.0 - beta2);
    p.eps = float(eps); p.neg_step = float(-step_size); p.step_size = float(step_size);
    dim3 grid((N + 127) / 128, G);
#define B200_LAUNCH_STEP(SD, CD) diodemix_mpq_kernel<SD, CD><<<grid, 128, 0, st>>>(p)
    if (storage_dtype == B200BIT_F16 && compute_dtype == B200BIT_F32) B200_LAUNCH_STEP(B200BIT_F16, B200BIT_F32);
    else if (storage_dtype == B200BIT_F16) B200_LAUNCH_STEP(B200BIT_F16, B200BIT_F16);
    else if (storage_dtype == B200BIT_BF16 && compute_dtype == B200BIT_F32) B200_LAUNCH_STEP(B200BIT_BF16, B200BIT_F32);
    else if (storage_dtype == B200BIT_BF16) B200_LAUNCH_STEP(B200BIT_BF16, B200BIT_BF16);
    else if (storage_dtype == B200BIT_F32) B200_LAUNCH_STEP(B200BIT_F32, B200BIT_F32);
    else return set_error(B200BIT_ERR_ARG, "diodemix_mpq_step: bad dtype codes %d/%d", storage_dtype, compute_dtype);
#undef B200_LAUNCH_STEP
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

int b200bit_diodemix_binary_step(int8_t* weight, const int8_t* grad_i8, const void* grad_c, void* exp_avg_l,
                                 void* exp_avg_s, size_t numel, int compute_dtype, double beta1, double beta2, double lr,
                                 void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(weight && (grad_i8 || grad_c) && exp_avg_l && exp_avg_s, B200BIT_ERR_ARG,
                 "diodemix_binary_step: null pointer argument");
    if (numel == 0) return B200BIT_OK;
    const unsigned blocks = unsigned((numel + 1023) / 1024);
    const float w1 = float(1.0 - beta1), w2 = float(1.0 - beta2), lrf = float(lr);
    if (compute_dtype == B200BIT_F32) diodemix_binary_kernel<B200BIT_F32><<<blocks, 256, 0, st>>>(weight, grad_i8, grad_c, exp_avg_l, exp_avg_s, numel, w1, w2, lrf);
    else if (compute_dtype == B200BIT_F16) diodemix_binary_kernel<B200BIT_F16><<<blocks, 256, 0, st>>>(weight, grad_i8, grad_c, exp_avg_l, exp_avg_s, numel, w1, w2, lrf);
    else if (compute_dtype == B200BIT_BF16) diodemix_binary_kernel<B200BIT_BF16><<<blocks, 256, 0, st>>>(weight, grad_i8, grad_c, exp_avg_l, exp_avg_s, numel, w1, w2, lrf);
    else return set_error(B200BIT_ERR_ARG, "diodemix_binary_step: bad dtype code %d", compute_dtype);
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}

}  // extern "C"
