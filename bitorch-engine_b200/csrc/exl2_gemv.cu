// exl2_gemv.cu -- fused mixed bit-width (exl2) GEMV / small-batch GEMM:  y[M,N] = x[:, perm] @ W,  M <= 32 per launch
// group (any M through the grid's y dimension), W never materialised.
//
// Format (mbwq_linear_cuda_kernel.cu:92-308, exl2/quant/qdq_*.cuh): the rows of W are sorted by bit-width 8,6,5,4,3,2
// ("sections"); inside a section of width b the codes of one column form an LSB-first bit stream over consecutive packed
// rows: 32 codes per b words.  Per-group fp16 scale / zero: w = q * s - z, group of weight row k = q_group_map[2k],
// groups are runs of 32 * i rows, so a 32-row block never straddles a group or a section.
//
// Kernel: lane = output column (a packed row is N contiguous words: every weight load is a full 128-byte line), warp =
// k-slice of the column strip, eight warps per CTA and up to four CTAs of a thread-block CLUSTER per strip (so that a
// 4096-column layer puts 512 CTAs on the 148 SMs instead of 128): the warps' partial sums meet in shared memory, the
// CTAs' through distributed shared memory in rank order -- y is written once, no atomics, deterministic.  A block of 32 codes costs b coalesced word loads; a code becomes the fp16 number 1024 + q with
// one funnel shift + one LOP3 ((v & mask) | 0x6400) and is multiplied with the activation by the mixed-precision FMA
// (fma.rn.f32.f16: exact product, fp32 accumulation).  The group affine is factored out:
//     y += s * sum((1024 + q) x) - (1024 s + z) * sum(x)
// Activations are gathered through q_perm into shared memory once per 256-row chunk, with the per-block sums of x.
// Replaces gemm_half_q_half_kernel (exl2/q_gemm_kernel.cuh:90-549; 64-row CTAs, K/64 half2 atomics per output, fp16
// accumulation) and, in this repo's round 1, a dequantise + cuBLAS pair.
#include "common.cuh"
#include <cooperative_groups.h>

namespace b200bit {

constexpr int EX_WARPS = 8;
constexpr int EX_CHUNK = 256;              // weight rows staged per warp at a time (8 blocks of 32)
struct Exl2Sections { int end[6]; int prow[6]; };   // section end (weight rows), first packed row; order 8,6,5,4,3,2

struct Exl2Params {
    const uint32_t* qw;        // [rows_packed, N]
    const __half* scales;      // [G, N]
    const __half* zeros;       // [G, N]
    const uint16_t* perm;      // [K] or null
    const uint16_t* gmap;      // [2K] (group, rows left)
    const __half* x;           // [M, K]
    __half* y;                 // [M, N]
    int M, K, N;
    int blocks_per_warp;       // 32-row blocks per warp (k-slice)
    Exl2Sections sec;
};

// 32 codes of width B from B consecutive words of one column -> fp16 bit patterns of 1024 + q, one per register
template <int B>
__device__ __forceinline__ void ex_unpack(const uint32_t (&w)[8], uint32_t (&h)[32]) {
    constexpr uint32_t mask = (1u << B) - 1u;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        constexpr int dummy = 0; (void)dummy;
        const int bit = j * B, i = bit >> 5, sh = bit & 31;
        uint32_t v;
        if (sh + B <= 32) v = w[i] >> sh;
        else v = __funnelshift_r(w[i], w[i + 1], sh);
        h[j] = (v & mask) | 0x6400u;
    }
}

template <int B, int MB>
__device__ __forceinline__ void ex_block(const Exl2Params& p, int n, bool col_ok, int prow, int group, const uint4* xs,
                                         const float* xsum, float (&yacc)[MB]) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = (i < B && col_ok) ? __ldg(p.qw + size_t(prow + i) * p.N + n) : 0u;
    float s = 0.f, z = 0.f;
    if (col_ok) {
        s = __half2float(p.scales[size_t(group) * p.N + n]);
        z = __half2float(p.zeros[size_t(group) * p.N + n]);
    }
    uint32_t h[32];
    ex_unpack<B>(w, h);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 xv = xs[m * (EX_CHUNK / 8) + q];        // 8 activations
            const uint32_t xr[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc = fhfma<false, false, false>(h[8 * q + 2 * r], xr[r], acc);
                acc = fhfma<false, false, true>(h[8 * q + 2 * r + 1], xr[r], acc);
            }
        }
        yacc[m] = fmaf(s, acc, yacc[m]);
        yacc[m] = fmaf(-(1024.f * s + z), xsum[m], yacc[m]);
    }
}

template <int MB>
__global__ void __launch_bounds__(EX_WARPS * 32) exl2_gemv_kernel(const Exl2Params p) {
    extern __shared__ __align__(16) unsigned char ex_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const bool col_ok = n < p.N;
    const int m0 = blockIdx.y * MB;
    // per warp: x chunk [MB][256] halfs, block sums [MB][8] f32
    __half* xs_w = reinterpret_cast<__half*>(ex_smem) + size_t(warp) * MB * EX_CHUNK;
    float* xsum_w = reinterpret_cast<float*>(ex_smem + size_t(EX_WARPS) * MB * EX_CHUNK * sizeof(__half)) + warp * MB * 8;
    float* red = reinterpret_cast<float*>(ex_smem + size_t(EX_WARPS) * MB * (EX_CHUNK * sizeof(__half) + 8 * sizeof(float)));

    float yacc[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) yacc[m] = 0.f;

    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = int(cluster.block_rank()), csize = int(cluster.num_blocks());
    const int blocks_total = p.K / 32;
    const int b_lo = (crank * EX_WARPS + warp) * p.blocks_per_warp;
    const int b_hi = min(blocks_total, b_lo + p.blocks_per_warp);
    for (int c0 = b_lo; c0 < b_hi; c0 += EX_CHUNK / 32) {
        const int nblk = min(EX_CHUNK / 32, b_hi - c0);
        // ---- stage the activations of the chunk (gathered through q_perm) + their per-block sums ----
        __syncwarp();
        for (int i = lane; i < MB * nblk * 32; i += 32) {
            const int m = i / (nblk * 32), kl = i % (nblk * 32);
            const int k = c0 * 32 + kl;
            const int src = p.perm ? int(p.perm[k]) : k;
            __half v = __float2half(0.f);
            if (m0 + m < p.M) v = p.x[size_t(m0 + m) * p.K + src];
            xs_w[m * EX_CHUNK + kl] = v;
        }
        __syncwarp();
        for (int i = lane; i < MB * nblk; i += 32) {
            const int m = i / nblk, b = i % nblk;
            float sum = 0.f;
            const __half2* src = reinterpret_cast<const __half2*>(xs_w + m * EX_CHUNK + b * 32);
#pragma unroll
            for (int q = 0; q < 16; ++q) { const float2 f = __half22float2(src[q]); sum += f.x + f.y; }
            xsum_w[m * 8 + b] = sum;
        }
        __syncwarp();
        for (int b = 0; b < nblk; ++b) {
            const int blk = c0 + b, k = blk * 32;
            int sec = 0;
            while (sec < 5 && k >= p.sec.end[sec]) ++sec;
            const int k_sec = sec == 0 ? 0 : p.sec.end[sec - 1];
            const int widths[6] = {8, 6, 5, 4, 3, 2};
            const int bits = widths[sec];
            const int prow = p.sec.prow[sec] + (k - k_sec) / 32 * bits;
            const int group = p.gmap[2 * k];
            const uint4* xs = reinterpret_cast<const uint4*>(xs_w + b * 32);
            float xsum[MB];
#pragma unroll
            for (int m = 0; m < MB; ++m) xsum[m] = xsum_w[m * 8 + b];
            switch (bits) {
                case 8: ex_block<8, MB>(p, n, col_ok, prow, group, xs, xsum, yacc); break;
                case 6: ex_block<6, MB>(p, n, col_ok, prow, group, xs, xsum, yacc); break;
                case 5: ex_block<5, MB>(p, n, col_ok, prow, group, xs, xsum, yacc); break;
                case 4: ex_block<4, MB>(p, n, col_ok, prow, group, xs, xsum, yacc); break;
                case 3: ex_block<3, MB>(p, n, col_ok, prow, group, xs, xsum, yacc); break;
                default: ex_block<2, MB>(p, n, col_ok, prow, group, xs, xsum, yacc); break;
            }
        }
    }
    // ---- the eight k-slices of the CTA meet in shared memory, then the CTAs of the cluster in rank order ----
#pragma unroll
    for (int m = 0; m < MB; ++m) red[(warp * MB + m) * 32 + lane] = yacc[m];
    __syncthreads();
    float* cta_sum = red + EX_WARPS * MB * 32;             // [MB][32]
    for (int i = threadIdx.x; i < MB * 32; i += EX_WARPS * 32) {
        const int m = i / 32, l = i % 32;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < EX_WARPS; ++w) t += red[(w * MB + m) * 32 + l];
        cta_sum[i] = t;
    }
    cluster.sync();
    if (crank == 0) {
        for (int i = threadIdx.x; i < MB * 32; i += EX_WARPS * 32) {
            const int m = i / 32, l = i % 32;
            float t = cta_sum[i];
            for (int r = 1; r < csize; ++r) t += cluster.map_shared_rank(cta_sum, r)[i];
            const int col = blockIdx.x * 32 + l;
            if (col < p.N && m0 + m < p.M) p.y[size_t(m0 + m) * p.N + col] = __float2half_rn(t);
        }
    }
    cluster.sync();          // the other CTAs' shared memory stays alive until rank 0 has read it
}

template <int MB>
static int launch_exl2(Exl2Params p, cudaStream_t st) {
    const size_t smem = size_t(EX_WARPS) * MB * (EX_CHUNK * sizeof(__half) + 8 * sizeof(float)) +
                        size_t(EX_WARPS + 1) * MB * 32 * sizeof(float);
    auto kern = exl2_gemv_kernel<MB>;
    if (smem > 48 * 1024) B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    // k-slices: 8 warps x cluster size; enough CTAs for ~3 per SM, at least 2 blocks of 32 rows per warp
    const int blocks = p.K / 32, strips = (p.N + 31) / 32, mchunks = (p.M + MB - 1) / MB;
    int csize = 1;
    while (csize < 4 && strips * mchunks * csize < 3 * sm_count() && blocks >= 2 * EX_WARPS * csize * 2) csize *= 2;
    p.blocks_per_warp = (blocks + EX_WARPS * csize - 1) / (EX_WARPS * csize);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(strips, mchunks, csize);
    cfg.blockDim = dim3(EX_WARPS * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = csize;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    return B200BIT_OK;
}

}  // namespace b200bit

using namespace b200bit;

extern "C" int b200bit_exl2_forward(const void* x, const int32_t* qweight, const void* scales, const void* zeros,
                                    const int16_t* perm, const int16_t* q_group_map, void* y, int M, int K, int N,
                                    const int* rows6, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(x && qweight && scales && zeros && q_group_map && y && rows6, B200BIT_ERR_ARG, "exl2_forward: null pointer argument");
    B200_REQUIRE(M >= 0 && K > 0 && N > 0 && K <= 65535 && K % 32 == 0, B200BIT_ERR_SHAPE,
                 "exl2_forward: bad sizes M=%d K=%d N=%d (K a multiple of 32, at most 65535)", M, K, N);
    if (M == 0) return B200BIT_OK;
    const int widths[6] = {8, 6, 5, 4, 3, 2};
    Exl2Params p{};
    int prev = 0, prow = 0;
    for (int i = 0; i < 6; ++i) {
        B200_REQUIRE(rows6[i] >= prev && rows6[i] <= K && (rows6[i] - prev) % 32 == 0, B200BIT_ERR_SHAPE,
                     "exl2_forward: rows[%d]=%d is not a cumulative multiple of 32 within K=%d", i, rows6[i], K);
        p.sec.end[i] = rows6[i];
        p.sec.prow[i] = prow;
        prow += (rows6[i] - prev) * widths[i] / 32;
        prev = rows6[i];
    }
    B200_REQUIRE(prev == K, B200BIT_ERR_SHAPE, "exl2_forward: rows cover %d of K=%d weight rows", prev, K);
    p.qw = reinterpret_cast<const uint32_t*>(qweight);
    p.scales = reinterpret_cast<const __half*>(scales);
    p.zeros = reinterpret_cast<const __half*>(zeros);
    p.perm = reinterpret_cast<const uint16_t*>(perm);
    p.gmap = reinterpret_cast<const uint16_t*>(q_group_map);
    p.x = reinterpret_cast<const __half*>(x);
    p.y = reinterpret_cast<__half*>(y);
    p.M = M; p.K = K; p.N = N;
    if (M == 1) return launch_exl2<1>(p, st);
    if (M == 2) return launch_exl2<2>(p, st);
    if (M <= 4) return launch_exl2<4>(p, st);
    return launch_exl2<8>(p, st);
}
