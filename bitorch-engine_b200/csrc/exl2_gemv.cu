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
#include "tma.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
#include <type_traits>

namespace b200bit {

constexpr int EX_WARPS = 8;
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
    int blocks_per_slice;      // 32-row blocks per k-slice
    Exl2Sections sec;
};

// MB rows of x per pass.  The eight warps of a CTA form WN column strips x WK k-slices: decode (MB <= 8) is latency-bound
// and wants many short k-slices (1 x 8); at 16 / 32 rows the activations cost as much L2 traffic as the weights (they are
// gathered through q_perm, one sector per value), so warps side by side share one staged copy (8 x 1 at 16 rows, 4 x 2 at
// 32 rows, where twice the k-slices are needed to put two CTAs on every SM).
// XB = 32-row blocks of x staged per k-slice at a time; D = blocks whose packed words are in flight per warp.
template <int MB> struct ExCfg {
    static constexpr int WN = MB <= 8 ? 1 : (MB == 16 ? 8 : 4);
    static constexpr int WK = EX_WARPS / WN;
    // (staging the whole k-slice of x at once + 4 blocks in flight was measured no better for decode: 11.6 / 24.3 / 30.2 us
    // against 11.3 / 24.9 / 26.5 us on the three Llama-7B shapes)
    static constexpr int XB = MB <= 8 ? 8 : 4;
    static constexpr int D = MB <= 8 ? 3 : (MB == 16 ? 2 : 1);
};

// 32 codes of width B from B consecutive words of one column -> fp16 bit patterns of 1024 + q, one per register
template <int B>
__device__ __forceinline__ void ex_unpack(const uint32_t (&w)[B], uint32_t (&h)[32]) {
    constexpr uint32_t mask = (1u << B) - 1u;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int bit = j * B, i = bit >> 5, sh = bit & 31;
        uint32_t v;
        if (sh + B <= 32) v = w[i] >> sh;
        else v = __funnelshift_r(w[i], w[i + 1 < B ? i + 1 : i], sh);
        h[j] = (v & mask) | 0x6400u;
    }
}

// issue the loads of one block (32 weight rows, B packed rows) of column n; only the group index is waited for
template <int B>
__device__ __forceinline__ void ex_fetch(const Exl2Params& p, const uint32_t* src, int k, int n, bool col_ok, uint32_t (&w)[B],
                                         unsigned short& s_raw, unsigned short& z_raw) {
    const size_t stride = size_t(p.N);
#pragma unroll
    for (int i = 0; i < B; ++i) { w[i] = col_ok ? __ldg(src) : 0u; src += stride; }
    const int group = __ldg(p.gmap + 2 * k);
    s_raw = z_raw = 0;          // raw halves, untouched until the block is computed: nothing here may wait for them
    if (col_ok) {
        const size_t at = size_t(group) * p.N + n;
        s_raw = __ldg(reinterpret_cast<const unsigned short*>(p.scales) + at);
        z_raw = __ldg(reinterpret_cast<const unsigned short*>(p.zeros) + at);
    }
}

// xs: the block's 32 activations of row m at xs + m * XSTRIDE halfs (16-byte aligned); xsum[m * XB]: their sum
template <int B, int MB, int XSTRIDE, int XB>
__device__ __forceinline__ void ex_block(const uint32_t (&w)[B], unsigned short s_raw, unsigned short z_raw, const __half* xs,
                                         const float* xsum, float (&yacc)[MB]) {
    uint32_t h[32];
    ex_unpack<B>(w, h);
    const float s = __half2float(__ushort_as_half(s_raw));
    const float z = __half2float(__ushort_as_half(z_raw));
    const float c = -(1024.f * s + z);
    // MR rows at a time, two chains per row: the mixed-precision FMA has a long latency (~20 clk measured) and at 16 / 32
    // rows only two warps share a scheduler, so the independent chains have to come from inside the warp
    constexpr int MR = MB >= 4 ? 4 : MB;
#pragma unroll
    for (int m = 0; m < MB; m += MR) {
        float a0[MR], a1[MR];
#pragma unroll
        for (int r = 0; r < MR; ++r) a0[r] = a1[r] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t xr[MR][4];
#pragma unroll
            for (int r = 0; r < MR; ++r) {
                // 8 activations (same address in every lane: one broadcast wavefront)
                const uint4 xv = reinterpret_cast<const uint4*>(xs + (m + r) * XSTRIDE)[q];
                xr[r][0] = xv.x; xr[r][1] = xv.y; xr[r][2] = xv.z; xr[r][3] = xv.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int r = 0; r < MR; ++r) {
                    a0[r] = fhfma<false, false, false>(h[8 * q + 2 * j], xr[r][j], a0[r]);
                    a1[r] = fhfma<false, false, true>(h[8 * q + 2 * j + 1], xr[r][j], a1[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            yacc[m + r] = fmaf(s, a0[r] + a1[r], yacc[m + r]);
            yacc[m + r] = fmaf(c, xsum[(m + r) * XB], yacc[m + r]);
        }
    }
}

__device__ __forceinline__ void ex_group_sync(int id, int threads) {
    if (threads == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int MB>
__global__ void __launch_bounds__(EX_WARPS * 32, MB <= 2 ? 3 : 2) exl2_gemv_kernel(const Exl2Params p) {
    constexpr int WN = ExCfg<MB>::WN, WK = ExCfg<MB>::WK, XB = ExCfg<MB>::XB, D = ExCfg<MB>::D, CH = XB * 32;
    constexpr int COLS = WN * 32;
    extern __shared__ __align__(16) unsigned char ex_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wn = warp % WN, wk = warp / WN;
    const int gl = wn * 32 + lane;                       // thread index inside the k-slice group
    const int n = blockIdx.x * COLS + gl;
    const bool col_ok = n < p.N;
    const int m0 = blockIdx.y * MB;
    // per k-slice: x chunk [MB][CH] halfs, block sums [MB][XB] f32; then red [WK][MB][COLS] f32 (WK > 1), cta_sum [MB][COLS]
    __half* xs_w = reinterpret_cast<__half*>(ex_smem) + size_t(wk) * MB * CH;
    float* xsum_w = reinterpret_cast<float*>(ex_smem + size_t(WK) * MB * CH * sizeof(__half)) + wk * MB * XB;
    float* red = reinterpret_cast<float*>(ex_smem + size_t(WK) * MB * (CH * sizeof(__half) + XB * sizeof(float)));
    float* cta_sum = red + (WK > 1 ? WK * MB * COLS : 0);

    float yacc[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) yacc[m] = 0.f;

    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = int(cluster.block_rank()), csize = int(cluster.num_blocks());
    const int blocks_total = p.K / 32;
    const int b_lo = min(blocks_total, (crank * WK + wk) * p.blocks_per_slice);
    const int b_hi = min(blocks_total, b_lo + p.blocks_per_slice);

    // activations of blocks [b, b + nblk) of this k-slice (gathered through q_perm) + their per-block sums; executed by
    // all WN warps of the slice
    auto stage_x = [&](int b, int nblk) {
        ex_group_sync(1 + wk, COLS);
        for (int kl = gl; kl < nblk * 32; kl += COLS) {
            const int k = b * 32 + kl;
            const int src = p.perm ? int(__ldg(p.perm + k)) : k;
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                __half v = __float2half(0.f);
                if (m0 + m < p.M) v = p.x[size_t(m0 + m) * p.K + src];
                xs_w[m * CH + kl] = v;
            }
        }
        ex_group_sync(1 + wk, COLS);
        for (int i = gl; i < MB * nblk; i += COLS) {
            const int m = i / nblk, bb = i % nblk;
            float sum = 0.f;
            const __half2* src = reinterpret_cast<const __half2*>(xs_w + m * CH + bb * 32);
#pragma unroll
            for (int q = 0; q < 16; ++q) { const float2 f = __half22float2(src[q]); sum += f.x + f.y; }
            xsum_w[m * XB + bb] = sum;
        }
        ex_group_sync(1 + wk, COLS);
    };

    // one section (constant bit width B) of the slice: software-pipelined, D blocks of packed words in flight
    auto run = [&](auto btag, int s_lo, int s_hi, int k_sec, int prow0) {
        constexpr int B = decltype(btag)::value;
        uint32_t w[D][B];
        unsigned short sr[D], zr[D];
        const uint32_t* ptr = p.qw + size_t(prow0 + ((s_lo * 32 - k_sec) >> 5) * B) * p.N + n;
        const size_t bstride = size_t(B) * p.N;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            sr[d] = zr[d] = 0;
#pragma unroll
            for (int i = 0; i < B; ++i) w[d][i] = 0u;
            if (s_lo + d < s_hi) ex_fetch<B>(p, ptr + d * bstride, (s_lo + d) * 32, n, col_ok, w[d], sr[d], zr[d]);
        }
        for (int base = s_lo; base < s_hi; base += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int b = base + d;
                if (b < s_hi) {
                    const int bl = (b - b_lo) % XB;
                    if (bl == 0) stage_x(b, min(XB, b_hi - b));
                    uint32_t cw[B];
#pragma unroll
                    for (int i = 0; i < B; ++i) cw[i] = w[d][i];
                    const unsigned short cs = sr[d], cz = zr[d];
                    if (b + D < s_hi)      // refill the slot: in flight during the FMAs below
                        ex_fetch<B>(p, ptr + size_t(b + D - s_lo) * bstride, (b + D) * 32, n, col_ok, w[d], sr[d], zr[d]);
                    ex_block<B, MB, CH, XB>(cw, cs, cz, xs_w + bl * 32, xsum_w + bl, yacc);
                }
            }
        }
    };
    {
        int k_sec = 0;
#pragma unroll
        for (int sec = 0; sec < 6; ++sec) {          // static indices: the section table stays in the constant bank
            const int s_lo = max(b_lo, k_sec >> 5), s_hi = min(b_hi, p.sec.end[sec] >> 5);
            if (s_lo < s_hi) {
                if (sec == 0) run(std::integral_constant<int, 8>{}, s_lo, s_hi, k_sec, p.sec.prow[sec]);
                else if (sec == 1) run(std::integral_constant<int, 6>{}, s_lo, s_hi, k_sec, p.sec.prow[sec]);
                else if (sec == 2) run(std::integral_constant<int, 5>{}, s_lo, s_hi, k_sec, p.sec.prow[sec]);
                else if (sec == 3) run(std::integral_constant<int, 4>{}, s_lo, s_hi, k_sec, p.sec.prow[sec]);
                else if (sec == 4) run(std::integral_constant<int, 3>{}, s_lo, s_hi, k_sec, p.sec.prow[sec]);
                else run(std::integral_constant<int, 2>{}, s_lo, s_hi, k_sec, p.sec.prow[sec]);
            }
            k_sec = p.sec.end[sec];
        }
    }
    // ---- the k-slices of the CTA meet in shared memory, then the CTAs of the cluster in rank order ----
    if constexpr (WK > 1) {
#pragma unroll
        for (int m = 0; m < MB; ++m) red[(wk * MB + m) * COLS + gl] = yacc[m];
        __syncthreads();
        for (int i = threadIdx.x; i < MB * COLS; i += EX_WARPS * 32) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < WK; ++w) t += red[w * MB * COLS + i];
            cta_sum[i] = t;
        }
    } else {
#pragma unroll
        for (int m = 0; m < MB; ++m) cta_sum[m * COLS + gl] = yacc[m];
    }
    cluster.sync();
    if (crank == 0) {
        for (int i = threadIdx.x; i < MB * COLS; i += EX_WARPS * 32) {
            const int m = i / COLS, l = i % COLS;
            float t = cta_sum[i];
            for (int r = 1; r < csize; ++r) t += cluster.map_shared_rank(cta_sum, r)[i];
            const int col = blockIdx.x * COLS + l;
            if (col < p.N && m0 + m < p.M) p.y[size_t(m0 + m) * p.N + col] = __float2half_rn(t);
        }
    }
    cluster.sync();          // the other CTAs' shared memory stays alive until rank 0 has read it
}

template <int MB>
static int launch_exl2(Exl2Params p, cudaStream_t st) {
    constexpr int WN = ExCfg<MB>::WN, WK = ExCfg<MB>::WK, XB = ExCfg<MB>::XB, COLS = WN * 32;
    const size_t smem = size_t(WK) * MB * (XB * 32 * sizeof(__half) + XB * sizeof(float)) +
                        size_t((WK > 1 ? WK : 0) + 1) * MB * COLS * sizeof(float);
    auto kern = exl2_gemv_kernel<MB>;
    if (smem > 48 * 1024) B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    // k-slices: WK per CTA x cluster size -- as many as still fit ONE wave of CTAs (a second wave doubles the latency-bound
    // run time: measured 11.3 vs 14.7 us at 4096^2, M = 1), at least 4 blocks of 32 rows per slice
    const int blocks = p.K / 32, strips = (p.N + COLS - 1) / COLS, mchunks = (p.M + MB - 1) / MB;
    int csize = 1;
    const int max_c = WK > 2 ? 4 : 8;
    while (csize < max_c && strips * mchunks * csize * 2 <= (12 * sm_count()) / 5 && blocks >= 2 * WK * csize * 2) csize *= 2;
    static const int force_c = getenv("B200BIT_EXL2_CSIZE") ? atoi(getenv("B200BIT_EXL2_CSIZE")) : 0;     // sweep hook
    if (force_c > 0 && force_c <= 8 && blocks >= WK * force_c) csize = force_c;
    p.blocks_per_slice = (blocks + WK * csize - 1) / (WK * csize);
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

// ---------------------------------------------------------------------------------------------------------------
// Decode flavour (M <= 4) on the TMA ring of the n-bit kernels: one producer warp streams the strip's packed rows --
// tiles of eight 32-row blocks of ONE bit-width section, i.e. 8 x B packed rows x 128 bytes, plus the eight group rows of
// scales / zeros they can touch -- into an 8-stage shared-memory ring (cp.async.bulk.tensor.2d, mbarrier completion);
// the eight consumer warps take one block of the tile each (lane = column: conflict-free 128-byte shared-memory rows)
// and never wait for a global load.  One tensor map per bit width (box = 8 x B rows).  k-slices over the cluster as above.
// ---------------------------------------------------------------------------------------------------------------
constexpr int ET_WARPS = 8;                 // consumer warps = blocks per tile
constexpr int ET_STAGES = 6;
constexpr int ET_W_BYTES = ET_WARPS * 8 * 128;          // widest tile (8-bit): 64 packed rows of 32 words
constexpr int ET_SZ_BYTES = ET_WARPS * 32 * 2;          // eight group rows of 32 halfs
constexpr int ET_SLOT_BYTES = ET_W_BYTES + 2 * ET_SZ_BYTES;

struct Exl2Maps { CUtensorMap w[6]; CUtensorMap s, z; };

__device__ __forceinline__ void et_wait(uint64_t* bar, unsigned parity) {       // bounded: a protocol bug traps, never hangs
    unsigned tries = 0;
    while (true) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (++tries > (1u << 22)) asm volatile("trap;");
    }
}

template <int B, int MB>
__device__ __forceinline__ void et_block(const uint32_t* wrow, int lane, unsigned short s_raw, unsigned short z_raw,
                                         const __half* xs, int xstride, const float* xsum, int sstride, float (&yacc)[MB]) {
    uint32_t w[B];
#pragma unroll
    for (int i = 0; i < B; ++i) w[i] = wrow[i * 32 + lane];
    uint32_t h[32];
    ex_unpack<B>(w, h);
    const float s = __half2float(__ushort_as_half(s_raw));
    const float z = __half2float(__ushort_as_half(z_raw));
    const float c = -(1024.f * s + z);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
        float a0 = 0.f, a1 = 0.f;
        const uint4* xm = reinterpret_cast<const uint4*>(xs + m * xstride);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 xv = xm[q];
            a0 = fhfma<false, false, false>(h[8 * q + 0], xv.x, a0);
            a1 = fhfma<false, false, true>(h[8 * q + 1], xv.x, a1);
            a0 = fhfma<false, false, false>(h[8 * q + 2], xv.y, a0);
            a1 = fhfma<false, false, true>(h[8 * q + 3], xv.y, a1);
            a0 = fhfma<false, false, false>(h[8 * q + 4], xv.z, a0);
            a1 = fhfma<false, false, true>(h[8 * q + 5], xv.z, a1);
            a0 = fhfma<false, false, false>(h[8 * q + 6], xv.w, a0);
            a1 = fhfma<false, false, true>(h[8 * q + 7], xv.w, a1);
        }
        yacc[m] = fmaf(s, a0 + a1, yacc[m]);
        yacc[m] = fmaf(c, xsum[m * sstride], yacc[m]);
    }
}

template <int MB>
__global__ void __launch_bounds__((ET_WARPS + 1) * 32) exl2_tma_kernel(const __grid_constant__ Exl2Maps maps, const Exl2Params p) {
    extern __shared__ __align__(1024) unsigned char et_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n0 = blockIdx.x * 32;
    const int m0 = blockIdx.y * MB;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = int(cluster.block_rank()), csize = int(cluster.num_blocks());
    const int blocks_total = p.K / 32;
    const int b_lo = min(blocks_total, crank * p.blocks_per_slice);
    const int b_hi = min(blocks_total, b_lo + p.blocks_per_slice);
    const int xlen = p.blocks_per_slice * 32;
    // carve-up: barriers (128 B) | ring | x [MB][xlen] halfs | red [8][MB][32] f32 | cta_sum [MB][32] | xsum [MB][bps] f32 |
    //           group index of every block of the slice [bps] u16
    uint64_t* full = reinterpret_cast<uint64_t*>(et_smem);
    uint64_t* empty = full + ET_STAGES;
    unsigned char* ring = et_smem + 128;
    __half* xs = reinterpret_cast<__half*>(ring + ET_STAGES * ET_SLOT_BYTES);
    float* red = reinterpret_cast<float*>(xs + size_t(MB) * xlen);
    float* cta_sum = red + ET_WARPS * MB * 32;
    float* xsum = cta_sum + MB * 32;
    unsigned short* gidx = reinterpret_cast<unsigned short*>(xsum + MB * p.blocks_per_slice);
    if (threadIdx.x == 0) {
        for (int i = 0; i < ET_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], ET_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // the CTA's tile sequence: per section the runs of up to eight blocks inside [b_lo, b_hi); f(sec, t0, nb, prow)
    auto for_tiles = [&](auto&& f) {
        int k_sec = 0;
#pragma unroll
        for (int sec = 0; sec < 6; ++sec) {
            const int bits = sec == 0 ? 8 : 7 - sec;
            const int s_lo = max(b_lo, k_sec >> 5), s_hi = min(b_hi, p.sec.end[sec] >> 5);
            for (int t0 = s_lo; t0 < s_hi; t0 += ET_WARPS)
                f(sec, t0, min(ET_WARPS, s_hi - t0), p.sec.prow[sec] + (t0 - (k_sec >> 5)) * bits);
            k_sec = p.sec.end[sec];
        }
    };

    if (warp == ET_WARPS) {
        // ===== producer warp =====
        const uint32_t leader = um_elect();
        // first group row of every tile: one load per lane, all issued before the first TMA (a load per tile in the loop
        // below would put an L2 round trip between any two tiles)
        int g0r[2] = {0, 0};
        {
            int ti = 0;
            for_tiles([&](int, int t0, int, int) {
                if (ti < 64 && (ti & 31) == lane) {
                    const int v = int(__ldg(p.gmap + 2 * (t0 * 32)));
                    if (ti < 32) g0r[0] = v; else g0r[1] = v;
                }
                ++ti;
            });
        }
        int slot = 0, ti = 0;
        unsigned eph = 1u;
        for_tiles([&](int sec, int t0, int /*nb*/, int prow) {
            const int bits = sec == 0 ? 8 : 7 - sec;
            int g0;
            if (ti < 64) g0 = __shfl_sync(0xffffffffu, ti < 32 ? g0r[0] : g0r[1], ti & 31);
            else g0 = int(__ldg(p.gmap + 2 * (t0 * 32)));
            ++ti;
            et_wait(&empty[slot], eph);
            unsigned char* dst = ring + size_t(slot) * ET_SLOT_BYTES;
            um_expect_tx(&full[slot], unsigned(ET_WARPS * bits * 128 + 2 * ET_SZ_BYTES), leader);
            um_tma_2d(dst, &maps.w[sec], n0, prow, &full[slot], leader);
            um_tma_2d(dst + ET_W_BYTES, &maps.s, n0, g0, &full[slot], leader);
            um_tma_2d(dst + ET_W_BYTES + ET_SZ_BYTES, &maps.z, n0, g0, &full[slot], leader);
            if (++slot == ET_STAGES) { slot = 0; eph ^= 1u; }
        });
    } else {
        // ===== eight consumer warps =====
        float yacc[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m) yacc[m] = 0.f;
        // the slice's activations, gathered through q_perm (the weight stream is already running)
        const int tid = threadIdx.x, nrows = (b_hi - b_lo) * 32;
        for (int kl = tid; kl < nrows; kl += ET_WARPS * 32) {
            const int k = b_lo * 32 + kl;
            const int src = p.perm ? int(__ldg(p.perm + k)) : k;
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                __half v = __float2half(0.f);
                if (m0 + m < p.M) v = p.x[size_t(m0 + m) * p.K + src];
                xs[m * xlen + kl] = v;
            }
        }
        for (int bb = tid; bb < b_hi - b_lo; bb += ET_WARPS * 32) gidx[bb] = __ldg(p.gmap + 2 * ((b_lo + bb) * 32));
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = tid; i < MB * (b_hi - b_lo); i += ET_WARPS * 32) {
            const int m = i / (b_hi - b_lo), bb = i % (b_hi - b_lo);
            float sum = 0.f;
            const __half2* src = reinterpret_cast<const __half2*>(xs + m * xlen + bb * 32);
#pragma unroll
            for (int q = 0; q < 16; ++q) { const float2 f = __half22float2(src[q]); sum += f.x + f.y; }
            xsum[m * p.blocks_per_slice + bb] = sum;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");

        int slot = 0;
        unsigned ph = 0u;
        for_tiles([&](int sec, int t0, int nb, int /*prow*/) {
            et_wait(&full[slot], ph);
            if (warp < nb) {
                const int b = t0 + warp;
                const int gi = int(gidx[b - b_lo]) - int(gidx[t0 - b_lo]);
                const unsigned char* src = ring + size_t(slot) * ET_SLOT_BYTES;
                const unsigned short s_raw = reinterpret_cast<const unsigned short*>(src + ET_W_BYTES)[gi * 32 + lane];
                const unsigned short z_raw = reinterpret_cast<const unsigned short*>(src + ET_W_BYTES + ET_SZ_BYTES)[gi * 32 + lane];
                const uint32_t* wsm = reinterpret_cast<const uint32_t*>(src);
                const __half* xb = xs + (b - b_lo) * 32;
                const float* xsb = xsum + (b - b_lo);
                switch (sec) {
                    case 0: et_block<8, MB>(wsm + warp * 8 * 32, lane, s_raw, z_raw, xb, xlen, xsb, p.blocks_per_slice, yacc); break;
                    case 1: et_block<6, MB>(wsm + warp * 6 * 32, lane, s_raw, z_raw, xb, xlen, xsb, p.blocks_per_slice, yacc); break;
                    case 2: et_block<5, MB>(wsm + warp * 5 * 32, lane, s_raw, z_raw, xb, xlen, xsb, p.blocks_per_slice, yacc); break;
                    case 3: et_block<4, MB>(wsm + warp * 4 * 32, lane, s_raw, z_raw, xb, xlen, xsb, p.blocks_per_slice, yacc); break;
                    case 4: et_block<3, MB>(wsm + warp * 3 * 32, lane, s_raw, z_raw, xb, xlen, xsb, p.blocks_per_slice, yacc); break;
                    default: et_block<2, MB>(wsm + warp * 2 * 32, lane, s_raw, z_raw, xb, xlen, xsb, p.blocks_per_slice, yacc); break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            if (++slot == ET_STAGES) { slot = 0; ph ^= 1u; }
        });
#pragma unroll
        for (int m = 0; m < MB; ++m) red[(warp * MB + m) * 32 + lane] = yacc[m];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = tid; i < MB * 32; i += ET_WARPS * 32) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < ET_WARPS; ++w) t += red[w * MB * 32 + i];
            cta_sum[i] = t;
        }
    }
    cluster.sync();
    if (crank == 0 && warp < ET_WARPS) {
        for (int i = threadIdx.x; i < MB * 32; i += ET_WARPS * 32) {
            const int m = i / 32, l = i % 32;
            float t = cta_sum[i];
            for (int r = 1; r < csize; ++r) t += cluster.map_shared_rank(cta_sum, r)[i];
            const int col = n0 + l;
            if (col < p.N && m0 + m < p.M) p.y[size_t(m0 + m) * p.N + col] = __float2half_rn(t);
        }
    }
    cluster.sync();
}

// tensor maps of a weight (six bit-width boxes over the packed matrix, scales, zeros): built once per (pointers, shape)
struct Exl2MapKey { const void* q; const void* s; const void* z; int N, rows, G; };
static int exl2_maps(const Exl2MapKey& key, Exl2Maps** out) {
    constexpr int CACHE = 256;
    static Exl2MapKey keys[CACHE];
    static Exl2Maps vals[CACHE];
    static int used = 0, next = 0;
    for (int i = 0; i < used; ++i)
        if (keys[i].q == key.q && keys[i].s == key.s && keys[i].z == key.z && keys[i].N == key.N && keys[i].rows == key.rows &&
            keys[i].G == key.G) { *out = &vals[i]; return B200BIT_OK; }
    const int at = used < CACHE ? used++ : (next = (next + 1) % CACHE);
    Exl2Maps& m = vals[at];
    const int widths[6] = {8, 6, 5, 4, 3, 2};
    for (int i = 0; i < 6; ++i) {
        const int rc = make_map_2d(&m.w[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, key.q, uint64_t(key.N), uint64_t(key.rows),
                                   uint64_t(key.N) * 4, 32, uint32_t(ET_WARPS * widths[i]), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
    }
    int rc = make_map_2d(&m.s, CU_TENSOR_MAP_DATA_TYPE_UINT16, key.s, uint64_t(key.N), uint64_t(key.G), uint64_t(key.N) * 2, 32,
                         ET_WARPS, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != B200BIT_OK) return rc;
    rc = make_map_2d(&m.z, CU_TENSOR_MAP_DATA_TYPE_UINT16, key.z, uint64_t(key.N), uint64_t(key.G), uint64_t(key.N) * 2, 32,
                     ET_WARPS, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != B200BIT_OK) return rc;
    keys[at] = key;
    *out = &m;
    return B200BIT_OK;
}

template <int MB>
static int launch_exl2_tma(Exl2Params p, const Exl2Maps& maps, cudaStream_t st) {
    const int blocks = p.K / 32, strips = (p.N + 31) / 32, mchunks = (p.M + MB - 1) / MB;
    int csize = 1;
    while (csize < 8 && strips * mchunks * csize * 2 <= (12 * sm_count()) / 5 && blocks >= csize * 2 * 16) csize *= 2;
    p.blocks_per_slice = (blocks + csize - 1) / csize;
    const size_t smem = 128 + size_t(ET_STAGES) * ET_SLOT_BYTES + size_t(MB) * p.blocks_per_slice * 32 * sizeof(__half) +
                        size_t(MB) * p.blocks_per_slice * sizeof(float) + size_t(ET_WARPS + 1) * MB * 32 * sizeof(float) +
                        size_t(p.blocks_per_slice) * sizeof(unsigned short) + 64;
    if (smem > 200 * 1024) return B200BIT_ERR_UNSUPPORTED;
    auto kern = exl2_tma_kernel<MB>;
    static bool configured[64] = {false};
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(strips, mchunks, csize);
    cfg.blockDim = dim3((ET_WARPS + 1) * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = csize;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, p));
    return B200BIT_OK;
}

}  // namespace b200bit

using namespace b200bit;

extern "C" int b200bit_exl2_forward(const void* x, const int32_t* qweight, const void* scales, const void* zeros,
                                    const int16_t* perm, const int16_t* q_group_map, void* y, int M, int K, int N, int G,
                                    const int* rows6, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(x && qweight && scales && zeros && q_group_map && y && rows6, B200BIT_ERR_ARG, "exl2_forward: null pointer argument");
    B200_REQUIRE(G > 0 && G <= K / 32, B200BIT_ERR_SHAPE, "exl2_forward: G=%d groups for K=%d", G, K);
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
    static const int use_tma = getenv("B200BIT_EXL2_TMA") ? atoi(getenv("B200BIT_EXL2_TMA")) : 1;     // sweep hook
    if (use_tma && M <= 4 && N % 8 == 0 && (reinterpret_cast<uintptr_t>(qweight) | reinterpret_cast<uintptr_t>(scales) |
                                           reinterpret_cast<uintptr_t>(zeros)) % 16 == 0) {
        Exl2Maps* maps = nullptr;
        const Exl2MapKey key{qweight, scales, zeros, N, prow, G};
        const int rc = exl2_maps(key, &maps);
        if (rc != B200BIT_OK) return rc;
        int r2;
        if (M == 1) r2 = launch_exl2_tma<1>(p, *maps, st);
        else if (M == 2) r2 = launch_exl2_tma<2>(p, *maps, st);
        else r2 = launch_exl2_tma<4>(p, *maps, st);
        if (r2 != B200BIT_ERR_UNSUPPORTED) return r2;
    }
    if (M == 1) return launch_exl2<1>(p, st);
    if (M == 2) return launch_exl2<2>(p, st);
    if (M <= 4) return launch_exl2<4>(p, st);
    if (M <= 8) return launch_exl2<8>(p, st);
    if (M <= 16) return launch_exl2<16>(p, st);
    return launch_exl2<32>(p, st);
}
