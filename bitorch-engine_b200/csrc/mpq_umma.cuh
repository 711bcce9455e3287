// mpq_umma.cuh -- tcgen05 / TMEM batched kernel (fp16 activations, 4-bit weights, up to 32 batch rows per pass), sm_100a.
//
// The decode kernels (mpq_gemv / mpq_stream) spend their compute on CUDA-core FHFMA or legacy mma.sync; both saturate near
// 1 us per 4096x4096 layer and scale linearly with the batch.  This kernel moves the multiply-accumulates to the
// 5th-generation tensor cores, where a 128 x N x 16 MMA costs max(13, N/2) cycles (tools/umma_rate.cu) and N is free up
// to the batch size, and leaves the CUDA cores only the bit-ops that turn a packed word into an operand:
//
//   * persistent CTA per SM owning whole 32-column strips (as mpq_stream.cuh); TMA streams 32x32-word weight tiles + scale
//     / zero rows into a shared-memory ring, starting BEFORE griddepcontrol.wait (PDL prefetch);
//   * split-K inside the MMA: the 128 TMEM lanes of one UMMA are (K-quarter s, column n) = lane 32*s + n; the B operand's N
//     dimension is (K-quarter s', batch row m).  Only the diagonal blocks s == s' are read back; the 4x redundant MACs are
//     free at N <= 128 and buy a 4-way K split with no cross-CTA reduction, so 128 strips x 148 SMs stay balanced;
//   * A operand in TMEM, written with tcgen05.st by 16 dequant warps: a masked packed word IS two fp16 K elements (fp16
//     subnormals q * 2^-24, honoured exactly by UTCHMMA -- tools/umma_probe.cu).  Fields at bit 4 carry 16*q; the B rows
//     they meet hold x/16, so both field classes accumulate into ONE accumulator.  The K permutation this implies
//     (codes 0,4,2,6 | 1,5,3,7 of every word) is applied to x once per call by umma_prepare_kernel, which writes the
//     "B image": x in canonical K-major core-matrix order, one contiguous block per step, plus the per-group sums of x;
//   * the B image streams through its own 2-stage ring with 1-D bulk copies (L2-resident: every CTA reads the same image);
//   * one elected lane of a converged warp issues tcgen05.mma.kind::f16 (M=128, N=4*MB, K=16); accumulators live in TMEM,
//     double buffered per quantisation group; tcgen05.commit -> mbarriers hand A / B / D buffers back;
//   * group affine factored out in fp32 at flush:  y += s * 2^24 * D - z * sum_g(x)   (tcgen05.ld of the thread's own lane).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include "common.cuh"
#include "mpq_mma.cuh"
#include "mpq_stream.cuh"   // mbarrier / TMA helpers, um_elect / um_expect_tx / um_tma_2d

namespace b200bit {

constexpr int UM_DW = 16;           // dequant warps: warp & 3 = TMEM lane quarter = K-quarter, warp >> 2 = sub-step it owns
constexpr int UM_THREADS = (UM_DW + 2) * 32;
constexpr int UM_TILE = 4096;       // 32 rows x 128 B
constexpr int UM_SZ = 512;          // scales (256) + zeros (256) per (K-quarter, stage)
constexpr int UM_BSTAGES = 2;

struct UmmaParams {
    const unsigned char* bimg;   // B image [steps][4 q][2 class][2 u][N_mma x 16 halves, canonical K-major tile]
    const float* xsum;           // [MB][4][steps * GPS] per-group sums of x
    uint16_t* y;                 // [M, N] f16
    int M, K, N;
    int strips;                  // N / 32
    int rps;                     // runs (32 packed rows) per strip
    int steps;                   // ceil(rps / 4): every K-quarter advances one run per step
    int ngr;                     // scale rows per run: 2 (g128), 4 (g64), 1 (group >= 256 k)
    int rpr, rpr_shift;          // runs per group when a group spans >= 1 run
    int asym;
    int S;                       // weight ring stages (steps in flight)
    unsigned long long* trace;
};

struct UmmaPrepParams {
    const uint16_t* x;           // [M, K] f16
    unsigned char* bimg;
    float* xsum;
    int M, K, MB, rps, steps, fj2;
    int cell_blocks;             // blocks [0, cell_blocks) write image cells, the rest compute xsum
};

__device__ __forceinline__ void um_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void um_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void um_commit(uint64_t* bar, uint32_t leader) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void um_mma(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc),
                 "r"(accumulate), "r"(leader) : "memory");
}
__device__ __forceinline__ void um_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint32_t leader) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                 "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void um_st16(uint32_t addr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
template <int MB> __device__ __forceinline__ void um_ld(uint32_t addr, float (&v)[MB]);
template <> __device__ __forceinline__ void um_ld<4>(uint32_t addr, float (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
template <> __device__ __forceinline__ void um_ld<8>(uint32_t addr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(addr));
}
template <> __device__ __forceinline__ void um_ld<16>(uint32_t addr, float (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                   "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "r"(addr));
}
template <> __device__ __forceinline__ void um_ld<32>(uint32_t addr, float (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                   "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
                   "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
                   "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31]) : "r"(addr));
}

#ifdef B200_UMMA_WATCHDOG
__device__ int um_prog[32];
#define UM_PROG(v_) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) um_prog[threadIdx.x >> 5] = (v_); } while (0)
__device__ __forceinline__ void um_wait(uint64_t* bar, unsigned parity, int id) {
    for (long long spin = 0;; ++spin) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (spin == (1ll << 21) && blockIdx.x == 0 && (threadIdx.x & 31) == 0)
            printf("umma watchdog: cta %d warp %d wait id %d parity %u | prog %d %d %d %d  %d %d %d %d  %d %d %d %d  %d %d %d %d  p%d m%d\n", blockIdx.x, threadIdx.x >> 5, id, parity,
                   um_prog[0], um_prog[1], um_prog[2], um_prog[3], um_prog[4], um_prog[5], um_prog[6], um_prog[7], um_prog[8], um_prog[9], um_prog[10],
                   um_prog[11], um_prog[12], um_prog[13], um_prog[14], um_prog[15], um_prog[16], um_prog[17]);
        if (spin > (1ll << 23)) __trap();
    }
}
#else
__device__ __forceinline__ void um_wait(uint64_t* bar, unsigned parity, int) { mbar_wait(bar, parity); }
#define UM_PROG(v_) do { } while (0)
#endif

// first run and run count of K-quarter s
__host__ __device__ __forceinline__ int um_slice_r0(int s, int rps) { return s * (rps >> 2) + (s < (rps & 3) ? s : (rps & 3)); }
__host__ __device__ __forceinline__ int um_slice_runs(int s, int rps) { return (rps >> 2) + (s < (rps & 3) ? 1 : 0); }

// byte offset of B element (row j, k index kk in [0,16)) inside one N_mma x 16 canonical K-major tile:
// core matrix = 8 rows x 16 B contiguous; K-adjacent cores 128 B apart (LBO), 8-row groups 256 B apart (SBO)
__host__ __device__ __forceinline__ int um_b_off(int j, int kk) { return (j >> 3) * 256 + (kk >> 3) * 128 + (j & 7) * 16 + (kk & 7) * 2; }

// ---------------------------------------------------------------------------------------------------------------------
// prepare: x [M, K] -> B image + per-group sums.  One thread per 8-byte image cell (4 K elements of one row of one tile).
// ---------------------------------------------------------------------------------------------------------------------
template <int UNUSED>   // (template only so the definition can live in this header)
__global__ void __launch_bounds__(256) umma_prepare_kernel(const UmmaPrepParams p) {
    pdl_launch_dependents();
    pdl_wait_primary();                       // x is produced by the previous kernel in the stream
    const int nmma = 4 * p.MB;
    if (int(blockIdx.x) < p.cell_blocks) {
        // cell index -> (step, q, class, u, row j, word-in-kstep wi)
        const int cells_per_step = 4 * 2 * 2 * nmma * 4;
        const int cell = blockIdx.x * 256 + threadIdx.x;
        if (cell >= p.steps * cells_per_step) return;
        const int pstep = cell / cells_per_step;
        int r = cell - pstep * cells_per_step;
        const int wi = r & 3; r >>= 2;
        const int j = r % nmma; r /= nmma;
        const int u = r & 1, c = (r >> 1) & 1, q = r >> 2;
        const int s = j / p.MB, m = j - s * p.MB;
        uint2 out = make_uint2(0u, 0u);
        if (m < p.M && pstep < um_slice_runs(s, p.rps)) {
            const int row = (um_slice_r0(s, p.rps) + pstep) * 32 + q * 8 + u * 4 + wi;       // packed row = 8 K elements
            const uint4 v = *reinterpret_cast<const uint4*>(p.x + size_t(m) * p.K + size_t(row) * 8);
            if (c == 0) {          // codes (0,4,2,6)
                out.x = __byte_perm(v.x, v.z, 0x5410); out.y = __byte_perm(v.y, v.w, 0x5410);
            } else {               // codes (1,5,3,7), scaled by 1/16: their A fields carry 16 * q
                const __half2 sc = __float2half2_rn(0.0625f);
                uint32_t ta = __byte_perm(v.x, v.z, 0x7632), tb = __byte_perm(v.y, v.w, 0x7632);
                const __half2 ha = __hmul2(*reinterpret_cast<__half2*>(&ta), sc);
                const __half2 hb = __hmul2(*reinterpret_cast<__half2*>(&tb), sc);
                out.x = *reinterpret_cast<const uint32_t*>(&ha); out.y = *reinterpret_cast<const uint32_t*>(&hb);
            }
        }
        unsigned char* dst = p.bimg + size_t(pstep) * (size_t(nmma) * 512) + size_t((q * 2 + c) * 2 + u) * (nmma * 32) + um_b_off(j, 4 * wi);
        *reinterpret_cast<uint2*>(dst) = out;
    } else {
        // one warp per (m, K-quarter, step): sums of x over segments of 8 * fj2 packed rows
        const int wid = (blockIdx.x - p.cell_blocks) * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
        const int total = p.MB * 4 * p.steps;
        if (wid >= total) return;
        const int pstep = wid % p.steps, s = (wid / p.steps) & 3, m = wid / (4 * p.steps);
        const int seg_rows = 8 * p.fj2, gps = 4 / p.fj2;
        float sum = 0.f;
        if (m < p.M && pstep < um_slice_runs(s, p.rps)) {
            const int row = (um_slice_r0(s, p.rps) + pstep) * 32 + lane;
            const uint4 v = *reinterpret_cast<const uint4*>(p.x + size_t(m) * p.K + size_t(row) * 8);
            sum = fhfma<false, false, false>(0x3C003C00u, v.x, sum); sum = fhfma<false, true, true>(0x3C003C00u, v.x, sum);
            sum = fhfma<false, false, false>(0x3C003C00u, v.y, sum); sum = fhfma<false, true, true>(0x3C003C00u, v.y, sum);
            sum = fhfma<false, false, false>(0x3C003C00u, v.z, sum); sum = fhfma<false, true, true>(0x3C003C00u, v.z, sum);
            sum = fhfma<false, false, false>(0x3C003C00u, v.w, sum); sum = fhfma<false, true, true>(0x3C003C00u, v.w, sum);
        }
        for (int off = 1; off < seg_rows; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if ((lane % seg_rows) == 0) p.xsum[(m * 4 + s) * (p.steps * gps) + pstep * gps + lane / seg_rows] = sum;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// main kernel.  FJ2 = sub-steps (8 packed rows) per quantisation group: 1 (g64), 2 (g128), 4 (group >= 256 k: flush per
// step);  MB = batch slots per K-quarter (4 / 8 / 16 / 32), UMMA N = 4 * MB.
// ---------------------------------------------------------------------------------------------------------------------
template <int FJ2, int MB>
__global__ void __launch_bounds__(UM_THREADS, 1) mpq_umma_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                                 const __grid_constant__ CUtensorMap tm_s,
                                                                 const __grid_constant__ CUtensorMap tm_z, const UmmaParams p) {
    constexpr int BITS = 4, NB = 8;
    constexpr int GPS = 4 / FJ2;                   // groups (flushes) per step
    constexpr int NMMA = 4 * MB;
    constexpr int BSTEP = NMMA * 512;              // B image bytes per step
    constexpr int TMEM_COLS = 512;                 // D: 2 x NMMA (<= 256) | A: 2 steps x 4 sub-steps x 32 = 256
    constexpr int NDB = GPS > 2 ? GPS : 2;         // accumulator barriers: one per group slot of a step, so every barrier
                                                   // is always waited on by the same warp group in order (no parity aliasing)
    constexpr int NFL = GPS;                       // flusher warp groups: FJ2=1 -> WG 0..3, FJ2=2 -> WG 0,2, FJ2=4 -> WG 0
    extern __shared__ unsigned char um_smem_raw[];
    unsigned char* smem_raw = um_smem_raw + ((128u - (smem_u32(um_smem_raw) & 127u)) & 127u);   // TMA destinations: 128 B
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, G = gridDim.x, S = p.S;
#undef ST_TRACE
#define ST_TRACE(slot_) do { if (p.trace && lane == 0 && (warp < 14 || warp >= UM_DW)) \
        p.trace[(size_t(blockIdx.x) * 16 + (warp >= UM_DW ? warp - 2 : warp)) * 8 + (slot_)] = st_gtime(); } while (0)
    ST_TRACE(0);

    // ---- shared memory ----
    unsigned char* bst = smem_raw;                                        // [UM_BSTAGES][BSTEP]
    unsigned char* wst = bst + UM_BSTAGES * BSTEP;                         // [S][4 K-quarters][4096]
    unsigned char* szst = wst + size_t(S) * 4 * UM_TILE;                   // [S][4][512]
    const int nseg = p.steps * GPS;
    float* xseg = reinterpret_cast<float*>(szst + size_t(S) * 4 * UM_SZ);  // [MB][4][nseg]
    float* part = xseg + ((MB * 4 * nseg + 3) & ~3);                       // [4][MB][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + 4 * MB * 32);
    uint64_t* wfull = bars;                    // [S][4]
    uint64_t* wempty = wfull + S * 4;          // [S][4]   count 4: one warp per warp group
    uint64_t* afull = wempty + S * 4;          // [2][4]   count 4: the 4 K-quarter warps of the warp group
    uint64_t* aempty = afull + 8;              // [2][4]   tcgen05.commit
    uint64_t* dfull = aempty + 8;              // [4]      tcgen05.commit
    uint64_t* dempty = dfull + 4;              // [4]      count 4
    uint64_t* bfull = dempty + 4;              // [UM_BSTAGES]
    uint64_t* bempty = bfull + UM_BSTAGES;     // [UM_BSTAGES] tcgen05.commit
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty + UM_BSTAGES);

    const int s_lo = int((long long)b * p.strips / G), s_hi = int((long long)(b + 1) * p.strips / G);
    const int nstrips = s_hi - s_lo;
    const int total_steps = nstrips * p.steps;

    if (tid < S * 4) { mbar_init(&wfull[tid], 1); mbar_init(&wempty[tid], 4); }
    if (tid < 8) { mbar_init(&afull[tid], 4); mbar_init(&aempty[tid], 1); }
    if (tid < 4) { mbar_init(&dfull[tid], 1); mbar_init(&dempty[tid], 4); }
    if (tid < UM_BSTAGES) { mbar_init(&bfull[tid], 1); mbar_init(&bempty[tid], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pdl_launch_dependents();
    __syncthreads();
    ST_TRACE(1);

    const unsigned tile_bytes = UM_TILE + unsigned(p.ngr) * 64u + (p.asym ? unsigned(p.ngr) * 16u : unsigned(p.ngr) * 64u);

    if (warp == UM_DW) {
        // =========================== producer (converged warp, elected lane issues) ===========================
        const uint32_t leader = um_elect();
        auto issue_w = [&](int t) {                       // weight tiles of global step t (all K-quarters that have a run)
            const int st = t / p.steps, pstep = t - st * p.steps;
            const int strip = s_lo + st;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int r0 = um_slice_r0(s, p.rps), nr = um_slice_runs(s, p.rps);
                if (pstep >= nr) continue;                 // padded step: no tile (the dequant warps write zeros)
                const int ts = st * nr + pstep;            // real steps of this K-quarter so far
                const int slot = ts % S;
                if (ts >= S) um_wait(&wempty[slot * 4 + s], ((ts / S) - 1) & 1, 1);
                const int kr = r0 + pstep;
                int g0;
                if (p.ngr > 1 || p.rpr == 1) g0 = kr * p.ngr;
                else g0 = (p.rpr_shift >= 0) ? (kr >> p.rpr_shift) : (kr / p.rpr);
                uint64_t* bar = &wfull[slot * 4 + s];
                um_expect_tx(bar, tile_bytes, leader);
                um_tma_2d(wst + (size_t(slot) * 4 + s) * UM_TILE, &tm_w, strip * 32, kr * 32, bar, leader);
                unsigned char* sz = szst + (size_t(slot) * 4 + s) * UM_SZ;
                um_tma_2d(sz, &tm_s, strip * 32, g0, bar, leader);
                um_tma_2d(sz + 256, &tm_z, p.asym ? strip * (32 / NB) : strip * 32, g0, bar, leader);
            }
        };
        const int ahead = S - 1 < total_steps ? S - 1 : total_steps;       // weight prefetch distance (ring holds S steps)
        for (int t = 0; t < ahead; ++t) issue_w(t);                         // weights do not depend on the previous kernel
        pdl_wait_primary();                                                 // the B image does
        ST_TRACE(2);
        for (int t = 0; t < total_steps; ++t) {
            const int bs = t % UM_BSTAGES;
            if (t >= UM_BSTAGES) um_wait(&bempty[bs], ((t / UM_BSTAGES) - 1) & 1, 2);
            um_expect_tx(&bfull[bs], BSTEP, leader);
            const int pstep = t % p.steps;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4)
                um_bulk_g2s(bst + bs * BSTEP + c4 * (BSTEP / 4), p.bimg + size_t(pstep) * BSTEP + c4 * (BSTEP / 4), BSTEP / 4,
                            &bfull[bs], leader);
            if (t < 4) ST_TRACE(3 + t);
            if (t + ahead < total_steps) issue_w(t + ahead);
        }
    } else if (warp == UM_DW + 1) {
        // =========================== MMA issuer (converged warp, elected lane issues) ===========================
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        um_fence_before();
        asm volatile("bar.sync 2, %0;" ::"n"((UM_DW + 1) * 32) : "memory");     // TMEM address visible to the dequant warps
        um_fence_after();
        const uint32_t tmem = *tmem_slot;
        const uint32_t tm_d = tmem, tm_a = tmem + 256;
        const uint32_t leader = um_elect();
        const uint32_t idesc = (1u << 4) | ((uint32_t(NMMA) >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t desc_hi = (uint64_t(128 >> 4) << 16) | (uint64_t(256 >> 4) << 32) | (uint64_t(1) << 46);
        const uint32_t bst_addr = smem_u32(bst);
        int cg = 0;
        for (int t = 0; t < total_steps; ++t) {
            const int bs = t % UM_BSTAGES;
            um_wait(&bfull[bs], (t / UM_BSTAGES) & 1, 3);
            if (t == 0) ST_TRACE(2);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int dbuf = cg & 1;
                const int ab = (t & 1) * 4 + q;
                const bool first = (q % FJ2) == 0, last = (q % FJ2) == FJ2 - 1;
                UM_PROG(cg * 10 + q);
                if (first && cg >= 2) um_wait(&dempty[(cg - 2) % NDB], ((cg - 2) / NDB) & 1, 4);
                um_wait(&afull[ab], (t >> 1) & 1, 5);
                um_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t boff = uint32_t(bs * BSTEP + ((q * 2 + c) * 2 + u) * (NMMA * 32));
                        const uint64_t bdesc = desc_hi | uint64_t(((bst_addr + boff) & 0x3FFFF) >> 4);
                        um_mma(tm_d + dbuf * NMMA, tm_a + ab * 32 + c * 16 + u * 8, bdesc, idesc,
                               (first && c == 0 && u == 0) ? 0u : 1u, leader);
                    }
                um_commit(&aempty[ab], leader);
                if (last) { um_commit(&dfull[cg % NDB], leader); ++cg; }
            }
            um_commit(&bempty[bs], leader);
            if (t < 4) ST_TRACE(4 + t);
        }
        // TMEM is released after the dequant warps have read the last accumulator (the CTA-wide barrier)
        um_fence_before();
        __syncthreads();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
        return;
    } else {
        // =========================== dequant + flush warps ===========================
        // thread <-> TMEM lane (K-quarter s = warp & 3, column n = lane); warp group wg = warp >> 2 owns sub-step wg
        const int s = warp & 3, wg = warp >> 2, n = lane;
        const int nr = um_slice_runs(s, p.rps);
        const bool flusher = ((wg + 3) & 3) % FJ2 == FJ2 - 1;        // the sub-step before mine closes a group ...
        const int fl_gi = ((wg + 3) & 3) / FJ2;                      // ... this group of its step

        pdl_wait_primary();                                          // xsum is produced by umma_prepare_kernel
        ST_TRACE(2);
        for (int i = tid; i < MB * 4 * nseg; i += UM_DW * 32) xseg[i] = p.xsum[i];
        asm volatile("bar.sync 2, %0;" ::"n"((UM_DW + 1) * 32) : "memory");      // TMEM address + xseg
        um_fence_after();
        const uint32_t tmem = *tmem_slot;
        const uint32_t tm_d = tmem, tm_a = tmem + 256;
        const uint32_t lane_addr = uint32_t(s * 32) << 16;
        ST_TRACE(3);

        int ts = 0;                       // real steps of my K-quarter so far (ring slot counter)
        int prev_slot = -1;               // WG 0: slot of the previous step (still needed by its lagging flush)
        bool prev_real = false;
        for (int st = 0; st < nstrips; ++st) {
            float yacc[MB];
#pragma unroll
            for (int m = 0; m < MB; ++m) yacc[m] = 0.f;

            auto flush = [&](int tg, int gi, int slot, int pstep, bool real) {
                const int cg = tg * GPS + gi;
                const int dbuf = cg & 1;
                UM_PROG(1000 + cg * 10 + 1);
                um_wait(&dfull[cg % NDB], (cg / NDB) & 1, 6);
                UM_PROG(1000 + cg * 10 + 2);
                um_fence_after();
                float d[MB];
                um_ld<MB>(tm_d + lane_addr + dbuf * NMMA + s * MB, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                um_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&dempty[cg % NDB]);
                UM_PROG(1000 + cg * 10 + 3);
                if (!real) return;                         // padded step of this K-quarter: A was zero
                const unsigned char* sz = szst + (size_t(slot) * 4 + s) * UM_SZ;
                const int gs = (p.ngr > 1) ? gi : 0;
                const float sf = __half2float(reinterpret_cast<const __half*>(sz + gs * 64)[n]);
                float zf;
                if (p.asym) {
                    const uint32_t zw = reinterpret_cast<const uint32_t*>(sz + 256 + gs * 16)[n / NB];
                    zf = sf * float(((zw >> ((n % NB) * BITS)) & 0xFu) + 1u);
                } else {
                    zf = __half2float(reinterpret_cast<const __half*>(sz + 256 + gs * 64)[n]);
                }
                const float smul = sf * 16777216.0f;
                const float* xs = xseg + s * nseg + pstep * GPS + gi;
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                    if (m < p.M) {
                        yacc[m] = fmaf(smul, d[m], yacc[m]);
                        yacc[m] = fmaf(-zf, xs[m * 4 * nseg], yacc[m]);
                    }
                }
            };

            for (int pstep = 0; pstep < p.steps; ++pstep) {
                const int tg = st * p.steps + pstep;               // global step
                const int slot = ts % S;
                const bool real = pstep < nr;
                uint32_t c0[16], c1[16];
                if (real) {
                    um_wait(&wfull[slot * 4 + s], (ts / S) & 1, 7);
                    if (tg == 0) ST_TRACE(4);
                    const unsigned char* wt = wst + (size_t(slot) * 4 + s) * UM_TILE + wg * 8 * 128 + n * 4;
#pragma unroll
                    for (int wd = 0; wd < 8; ++wd) {
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(wt + wd * 128);
                        const uint32_t tt = w >> 8;
                        c0[2 * wd] = w & 0x000F000Fu;  c0[2 * wd + 1] = tt & 0x000F000Fu;
                        c1[2 * wd] = w & 0x00F000F0u;  c1[2 * wd + 1] = tt & 0x00F000F0u;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) c0[e] = c1[e] = 0u;
                }
                UM_PROG(2000 + tg * 10 + 1);
                const int ab = (tg & 1) * 4 + wg;
                if (tg >= 2) { um_wait(&aempty[ab], ((tg >> 1) - 1) & 1, 8); um_fence_after(); }
                UM_PROG(2000 + tg * 10 + 2);
                um_st16(tm_a + lane_addr + ab * 32, c0);
                um_st16(tm_a + lane_addr + ab * 32 + 16, c1);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                um_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&afull[ab]);
                UM_PROG(2000 + tg * 10 + 3);

                // lagging flush: the group closed by the sub-step before mine (previous step's last group for WG 0)
                if (flusher) {
                    if (wg > 0) flush(tg, fl_gi, slot, pstep, real);
                    else if (pstep > 0) flush(tg - 1, fl_gi, prev_slot, pstep - 1, prev_real);
                }
                // ring slot hand-back: WG 1..3 are done with this step's slot; WG 0 is done with the previous step's
                __syncwarp();
                if (wg > 0) {
                    if (real && lane == 0) mbar_arrive(&wempty[slot * 4 + s]);
                } else {
                    if (prev_slot >= 0 && prev_real && lane == 0) mbar_arrive(&wempty[prev_slot * 4 + s]);
                    prev_slot = slot; prev_real = real;
                }
                if (real) ++ts;
                if (tg == 0) ST_TRACE(5);
            }
            // strip done: WG 0 flushes the strip's last group and hands its slot back
            if (wg == 0) {
                flush(st * p.steps + p.steps - 1, fl_gi, prev_slot, p.steps - 1, prev_real);
                __syncwarp();
                if (prev_slot >= 0 && prev_real && lane == 0) mbar_arrive(&wempty[prev_slot * 4 + s]);
                prev_slot = -1; prev_real = false;
            }
            ST_TRACE(6);
            UM_PROG(3000 + st);
            // ---- combine: flusher warp groups in fixed order, then the 4 K-quarters in fixed order ----
#pragma unroll
            for (int f = 0; f < NFL; ++f) {
                if (flusher && wg == (f * FJ2) % 4) {          // FJ2=1: 0,1,2,3; FJ2=2: 0,2; FJ2=4: 0
#pragma unroll
                    for (int m = 0; m < MB; ++m)
                        if (m < p.M) {
                            float* dst = &part[(s * MB + m) * 32 + n];
                            *dst = (f == 0) ? yacc[m] : (*dst + yacc[m]);
                        }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(UM_DW * 32) : "memory");
            }
            for (int o = tid; o < p.M * 32; o += UM_DW * 32) {
                const int m = o >> 5, nn = o & 31;
                const float v = (part[(0 * MB + m) * 32 + nn] + part[(1 * MB + m) * 32 + nn]) +
                                (part[(2 * MB + m) * 32 + nn] + part[(3 * MB + m) * 32 + nn]);
                p.y[size_t(m) * p.N + (s_lo + st) * 32 + nn] = f32_to_16<false>(v);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(UM_DW * 32) : "memory");
        }
    }
    um_fence_before();
    __syncthreads();
    ST_TRACE(7);
}

struct UmmaLaunch {
    int FJ2, MB, grid;
    size_t smem;
    unsigned flags;
    cudaStream_t stream;
};
int launch_umma(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const UmmaParams& p, const UmmaLaunch& l);
int launch_umma_prepare(const UmmaPrepParams& p, unsigned flags, cudaStream_t stream);

}  // namespace b200bit
