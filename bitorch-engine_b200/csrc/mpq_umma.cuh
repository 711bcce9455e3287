// mpq_umma.cuh -- tcgen05 / TMEM small-batch kernel (fp16 activations, 4-bit weights, M <= 4 per pass) for sm_100a.
//
// The measured bound of the CUDA-core / mma.sync decode kernels is the compute tail after the weights have landed
// (~1 us per 4096x4096 layer: FHFMA runs at ~89 lanes/clk/SM, legacy HMMA at 0.5 instr/clk/SM).
// This kernel moves the multiply-accumulates to the 5th-generation tensor cores (8x the legacy MMA rate, issued by ONE
// thread) and leaves the CUDA cores only the 5 bit-ops per packed word that turn it into an operand:
//
//   * work decomposition as in mpq_stream.cuh: a persistent CTA owns whole 32-column strips; TMA streams 32x32-word
//     tiles + scale / zero rows into a shared-memory ring BEFORE griddepcontrol.wait (PDL prefetch);
//   * split-K inside the MMA:  the 128 TMEM lanes of one UMMA are (K-quarter s, column n): lane 32*s + n.  The B
//     operand's N dimension is (K-quarter s', batch m): only the diagonal blocks s == s' are used; the 4x redundant
//     MACs are free on tcgen05 and buy a 4-way K split with NO cross-CTA (or cross-MMA) reduction;
//   * A operand in TMEM, written by the dequant warps with tcgen05.st: a masked packed word IS two fp16 K elements
//     (fp16 subnormals q * 2^-24, honoured exactly by UTCHMMA -- tools/umma_probe.cu); fields at bit 4 (16*q) go to a
//     second A/D pair ("class 1") and are folded in at flush time; the K permutation this implies (0,4,2,6 | 1,5,3,7
//     per word) is applied once to x when it is staged into the canonical K-major core-matrix layout for B;
//   * one thread issues tcgen05.mma.kind::f16 (M=128, N=16, K=16) per 16 K elements, accumulators in TMEM, double
//     buffered per quantisation group; tcgen05.commit -> mbarriers hand buffers back;
//   * group affine factored out in fp32 at flush (tcgen05.ld of the thread's own lane / batch columns).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "mpq_mma.cuh"
#include "mpq_stream.cuh"   // mbarrier / TMA helpers

namespace b200bit {

constexpr int UM_NBUF = 4;          // A buffers (sub-steps in flight)
constexpr int UM_NMMA = 16;         // UMMA N = 4 K-quarters x 4 batch slots
constexpr int UM_MB = 4;            // batch slots per K-quarter
constexpr int UM_TILE = 4096;       // 32 rows x 128 B
constexpr int UM_SZ = 512;          // scales (256) + zeros (256) per (slice, stage)
constexpr int UM_TMEM_COLS = 256;   // D: [2][2][16] = 64 | A: [NBUF][2][16] = 128

struct UmmaParams {
    const uint16_t* x;   // [M, K] f16
    uint16_t* y;         // [M, N] f16
    int M, K, N;
    int strips;          // N / 32
    int rps;             // runs (32 packed rows) per strip
    int steps;           // ceil(rps / 4): steps per strip (every K-quarter advances one run per step)
    int ngr;             // scale rows per run: 2 (g128), 4 (g64), 1 (group >= 256 k)
    int rpr, rpr_shift;  // runs per group when a group spans >= 1 run
    int asym;
    int S;               // ring stages (steps in flight)
    unsigned long long* trace;
};

__device__ __forceinline__ void um_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void um_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// The MMA warp runs its loop CONVERGED (all 32 lanes wait on the barriers) and only the tcgen05 instruction itself is
// predicated on the elected lane: issued from a divergent `if (lane == 0)` region the compiler wraps every UTCHMMA in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop that costs ~110 cycles per instruction (tools/umma_rate.cu).
__device__ __forceinline__ uint32_t um_elect() {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
    return leader;
}
__device__ __forceinline__ void um_expect_tx(uint64_t* bar, unsigned bytes, uint32_t leader) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
                 "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes), "r"(leader) : "memory");
}
__device__ __forceinline__ void um_tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint32_t leader) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void um_commit(uint64_t* bar, uint32_t leader) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void um_mma(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc),
                 "r"(accumulate), "r"(leader) : "memory");
}
__device__ __forceinline__ void um_st16(uint32_t addr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void um_ld4(uint32_t addr, float (&v)[4]) {
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
    v[0] = __uint_as_float(a); v[1] = __uint_as_float(b); v[2] = __uint_as_float(c); v[3] = __uint_as_float(d);
}

// first run and run count of K-quarter s
__device__ __forceinline__ int um_slice_r0(int s, int rps) { return s * (rps >> 2) + min(s, rps & 3); }
__device__ __forceinline__ int um_slice_runs(int s, int rps) { return (rps >> 2) + (s < (rps & 3) ? 1 : 0); }

// byte offset of B element (row j, k index kk in [0,16)) inside one 16 x 16 canonical K-major tile (512 B):
// core matrix = 8 rows x 16 B contiguous; K-adjacent cores 128 B apart (LBO), 8-row groups 256 B apart (SBO)
__device__ __forceinline__ int um_b_off(int j, int kk) { return (j >> 3) * 256 + (kk >> 3) * 128 + (j & 7) * 16 + (kk & 7) * 2; }

// FJ2 = sub-steps (8 packed rows) per quantisation group: 1 (g64), 2 (g128), 4 (group >= 256 k: flush per step)
template <int FJ2>
__global__ void __launch_bounds__(192, 1) mpq_umma_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                          const __grid_constant__ CUtensorMap tm_s,
                                                          const __grid_constant__ CUtensorMap tm_z, const UmmaParams p) {
    constexpr int BITS = 4, NB = 8;
    constexpr int GPS = 4 / FJ2;     // groups (flushes) per step
    extern __shared__ unsigned char um_smem_raw[];
    unsigned char* smem_raw = um_smem_raw + ((128u - (smem_u32(um_smem_raw) & 127u)) & 127u);   // TMA destinations: 128 B
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, G = gridDim.x, S = p.S;
    ST_TRACE(0);

    // ---- shared memory ----
    unsigned char* wst = smem_raw;                                        // [S][4 slices][4096]
    unsigned char* szst = wst + size_t(S) * 4 * UM_TILE;                   // [S][4][512]
    unsigned char* bt = szst + size_t(S) * 4 * UM_SZ;                      // B tiles: [steps][4 q][2 class][2 u][512]
    const int b_bytes = p.steps * 4 * 2 * 2 * 512;
    float* xseg = reinterpret_cast<float*>(bt + b_bytes);                 // [M][4 slices][steps * GPS]
    const int nseg_slice = p.steps * GPS;
    float* part = xseg + ((p.M * 4 * nseg_slice + 3) & ~3);               // [4][M][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + 4 * UM_MB * 32);
    uint64_t* wfull = bars;                    // [S][4]
    uint64_t* wempty = wfull + S * 4;          // [S][4]
    uint64_t* afull = wempty + S * 4;          // [NBUF]
    uint64_t* aempty = afull + UM_NBUF;        // [NBUF]
    uint64_t* dfull = aempty + UM_NBUF;        // [2]
    uint64_t* dempty = dfull + 2;              // [2]
    uint64_t* bready = dempty + 2;             // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bready + 1);

    const int s_lo = int((long long)b * p.strips / G), s_hi = int((long long)(b + 1) * p.strips / G);
    const int nstrips = s_hi - s_lo;
    const int total_steps = nstrips * p.steps;

    if (tid < S * 8) mbar_init(&wfull[tid], 1);                          // wfull + wempty are contiguous
    if (tid < UM_NBUF) { mbar_init(&afull[tid], 128); mbar_init(&aempty[tid], 1); }
    if (tid < 2) { mbar_init(&dfull[tid], 1); mbar_init(&dempty[tid], 128); }
    if (tid == 0) mbar_init(bready, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    um_fence_before();
    __syncthreads();
    um_fence_after();
    ST_TRACE(1);
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_d = tmem;                  // + dbuf*32 + class*16
    const uint32_t tm_a = tmem + 64;             // + abuf*32 + class*16

    const unsigned tile_bytes = UM_TILE + unsigned(p.ngr) * 64u + (p.asym ? unsigned(p.ngr) * 16u : unsigned(p.ngr) * 64u);

    if (warp == 4) {
        // =========================== producer (converged warp, elected lane issues the TMA) ===========================
        const uint32_t leader = um_elect();
        for (int st = 0; st < nstrips; ++st) {
            const int strip = s_lo + st;
            for (int pstep = 0; pstep < p.steps; ++pstep) {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int r0 = um_slice_r0(s, p.rps), nr = um_slice_runs(s, p.rps);
                    if (pstep >= nr) continue;                     // padded step: no tile (the consumer writes zeros)
                    const int ts = st * nr + pstep;                // real steps of this K-quarter so far
                    const int slot = ts % S;
                    if (ts >= S) mbar_wait(&wempty[slot * 4 + s], ((ts / S) - 1) & 1);
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
            }
        }
    } else if (warp == 5) {
        // =========================== MMA issuer ===========================
        {
            const uint32_t leader = um_elect();
            mbar_wait(bready, 0);
            um_fence_after();
            ST_TRACE(2);
            const uint32_t idesc = (1u << 4) | ((uint32_t(UM_NMMA) >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t desc_hi = (uint64_t(128 >> 4) << 16) | (uint64_t(256 >> 4) << 32) | (uint64_t(1) << 46);
            const uint32_t bt_addr = smem_u32(bt);
            int css = 0, cg = 0;
            for (int t = 0; t < total_steps; ++t) {
                const int pstep = t % p.steps;
#pragma unroll 1
                for (int q = 0; q < 4; ++q, ++css) {
                    const int abuf = css % UM_NBUF;
                    const int dbuf = cg & 1;
                    const bool first = (q % FJ2) == 0, last = (q % FJ2) == FJ2 - 1;
                    if (first && cg >= 2) mbar_wait(&dempty[dbuf], ((cg >> 1) - 1) & 1);
                    mbar_wait(&afull[abuf], (css / UM_NBUF) & 1);
                    um_fence_after();
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const uint32_t boff = uint32_t((((pstep * 4 + q) * 2 + c) * 2 + u) * 512);
                            const uint64_t bdesc = desc_hi | uint64_t(((bt_addr + boff) & 0x3FFFF) >> 4);
                            um_mma(tm_d + dbuf * 32 + c * 16, tm_a + abuf * 32 + c * 16 + u * 8, bdesc, idesc,
                                   (first && u == 0) ? 0u : 1u, leader);
                        }
                    um_commit(&aempty[abuf], leader);
                    if (last) { um_commit(&dfull[dbuf], leader); ++cg; }
                }
                if (t == 0) ST_TRACE(4);
            }
            ST_TRACE(5);
        }
    } else {
        // =========================== dequant + epilogue warps: thread <-> TMEM lane (K-quarter s = warp, column n = lane) =====
        const int s = warp, n = lane;
        const int r0 = um_slice_r0(s, p.rps), nr = um_slice_runs(s, p.rps);

        pdl_wait_primary();   // x is produced by the previous kernel
        ST_TRACE(2);

        // ---- stage B (x in UMMA core-matrix order, class split, K permuted) and the per-group sums of x ----
        {
            const uint4 z4 = make_uint4(0, 0, 0, 0);
            for (int i = tid * 16; i < b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(bt + i) = z4;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int krows = p.K / NB;                         // packed rows over K
            const int total = p.M * krows;                      // multiple of 32
            constexpr int SEG_ROWS = 8 * FJ2;
            auto stage_one = [&](int i, const uint4 v) {
                const int m = i / krows, row = i - m * krows;
                const int run = row >> 5, rl = row & 31;
                int sl = 0;                                      // K-quarter of this run
#pragma unroll
                for (int q = 1; q < 4; ++q) sl += (run >= um_slice_r0(q, p.rps)) ? 1 : 0;
                const int pstep = run - um_slice_r0(sl, p.rps);
                const int q = rl >> 3, wd = rl & 7, u = wd >> 2, wi = wd & 3;
                const int j = sl * UM_MB + m;
                // class 0: codes (0,4,2,6); class 1: codes (1,5,3,7)
                const uint32_t c0a = __byte_perm(v.x, v.z, 0x5410), c0b = __byte_perm(v.y, v.w, 0x5410);
                const uint32_t c1a = __byte_perm(v.x, v.z, 0x7632), c1b = __byte_perm(v.y, v.w, 0x7632);
                unsigned char* t0 = bt + ((((pstep * 4 + q) * 2 + 0) * 2 + u) * 512) + um_b_off(j, 4 * wi);
                *reinterpret_cast<uint2*>(t0) = make_uint2(c0a, c0b);
                *reinterpret_cast<uint2*>(t0 + 1024) = make_uint2(c1a, c1b);      // class 1 tile = +2 tiles
                float sum = 0.f;
                sum = fhfma<false, false, false>(0x3C003C00u, v.x, sum); sum = fhfma<false, true, true>(0x3C003C00u, v.x, sum);
                sum = fhfma<false, false, false>(0x3C003C00u, v.y, sum); sum = fhfma<false, true, true>(0x3C003C00u, v.y, sum);
                sum = fhfma<false, false, false>(0x3C003C00u, v.z, sum); sum = fhfma<false, true, true>(0x3C003C00u, v.z, sum);
                sum = fhfma<false, false, false>(0x3C003C00u, v.w, sum); sum = fhfma<false, true, true>(0x3C003C00u, v.w, sum);
#pragma unroll
                for (int off = 1; off < SEG_ROWS; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                if ((rl % SEG_ROWS) == 0) xseg[(m * 4 + sl) * nseg_slice + pstep * GPS + rl / SEG_ROWS] = sum;
            };
            const uint4* xv = reinterpret_cast<const uint4*>(p.x);       // chunk i = 8 halves; [M, K] is contiguous
            int i0 = warp * 32 + lane;
            for (; i0 + 3 * 128 < total; i0 += 4 * 128) {               // 4 loads in flight per thread
                const uint4 v0 = xv[i0], v1 = xv[i0 + 128], v2 = xv[i0 + 256], v3 = xv[i0 + 384];
                stage_one(i0, v0); stage_one(i0 + 128, v1); stage_one(i0 + 256, v2); stage_one(i0 + 384, v3);
            }
            for (; i0 < total; i0 += 128) stage_one(i0, xv[i0]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // B is read by the tensor core (async proxy)
            mbar_arrive(bready);
            ST_TRACE(3);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // xseg visible to all dequant warps

        const uint32_t lane_addr = uint32_t(warp * 32) << 16;
        int css = 0, cg_flushed = 0, ts = 0, rel_slot = -1;
        // pending group bookkeeping for the lagging flush
        int pend_slot = 0, pend_gi = 0, pend_pstep = 0;
        bool pend_valid = false, pend_real = false;

        for (int st = 0; st < nstrips; ++st) {
            float yacc[UM_MB] = {0.f, 0.f, 0.f, 0.f};
            auto flush = [&](int slot, int gi, int pstep, bool real) {
                const int dbuf = cg_flushed & 1;
                mbar_wait(&dfull[dbuf], (cg_flushed >> 1) & 1);
                um_fence_after();
                float d0[4], d1[4];
                um_ld4(tm_d + lane_addr + dbuf * 32 + 0 + s * UM_MB, d0);
                um_ld4(tm_d + lane_addr + dbuf * 32 + 16 + s * UM_MB, d1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                um_fence_before();
                mbar_arrive(&dempty[dbuf]);
                ++cg_flushed;
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
#pragma unroll
                for (int m = 0; m < UM_MB; ++m) {
                    if (m < p.M) {
                        const float tsum = fmaf(d1[m], 1.0f / 16.0f, d0[m]);
                        const float xs = xseg[(m * 4 + s) * nseg_slice + pstep * GPS + gi];
                        yacc[m] = fmaf(smul, tsum, yacc[m]);
                        yacc[m] = fmaf(-zf, xs, yacc[m]);
                    }
                }
            };

            for (int pstep = 0; pstep < p.steps; ++pstep) {
                const int slot = ts % S;
                const bool real = pstep < nr;
                const unsigned char* wt = wst + (size_t(slot) * 4 + s) * UM_TILE;
                if (real) mbar_wait(&wfull[slot * 4 + s], (ts / S) & 1);
                if (st == 0 && pstep == 0) ST_TRACE(4);
#pragma unroll 1
                for (int q = 0; q < 4; ++q, ++css) {
                    const int abuf = css % UM_NBUF;
                    uint32_t c0[16], c1[16];
                    if (real) {
#pragma unroll
                        for (int wd = 0; wd < 8; ++wd) {
                            const uint32_t w = *reinterpret_cast<const uint32_t*>(wt + (q * 8 + wd) * 128 + n * 4);
                            const uint32_t tt = w >> 8;
                            c0[2 * wd] = w & 0x000F000Fu;  c0[2 * wd + 1] = tt & 0x000F000Fu;
                            c1[2 * wd] = w & 0x00F000F0u;  c1[2 * wd + 1] = tt & 0x00F000F0u;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) c0[e] = c1[e] = 0u;
                    }
                    if (css >= UM_NBUF) { mbar_wait(&aempty[abuf], ((css / UM_NBUF) - 1) & 1); um_fence_after(); }
                    um_st16(tm_a + lane_addr + abuf * 32, c0);
                    um_st16(tm_a + lane_addr + abuf * 32 + 16, c1);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    um_fence_before();
                    mbar_arrive(&afull[abuf]);
                    if ((q % FJ2) == FJ2 - 1) {
                        // group complete on the write side: flush the PREVIOUS group (its MMAs overlap these writes)
                        if (pend_valid) flush(pend_slot, pend_gi, pend_pstep, pend_real);
                        pend_valid = true; pend_slot = slot; pend_gi = q / FJ2; pend_pstep = pstep; pend_real = real;
                    }
                }
                // the scale rows of a step are still needed by the lagging flush of its last group, which has happened by
                // now for the PREVIOUS real step -> hand that ring slot back to the producer
                __syncwarp();
                if (rel_slot >= 0 && lane == 0) mbar_arrive(&wempty[rel_slot * 4 + s]);
                rel_slot = -1;
                if (real) { rel_slot = slot; ++ts; }
                if (st == 0 && pstep == 0) ST_TRACE(5);
            }
            // strip done: flush the last pending group before the epilogue of this strip
            if (pend_valid) { flush(pend_slot, pend_gi, pend_pstep, pend_real); pend_valid = false; }
            ST_TRACE(6);
            // ---- combine the 4 K-quarters (fixed order) and write y ----
#pragma unroll
            for (int m = 0; m < UM_MB; ++m)
                if (m < p.M) part[(s * UM_MB + m) * 32 + n] = yacc[m];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 0) {
                for (int m = 0; m < p.M; ++m) {
                    const float v = (part[(0 * UM_MB + m) * 32 + n] + part[(1 * UM_MB + m) * 32 + n]) +
                                    (part[(2 * UM_MB + m) * 32 + n] + part[(3 * UM_MB + m) * 32 + n]);
                    p.y[size_t(m) * p.N + (s_lo + st) * 32 + n] = f32_to_16<false>(v);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
    }
    um_fence_before();
    __syncthreads();
    ST_TRACE(7);
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(UM_TMEM_COLS) : "memory");
    }
}

struct UmmaLaunch {
    int FJ2, grid;
    size_t smem;
    unsigned flags;
    cudaStream_t stream;
};
int launch_umma(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const UmmaParams& p, const UmmaLaunch& l);

}  // namespace b200bit
