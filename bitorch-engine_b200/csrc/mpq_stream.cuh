// mpq_stream.cuh -- TMA-streamed small-batch kernel (fp16 activations, W{2,4,8}, 1 <= M <= 32) for sm_100a.
//
// The decode-shaped Linear is a pure HBM stream (8.9 - 24 MB per Llama-7B layer) that lasts only 1.4 - 3.7 us at
// the roofline, so the design is organised around keeping HBM requests in flight ACROSS kernel boundaries:
//   * grid = min(#strips, #SMs) persistent CTAs; a CTA owns a contiguous range of 32-column strips over the full K
//     (no cross-CTA reduction, no workspace, no atomics, no fences); its work items -- 32 columns x 32 packed rows
//     = 4 KB of packed weights -- are split over the consumer warps;
//   * a producer thread issues TMA tile loads (cp.async.bulk.tensor.2d, 128B-swizzled, mbarrier completion) of the
//     CTA's packed-weight tiles plus the matching scale / zero rows into a shared-memory ring.  None of this depends
//     on the previous kernel in the stream, so with programmatic dependent launch it is issued BEFORE
//     griddepcontrol.wait: while layer i computes, layer i+1 (and i+2: the footprint is < 1/3 of an SM) already has
//     its weights landing in shared memory;
//   * consumer warps own contiguous run ranges; per run: 8 conflict-free LDS.128 of packed words, field masks that
//     turn the words into fp16-subnormal A fragments (no int->float conversion, see mpq_mma.cuh), x fragments straight
//     from global/L1 (byte-permuted in registers), mma.sync.m16n8k16 with fp32 accumulation, an all-ones A fragment
//     delivering the per-group sum of x for the zero-point term; no block-wide barrier before the final reduction;
//   * group affine factored out (exact fp32 evaluation of the quantised model, DESIGN.md "numerics");
//   * deterministic output: per-warp partials -> fixed-order CTA sum.  y is written exactly once.
// Replaces quant_mm_kernel{,_asym} (mpq_linear_cuda_kernel.cu:67-451) + the torch::zeros memset (:618).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "tma.cuh"

namespace b200bit {

constexpr int ST_RUN_ROWS = 32;      // packed rows per work item
constexpr int ST_TILE_BYTES = 4096;  // 32 rows x 128 B
constexpr int ST_SZ_BYTES = 512;     // per-stage scale + zero rows (<= 4 groups x 64 B each)
constexpr int ST_MAXSEG = 4;         // strip segments a warp's run range may touch
constexpr int ST_MAX_WARPS = 16;

struct StreamParams {
    const uint16_t* x;   // [M, K] f16
    const uint16_t* zero_page;   // >= 2 KB of zeros (lanes whose batch row is >= M read their x fragments here)
    uint16_t* y;         // [M, N] f16
    int M, K, N;
    int strips;          // N / 32
    int rps;             // runs per strip = K / (32 * NB)
    int ngr;             // scale rows per run (1, 2 or 4)
    int rpr, rpr_shift;  // runs per group when a group spans >= 1 run
    int asym;
    int S;               // ring stages
    int maxseg;          // partial-sum slots per consumer warp (<= ST_MAXSEG)
    int debug_no_x;      // diagnostics: read x fragments from the zero page (timing experiments only)
    int producer_mode;           // 0: lane w feeds warp w (independent waits); 1: converged warp, elected-lane issue
    unsigned long long* trace;   // optional [grid][16 warps][8] globaltimer stamps (diagnostics; nullptr = off)
};

#define ST_TRACE(slot_) do { if (p.trace && lane == 0) p.trace[(size_t(blockIdx.x) * 16 + warp) * 8 + (slot_)] = st_gtime(); } while (0)

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// XS = true : x is staged ONCE per CTA into shared memory, already permuted into fragment order, together with its
//             per-segment sums (<= 8 batch rows, M*K*2 <= ~64 KB); up to 16 consumer warps, no x registers, no
//             ones-MMA.  This is the decode (M = 1) configuration.
// XS = false: x fragments are pulled from global through a register ring (larger M*K); <= 8 consumer warps.
template <int BITS, int MT, int FJ, bool XS>
__global__ void __launch_bounds__(XS ? 544 : 288, 1) mpq_stream_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                         const __grid_constant__ CUtensorMap tm_s,
                                                         const __grid_constant__ CUtensorMap tm_z,
                                                         const StreamParams p) {
    constexpr int NB = 32 / BITS;
    constexpr int NF = 16 / BITS;
    constexpr int NACC = BITS >= 8 ? 1 : 8 / BITS;
    constexpr int XR = NB / 2;
    constexpr int NSEG = 8 / FJ;
    constexpr uint32_t FM = (1u << BITS) - 1u;
    (void)NF;

    extern __shared__ __align__(1024) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NW = (blockDim.x >> 5) - 1;          // consumer warps; warp NW is the producer
    const int r = lane >> 2, c = lane & 3;
    const int b = blockIdx.x, G = gridDim.x;
    const int S = p.S;

    // ---- shared memory carve-up ----
    unsigned char* wst = smem_raw;                                   // S x 4096 (1024-aligned)
    unsigned char* szst = wst + size_t(S) * ST_TILE_BYTES;           // S x 512
    uint64_t* full = reinterpret_cast<uint64_t*>(szst + size_t(S) * ST_SZ_BYTES);
    uint64_t* empty = full + S;
    int* seg_strip = reinterpret_cast<int*>(empty + S);             // [NW][ST_MAXSEG]
    int* seg_count = seg_strip + ST_MAX_WARPS * ST_MAXSEG;           // [NW]
    int* flags = seg_count + ST_MAX_WARPS;                           // [4]
    float* part = reinterpret_cast<float*>(flags + 4);               // [NW][ST_MAXSEG][M][32]
    // XS only: permuted x [M][xs_mstride] u16 (16-byte aligned, per-row stride == 64 mod 128 bytes), a zero row,
    // and the per-segment sums [M][K rows / SEG_ROWS] f32
    constexpr int SEG_ROWS = 4 * FJ;
    const int krows = p.K / NB;                                      // packed rows over the full K
    const int xs_mstride = krows * NB + 32;                          // halves
    uint16_t* xs = reinterpret_cast<uint16_t*>(part + size_t(NW) * p.maxseg * p.M * 32);
    uint16_t* xzero = xs + size_t(p.M) * xs_mstride;                 // ST_RUN_ROWS * NB halves of zeros
    float* xseg = reinterpret_cast<float*>(xzero + ST_RUN_ROWS * NB);

    // ---- this CTA's strips (whole K each), its run range and the per-warp split ----
    const int s_lo = int((long long)b * p.strips / G), s_hi = int((long long)(b + 1) * p.strips / G);   // [s_lo, s_hi)
    const int lo = s_lo * p.rps, hi = s_hi * p.rps;
    const int Rc = hi - lo;
    const int q0 = Rc / NW, rem = Rc - q0 * NW;

    ST_TRACE(0);
    if (tid < 2 * S) mbar_init(&full[tid], 1);        // full[0..S) and empty[0..S) are contiguous
    if (tid < ST_MAX_WARPS) seg_count[tid] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pdl_launch_dependents();
    __syncthreads();

    ST_TRACE(1);
    const unsigned stage_bytes = ST_TILE_BYTES + unsigned(p.ngr) * 64u + (p.asym ? unsigned(p.ngr) * (128u / NB) : unsigned(p.ngr) * 64u);

    if (warp == NW) {
        if (p.producer_mode == 0 || p.producer_mode == 2) {
            // =========================== producer: lane w feeds consumer warp w ===========================
            // warp w's k-th run lives in ring slot (k % depth) * NW + w, depth = S / NW: every slot has exactly one producer
            // lane and one consumer warp, so the empty/full phases of a slot are always used in order.
            if (lane < NW) {
                const int w = lane;
                const int cnt_w = q0 + (w < rem ? 1 : 0);
                const int wlo = lo + w * q0 + min(w, rem);
                int strip = (cnt_w > 0) ? wlo / p.rps : 0;
                int kr = wlo - strip * p.rps;
                const int depth = S / NW;                 // ring slots owned by this lane: (k % depth) * NW + w
                for (int k = 0; k < cnt_w; ++k) {
                    const int slot = (k % depth) * NW + w;
                    if (k >= depth) mbar_wait(&empty[slot], ((k / depth) - 1) & 1);
                    int g0;
                    if (p.ngr > 1 || p.rpr == 1) g0 = kr * p.ngr;
                    else g0 = (p.rpr_shift >= 0) ? (kr >> p.rpr_shift) : (kr / p.rpr);
                    if (p.producer_mode == 2) {      // timing experiment: weight tile only (results are wrong)
                        mbar_expect_tx(&full[slot], ST_TILE_BYTES);
                        tma_load_2d(wst + size_t(slot) * ST_TILE_BYTES, &tm_w, strip * 32, kr * ST_RUN_ROWS, &full[slot]);
                    } else {
                    mbar_expect_tx(&full[slot], stage_bytes);
                    tma_load_2d(wst + size_t(slot) * ST_TILE_BYTES, &tm_w, strip * 32, kr * ST_RUN_ROWS, &full[slot]);
                    unsigned char* sz = szst + size_t(slot) * ST_SZ_BYTES;
                    tma_load_2d(sz, &tm_s, strip * 32, g0, &full[slot]);
                    tma_load_2d(sz + 256, &tm_z, p.asym ? strip * (32 / NB) : strip * 32, g0, &full[slot]);
                    }
                    if (++kr == p.rps) { kr = 0; ++strip; }
                }
            }
        } else {
            // =========================== alternative producer (converged warp; the elected lane issues every TMA) ===========================
            // warp w's k-th run lives in ring slot (k % depth) * NW + w, depth = S / NW: every slot has exactly one consumer
            // warp, so the empty/full phases of a slot are always used in order.  Issue order is k-major (the order of use).
            const uint32_t leader = um_elect();
            const int depth = S / NW;
            const int kmax = q0 + (rem > 0 ? 1 : 0);
            for (int k = 0; k < kmax; ++k) {
                for (int w = 0; w < NW; ++w) {
                    if (k >= q0 + (w < rem ? 1 : 0)) continue;
                    const int run = lo + w * q0 + min(w, rem) + k;
                    const int strip = run / p.rps;
                    const int kr = run - strip * p.rps;
                    const int slot = (k % depth) * NW + w;
                    if (k >= depth) mbar_wait(&empty[slot], ((k / depth) - 1) & 1);
                    int g0;
                    if (p.ngr > 1 || p.rpr == 1) g0 = kr * p.ngr;
                    else g0 = (p.rpr_shift >= 0) ? (kr >> p.rpr_shift) : (kr / p.rpr);
                    um_expect_tx(&full[slot], stage_bytes, leader);
                    um_tma_2d(wst + size_t(slot) * ST_TILE_BYTES, &tm_w, strip * 32, kr * ST_RUN_ROWS, &full[slot], leader);
                    unsigned char* sz = szst + size_t(slot) * ST_SZ_BYTES;
                    um_tma_2d(sz, &tm_s, strip * 32, g0, &full[slot], leader);
                    um_tma_2d(sz + 256, &tm_z, p.asym ? strip * (32 / NB) : strip * 32, g0, &full[slot], leader);
                }
            }
        }
    } else {
        // =========================== consumers ===========================
        const int w = warp;
        const int cnt = q0 + (w < rem ? 1 : 0);
        const int wlo = lo + w * q0 + min(w, rem);
        int strip = (cnt > 0) ? wlo / p.rps : 0;
        int kr = wlo - strip * p.rps;
        int seg = 0;

        float yacc[MT][4][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int e = 0; e < 4; ++e) yacc[mt][e][0] = yacc[mt][e][1] = 0.f;

        auto store_partial = [&]() {
            float* dst = part + size_t(w * p.maxseg + seg) * p.M * 32;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int m = mt * 8 + 2 * c + h;
                        if (m < p.M) dst[m * 32 + 4 * r + e] = yacc[mt][e][h];
                        yacc[mt][e][h] = 0.f;
                    }
            if (lane == 0) seg_strip[w * p.maxseg + seg] = strip;
            ++seg;
        };

        pdl_wait_primary();   // x is produced by the previous kernel; y / workspace may still be read by it
        ST_TRACE(2);

        constexpr int PF = XS ? 1 : ((MT <= 2) ? 8 : 4);
        uint32_t xq[XS ? 1 : MT][PF][XR];
        const uint16_t* xbase[MT];
        uint32_t xrun_stride[MT];
        (void)xq;
        if constexpr (XS) {
            // ---- stage x once per CTA: thread <-> (m, packed row): permute into fragment order, row sums,
            //      SEG_ROWS consecutive rows (== consecutive lanes) reduced by shuffle ----
            const int total = p.M * krows;                           // multiple of 32
            const int nsegx = krows / SEG_ROWS;
            for (int i0 = w * 32; i0 < total; i0 += NW * 32) {
                const int i = i0 + lane;
                const int m = i / krows, row = i - m * krows;
                uint32_t in[XR];
                const uint16_t* xg = p.x + size_t(m) * p.K + size_t(row) * NB;
                if constexpr (XR >= 4) {
#pragma unroll
                    for (int v4 = 0; v4 < XR / 4; ++v4) {
                        const uint4 v = *reinterpret_cast<const uint4*>(xg + v4 * 8);
                        in[v4 * 4 + 0] = v.x; in[v4 * 4 + 1] = v.y; in[v4 * 4 + 2] = v.z; in[v4 * 4 + 3] = v.w;
                    }
                } else {
                    const uint2 v = *reinterpret_cast<const uint2*>(xg);
                    in[0] = v.x; in[1] = v.y;
                }
                uint32_t outv[XR];
                float sum = 0.f;
#pragma unroll
                for (int v = 0; v < XR; ++v) {
                    const int ka = mma_kperm<BITS>(2 * v), kb = mma_kperm<BITS>(2 * v + 1);
                    const uint32_t sel = ((ka & 1) ? 0x32u : 0x10u) | (((kb & 1) ? 0x76u : 0x54u) << 8);
                    outv[v] = __byte_perm(in[ka >> 1], in[kb >> 1], sel);
                    sum = fhfma<false, false, false>(0x3C003C00u, in[v], sum);
                    sum = fhfma<false, true, true>(0x3C003C00u, in[v], sum);
                }
                uint16_t* dst = xs + m * xs_mstride + row * NB;
                if constexpr (XR >= 4) {
#pragma unroll
                    for (int v4 = 0; v4 < XR / 4; ++v4)
                        *reinterpret_cast<uint4*>(dst + v4 * 8) = make_uint4(outv[v4 * 4], outv[v4 * 4 + 1], outv[v4 * 4 + 2], outv[v4 * 4 + 3]);
                } else {
                    *reinterpret_cast<uint2*>(dst) = make_uint2(outv[0], outv[1]);
                }
#pragma unroll
                for (int off = 1; off < SEG_ROWS; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                if ((row % SEG_ROWS) == 0) xseg[m * nsegx + row / SEG_ROWS] = sum;
            }
            for (int z = w * 32 + lane; z < ST_RUN_ROWS * NB / 2; z += NW * 32) reinterpret_cast<uint32_t*>(xzero)[z] = 0u;
            asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");   // consumer warps only
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const bool valid = (mt * 8 + r) < p.M;
                xbase[mt] = valid ? xs + (mt * 8 + r) * xs_mstride + (2 * c) * NB : xzero + (2 * c) * NB;
                xrun_stride[mt] = valid ? uint32_t(ST_RUN_ROWS * NB) : 0u;
            }
        }
        // ---- x fragments: a register ring of PF j-steps, refilled right after use (the refill for the next run
        //      flies while the rest of the current run computes); every x access is an L2 hit of ~700 cycles ----
        // per-(lane, mt) base pointer and run stride: lanes without a batch row read zeros (stride 0), so every
        // load below is unconditional -- no divergent branch in front of the warp-synchronous MMAs
        if constexpr (!XS) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const bool valid = ((mt * 8 + r) < p.M) && !p.debug_no_x;
            xbase[mt] = valid ? p.x + size_t(mt * 8 + r) * p.K + (2 * c) * NB : p.zero_page + (2 * c) * NB;
            xrun_stride[mt] = valid ? uint32_t(ST_RUN_ROWS * NB) : 0u;
        }
        }
        auto load_x = [&](int e, int krun, int j) {
            if constexpr (XS) { (void)e; (void)krun; (void)j; return; } else {
            // j-step j of the run whose k index is krun -> ring entry e
            const int rl_base = 8 * (j >> 1) + (j & 1);                    // + 2c is folded into xbase
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const uint16_t* xp = xbase[mt] + size_t(uint32_t(krun) * xrun_stride[mt]) + rl_base * NB;
                if constexpr (XR >= 4) {
#pragma unroll
                    for (int v4 = 0; v4 < XR / 4; ++v4) {
                        const uint4 v = *reinterpret_cast<const uint4*>(xp + v4 * 8);
                        xq[mt][e][v4 * 4 + 0] = v.x; xq[mt][e][v4 * 4 + 1] = v.y;
                        xq[mt][e][v4 * 4 + 2] = v.z; xq[mt][e][v4 * 4 + 3] = v.w;
                    }
                } else {
                    const uint2 v = *reinterpret_cast<const uint2*>(xp);
                    xq[mt][e][0] = v.x; xq[mt][e][1] = v.y;
                }
            }
            }
        };
        if constexpr (!XS) {
            if (cnt > 0) {
#pragma unroll
                for (int j = 0; j < PF; ++j) load_x(j, kr, j);
            }
        }

        for (int k = 0; k < cnt; ++k) {
            const int depth = S / NW;
            const int slot = (k % depth) * NW + w;
            // k index of this warp's next run (the last run refills with its own data: harmless, keeps loads uniform)
            const int kr_next = (k + 1 < cnt) ? ((kr + 1 == p.rps) ? 0 : kr + 1) : kr;
            mbar_wait(&full[slot], (k / depth) & 1);
            if (k == 0) ST_TRACE(3);
            const unsigned char* wt = wst + size_t(slot) * ST_TILE_BYTES;
            const unsigned char* sz = szst + size_t(slot) * ST_SZ_BYTES;
#pragma unroll
            for (int f = 0; f < NSEG; ++f) {
                float D[MT][2][NACC][4];
                float D1[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) D1[mt][q] = 0.f;
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int a = 0; a < NACC; ++a)
#pragma unroll
                            for (int q = 0; q < 4; ++q) D[mt][t][a][q] = 0.f;
                }
#pragma unroll
                for (int jj = 0; jj < FJ; ++jj) {
                    const int j = f * FJ + jj;
                    const int e = j % PF;
                    const int rl = 8 * (j >> 1) + 2 * c + (j & 1);        // local packed row of this lane
                    const uint4 wv = *reinterpret_cast<const uint4*>(wt + rl * 128 + ((r ^ (rl & 7)) << 4));
                    uint32_t xb[MT][XR];
                    if constexpr (XS) {
                        (void)e;
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint16_t* xp = xbase[mt] + size_t(uint32_t(kr) * xrun_stride[mt]) +
                                                 (8 * (j >> 1) + (j & 1)) * NB;
                            if constexpr (XR >= 4) {
#pragma unroll
                                for (int v4 = 0; v4 < XR / 4; ++v4) {
                                    const uint4 v = *reinterpret_cast<const uint4*>(xp + v4 * 8);
                                    xb[mt][v4 * 4 + 0] = v.x; xb[mt][v4 * 4 + 1] = v.y;
                                    xb[mt][v4 * 4 + 2] = v.z; xb[mt][v4 * 4 + 3] = v.w;
                                }
                            } else {
                                const uint2 v = *reinterpret_cast<const uint2*>(xp);
                                xb[mt][0] = v.x; xb[mt][1] = v.y;
                            }
                        }
                    } else {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                            for (int v = 0; v < XR; ++v) {
                                const int ka = mma_kperm<BITS>(2 * v), kb = mma_kperm<BITS>(2 * v + 1);
                                const uint32_t sel = ((ka & 1) ? 0x32u : 0x10u) | (((kb & 1) ? 0x76u : 0x54u) << 8);
                                xb[mt][v] = __byte_perm(xq[mt][e][ka >> 1], xq[mt][e][kb >> 1], sel);
                            }
                        // refill the ring entry: same run (j + PF < 8) or the next run of this warp
                        if (j + PF < 8) load_x(e, kr, j + PF);
                        else load_x(e, kr_next, j + PF - 8);
                    }
                    const uint32_t tx = wv.x >> 8, ty = wv.y >> 8, tz = wv.z >> 8, tw = wv.w >> 8;
#pragma unroll
                    for (int a = 0; a < NACC; ++a) {
                        const uint32_t m2 = (FM << (a * BITS)) | (FM << (a * BITS + 16));
                        const uint32_t a0 = wv.x & m2, a1 = wv.y & m2, a2 = tx & m2, a3 = ty & m2;   // tile 0
                        const uint32_t c0 = wv.z & m2, c1 = wv.w & m2, c2 = tz & m2, c3 = tw & m2;   // tile 1
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            mma_16816(D[mt][0][a], a0, a1, a2, a3, xb[mt][2 * a], xb[mt][2 * a + 1]);
                            mma_16816(D[mt][1][a], c0, c1, c2, c3, xb[mt][2 * a], xb[mt][2 * a + 1]);
                            if constexpr (!XS)
                                mma_16816(D1[mt], 0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u, xb[mt][2 * a],
                                          xb[mt][2 * a + 1]);   // sum of x over the same k slots
                        }
                    }
                }
                // ---- flush segment f through its group's affine parameters ----
                const int gs = (p.ngr > 1) ? f : 0;
                const uint2 s4 = *reinterpret_cast<const uint2*>(sz + gs * 64 + r * 8);
                uint2 z4;
                if (p.asym) {
                    const uint32_t zw = *reinterpret_cast<const uint32_t*>(sz + 256 + gs * (128 / NB) + ((4 * r) / NB) * 4);
                    z4 = make_uint2(zw >> (((4 * r) % NB) * BITS), 0u);
                } else {
                    z4 = *reinterpret_cast<const uint2*>(sz + 256 + gs * 64 + r * 8);
                }
                const uint32_t s2[2] = {s4.x, s4.y}, z2[2] = {z4.x, z4.y};
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const float sf = (e4 & 1) ? cvt16_hi<false>(s2[e4 >> 1]) : cvt16_lo<false>(s2[e4 >> 1]);
                    float zf;
                    if (p.asym) zf = sf * float(((z2[0] >> (e4 * BITS)) & FM) + 1u);
                    else zf = (e4 & 1) ? cvt16_hi<false>(z2[e4 >> 1]) : cvt16_lo<false>(z2[e4 >> 1]);
                    const float smul = sf * 16777216.0f;   // codes carry 2^-24
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float t = D[mt][e4 >> 1][NACC - 1][(e4 & 1) * 2 + h];
#pragma unroll
                            for (int a = NACC - 2; a >= 0; --a)
                                t = fmaf(t, 1.0f / float(1 << BITS), D[mt][e4 >> 1][a][(e4 & 1) * 2 + h]);
                            float xsum;
                            if constexpr (XS) {
                                const int mrow = mt * 8 + 2 * c + h;
                                xsum = (mrow < p.M) ? xseg[mrow * (krows / SEG_ROWS) + kr * NSEG + f] : 0.f;
                            } else {
                                xsum = D1[mt][h];
                            }
                            yacc[mt][e4][h] = fmaf(smul, t, yacc[mt][e4][h]);
                            yacc[mt][e4][h] = fmaf(-zf, xsum, yacc[mt][e4][h]);
                        }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            if (k == 0) ST_TRACE(4);
            // ---- advance; close the partial when the strip ends ----
            if (++kr == p.rps) {
                store_partial();
                kr = 0;
                ++strip;
            }
        }
        if (cnt > 0 && kr != 0) store_partial();
        if (lane < p.maxseg && lane >= seg) seg_strip[w * p.maxseg + lane] = -1;   // unused slots
        ST_TRACE(5);
    }
    __syncthreads();
    ST_TRACE(6);

    // =========================== CTA-level fixed-order reduction and output ===========================
    const int ns = s_hi - s_lo;
    const int per_strip = p.M * 32;
    for (int o = tid; o < ns * per_strip; o += blockDim.x) {
        const int srel = o / per_strip, rm = o - srel * per_strip;
        const int s = s_lo + srel;
        float sum = 0.f;
        for (int ws = 0; ws < NW * p.maxseg; ++ws)       // fixed order, independent loads
            if (seg_strip[ws] == s) sum += part[size_t(ws) * per_strip + rm];
        p.y[size_t(rm >> 5) * p.N + s * 32 + (rm & 31)] = f32_to_16<false>(sum);
    }
    ST_TRACE(7);
}

struct StreamLaunch {
    int MT, FJ, warps, grid, xs;
    size_t smem;
    unsigned flags;
    cudaStream_t stream;
};
template <int BITS>
int launch_stream_family(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const StreamParams& p,
                         const StreamLaunch& l);

}  // namespace b200bit
