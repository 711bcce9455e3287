// mpq_pipe_mma.cuh -- the cross-kernel pipelined decode kernel (same idea as mpq_pipe.cuh: every weight byte of the CTA
// is requested by TMA BEFORE griddepcontrol.wait, three layers resident per SM) with the consumer math on mma.sync
// and NO global load inside the math loop.  4-bit, fp16, M == 1.
//
// Why (measured, profiles/r40_pipe_timeline.txt, r41_pipe_timeline.txt): everything after griddepcontrol.wait sits on
// the token's dependent chain.  The CUDA-core loop needs ~2.5 issue slots per weight (2.2 - 3.6 us per 4096x4096
// layer); and any global load on the chain costs an L2 round trip of ~0.45 us while HBM is saturated by the prefetch
// of the next layers -- a per-unit register ring of activations paid that latency once per 16 packed rows.  So:
//   * strips are 28 columns wide (112-byte tile rows): 4096 / 28 -> 147 CTAs for 148 SMs, 11008 / 28 -> 394 <= 3 x 148,
//     and 4 stages x 14 KB of packed words + scales/zeros + the CTA's activations fit 75 KB, i.e. 3 CTAs per SM;
//   * warp 0 issues every TMA of the CTA up front (the K range always fits the ring: no refill, no empty barriers,
//     no producer warp: 256 threads x 80 registers x 3 CTAs);
//   * after the wait each warp loads the activations of ITS OWN units with two 128-bit loads per lane (one L2 round
//     trip for the whole CTA), writes them to shared memory already permuted into B-fragment order, and reduces the
//     per-group sums of x with shuffles -- warp-private, so no block barrier in front of the math;
//   * masked words go to the legacy tensor pipe as fp16-subnormal A fragments (exact products, fp32 accumulation:
//     numerically identical to the FHFMA path): 0.5 LOP3 + 1/8 SHF per weight, one HMMA per 256 weights.  Only column
//     0 of the 8-wide B operand carries the activation vector; the other seven are don't-care.
//
// Fragment mapping (m16n8k16, g = lane >> 2, c = lane & 3; half-unit = 8 packed rows of the strip):
//   thread loads W1 = words (row base + c, cols 4g .. 4g+3), W2 = words (row base + 4 + c, same cols)   [2 LDS.128]
//   for nibble pair f (k = f and f + 4 of a packed row; mask 0x000f000f << 4(f&1) on w or w >> 8), i in {0, 1}:
//     A row g   <- column 4g + i      a0 = W1[i]   & m  (k-slots 2c, 2c+1  = row base+c,   nibbles f, f+4)
//     A row g+8 <- column 4g + 2 + i  a1 = W1[i+2] & m
//                                     a2 = W2[i]   & m  (k-slots 2c+8,2c+9 = row base+4+c, nibbles f, f+4)
//                                     a3 = W2[i+2] & m
//     B column 0: b0 = {x[8(base+c) + f], x[8(base+c) + f + 4]}, b1 = same for row base + 4 + c   [1 LDS.128 each]
//   D[g][0] (d0) and D[g+8][0] (d2) of the lanes with c == 0 are the partial outputs of columns 4g+i and 4g+2+i;
//   lanes with g == 7 work on columns 28..31, which do not belong to the strip (ignored).
// Replaces quant_mm_kernel{,_asym} (bitorch_engine/layers/qlinear/nbit/cuda/mpq_linear_cuda_kernel.cu:67-451).
#pragma once
#include "mpq_pipe.cuh"

namespace b200bit {

constexpr int PGM_THREADS = PG_WARPS * 32;   // 8 consumer warps, warp 0 doubles as the TMA issuer
constexpr int PGM_COLS = 28;                 // strip width
constexpr int PGM_PITCH = PGM_COLS * 4;      // bytes per tile row
constexpr int PGM_TILE_BYTES = PG_STAGE_ROWS * PGM_PITCH;            // 14336
constexpr int PGM_X_BYTES = PG_MAX_STAGES * PG_STAGE_ROWS * 16;     // activations of the CTA's K range: 8 KB

#define PGM_TRACE(slot_) do { if constexpr (TRACE) { PG_TRACE(slot_); } } while (0)

__device__ __forceinline__ void pg_mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 4-bit, fp16.  FS2 = half-units (8 packed rows) per flush through the group's affine parameters: 1 (rpg == 8) or 2.
template <int FS2, bool ASYM, bool TRACE>
__global__ void __launch_bounds__(PGM_THREADS, 3) mpq_pipe_mma_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                                      const __grid_constant__ CUtensorMap tm_s,
                                                                      const __grid_constant__ CUtensorMap tm_z,
                                                                      const PipeParams p) {
    constexpr int NB = 8;
    constexpr uint32_t ONES = 0x3C003C00u;

    extern __shared__ __align__(1024) unsigned char pg_smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.S;
    // carve-up: x [8 warps][4 units][16 rows][16 B] | W ring S x 14336 | scale/zero tiles S x 2 x sz_bytes | mbarriers | red
    unsigned char* xs = pg_smem;
    unsigned char* wst = xs + PGM_X_BYTES;
    unsigned char* szst = wst + size_t(S) * PGM_TILE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(szst + size_t(S) * 2 * p.sz_bytes);
    float* red = reinterpret_cast<float*>(full + PG_MAX_STAGES);         // [PG_WARPS][32]

    const int strip = blockIdx.x;
    const int n0 = strip * PGM_COLS;
    const int st_lo = blockIdx.y * p.stages_per_split;
    const int nst = min(p.stages_per_split, p.stages_total - st_lo);
    const int r0 = st_lo * PG_STAGE_ROWS;
    const int rows_cta = min(p.R - r0, nst * PG_STAGE_ROWS);
    const int units_cta = rows_cta / PG_UNIT_ROWS;

    PGM_TRACE(0);
    if constexpr (TRACE) {      // slot 6 (the ticket stamp of split-K launches) carries the SM id otherwise
        if (p.trace && tid == 0 && gridDim.y == 1) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[size_t(blockIdx.x) * 8 + 6] = smid + 1;
        }
    }
    if (tid < S) mbar_init(&full[tid], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.early) pdl_launch_dependents();     // see PipeParams::early
    __syncthreads();

    if (warp == 0) {
        // ---- the whole K range of the CTA fits the ring (host: stages_per_split <= S): warp 0 requests every stage now,
        //      before griddepcontrol.wait; converged warp, the elected lane issues ----
        const uint32_t leader = um_elect();
        const unsigned bytes = unsigned(PGM_TILE_BYTES) + unsigned(p.s_tile_bytes) + unsigned(p.z_tile_bytes);
        for (int it = 0; it < nst; ++it) {
            const int row = r0 + it * PG_STAGE_ROWS;
            const int g0 = (p.rpg_shift >= 0) ? (row >> p.rpg_shift) : (row / p.rpg);
            unsigned char* sz = szst + size_t(it) * 2 * p.sz_bytes;
            um_expect_tx(&full[it], bytes, leader);
            um_tma_2d(wst + size_t(it) * PGM_TILE_BYTES, &tm_w, n0, row, &full[it], leader);
            // TMA needs the box to start on a 16-byte boundary in global memory: the fp16 rows start at column n0 & ~7
            // (the strip then sits at column offset n0 & 7 in {0, 4} of the 32-column box), the packed zero words at
            // word (n0 >> 3) & ~3 (8-word box)
            um_tma_2d(sz, &tm_s, n0 & ~7, g0, &full[it], leader);
            um_tma_2d(sz + p.sz_bytes, &tm_z, ASYM ? ((n0 >> 3) & ~3) : (n0 & ~7), g0, &full[it], leader);
        }
    }

    const int g = lane >> 2, c = lane & 3;
    if (!p.early) {
        pdl_wait_primary();      // x is produced by the previous kernel; y / workspace may still be in use by it
        pdl_launch_dependents(); // only now: everything in front of this kernel is complete when its dependents start
    }
    PGM_TRACE(1);

    // ---- this warp's activations: units it = 0..3 are packed rows it*128 + warp*16 + (0..15) of the CTA's K range.
    //      lanes 0-15 take unit 2q, lanes 16-31 unit 2q+1 (q = 0, 1), one 16-byte row each; both loads are issued before
    //      either is used.  Rows go to shared memory permuted into B-fragment order {x0,x4,x1,x5,x2,x6,x3,x7}; the sum of
    //      each flush segment (FS2*8 rows) is reduced over the lanes that loaded it and stays in a register. ----
    const uint16_t* xg = p.x + size_t(r0) * NB;
    unsigned char* xs_w = xs + warp * (PG_MAX_STAGES * PG_UNIT_ROWS * 16);
    float xsum_pair[2];
    {
        const int hl = lane >> 4, r16 = lane & 15;
        uint4 v[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int it = 2 * q + hl;
            const bool valid = it < nst && it * PG_WARPS + warp < units_cta;
            // units this warp does not own are never read back: any valid address keeps the load unconditional
            const uint16_t* src = valid ? xg + size_t(it * PG_STAGE_ROWS + warp * PG_UNIT_ROWS + r16) * NB : xg;
            v[q] = ld_global_v4(src);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int it = 2 * q + hl;
            const uint32_t w4[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s0 = fhfma<false, false, false>(ONES, w4[e], s0);
                s1 = fhfma<false, false, true>(ONES, w4[e], s1);
            }
            float sum = s0 + s1;
#pragma unroll
            for (int off = 1; off < FS2 * 8; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            xsum_pair[q] = sum;
            uint4 o;
            o.x = __byte_perm(v[q].x, v[q].z, 0x5410); o.y = __byte_perm(v[q].x, v[q].z, 0x7632);
            o.z = __byte_perm(v[q].y, v[q].w, 0x5410); o.w = __byte_perm(v[q].y, v[q].w, 0x7632);
            *reinterpret_cast<uint4*>(xs_w + (it * PG_UNIT_ROWS + r16) * 16) = o;
        }
    }
    __syncwarp();
    PGM_TRACE(2);

    float yacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int it = 0; it < PG_MAX_STAGES; ++it) {       // fully unrolled: stage offsets become immediates
        if (it < nst && it * PG_WARPS + warp < units_cta) {
            mbar_wait(&full[it], 0);
            if (it == 0) PGM_TRACE(3);
            const unsigned char* wt = wst + it * PGM_TILE_BYTES + (warp * PG_UNIT_ROWS + c) * PGM_PITCH + g * 16;
            const unsigned char* xt = xs_w + (it * PG_UNIT_ROWS + c) * 16;
            const unsigned char* sz = szst + size_t(it) * 2 * p.sz_bytes;
            float D[2][2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int q = 0; q < 4; ++q) D[i][a][q] = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint4 w1 = *reinterpret_cast<const uint4*>(wt + h * 8 * PGM_PITCH);
                const uint4 w2 = *reinterpret_cast<const uint4*>(wt + (h * 8 + 4) * PGM_PITCH);
                const uint4 xa = *reinterpret_cast<const uint4*>(xt + h * 8 * 16);
                const uint4 xb = *reinterpret_cast<const uint4*>(xt + (h * 8 + 4) * 16);
                const uint32_t b0[4] = {xa.x, xa.y, xa.z, xa.w}, b1[4] = {xb.x, xb.y, xb.z, xb.w};
                const uint32_t W1[4] = {w1.x, w1.y, w1.z, w1.w}, W2[4] = {w2.x, w2.y, w2.z, w2.w};
                const uint32_t T1[4] = {w1.x >> 8, w1.y >> 8, w1.z >> 8, w1.w >> 8};
                const uint32_t T2[4] = {w2.x >> 8, w2.y >> 8, w2.z >> 8, w2.w >> 8};
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const uint32_t m = 0x000f000fu << (4 * (f & 1));
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint32_t a0 = (f < 2 ? W1[i] : T1[i]) & m, a1 = (f < 2 ? W1[i + 2] : T1[i + 2]) & m;
                        const uint32_t a2 = (f < 2 ? W2[i] : T2[i]) & m, a3 = (f < 2 ? W2[i + 2] : T2[i + 2]) & m;
                        pg_mma_16816(D[i][f & 1], a0, a1, a2, a3, b0[f], b1[f]);
                    }
                }
                if ((h + 1) % FS2 == 0) {
                    // ---- flush FS2 half-units (rows of one group) through the group's affine parameters ----
                    const int rs = warp * PG_UNIT_ROWS + (h + 1 - FS2) * 8;           // first row of the segment in the stage
                    const int gl = (p.rpg_shift >= 0) ? (rs >> p.rpg_shift) : 0;
                    const float xsum = __shfl_sync(0xffffffffu, xsum_pair[it >> 1], (it & 1) * 16 + (FS2 == 1 ? h * 8 : 0));
                    const int co = (n0 & 7) * 2;                                      // byte offset of the strip inside the fp16 tile rows
                    const uint2 s4 = *reinterpret_cast<const uint2*>(sz + gl * 64 + g * 8 + co);
                    uint2 z4;
                    if constexpr (ASYM) {
                        // zero tile row = 8 packed words from word (n0 >> 3) & ~3; the strip starts at nibble (n0 & 7) in {0, 4}
                        // of word (n0 >> 3) & 3 of the box
                        const int gq = g + ((n0 & 7) >> 2);
                        const uint32_t zw = *reinterpret_cast<const uint32_t*>(sz + p.sz_bytes + gl * 32 + (((n0 >> 3) & 3) + (gq >> 1)) * 4);
                        z4 = make_uint2(zw >> ((gq & 1) * 16), 0u);
                    } else {
                        z4 = *reinterpret_cast<const uint2*>(sz + p.sz_bytes + gl * 64 + g * 8 + co);
                    }
                    const uint32_t s2[2] = {s4.x, s4.y}, z2[2] = {z4.x, z4.y};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float sf = (j & 1) ? cvt16_hi<false>(s2[j >> 1]) : cvt16_lo<false>(s2[j >> 1]);
                        float zf;
                        if constexpr (ASYM) zf = sf * float(((z2[0] >> (j * 4)) & 15u) + 1u);
                        else zf = (j & 1) ? cvt16_hi<false>(z2[j >> 1]) : cvt16_lo<false>(z2[j >> 1]);
                        const int i = j & 1, q = (j >> 1) * 2;          // column 4g + j  <-  MMA i, D row g (q = 0) / g + 8 (q = 2)
                        const float t = fmaf(D[i][1][q], 1.0f / 16.0f, D[i][0][q]);
                        yacc[j] = fmaf(sf * 16777216.0f, t, yacc[j]);   // codes carry 2^-24
                        yacc[j] = fmaf(-zf, xsum, yacc[j]);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int q = 0; q < 4; ++q) D[i][a][q] = 0.f;
                }
            }
        }
    }
    PGM_TRACE(4);
    // lanes with c == 0 hold this warp's sums of columns 4g .. 4g+3
    if (c == 0) *reinterpret_cast<float4*>(red + warp * 32 + g * 4) = make_float4(yacc[0], yacc[1], yacc[2], yacc[3]);
    __syncthreads();

    // =========================== fixed-order CTA sum, output ===========================
    if (p.early) pdl_wait_primary();          // y / workspace may still be in use by the previous kernel
    PGM_TRACE(7);
    const int splitk = gridDim.y;
    const bool owner = tid < PGM_COLS && n0 + tid < p.N;
    if (owner) {
        float total = 0.f;
#pragma unroll
        for (int w = 0; w < PG_WARPS; ++w) total += red[w * 32 + tid];
        if (splitk == 1) p.y[n0 + tid] = f32_to_16<false>(total);
        else p.ws_part[size_t(blockIdx.y) * p.N + n0 + tid] = total;
    }
    PGM_TRACE(5);
    if (splitk == 1) return;
    // ---- deterministic split-K: the last CTA of the strip (ticket) sums the partials in split order ----
    __shared__ int s_last;
    if (tid < 32) __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&p.tickets[strip], 1u) == unsigned(splitk - 1));
    __syncthreads();
    if (!s_last) return;
    if (tid < 32) {
        __threadfence();
        if (owner) {
            float v = 0.f;
            for (int sp = 0; sp < splitk; ++sp) v += __ldcg(p.ws_part + size_t(sp) * p.N + n0 + tid);
            p.y[n0 + tid] = f32_to_16<false>(v);
        }
        if (tid == 0) p.tickets[strip] = 0u;
    }
    PGM_TRACE(6);
}

int launch_pipe_mma(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const PipeParams& p,
                    const PipeLaunch& l);   // l.FS = half-units per flush (1 or 2)

}  // namespace b200bit
