// mpq_pipe.cuh -- decode GEMV (M == 1) built for CROSS-KERNEL pipelining on sm_100a.  CUDA cores only (FHFMA),
// warp-shuffle + shared-memory reduction, fp32 accumulation: the batch-1 shape never touches the tensor cores.
//
// Why another GEMV: at batch 1 a Llama-7B linear layer is 8.9 - 24 MB of packed weights, i.e. 1.4 - 3.7 us of HBM
// time.  A kernel that loads, then computes, then reduces inside that window leaves HBM idle for most of it, and the
// previous kernels of this library (mpq_gemv.cuh, mpq_stream.cuh) occupy a whole SM per CTA, so the next layer's CTAs
// cannot even start before the current layer's have left (profiles/r15_stream_kernel_timeline.txt).  This kernel is
// sized so that THREE layers are resident on every SM at once (288 threads x <= 72 registers, <= 75 KB of shared
// memory) and each CTA's packed weights fit its shared-memory ring whole:
//   * CTA = one 32-column strip x up to `stages_per_split` stages of 128 packed rows (16 KB each).  A producer warp
//     issues the TMA tile loads (cp.async.bulk.tensor.2d, mbarrier completion) for the weight tile and the matching
//     scale / zero rows of every stage BEFORE griddepcontrol.wait -- weights do not depend on the previous kernel --
//     so with programmatic dependent launch layer i+1 and i+2 stream their weights into shared memory while layer i
//     is still waiting for its activations, computing and reducing.  HBM never waits for the dependent chain.
//   * after the wait, 8 consumer warps each own one 16-row unit of every stage: LDS.128 of packed words (conflict-free,
//     512 contiguous bytes per warp), activations straight from global/L2 through a 4-deep register ring (no shared
//     memory staging, no block barrier in front of the math), the b-bit fields masked in place and used as fp16
//     subnormals by FHFMA (WordDot, mpq_gemv.cuh), group affine factored out:  y = sum_g s_g * sum x q - z_g * sum x.
//   * y is written exactly once; optional split-K is reduced deterministically (partials + ticket, fixed order).
// Replaces quant_mm_kernel{,_asym} (bitorch_engine/layers/qlinear/nbit/cuda/mpq_linear_cuda_kernel.cu:67-451) and
// the torch::zeros memset in front of it (:618).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "mpq_gemv.cuh"
#include "tma.cuh"

namespace b200bit {

constexpr int PG_STAGE_ROWS = 128;
constexpr int PG_TILE_BYTES = PG_STAGE_ROWS * 128;   // 32 columns x 4 B x 128 rows
constexpr int PG_UNIT_ROWS = 16;                     // rows one consumer warp takes from a stage (4 steps x 4 rows)
constexpr int PG_WARPS = 8;                          // consumer warps == units per stage
constexpr int PG_THREADS = (PG_WARPS + 1) * 32;      // + producer warp
constexpr int PG_MAX_STAGES = 4;                     // ring depth: 4 x 16 KB of packed weights per CTA

struct PipeParams {
    const uint16_t* x;       // [K] f16 / bf16 bits
    uint16_t* y;             // [N]
    float* ws_part;          // [splitk, N]   (splitk > 1)
    unsigned* tickets;       // [N / 32]      (splitk > 1)
    int K, N, R;             // R = K / nb packed rows
    int stages_total;        // ceil(R / 128)
    int stages_per_split;
    int S;                   // ring stages (<= PG_MAX_STAGES)
    int rpg;                 // packed rows per group
    int rpg_shift;           // log2(rpg) when rpg <= 128 (a power of two), else -1 (rpg % 128 == 0: one group per stage)
    int sz_bytes;            // bytes reserved per stage for the scale tile (same again for the zero tile), 128-aligned
    int s_tile_bytes, z_tile_bytes;   // bytes the two TMA boxes deliver
    int asym;
    // Programmatic-dependent-launch protocol (set by the host, mpq_forward.cu):
    //   early == 0 (x may be produced by the kernel in front): prefetch weights, griddepcontrol.wait, THEN
    //              launch_dependents, read x, compute, write y.  Triggering only after the wait guarantees that every
    //              kernel in front of this one is complete and flushed when this kernel's dependents start.
    //   early == 1 (x was complete before the previous b200bit kernel of this stream was launched -- a "sibling" such as
    //              k_proj after q_proj): launch_dependents at once, read x and compute BEFORE griddepcontrol.wait
    //              (overlapping the kernels in front), wait, then write y.  The final wait also keeps completion
    //              transitive: this kernel cannot finish before the one in front of it.
    int early;
    unsigned long long* trace;   // optional [grid][8] globaltimer stamps (diagnostics; nullptr = off)
};

__device__ __forceinline__ uint4 ld_global_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint2 ld_global_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}

// activations of one packed row (nb k-values) -> XREGS packed registers
template <int XREGS>
__device__ __forceinline__ void pg_load_x(uint32_t (&dst)[XREGS], const uint16_t* src) {
    if constexpr (XREGS >= 4) {
#pragma unroll
        for (int q = 0; q < XREGS / 4; ++q) {
            const uint4 v = ld_global_v4(src + q * 8);
            dst[q * 4 + 0] = v.x; dst[q * 4 + 1] = v.y; dst[q * 4 + 2] = v.z; dst[q * 4 + 3] = v.w;
        }
    } else {
        const uint2 v = ld_global_v2(src);
        dst[0] = v.x; dst[1] = v.y;
    }
}

#define PG_TRACE(slot_) do { if (p.trace && tid == 0) p.trace[(size_t(blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot_)] = st_gtime(); } while (0)

// FS = steps (of 4 packed rows) between flushes through the group's affine parameters: min(rpg, 16) / 4.
template <int BITS, bool BF16, int FS>
__global__ void __launch_bounds__(PG_THREADS, 3) mpq_pipe_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                                 const __grid_constant__ CUtensorMap tm_s,
                                                                 const __grid_constant__ CUtensorMap tm_z,
                                                                 const PipeParams p) {
    using WD = WordDot<BITS, BF16>;
    constexpr int NB = WD::NB;
    constexpr int XREGS = WD::XREGS;
    constexpr int NACC = WD::NACC;
    constexpr int D = (BITS == 2) ? 2 : 4;                // activation register ring depth (steps)
    constexpr int XCNT = FS * NB / 4;                     // 16-element x segments per flush segment (1..16)
    constexpr uint32_t FM = (1u << BITS) - 1u;
    constexpr uint32_t ONES = BF16 ? 0x3F803F80u : 0x3C003C00u;

    extern __shared__ __align__(1024) unsigned char pg_smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.S;
    unsigned char* wst = pg_smem;                                        // S x 16 KB
    unsigned char* szst = wst + size_t(S) * PG_TILE_BYTES;                // S x 2 x sz_bytes
    uint64_t* full = reinterpret_cast<uint64_t*>(szst + size_t(S) * 2 * p.sz_bytes);
    uint64_t* empty = full + PG_MAX_STAGES;
    float* red = reinterpret_cast<float*>(empty + PG_MAX_STAGES);         // [PG_WARPS][32]
    float* xseg = red + PG_WARPS * 32;                                    // [rows_cta * NB / 16]

    const int strip = blockIdx.x;
    const int st_lo = blockIdx.y * p.stages_per_split;
    const int nst = min(p.stages_per_split, p.stages_total - st_lo);
    const int r0 = st_lo * PG_STAGE_ROWS;
    const int rows_cta = min(p.R - r0, nst * PG_STAGE_ROWS);
    const int units_cta = rows_cta / PG_UNIT_ROWS;

    PG_TRACE(0);
    if (tid < S) mbar_init(&full[tid], 1);
    else if (tid >= 32 && tid < 32 + S) mbar_init(&empty[tid - 32], PG_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.early) pdl_launch_dependents();     // see PipeParams::early
    __syncthreads();

    if (warp == PG_WARPS) {
        // =========================== producer: one converged warp, the elected lane issues every TMA ===========================
        const uint32_t leader = um_elect();
        const unsigned bytes = unsigned(PG_TILE_BYTES) + unsigned(p.s_tile_bytes) + unsigned(p.z_tile_bytes);
        int s = 0, round = 0;
        for (int it = 0; it < nst; ++it) {
            if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
            const int row = r0 + it * PG_STAGE_ROWS;
            const int g0 = (p.rpg_shift >= 0) ? (row >> p.rpg_shift) : (row / p.rpg);
            unsigned char* sz = szst + size_t(s) * 2 * p.sz_bytes;
            um_expect_tx(&full[s], bytes, leader);
            um_tma_2d(wst + size_t(s) * PG_TILE_BYTES, &tm_w, strip * 32, row, &full[s], leader);
            um_tma_2d(sz, &tm_s, strip * 32, g0, &full[s], leader);
            um_tma_2d(sz + p.sz_bytes, &tm_z, p.asym ? strip * (32 / NB) : strip * 32, g0, &full[s], leader);
            if (++s == S) { s = 0; ++round; }
            if (!p.early && it == min(nst, S) - 1) {     // ring requested once: join the CTA's late trigger
                pdl_wait_primary();
                pdl_launch_dependents();
            }
        }
    } else {
        // =========================== consumers ===========================
        const int cq = lane & 7, rl = lane >> 3;
        if (!p.early) {
            pdl_wait_primary();      // x is produced by the previous kernel; y / workspace may still be in use by it
            pdl_launch_dependents();
        }
        PG_TRACE(1);

        const uint16_t* xg = p.x + size_t(r0) * NB;
        // this warp's x rows: unit (it * 8 + warp), step j -> packed row it*128 + warp*16 + j*4 + rl (relative to r0)
        const uint16_t* xrow = xg + size_t(warp * PG_UNIT_ROWS + rl) * NB;
        uint32_t XR[D][XREGS];
        if (warp < units_cta) {
#pragma unroll
            for (int j = 0; j < D; ++j) pg_load_x<XREGS>(XR[j], xrow + j * 4 * NB);
        }
        // ---- sums of x over 16-element segments (fp32, exact products with 1.0) -> shared memory ----
        {
            const int nseg = rows_cta * NB / 16;
            for (int i = tid; i < nseg; i += PG_WARPS * 32) {
                const uint4 a = ld_global_v4(xg + size_t(i) * 16);
                const uint4 b = ld_global_v4(xg + size_t(i) * 16 + 8);
                const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    s0 = fhfma<BF16, false, false>(ONES, v[q], s0);
                    s1 = fhfma<BF16, false, true>(ONES, v[q], s1);
                }
                xseg[i] = s0 + s1;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(PG_WARPS * 32) : "memory");   // consumer warps only
        PG_TRACE(2);

        float yacc[4] = {0.f, 0.f, 0.f, 0.f};
        int s = 0, ph = 0;
        for (int it = 0; it < nst; ++it) {
            if (it * PG_WARPS + warp < units_cta) {
                mbar_wait(&full[s], ph);
                if (it == 0) PG_TRACE(3);
                const unsigned char* wt = wst + size_t(s) * PG_TILE_BYTES + (warp * PG_UNIT_ROWS + rl) * 128 + cq * 16;
                const unsigned char* sz = szst + size_t(s) * 2 * p.sz_bytes;
                float acc[4][NACC];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int a = 0; a < NACC; ++a) acc[c][a] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 w = *reinterpret_cast<const uint4*>(wt + j * 512);
                    WD::run(w.x, XR[j % D], acc[0]);
                    WD::run(w.y, XR[j % D], acc[1]);
                    WD::run(w.z, XR[j % D], acc[2]);
                    WD::run(w.w, XR[j % D], acc[3]);
                    {   // refill the ring slot with the activations of step j + D (same unit or this warp's next unit)
                        const int it2 = it + ((j + D) >> 2), j2 = (j + D) & 3;
                        if (it2 < nst && it2 * PG_WARPS + warp < units_cta)
                            pg_load_x<XREGS>(XR[j % D], xrow + (size_t(it2) * PG_STAGE_ROWS + j2 * 4) * NB);
                    }
                    if ((j + 1) % FS == 0) {
                        // ---- flush the segment (FS*4 packed rows of one group) through the group's affine parameters ----
                        const int rs = warp * PG_UNIT_ROWS + (j + 1 - FS) * 4;      // first row of the segment in the stage
                        const int gl = (p.rpg_shift >= 0) ? (rs >> p.rpg_shift) : 0;
                        float xsum = xseg[(((it * PG_STAGE_ROWS + rs) * NB) >> 4) + (lane & (XCNT - 1))];
#pragma unroll
                        for (int off = 1; off < XCNT; off <<= 1) xsum += __shfl_xor_sync(0xffffffffu, xsum, off);
                        if (rl != 0) xsum = 0.f;     // the four row lanes are summed below: the z * sum(x) term counts once
                        const uint2 s4 = *reinterpret_cast<const uint2*>(sz + gl * 64 + cq * 8);
                        uint2 z4;
                        if (p.asym) {
                            const uint32_t zw = *reinterpret_cast<const uint32_t*>(sz + p.sz_bytes + gl * (128 / NB) + ((cq * 4) / NB) * 4);
                            z4 = make_uint2(zw >> (((cq * 4) % NB) * BITS), 0u);
                        } else {
                            z4 = *reinterpret_cast<const uint2*>(sz + p.sz_bytes + gl * 64 + cq * 8);
                        }
                        const uint32_t s2[2] = {s4.x, s4.y}, z2[2] = {z4.x, z4.y};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float sf = (c & 1) ? cvt16_hi<BF16>(s2[c >> 1]) : cvt16_lo<BF16>(s2[c >> 1]);
                            float zf;
                            if (p.asym) zf = sf * float(((z2[0] >> (c * BITS)) & FM) + 1u);
                            else zf = (c & 1) ? cvt16_hi<BF16>(z2[c >> 1]) : cvt16_lo<BF16>(z2[c >> 1]);
                            const float smul = BF16 ? sf : sf * 16777216.0f;     // fp16: codes carry 2^-24
                            const float zmul = BF16 ? fmaf(128.0f, sf, zf) : zf; // bf16: codes carry +128
                            const float t = WD::combine(acc[c]);
                            yacc[c] = fmaf(smul, t, yacc[c]);
                            yacc[c] = fmaf(-zmul, xsum, yacc[c]);
#pragma unroll
                            for (int a = 0; a < NACC; ++a) acc[c][a] = 0.f;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == S) { s = 0; ph ^= 1; }
        }
        PG_TRACE(4);
        // ---- reduce the 4 row lanes of the warp, park the warp's 32 column sums ----
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            yacc[c] += __shfl_xor_sync(0xffffffffu, yacc[c], 8);
            yacc[c] += __shfl_xor_sync(0xffffffffu, yacc[c], 16);
        }
        if (rl == 0) *reinterpret_cast<float4*>(red + warp * 32 + cq * 4) = make_float4(yacc[0], yacc[1], yacc[2], yacc[3]);
    }
    __syncthreads();

    // =========================== fixed-order CTA sum, output ===========================
    if (p.early) pdl_wait_primary();          // y / workspace may still be in use by the previous kernel
    const int splitk = gridDim.y;
    const int n0 = strip * 32;
    if (tid < 32) {
        float total = 0.f;
#pragma unroll
        for (int w = 0; w < PG_WARPS; ++w) total += red[w * 32 + tid];
        if (splitk == 1) p.y[n0 + tid] = f32_to_16<BF16>(total);
        else p.ws_part[size_t(blockIdx.y) * p.N + n0 + tid] = total;
    }
    PG_TRACE(5);
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
        float v = 0.f;
        for (int sp = 0; sp < splitk; ++sp) v += __ldcg(p.ws_part + size_t(sp) * p.N + n0 + tid);
        p.y[n0 + tid] = f32_to_16<BF16>(v);
        if (tid == 0) p.tickets[strip] = 0u;
    }
    PG_TRACE(6);
}

struct PipeLaunch {
    int FS, splitk, strips;
    size_t smem;
    unsigned flags;
    cudaStream_t stream;
};
template <int BITS, bool BF16>
int launch_pipe_family(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const PipeParams& p,
                       const PipeLaunch& l);

}  // namespace b200bit
