// mpq_gemv.cuh -- decode-shaped (M <= 4) W{1,2,4,8} x A16 GEMV for sm_100a.  CUDA cores only (no tensor cores),
// warp-shuffle + shared-memory reduction, fp32 accumulation.
//
// Replaces the reference's quant_mm_kernel / quant_mm_kernel_asym (bitorch_engine/layers/qlinear/nbit/cuda/
// mpq_linear_cuda_kernel.cu:67-451): one thread per column, 256-deep scalar dequant into a local array, fp16 HFMA
// accumulation and 16 half atomics per output.  This kernel instead
//   * reads the packed matrix with 128-bit streaming loads (4 adjacent columns x nb k-values per load, a lane group
//     of L lanes covers one 16*L-byte row segment, 32/L lane groups of a warp take different row runs);
//   * never converts a code to a float: the b-bit fields are masked IN PLACE inside each 16-bit lane of the packed
//     word and reinterpreted as fp16 *subnormals* (value = field * 2^-24 * 2^pos), then fed to Blackwell's
//     mixed-precision FMA  fma.rn.f32.f16 (SASS FHFMA: f32 += f16 * f16, product exact) together with the fp16
//     activation.  4-bit costs 1 shift + 4 LOP3 + 8 FHFMA per 8 weights.  bf16 activations use the bf16 form with a
//     128-biased normal number (bf16 subnormals would underflow fp32) and de-bias with the row sums of x;
//   * factors the group affine out of the inner loop:  y = sum_g [ s_g * sum_{k in g} x_k q_k  -  z_g * sum_{k in g} x_k ],
//     i.e. evaluates the *exact* quantised model in fp32 (no fp16 rounding of W; DESIGN.md "numerics");
//   * writes y exactly once (no memset, no atomics on y).  Optional split-K across CTAs is reduced
//     deterministically: partials to a workspace, last CTA of a column strip (ticket) sums them in fixed order.
#pragma once
#include "common.cuh"

namespace b200bit {

constexpr int GEMV_RUN = 8;        // packed rows per lane-group run == 128-bit loads in flight per thread
constexpr int GEMV_MAX_M = 4;

struct GemvParams {
    const uint16_t* x;       // [M, K] f16 / bf16 bits
    const uint32_t* qw;      // [R, N]
    const uint16_t* scales;  // [G, N]
    const void* zeros;       // sym: u16 [G, N]; asym: u32 [G, N*bits/32]
    uint16_t* y;             // [M, N]
    float* ws_part;          // [splitk, M, N] (splitk > 1)
    unsigned* tickets;       // [gridDim.x]    (splitk > 1)
    int K, N, R, G;
    int rpr;                 // runs per group (rpg / RUN) when FR == RUN, else unused
    int rpr_shift;           // log2(rpr) when it is a power of two, else -1
    int runs_total;          // ceil(R / GEMV_RUN)
    int runs_per_split;
    int L_log2;              // lanes per row segment = 1 << L_log2 (3, 4 or 5)
    int asym;
};

// ---------------------------------------------------------------------------------------------------------------
// field extraction + dot product of one packed word (one column, nb consecutive k) with the row's activations.
//   XH[j] holds activations (2j, 2j+1) of the row.  acc[a] accumulates fields whose in-lane position p has
//   (p % 8) / BITS == a; their weight is 2^(a*BITS) (folded back in combine()).
// ---------------------------------------------------------------------------------------------------------------
template <int BITS, bool BF16>
struct WordDot {
    static constexpr int NB = 32 / BITS;                       // k-values per word
    static constexpr int NF = 16 / BITS;                       // fields per 16-bit lane
    static constexpr int NACC = BF16 ? 1 : (BITS >= 8 ? 1 : 8 / BITS);
    static constexpr int XREGS = NB / 2;

    __device__ __forceinline__ static void run(uint32_t w, const uint32_t (&XH)[XREGS], float (&acc)[NACC]) {
        constexpr uint32_t FM = (1u << BITS) - 1u;
        if constexpr (!BF16) {
            const uint32_t t = w >> 8;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const int p = f * BITS;              // position inside the 16-bit lane
                const int pp = p & 7;                // position after the optional >> 8
                const uint32_t src = (p < 8) ? w : t;
                const uint32_t m2 = (FM << pp) | (FM << (pp + 16));
                const uint32_t a2 = src & m2;        // {field_f, field_(f+NF)} as fp16 subnormals * 2^pp
                const int a = pp / BITS;
                const int klo = f, khi = f + NF;     // k index (inside the row) of the low / high lane
                acc[a] = (klo & 1) ? fhfma<false, false, true>(a2, XH[klo >> 1], acc[a])
                                   : fhfma<false, false, false>(a2, XH[klo >> 1], acc[a]);
                acc[a] = (khi & 1) ? fhfma<false, true, true>(a2, XH[khi >> 1], acc[a])
                                   : fhfma<false, true, false>(a2, XH[khi >> 1], acc[a]);
            }
        } else {
            static_assert(!BF16 || BITS <= 4, "bf16 fast path needs the code to fit a 7-bit mantissa");
            constexpr uint32_t m2 = FM | (FM << 16);
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const uint32_t a2 = ((w >> (f * BITS)) & m2) | 0x43004300u;   // {128 + q_f, 128 + q_(f+NF)} bf16
                const int klo = f, khi = f + NF;
                acc[0] = (klo & 1) ? fhfma<true, false, true>(a2, XH[klo >> 1], acc[0])
                                   : fhfma<true, false, false>(a2, XH[klo >> 1], acc[0]);
                acc[0] = (khi & 1) ? fhfma<true, true, true>(a2, XH[khi >> 1], acc[0])
                                   : fhfma<true, true, false>(a2, XH[khi >> 1], acc[0]);
            }
        }
    }
    // sum_k x_k * q_k  (fp16: still scaled by 2^-24, folded into the scale; bf16: still biased by 128 * sum x)
    __device__ __forceinline__ static float combine(const float (&acc)[NACC]) {
        float t = acc[NACC - 1];
#pragma unroll
        for (int a = NACC - 2; a >= 0; --a) t = fmaf(t, 1.0f / float(1 << BITS), acc[a]);
        return t;
    }
};

// Shared-memory image of the x chunk: element e of the chunk lives at e + 8 * (e / (RUN*NB)), i.e. one 16-byte pad
// after every run, which keeps the lane groups of a warp (consecutive runs) on different banks.
//
// Host-side guarantees (plan_gemv): N % (4L) == 0, R % RUN == 0, group boundaries fall on run boundaries
// (rpg % RUN == 0) or FR = rpg divides RUN.
template <int BITS, bool BF16, int M, int FR>
__global__ void __launch_bounds__(512) mpq_gemv_kernel(const GemvParams p) {
    using WD = WordDot<BITS, BF16>;
    constexpr int NB = WD::NB;
    constexpr int RUN = GEMV_RUN;
    constexpr int NSEG = RUN / FR;
    constexpr int XREGS = WD::XREGS;
    constexpr int NACC = WD::NACC;
    constexpr int RUN_ELEMS = RUN * NB;          // activations per run
    constexpr int RUN_STRIDE = RUN_ELEMS + 8;    // + 16-byte pad

    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int L = 1 << p.L_log2, LG = 32 >> p.L_log2;
    const int li = lane & (L - 1), lg = lane >> p.L_log2;
    const int strip_shift = p.L_log2 + 2, strip_cols = 1 << strip_shift;
    const int col = (blockIdx.x << strip_shift) + li * 4;
    const int slots = nwarps * LG, slot = warp * LG + lg;

    const int run_lo = blockIdx.y * p.runs_per_split;
    const int nruns = min(p.runs_per_split, p.runs_total - run_lo);   // runs in this CTA's K chunk
    const int nsegs = nruns * NSEG;

    // smem carve-up: x [M][nruns*RUN_STRIDE] u16 | xseg [M][nsegs] f32 | red [nwarps][M][strip_cols] f32
    const int xs_stride = nruns * RUN_STRIDE;
    uint16_t* xs = reinterpret_cast<uint16_t*>(smem_raw);
    float* xseg = reinterpret_cast<float*>(smem_raw + size_t(M) * xs_stride * 2);
    float* red = xseg + ((M * nsegs + 3) & ~3);

    const size_t row_bytes = size_t(p.N) * 4;
    const char* wbase = reinterpret_cast<const char*>(p.qw + size_t(run_lo) * RUN * p.N + col);

    uint4 W[RUN];
    uint2 S[NSEG];
    uint2 Z[NSEG];

    auto issue_loads = [&](int run_local) {
        const char* wp = wbase + size_t(run_local) * RUN * row_bytes;
#pragma unroll
        for (int i = 0; i < RUN; ++i) W[i] = ldg_stream_v4(wp + i * row_bytes);
        const int run_g = run_lo + run_local;
        int g;
        if constexpr (FR == RUN) g = (p.rpr_shift >= 0) ? (run_g >> p.rpr_shift) : (run_g / p.rpr);
        else g = run_g * NSEG;
        const uint16_t* sp = p.scales + size_t(g) * p.N + col;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) S[s] = ldg_nc_v2(sp + size_t(s) * p.N);
        if (p.asym) {
            const int zw_n = p.N / NB;
            const uint32_t* zp = reinterpret_cast<const uint32_t*>(p.zeros) + size_t(g) * zw_n + col / NB;
#pragma unroll
            for (int s = 0; s < NSEG; ++s)
                Z[s] = make_uint2(ldg_nc_u32(zp + size_t(s) * zw_n) >> ((col % NB) * BITS), 0u);
        } else {
            const uint16_t* zp = reinterpret_cast<const uint16_t*>(p.zeros) + size_t(g) * p.N + col;
#pragma unroll
            for (int s = 0; s < NSEG; ++s) Z[s] = ldg_nc_v2(zp + size_t(s) * p.N);
        }
    };

    // ---- weights first: they do not depend on the previous kernel in the stream (PDL overlap) ----
    pdl_launch_dependents();
    if (slot < nruns) issue_loads(slot);

    pdl_wait_primary();   // x (and y / workspace) may be produced / consumed by the previous kernel

    // ---- stage the x chunk into shared memory (16-byte chunks) ----
    {
        const int chunk_elems = nruns * RUN_ELEMS;
        const uint16_t* xg = p.x + size_t(run_lo) * RUN_ELEMS;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            for (int e = tid * 8; e < chunk_elems; e += blockDim.x * 8) {
                const uint4 v = *reinterpret_cast<const uint4*>(xg + size_t(m) * p.K + e);
                *reinterpret_cast<uint4*>(xs + m * xs_stride + e + ((e / RUN_ELEMS) << 3)) = v;
            }
        }
    }
    __syncthreads();
    // ---- per-segment sums of x (fp32): one lane per packed row, FR consecutive lanes combine by shuffle ----
    {
        const int chunk_rows = nruns * RUN;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            for (int r0 = warp * 32; r0 < chunk_rows; r0 += blockDim.x) {
                const int r = r0 + lane;
                float sum = 0.f;
                if (r < chunk_rows) {
                    const uint16_t* xr = xs + m * xs_stride + (r / RUN) * RUN_STRIDE + (r % RUN) * NB;
#pragma unroll
                    for (int j = 0; j < NB; j += 2) {
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(xr + j);
                        sum += cvt16_lo<BF16>(v);
                        sum += cvt16_hi<BF16>(v);
                    }
                }
#pragma unroll
                for (int off = 1; off < FR; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                if (r < chunk_rows && (r % FR) == 0) xseg[m * nsegs + r / FR] = sum;
            }
        }
    }
    __syncthreads();

    float yacc[M][4];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) yacc[m][c] = 0.f;

    for (int run_local = slot; run_local < nruns; run_local += slots) {
        if (run_local != slot) issue_loads(run_local);
        const uint16_t* xrun = xs + run_local * RUN_STRIDE;
        const float* xsegp = xseg + run_local * NSEG;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
            float acc[M][4][NACC];
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int a = 0; a < NACC; ++a) acc[m][c][a] = 0.f;
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                const uint4 w = W[s * FR + i];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    uint32_t XH[XREGS];
                    const uint16_t* xr = xrun + m * xs_stride + (s * FR + i) * NB;
                    if constexpr (XREGS >= 4) {
#pragma unroll
                        for (int q = 0; q < XREGS / 4; ++q) {
                            const uint4 v = *reinterpret_cast<const uint4*>(xr + q * 8);
                            XH[q * 4 + 0] = v.x; XH[q * 4 + 1] = v.y; XH[q * 4 + 2] = v.z; XH[q * 4 + 3] = v.w;
                        }
                    } else {
                        const uint2 v = *reinterpret_cast<const uint2*>(xr);
                        XH[0] = v.x; XH[1] = v.y;
                    }
                    WD::run(w.x, XH, acc[m][0]);
                    WD::run(w.y, XH, acc[m][1]);
                    WD::run(w.z, XH, acc[m][2]);
                    WD::run(w.w, XH, acc[m][3]);
                }
            }
            // ---- flush the segment through its group's affine parameters ----
            float xsum[M];
#pragma unroll
            for (int m = 0; m < M; ++m) xsum[m] = xsegp[m * nsegs + s];
            const uint32_t s2[2] = {S[s].x, S[s].y};
            const uint32_t z2[2] = {Z[s].x, Z[s].y};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float sf = (c & 1) ? cvt16_hi<BF16>(s2[c >> 1]) : cvt16_lo<BF16>(s2[c >> 1]);
                float zf;
                if (p.asym) zf = sf * float(((z2[0] >> (c * BITS)) & ((1u << BITS) - 1u)) + 1u);
                else zf = (c & 1) ? cvt16_hi<BF16>(z2[c >> 1]) : cvt16_lo<BF16>(z2[c >> 1]);
                // fp16: codes carry 2^-24; bf16: codes carry +128
                const float smul = BF16 ? sf : sf * 16777216.0f;
                const float zmul = BF16 ? fmaf(128.0f, sf, zf) : zf;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float t = WD::combine(acc[m][c]);
                    yacc[m][c] = fmaf(smul, t, yacc[m][c]);
                    yacc[m][c] = fmaf(-zmul, xsum[m], yacc[m][c]);
                }
            }
        }
    }

    // ---- reduce: lane groups (shuffles) -> warps (shared memory, fixed order) ----
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = yacc[m][c];
            for (int off = L; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            yacc[m][c] = v;
        }
    if (lg == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m)
            *reinterpret_cast<float4*>(red + (((warp * M + m) << strip_shift) + li * 4)) =
                make_float4(yacc[m][0], yacc[m][1], yacc[m][2], yacc[m][3]);
    }
    __syncthreads();

    const int splitk = gridDim.y;
    const int nout = M << strip_shift;
    const int n0 = blockIdx.x << strip_shift;
    // fixed-order sum over warps; thread o owns output (m = o / strip_cols, n = n0 + o % strip_cols)
    for (int o = tid; o < nout; o += blockDim.x) {
        const int om = o >> strip_shift, oc = o & (strip_cols - 1);
        float total = 0.f;
        for (int w = 0; w < nwarps; ++w) total += red[((w * M + om) << strip_shift) + oc];
        if (splitk == 1) p.y[size_t(om) * p.N + n0 + oc] = f32_to_16<BF16>(total);
        else p.ws_part[(size_t(blockIdx.y) * M + om) * p.N + n0 + oc] = total;
    }
    if (splitk == 1) return;
    // ---- deterministic split-K: last CTA of the strip (ticket) sums the partials in split order ----
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&p.tickets[blockIdx.x], 1u) == unsigned(splitk - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int o = tid; o < nout; o += blockDim.x) {
        const int om = o >> strip_shift, oc = o & (strip_cols - 1);
        float v = 0.f;
        for (int sp = 0; sp < splitk; ++sp) v += __ldcg(p.ws_part + (size_t(sp) * M + om) * p.N + n0 + oc);
        p.y[size_t(om) * p.N + n0 + oc] = f32_to_16<BF16>(v);
    }
    if (tid == 0) p.tickets[blockIdx.x] = 0u;
}

// host-side launch of one (BITS, BF16) family; defined in mpq_gemv_b{1,2,4,8}.cu
struct GemvLaunch {
    int M, FR, warps, splitk;
    size_t smem;
    unsigned flags;
    cudaStream_t stream;
};
template <int BITS, bool BF16>
int launch_gemv_family(const GemvParams& p, const GemvLaunch& l);

}  // namespace b200bit
