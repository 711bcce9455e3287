// mpq_mma.cuh -- small-batch (1 <= M <= 32) W{2,4,8} x A(fp16) kernel for sm_100a: same 128-bit streaming loads of
// the packed matrix as the CUDA-core GEMV, but the dot products go through warp-level mma.sync.m16n8k16 (f32
// accumulate), with the weights fed to the tensor core WITHOUT any int->fp conversion:
//
//   * thread (r = lane/4, c = lane%4) loads one uint4 = 4 adjacent columns (4r..4r+3 of the 32-column strip) of packed
//     row 4j+c.  Columns 4r, 4r+1 are rows r, r+8 of MMA tile 0; columns 4r+2, 4r+3 are rows r, r+8 of tile 1.
//   * a masked packed word IS an A-fragment register: (w & 0x000F000F) holds fields f and f+NF of the word as two
//     fp16 subnormals (value q * 2^-24); the K order inside the instruction is therefore a fixed permutation of k,
//     which is applied once to x when it is staged into shared memory (B fragments are then plain LDS.128).
//     Fields at bit position 4 (value 16*q*2^-24) use a second accumulator set and are folded in at flush time
//     (exact: no scaling of x is needed).  4-bit: 4 SHF + 16 LOP3 + 4 HMMA per 32 weights per thread.
//   * group affine factored out as in the GEMV:  y += s_g * (sum x q) - z_g * (sum x), per group, in fp32.
//   * M <= 8 rides in the n8 dimension for free; MT batch tiles cover M <= 32.
// Replaces, for 1 <= M <= 32, the reference's quant_mm_kernel (mpq_linear_cuda_kernel.cu:393-451) whose caller
// stops at 32 rows (mpq_layer.py:59).
#pragma once
#include "common.cuh"

namespace b200bit {

constexpr int MMA_U = 8;                 // j-steps (4 packed rows each) per warp run == uint4 loads in flight
constexpr int MMA_RUN_ROWS = 4 * MMA_U;  // 32 packed rows per warp run

struct MmaParams {
    const uint16_t* x;       // [M, K] f16
    const uint32_t* qw;      // [R, N]
    const uint16_t* scales;  // [G, N]
    const void* zeros;       // sym: u16 [G, N]; asym: u32 [G, N*bits/32]
    uint16_t* y;             // [M, N]
    float* ws_part;          // [splitk, M, N]
    unsigned* tickets;       // [N/32]
    int M, K, N, R, G;
    int runs_total;          // R / 32
    int runs_per_split;
    int rpr;                 // runs per group when the group spans >= 1 run (FJ == 8), else unused
    int rpr_shift;           // log2(rpr) or -1
    int asym;
};

__device__ __forceinline__ void mma_m16n8k16_f16f32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                    uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// k index (inside a packed row) of half `j` of the permuted x row: class a -> [a, a+NF, NACC+a, NACC+a+NF]
template <int BITS>
__host__ __device__ constexpr int mma_kperm(int j) {
    constexpr int NF = 16 / BITS;
    constexpr int NACC = BITS >= 8 ? 1 : 8 / BITS;
    const int a = j >> 2, q = j & 3;
    return (q == 0) ? a : (q == 1) ? a + NF : (q == 2) ? NACC + a : NACC + a + NF;
}

template <int BITS, int MT, int FJ>
__global__ void __launch_bounds__(256) mpq_mma_kernel(const MmaParams p) {
    constexpr int NB = 32 / BITS;
    constexpr int NACC = BITS >= 8 ? 1 : 8 / BITS;
    constexpr int XR = NB / 2;               // 32-bit x registers per packed row
    constexpr int U = MMA_U;
    constexpr int NSEG = U / FJ;
    constexpr int SEG_ROWS = 4 * FJ;          // packed rows per flush segment
    constexpr uint32_t FM = (1u << BITS) - 1u;

    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int r = lane >> 2, c = lane & 3;
    const int n0 = blockIdx.x * 32;
    const int col = n0 + 4 * r;

    const int run_lo = blockIdx.y * p.runs_per_split;
    const int nruns = min(p.runs_per_split, p.runs_total - run_lo);
    const int chunk_rows = nruns * MMA_RUN_ROWS;
    const int nsegs = nruns * NSEG;

    // smem: xs [M][m_stride] u16 (permuted x, per-m stride == 64 mod 128 bytes) | xseg [M][nsegs] f32 |
    //       red [nwarps][M][32] f32
    const int m_stride = chunk_rows * NB + 32;        // in halves; chunk_rows*NB*2 is a multiple of 128 bytes
    uint16_t* xs = reinterpret_cast<uint16_t*>(smem_raw);
    float* xseg = reinterpret_cast<float*>(smem_raw + size_t(p.M) * m_stride * 2);
    float* red = xseg + ((p.M * nsegs + 3) & ~3);

    const uint32_t row_bytes = uint32_t(p.N) * 4u;
    const char* wbase = reinterpret_cast<const char*>(p.qw) + (size_t(run_lo) * MMA_RUN_ROWS + c) * row_bytes +
                        size_t(col) * 4;

    uint4 W[U];
    uint2 S[NSEG], Z[NSEG];

    auto issue_loads = [&](int run_local) {
        const char* wp = wbase + size_t(run_local) * (MMA_RUN_ROWS * size_t(row_bytes));
#pragma unroll
        for (int j = 0; j < U; ++j) W[j] = ldg_stream_v4(wp + size_t(uint32_t(j) * 4u * row_bytes));
        const int run_g = run_lo + run_local;
        int g;
        if constexpr (FJ == U) g = (p.rpr_shift >= 0) ? (run_g >> p.rpr_shift) : (run_g / p.rpr);
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

    pdl_launch_dependents();
    if (warp < nruns) issue_loads(warp);
    pdl_wait_primary();

    // ---- stage x: one thread per (m, packed row): permute the row's NB activations into fragment order, and
    //      reduce the row sums over each flush segment (SEG_ROWS consecutive rows == consecutive lanes) ----
    {
        const int total = p.M * chunk_rows;                   // multiple of 32
        for (int i0 = warp * 32; i0 < total; i0 += blockDim.x) {
            const int i = i0 + lane;
            const int m = i / chunk_rows, row = i - m * chunk_rows;
            uint32_t in[XR];
            {
                const uint16_t* xg = p.x + size_t(m) * p.K + (size_t(run_lo) * MMA_RUN_ROWS + row) * NB;
                if constexpr (XR >= 4) {
#pragma unroll
                    for (int q = 0; q < XR / 4; ++q) {
                        const uint4 v = *reinterpret_cast<const uint4*>(xg + q * 8);
                        in[q * 4 + 0] = v.x; in[q * 4 + 1] = v.y; in[q * 4 + 2] = v.z; in[q * 4 + 3] = v.w;
                    }
                } else {
                    const uint2 v = *reinterpret_cast<const uint2*>(xg);
                    in[0] = v.x; in[1] = v.y;
                }
            }
            uint32_t out[XR];
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < XR; ++j) {
                const int ka = mma_kperm<BITS>(2 * j), kb = mma_kperm<BITS>(2 * j + 1);
                const uint32_t lo = (ka & 1) ? (in[ka >> 1] >> 16) : (in[ka >> 1] & 0xffffu);
                const uint32_t hi = (kb & 1) ? (in[kb >> 1] & 0xffff0000u) : (in[kb >> 1] << 16);
                out[j] = lo | hi;
                sum = fhfma<false, false, false>(0x3C003C00u, in[j], sum);
                sum = fhfma<false, true, true>(0x3C003C00u, in[j], sum);
            }
            uint16_t* dst = xs + m * m_stride + row * NB;
            if constexpr (XR >= 4) {
#pragma unroll
                for (int q = 0; q < XR / 4; ++q)
                    *reinterpret_cast<uint4*>(dst + q * 8) = make_uint4(out[q * 4], out[q * 4 + 1], out[q * 4 + 2], out[q * 4 + 3]);
            } else {
                *reinterpret_cast<uint2*>(dst) = make_uint2(out[0], out[1]);
            }
#pragma unroll
            for (int off = 1; off < SEG_ROWS; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if ((row % SEG_ROWS) == 0) xseg[m * nsegs + row / SEG_ROWS] = sum;
        }
    }
    __syncthreads();

    float yacc[MT][4][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) yacc[mt][e][0] = yacc[mt][e][1] = 0.f;

    for (int run_local = warp; run_local < nruns; run_local += nwarps) {
        if (run_local != warp) issue_loads(run_local);
        const uint16_t* xrun = xs + r * m_stride + (run_local * MMA_RUN_ROWS + c) * NB;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
            float D[MT][2][NACC][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int a = 0; a < NACC; ++a)
#pragma unroll
                        for (int q = 0; q < 4; ++q) D[mt][t][a][q] = 0.f;
#pragma unroll
            for (int jj = 0; jj < FJ; ++jj) {
                const int j = s * FJ + jj;
                const uint4 w = W[j];
                uint32_t xb[MT][XR];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    if (mt * 8 + r < p.M) {
                        const uint16_t* xr = xrun + mt * 8 * m_stride + j * 4 * NB;
                        if constexpr (XR >= 4) {
#pragma unroll
                            for (int q = 0; q < XR / 4; ++q) {
                                const uint4 v = *reinterpret_cast<const uint4*>(xr + q * 8);
                                xb[mt][q * 4 + 0] = v.x; xb[mt][q * 4 + 1] = v.y;
                                xb[mt][q * 4 + 2] = v.z; xb[mt][q * 4 + 3] = v.w;
                            }
                        } else {
                            const uint2 v = *reinterpret_cast<const uint2*>(xr);
                            xb[mt][0] = v.x; xb[mt][1] = v.y;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < XR; ++q) xb[mt][q] = 0u;
                    }
                }
                const uint32_t tx = w.x >> 8, ty = w.y >> 8, tz = w.z >> 8, tw = w.w >> 8;
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const uint32_t m2 = (FM << (a * BITS)) | (FM << (a * BITS + 16));
                    const uint32_t a0 = w.x & m2, a1 = w.y & m2, a2 = tx & m2, a3 = ty & m2;   // tile 0
                    const uint32_t c0 = w.z & m2, c1 = w.w & m2, c2 = tz & m2, c3 = tw & m2;   // tile 1
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_m16n8k16_f16f32(D[mt][0][a], a0, a1, a2, a3, xb[mt][2 * a], xb[mt][2 * a + 1]);
                        mma_m16n8k16_f16f32(D[mt][1][a], c0, c1, c2, c3, xb[mt][2 * a], xb[mt][2 * a + 1]);
                    }
                }
            }
            // ---- flush through the group's affine parameters ----
            const uint32_t s2[2] = {S[s].x, S[s].y};
            const uint32_t z2[2] = {Z[s].x, Z[s].y};
            const int seg = run_local * NSEG + s;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float sf = (e & 1) ? cvt16_hi<false>(s2[e >> 1]) : cvt16_lo<false>(s2[e >> 1]);
                float zf;
                if (p.asym) zf = sf * float(((z2[0] >> (e * BITS)) & FM) + 1u);
                else zf = (e & 1) ? cvt16_hi<false>(z2[e >> 1]) : cvt16_lo<false>(z2[e >> 1]);
                const float smul = sf * 16777216.0f;   // codes carry 2^-24
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // column e: tile e>>1, MMA row r (+8 if e&1) -> accumulator element (e&1)*2 + h
                        float t = D[mt][e >> 1][NACC - 1][(e & 1) * 2 + h];
#pragma unroll
                        for (int a = NACC - 2; a >= 0; --a)
                            t = fmaf(t, 1.0f / float(1 << BITS), D[mt][e >> 1][a][(e & 1) * 2 + h]);
                        const int mrow = mt * 8 + 2 * c + h;
                        const float xsum = (mrow < p.M) ? xseg[mrow * nsegs + seg] : 0.f;
                        yacc[mt][e][h] = fmaf(smul, t, yacc[mt][e][h]);
                        yacc[mt][e][h] = fmaf(-zf, xsum, yacc[mt][e][h]);
                    }
            }
        }
    }

    // ---- cross-warp reduction (fixed order) ----
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int mrow = mt * 8 + 2 * c + h;
                if (mrow < p.M) red[(warp * p.M + mrow) * 32 + 4 * r + e] = yacc[mt][e][h];
            }
    __syncthreads();

    const int splitk = gridDim.y;
    const int nout = p.M * 32;
    for (int o = tid; o < nout; o += blockDim.x) {
        const int om = o >> 5, oc = o & 31;
        float total = 0.f;
        for (int w = 0; w < nwarps; ++w) total += red[(w * p.M + om) * 32 + oc];
        if (splitk == 1) p.y[size_t(om) * p.N + n0 + oc] = f32_to_16<false>(total);
        else p.ws_part[(size_t(blockIdx.y) * p.M + om) * p.N + n0 + oc] = total;
    }
    if (splitk == 1) return;
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&p.tickets[blockIdx.x], 1u) == unsigned(splitk - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int o = tid; o < nout; o += blockDim.x) {
        const int om = o >> 5, oc = o & 31;
        float v = 0.f;
        for (int sp = 0; sp < splitk; ++sp) v += __ldcg(p.ws_part + (size_t(sp) * p.M + om) * p.N + n0 + oc);
        p.y[size_t(om) * p.N + n0 + oc] = f32_to_16<false>(v);
    }
    if (tid == 0) p.tickets[blockIdx.x] = 0u;
}

struct MmaLaunch {
    int MT, FJ, warps, splitk;
    size_t smem;
    unsigned flags;
    cudaStream_t stream;
};
template <int BITS>
int launch_mma_family(const MmaParams& p, const MmaLaunch& l);

}  // namespace b200bit
