// mpq_tc.cu -- batched forward on the 5th-generation tensor cores (tcgen05 / TMEM), ONE pass over the packed weight for
// any number of rows:   y[M,N] = x[M,K] @ W,   W = fp16(s*q - z) dequantised on the fly, fp32 accumulation in TMEM.
//
//   tile      : 128 output columns (UMMA M: one TMEM lane per column n) x 128 tokens (UMMA N), whole K; grid (N/128, M/128)
//   A operand : the weights, written STRAIGHT INTO TENSOR MEMORY by the dequant warps (tcgen05.st): a packed word of
//               column n (8 consecutive k) becomes four half2 registers = four TMEM columns of lane n.  No shared-memory
//               round trip for the 2 bytes a dequantised weight occupies (that round trip, not the MMA, bounds an
//               SS-mode kernel: 5 B of shared-memory traffic per weight against 128 B/clk).  Packed words come from
//               global memory with coalesced 128-byte warp loads (lane = column), prefetched eight stages ahead
//               in a register ring.
//   B operand : the activations, K-major as they lie in memory: one TMA box of 64 k x 128 tokens per stage, 128-byte
//               swizzled (the UMMA shared-memory descriptor names the same swizzle), 4-stage ring.
//   MMA       : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x 128 x 16), four per 64-k stage;
//               tcgen05.commit hands the A stage (TMEM) and the x stage (smem) back.
//   epilogue  : tcgen05.ld of the thread's own lane (= column n), fp16 stores (32 lanes = 64 contiguous bytes per token).
//   dequant   : a 4-bit code byte read as fp8 e4m3 IS q * 2^-9: cvt.f16x2.e4m3x2 makes two exact halves per instruction,
//               one HFMA2 with (512 s, -z) gives fp16(s*q - z) with a single rounding, i.e. exactly the operand the
//               reference's large-batch path feeds to cuBLAS (unpack_qweight + matmul, mpq_layer.py:59-63) -- which this
//               kernel replaces without ever writing the fp16 matrix to HBM.
// Conventions (instruction / shared-memory descriptors, TMEM addressing, tcgen05.st / ld shapes) are the ones verified on
// hardware by tools/umma_probe.cu (profiles/r21_*).  Every wait is bounded: a protocol error traps instead of hanging.
#include <cuda.h>
#include "tma.cuh"

namespace b200bit {

constexpr int TC_BN = 128;             // output columns per CTA (TMEM lanes)
constexpr int TC_KS = 64;              // k per stage
constexpr int TC_XSTAGES = 4;
constexpr int TC_ASTAGES = 8;         // two TMEM A slots per dequant warp group: it fills one while the MMAs read the other
constexpr int TC_DQ_WARPS = 16;        // warp w: TMEM lane quarter w & 3, group w >> 2 owns the stages s % 4 == group: four
                                       // stages are being dequantised at any time, so the tcgen05.st latency of one hides
                                       // behind the arithmetic of the others
constexpr int TC_THREADS = (TC_DQ_WARPS + 2) * 32;
constexpr int TC_TMEM_COLS = 512;      // D: BM columns (fp32 x BM tokens: 128 or 256) | A: 8 stages x 32 columns

struct TcParams {
    const uint32_t* qw;       // [K * bits / 32, N]
    const uint16_t* scales;   // [G, N] f16
    const void* zeros;        // sym: f16 [G, N]; asym: packed int32 [G, N * bits / 32]
    uint16_t* y;              // [M, N] f16
    int M, K, N, G;
    int gs, gs_shift;         // group size (k) = 1 << gs_shift
    int Gs;                   // rows of the shared-memory group table: most groups any one k-slice touches
    int splits;               // split-K factor (grid.z): > 1 when the tiles alone would leave most SMs idle (small M)
    float* part;              // [splits][M][N] fp32 partial results (splits > 1)
    unsigned* tickets;        // [tiles] zero-initialised, self-resetting (splits > 1)
    int asym;
};

__device__ __forceinline__ void tc_wait(uint64_t* bar, unsigned parity) {
    unsigned tries = 0;
    while (true) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (++tries > (1u << 24)) asm volatile("trap;");
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar, uint32_t leader) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc),
                 "r"(accumulate), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t addr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t addr, float (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                   "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
                   "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
                   "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31]) : "r"(addr));
}

// one packed word (8 four-bit codes, LSB first) -> four half2 registers (k0,k1) (k2,k3) (k4,k5) (k6,k7) of the weight
// s*q - z.  A code byte 0000qqqq read as an fp8 e4m3 number IS q * 2^-9 (gradual underflow: 0..7 subnormal, 8..15 in the
// first binade), so cvt.f16x2.e4m3x2 turns two codes into two exact halves in one instruction; one HFMA2 with
// s512 = (512 s, 512 s), mz = (-z, -z) gives fp16(s*q - z) with a single rounding.
__device__ __forceinline__ void tc_cvt_e4m3x4(uint32_t four, uint32_t& lo2, uint32_t& hi2) {
    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.rn.f16x2.e4m3x2 %0, a;\n\tcvt.rn.f16x2.e4m3x2 %1, b;\n\t}"
        : "=r"(lo2), "=r"(hi2) : "r"(four));
}
__device__ __forceinline__ void tc_dequant4(uint32_t w, uint32_t s512, uint32_t mz, uint32_t* out) {
    const uint32_t ev = w & 0x0f0f0f0fu;            // codes 0, 2, 4, 6 in bytes 0..3
    const uint32_t od = (w >> 4) & 0x0f0f0f0fu;     // codes 1, 3, 5, 7
    const uint32_t r01 = __byte_perm(ev, od, 0x5140);      // bytes (q0, q1, q2, q3)
    const uint32_t r23 = __byte_perm(ev, od, 0x7362);      // bytes (q4, q5, q6, q7)
    uint32_t h[4];
    tc_cvt_e4m3x4(r01, h[0], h[1]);
    tc_cvt_e4m3x4(r23, h[2], h[3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(out[c]) : "r"(h[c]), "r"(s512), "r"(mz));
}

// sixteen two-bit codes (LSB first) -> eight half2 registers (k0,k1) ... (k14,k15); same fp8 reinterpretation (0..3 are e4m3
// subnormals)
__device__ __forceinline__ void tc_dequant2(uint32_t w, uint32_t s512, uint32_t mz, uint32_t* out) {
    const uint32_t t0 = w & 0x03030303u, t1 = (w >> 2) & 0x03030303u, t2 = (w >> 4) & 0x03030303u, t3 = (w >> 6) & 0x03030303u;
    const uint32_t a = __byte_perm(t0, t1, 0x5140), b = __byte_perm(t0, t1, 0x7362);     // (q0,q1,q4,q5), (q8,q9,q12,q13)
    const uint32_t c = __byte_perm(t2, t3, 0x5140), d = __byte_perm(t2, t3, 0x7362);     // (q2,q3,q6,q7), (q10,q11,q14,q15)
    uint32_t h[8];
    tc_cvt_e4m3x4(a, h[0], h[2]);
    tc_cvt_e4m3x4(c, h[1], h[3]);
    tc_cvt_e4m3x4(b, h[4], h[6]);
    tc_cvt_e4m3x4(d, h[5], h[7]);
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(out[i]) : "r"(h[i]), "r"(s512), "r"(mz));
}

// TC_BM = tokens per CTA (UMMA N): 128, or 256 for large batches (a dequantised weight then feeds twice the MMA work)
template <int TC_BM, int BITS>
__global__ void __launch_bounds__(TC_THREADS, 1) mpq_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const TcParams p) {
    constexpr int WPS = 2 * BITS;                           // packed words of one column per 64-k stage
    constexpr int NB = 32 / BITS;                           // codes per word
    constexpr int TC_X_STAGE_BYTES = TC_BM * TC_KS * 2;     // 16 / 32 KB
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char* xst = tc_smem;                                             // x ring: 4 x 16 KB
    uint64_t* x_full = reinterpret_cast<uint64_t*>(xst + TC_XSTAGES * TC_X_STAGE_BYTES);
    uint64_t* x_empty = x_full + TC_XSTAGES;
    uint64_t* a_full = x_empty + TC_XSTAGES;
    uint64_t* a_empty = a_full + TC_ASTAGES;
    uint64_t* d_full = a_empty + TC_ASTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);
    // group parameters of the CTA's 128 columns, all groups: scales [G][128] f16 | zeros [G][128] f16 (sym) or [G][16] packed
    // words (asym).  Staged once: a global load per group change sat on the dequant warps' critical path (~1 us each).
    uint16_t* s_sm = reinterpret_cast<uint16_t*>(tc_smem + TC_XSTAGES * TC_X_STAGE_BYTES + 256);
    unsigned char* z_sm = reinterpret_cast<unsigned char*>(s_sm + size_t(p.Gs) * TC_BN);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
    // this CTA's share of the k-stages (split-K over grid.z)
    const int stages_all = p.K / TC_KS;
    const int s_beg = int((long long)stages_all * blockIdx.z / p.splits), s_end = int((long long)stages_all * (blockIdx.z + 1) / p.splits);
    const int stages = s_end - s_beg;
    // only the groups this CTA's k-range touches are staged (split-K keeps the table small: 344 groups of a 2-bit g32
    // 11008-row layer would not fit as a whole)
    const int g_lo = (s_beg * TC_KS) >> p.gs_shift;
    const int g_cnt = stages > 0 ? (((s_end * TC_KS - 1) >> p.gs_shift) - g_lo + 1) : 0;

    if (tid == 0) {
        for (int i = 0; i < TC_XSTAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < TC_ASTAGES; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
        mbar_init(d_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_DQ_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < g_cnt * TC_BN; i += TC_THREADS) {
        const int g = g_lo + i / TC_BN, c = i % TC_BN;
        const int col = n0 + c;
        s_sm[i] = col < p.N ? p.scales[size_t(g) * p.N + col] : uint16_t(0);
        if (!p.asym) reinterpret_cast<uint16_t*>(z_sm)[i] = col < p.N ? reinterpret_cast<const uint16_t*>(p.zeros)[size_t(g) * p.N + col] : uint16_t(0);
    }
    if (p.asym)
        for (int i = tid; i < g_cnt * (TC_BN / NB); i += TC_THREADS) {
            const int g = g_lo + i / (TC_BN / NB), c = i % (TC_BN / NB);
            const int wcol = n0 / NB + c;
            reinterpret_cast<uint32_t*>(z_sm)[i] = wcol < p.N / NB ? reinterpret_cast<const uint32_t*>(p.zeros)[size_t(g) * (p.N / NB) + wcol] : 0u;
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_d = tmem, tm_a = tmem + TC_BM;      // D: columns [0, BM); A stage s: columns BM + 32 s

    if (warp == TC_DQ_WARPS) {
        // ===================== TMA producer: x tiles =====================
        const uint32_t leader = um_elect();
        for (int s = 0; s < stages; ++s) {
            const int slot = s % TC_XSTAGES;
            if (s >= TC_XSTAGES) tc_wait(&x_empty[slot], ((s / TC_XSTAGES) - 1) & 1);
            um_expect_tx(&x_full[slot], TC_X_STAGE_BYTES, leader);
            // one box = 64 k (128 bytes) x 128 tokens, 128-byte swizzled: rows of 128 B, 8-row atoms of 1 KB.  (Boxes of 8 k
            // = 16-byte rows in the no-swizzle layout cost 1024 sixteen-byte requests per stage: 1800 clk per stage
            // measured, 7x the tensor time.)
            um_tma_2d(xst + slot * TC_X_STAGE_BYTES, &tm_x, (s_beg + s) * TC_KS, m0, &x_full[slot], leader);
        }
    } else if (warp == TC_DQ_WARPS + 1) {
        // ===================== MMA issuer =====================
        const uint32_t leader = um_elect();
        const uint32_t idesc = (1u << 4) | ((uint32_t(TC_BM) >> 3) << 17) | ((uint32_t(TC_BN) >> 4) << 24);
        // B tile of one MMA: 128 tokens x 16 k inside the 128-byte-swizzled stage: 8-token atoms 1024 B apart (SBO), the
        // k-step advances the start address by 32 bytes inside the swizzle atom; layout type 2 = SWIZZLE_128B
        const uint64_t desc_hi = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
        for (int s = 0; s < stages; ++s) {
            const int xs = s % TC_XSTAGES, as = s % TC_ASTAGES;
            tc_wait(&a_full[as], (s / TC_ASTAGES) & 1);
            tc_wait(&x_full[xs], (s / TC_XSTAGES) & 1);
            tc_fence_after();
            const uint32_t xbase = smem_u32(xst + xs * TC_X_STAGE_BYTES);
#pragma unroll
            for (int kk = 0; kk < TC_KS / 16; ++kk) {
                const uint64_t bdesc = desc_hi | uint64_t(((xbase + kk * 32) & 0x3FFFF) >> 4);
                tc_mma(tm_d, tm_a + as * 32 + kk * 8, bdesc, idesc, (s | kk) != 0 ? 1u : 0u, leader);
            }
            tc_commit(&a_empty[as], leader);
            tc_commit(&x_empty[xs], leader);
        }
        tc_commit(d_full, leader);
    } else {
        // ===================== dequant warps (then epilogue) =====================
        const int q4 = warp & 3, grp = warp >> 2;         // TMEM lane quarter; stages grp, grp + 4, ... (A slot = grp)
        static_assert(TC_ASTAGES == 8 && TC_DQ_WARPS == 16, "two A slots per dequant warp group");
        const int n = n0 + q4 * 32 + lane;
        const bool col_ok = n < p.N;
        const uint32_t* wp = p.qw + (col_ok ? n : 0);
        const uint32_t a_lane = tm_a + (uint32_t(q4 * 32) << 16);
        // packed words of the group's next two stages (8 global stages ahead) wait in registers: a stage is ~0.13 us of
        // tensor time, a global load ~1 us away
        constexpr int TC_PF = 2;
        uint32_t wr[TC_PF][WPS];
        auto load_words = [&](int s, uint32_t (&w)[WPS]) {
#pragma unroll
            for (int i = 0; i < WPS; ++i) w[i] = (col_ok && s < stages) ? ldg_nc_u32(wp + size_t((s_beg + s) * WPS + i) * p.N) : 0u;
        };
#pragma unroll
        for (int u = 0; u < TC_PF; ++u) load_words(grp + 4 * u, wr[u]);
        int g_prev = -1;
        uint32_t s512 = 0u, mz = 0u;
        const int cl = q4 * 32 + lane;                 // column inside the CTA's tile
        int use = 0;                                   // how often the group's A slot has been filled
        auto group_params = [&](int g_abs, uint32_t& s2, uint32_t& z2) {
            const int g = g_abs - g_lo;
            const float sf = __half2float(__ushort_as_half(s_sm[g * TC_BN + cl]));
            float zf;
            if (p.asym) zf = sf * float(((reinterpret_cast<const uint32_t*>(z_sm)[g * (TC_BN / NB) + cl / NB] >> ((cl % NB) * BITS)) & ((1u << BITS) - 1u)) + 1u);
            else zf = __half2float(__ushort_as_half(reinterpret_cast<const uint16_t*>(z_sm)[g * TC_BN + cl]));
            const uint32_t sh = __half_as_ushort(__float2half_rn(512.f * sf));
            const uint32_t zh = __half_as_ushort(__float2half_rn(-zf));
            s2 = sh | (sh << 16);
            z2 = zh | (zh << 16);
        };
        for (int s0 = grp; s0 < stages; s0 += 4 * TC_PF) {
#pragma unroll
            for (int u = 0; u < TC_PF; ++u) {
                const int s = s0 + 4 * u;
                if (s < stages) {
                    uint32_t regs[32];
                    // group parameters once per stage, not per word: a 64-k stage lies in one group (groups >= 64) or in
                    // two (32-k groups); groups are 32 * 2^i wide
                    const int gA = ((s_beg + s) * TC_KS) >> p.gs_shift, gB = ((s_beg + s) * TC_KS + 32) >> p.gs_shift;
                    if (gA != g_prev) {
                        g_prev = gA;
                        group_params(gA, s512, mz);
                    }
                    uint32_t s512b = s512, mzb = mz;
                    if (gB != gA) {
                        group_params(gB, s512b, mzb);
                        g_prev = -1;
                    }
#pragma unroll
                    for (int i = 0; i < WPS; ++i) {
                        if constexpr (BITS == 4) tc_dequant4(wr[u][i], i < 4 ? s512 : s512b, i < 4 ? mz : mzb, regs + 4 * i);
                        else tc_dequant2(wr[u][i], i < 2 ? s512 : s512b, i < 2 ? mz : mzb, regs + 8 * i);
                    }
                    load_words(s + 4 * TC_PF, wr[u]);
                    const int as = s % TC_ASTAGES;                 // grp or grp + 4
                    if (use >= 2) { tc_wait(&a_empty[as], ((use >> 1) - 1) & 1); tc_fence_after(); }
                    ++use;
                    tc_st16(a_lane + as * 32, reinterpret_cast<const uint32_t(&)[16]>(regs[0]));
                    tc_st16(a_lane + as * 32 + 16, reinterpret_cast<const uint32_t(&)[16]>(regs[16]));
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[as]);
                }
            }
        }
        // ---- epilogue: the thread's own TMEM lane = column n; warp group grp takes a quarter of the tokens ----
        tc_wait(d_full, 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < TC_BM / 128; ++c) {
            const int t0 = grp * (TC_BM / 4) + c * 32;
            float v[32];
            tc_ld32(tm_d + (uint32_t(q4 * 32) << 16) + t0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (col_ok) {
                if (p.splits == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int m = m0 + t0 + j;
                        if (m < p.M) p.y[size_t(m) * p.N + n] = __half_as_ushort(__float2half_rn(v[j]));
                    }
                } else {
                    float* dst = p.part + size_t(blockIdx.z) * p.M * p.N;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int m = m0 + t0 + j;
                        if (m < p.M) __stcg(dst + size_t(m) * p.N + n, v[j]);
                    }
                }
            }
        }
        if (p.splits > 1) {
            // split-K: the CTA that draws the last ticket of the tile sums the partial tiles in split order (deterministic)
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(TC_DQ_WARPS * 32) : "memory");
            unsigned* flag = reinterpret_cast<unsigned*>(tmem_slot) + 1;
            const int tile = blockIdx.y * gridDim.x + blockIdx.x;
            if (tid == 0) *flag = atomicAdd(p.tickets + tile, 1u);
            asm volatile("bar.sync 1, %0;" ::"n"(TC_DQ_WARPS * 32) : "memory");
            if (*flag == unsigned(p.splits) - 1u) {
                __threadfence();
                if (tid == 0) p.tickets[tile] = 0u;
                const int cols = min(TC_BN, p.N - n0), rows = min(TC_BM, p.M - m0);
                for (int i = tid; i < rows * (TC_BN / 2); i += TC_DQ_WARPS * 32) {
                    const int r = i / (TC_BN / 2), c2 = (i % (TC_BN / 2)) * 2;
                    if (c2 < cols) {
                        const size_t off = size_t(m0 + r) * p.N + n0 + c2;
                        float2 acc = make_float2(0.f, 0.f);
                        for (int z = 0; z < p.splits; ++z) {
                            const float2 t = __ldcg(reinterpret_cast<const float2*>(p.part + size_t(z) * p.M * p.N + off));
                            acc.x += t.x; acc.y += t.y;
                        }
                        *reinterpret_cast<__half2*>(p.y + off) = __floats2half2_rn(acc.x, acc.y);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_DQ_WARPS + 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
}

}  // namespace b200bit

using namespace b200bit;

namespace b200bit {
struct TcPlan { int BM, tiles, stages, splits, Gs; size_t smem; bool fits; };
// tile height, split-K factor, rows of the group table and shared memory of a launch (pure host arithmetic)
static TcPlan tc_plan(int M, int K, int N, int G, int w_bit, int asym, size_t workspace_bytes) {
    TcPlan pl{};
    const int tiles128 = ((N + TC_BN - 1) / TC_BN) * ((M + 127) / 128);
    // tokens per CTA: 128 while that still gives every SM at most one tile, 256 beyond (half the dequant work per MMA)
    pl.BM = (M > 128 && tiles128 > sm_count()) ? 256 : 128;
    pl.tiles = ((N + TC_BN - 1) / TC_BN) * ((M + pl.BM - 1) / pl.BM);
    pl.stages = K / TC_KS;
    // split-K when the tiles alone leave most SMs idle (M <= 128 on a 4096-column layer is 32 tiles): partial tiles in the
    // caller's workspace, the CTA that draws the last ticket of a tile adds them up in split order
    int splits = 1;
    while (splits < 8 && pl.tiles * splits * 2 <= sm_count() && pl.stages / (splits * 2) >= 8) splits *= 2;
    if (splits > 1) {
        const size_t need = size_t(B200BIT_WS_TICKET_BYTES) + size_t(splits) * M * N * sizeof(float);
        if (workspace_bytes < need || size_t(pl.tiles) * sizeof(unsigned) > B200BIT_WS_ZERO_OFFSET) splits = 1;
    }
    pl.splits = splits;
    const int gs = K / G;
    const int per_slice = (pl.stages + splits - 1) / splits;                   // most stages of one k-slice
    const int span = (per_slice * TC_KS + gs - 1) / gs + 1;                    // groups such a range can touch
    pl.Gs = span < G ? span : G;
    pl.smem = size_t(TC_XSTAGES) * pl.BM * TC_KS * 2 + 256 + size_t(pl.Gs) * TC_BN * 2 +
              (asym ? size_t(pl.Gs) * (TC_BN / (32 / w_bit)) * 4 : size_t(pl.Gs) * TC_BN * 2);
    pl.fits = pl.smem <= size_t(227) * 1024;
    return pl;
}
}  // namespace b200bit

// 1 when b200bit_mpq_forward_tc runs this problem (shape rules + the group table of a k-slice fits shared memory)
extern "C" int b200bit_mpq_forward_tc_supported(int M, int K, int N, int G, int w_bit, int asym, int dtype, size_t workspace_bytes) {
    if (!((w_bit == 4 || w_bit == 2) && dtype == B200BIT_F16)) return 0;
    if (!(M > 0 && K > 0 && N > 0 && G > 0 && K % G == 0 && K % TC_KS == 0 && N % 8 == 0 && (K / G) % 32 == 0)) return 0;
    const int gs = K / G;
    if ((gs & (gs - 1)) != 0) return 0;
    if (asym && N % (32 / w_bit) != 0) return 0;
    return tc_plan(M, K, N, G, w_bit, asym, workspace_bytes).fits ? 1 : 0;
}

// y[M,N] = x[M,K] @ dequant(qweight): 4-bit, f16, contiguous groups of 32*i values, K % 64 == 0, N % 8 == 0.
extern "C" int b200bit_mpq_forward_tc(const void* x, const int32_t* qweight, const void* scales, const void* zeros, void* y,
                                      int M, int K, int N, int G, int w_bit, int asym, int dtype, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(x && qweight && scales && zeros && y, B200BIT_ERR_ARG, "mpq_forward_tc: null pointer argument");
    B200_REQUIRE((w_bit == 4 || w_bit == 2) && dtype == B200BIT_F16, B200BIT_ERR_UNSUPPORTED, "mpq_forward_tc: w_bit=%d dtype code %d (2- / 4-bit, f16)", w_bit, dtype);
    B200_REQUIRE(!asym || N % (32 / w_bit) == 0, B200BIT_ERR_SHAPE, "mpq_forward_tc: asym needs N %% %d == 0", 32 / w_bit);
    B200_REQUIRE(M > 0 && K > 0 && N > 0 && G > 0 && K % G == 0 && K % TC_KS == 0 && N % 8 == 0 && (K / G) % 32 == 0, B200BIT_ERR_SHAPE,
                 "mpq_forward_tc: bad sizes M=%d K=%d N=%d G=%d (K %% 64 == 0, N %% 8 == 0, groups of 32*i)", M, K, N, G);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, B200BIT_ERR_ARG, "mpq_forward_tc: x must be 16-byte aligned");
    const TcPlan pl = tc_plan(M, K, N, G, w_bit, asym, workspace ? workspace_bytes : 0);
    B200_REQUIRE(pl.fits, B200BIT_ERR_UNSUPPORTED, "mpq_forward_tc: the %d groups of a k-slice do not fit the shared-memory parameter table", pl.Gs);
    const int BM = pl.BM;
    CUtensorMap tm_x;
    int rc = make_map_2d(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, x, uint64_t(K), uint64_t(M), uint64_t(K) * 2, TC_KS, uint32_t(BM),
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != B200BIT_OK) return rc;
    TcParams p{};
    p.qw = reinterpret_cast<const uint32_t*>(qweight);
    p.scales = reinterpret_cast<const uint16_t*>(scales);
    p.zeros = zeros;
    p.y = reinterpret_cast<uint16_t*>(y);
    p.M = M; p.K = K; p.N = N; p.G = G; p.gs = K / G; p.asym = asym;
    p.gs_shift = 0;
    while ((1 << p.gs_shift) < p.gs) ++p.gs_shift;
    B200_REQUIRE((1 << p.gs_shift) == p.gs, B200BIT_ERR_UNSUPPORTED, "mpq_forward_tc: group size %d is not a power of two", p.gs);
    const size_t smem = pl.smem;
    p.Gs = pl.Gs;
    static bool configured_dev[64] = {false};
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    if (!configured_dev[dev & 63]) {
        B200_CUDA_OK(cudaFuncSetAttribute(mpq_tc_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B200_CUDA_OK(cudaFuncSetAttribute(mpq_tc_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B200_CUDA_OK(cudaFuncSetAttribute(mpq_tc_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B200_CUDA_OK(cudaFuncSetAttribute(mpq_tc_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured_dev[dev & 63] = true;
    }
    const int splits = pl.splits;
    p.splits = splits;
    p.tickets = reinterpret_cast<unsigned*>(workspace);
    p.part = splits > 1 ? reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + B200BIT_WS_TICKET_BYTES) : nullptr;
    dim3 grid((N + TC_BN - 1) / TC_BN, (M + BM - 1) / BM, splits);
    if (w_bit == 4) {
        if (BM == 256) mpq_tc_kernel<256, 4><<<grid, TC_THREADS, smem, st>>>(tm_x, p);
        else mpq_tc_kernel<128, 4><<<grid, TC_THREADS, smem, st>>>(tm_x, p);
    } else {
        if (BM == 256) mpq_tc_kernel<256, 2><<<grid, TC_THREADS, smem, st>>>(tm_x, p);
        else mpq_tc_kernel<128, 2><<<grid, TC_THREADS, smem, st>>>(tm_x, p);
    }
    B200_CUDA_OK(cudaGetLastError());
    return B200BIT_OK;
}
