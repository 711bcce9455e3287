// mpq_imma.cuh -- decode GEMV (M == 1), 4-bit weights, f16 / bf16 activations: persistent, one CTA per SM, the math
// on the INTEGER tensor pipe (IMMA.16832.U8.S8) with the activations expanded into signed base-256 digits.
//
// Why (measured on B200, profiles/r50_*, r51_*):
//   * Every layer's work after griddepcontrol.wait sits on the token's dependent chain, and the fp16-subnormal kernels
//     (mpq_pipe.cuh: FHFMA, mpq_pipe_mma.cuh: HMMA.16816) spend ~224 issue slots per 4096 weights (4 LOP3 + 1 SHF per
//     packed word just to isolate the nibbles, 256 weights per HMMA): 40 - 52 GB/s per SM, i.e. the whole GPU computes
//     at about the speed HBM delivers, so compute cannot hide behind the weight stream and the chain is 2x HBM time.
//     A packed word is already a valid u8 A-fragment register -- as it is (byte = 16 x odd code + even code) and ANDed
//     with 0x0f0f0f0f (byte = even code): 1 LOP3 per word, 512 weights per IMMA, and IMMA.16832 issues at the same
//     0.5 / clk / SM as HMMA.16816 (tools/microbench4.cu) -- half the tensor time and a third of the issue slots.
//   * CTAs of one layer that land on the same SM finish 1.5x later than the median (r51_timeline_dump.txt: 52 of the
//     147 CTAs of a 4096x4096 layer shared an SM while 27 SMs had none).  grid = #SMs persistent CTAs of 512 threads
//     (4096 columns = 136 strips of 28 + 12 of 24 = 148), two resident per SM (the layer computing + the next one
//     prefetching): every SM always has exactly one free slot when the next layer launches, so placement stays one
//     CTA per layer per SM.  No split-K, no tickets, no workspace: a CTA walks whole 28-column strips (dealt out
//     cyclically, so that the CTAs of a layer read whole contiguous rows together) through a 3-slot TMA ring, refilled
//     by whichever warp releases a slot last, and writes y once.
//
// Exactness.  x (f16 / bf16) of a 128-value unit is scaled by a power of two to a 31-bit fixed-point integer X
// (odd k: 27 bits, the odd nibbles carry a factor 16) and split into four balanced base-256 digits in [-128, 127];
// digit d goes to column d of the 8-wide B operand.  Products and sums are integers (|sum| < 2^23 per 16 packed rows),
// the digits are recombined as integers (< 2^31) and converted to fp32 once per 16 rows: the only roundings are that
// conversion (2^-24) and the fp32 accumulation over the groups, as in the fp16 kernels.  An activation smaller than
// 2^-15 of its unit's maximum loses low bits (absolute error <= 2^-27 of that maximum).
//
// Fragment mapping (m16n8k32, g = lane >> 2, c = lane & 3), 8-row block b of the warp's 16-row unit, k-step j in {0,1}:
//   lane loads two LDS.64 of packed row 8b + 2c + j: columns 2g, 2g+1 (IMMA alpha) and 16+2g, 17+2g (IMMA beta) of the
//   strip.  With the 112-byte pitch the sixteen lanes of a half warp (g = 0..3 or 4..7, c = 0..3) touch all 32 banks
//   exactly once: shared memory moves every weight byte twice (TMA write + this read) and is the busiest unit of the SM.
//   IMMA alpha: A row g <- column 2g: a0 = word UNMASKED (byte = 16 x code of k = 1,3,5,7 + code of k = 0,2,4,6),
//               a2 = word & 0x0f0f0f0f (codes of k = 0,2,4,6); A row g+8 <- column 2g+1 (a1, a3).  One LOP3 per packed
//               word, and {raw, raw, masked, masked} is a register quad without moves.
//   B column (g & 3) = digit; with Xe = X[k even], Xo = X[k odd] / 16:  b0 = digit bytes of Xo, b1 = of (Xe - Xo), so
//               (16 code_o + code_e) Xo + code_e (Xe - Xo) = 16 code_o Xo + code_e Xe: the unmasked low nibbles cancel
//               Only lanes g < 4 load B (columns 4..7 of B are don't-care: their results are never read).
//   D: lanes c = 0 hold digit columns 0,1, lanes c = 1 digit columns 2,3 (c = 2,3: don't-care columns, weight 0).
// Replaces quant_mm_kernel{,_asym} (bitorch_engine/layers/qlinear/nbit/cuda/mpq_linear_cuda_kernel.cu:67-451) and the
// torch::zeros memset in front of it (:618).
#pragma once
#include "mpq_pipe.cuh"

namespace b200bit {

constexpr int IM_WARPS = 16;
constexpr int IM_THREADS = IM_WARPS * 32;
constexpr int IM_COLS = 28;                           // strip width
constexpr int IM_PITCH = IM_COLS * 4;                 // bytes per tile row
constexpr int IM_TILE_ROWS = 256;                     // packed rows per tile: 16 warps x 16-row unit
constexpr int IM_TILE_BYTES = IM_TILE_ROWS * IM_PITCH;   // 28672
constexpr int IM_UNIT_ROWS = 16;
constexpr int IM_MAX_STAGES = 3;
constexpr int IM_WP_BYTES = 1024 + 32 + 32;           // per warp: digit image of its unit in two consecutive tiles (2 x 512 B),
                                                      // sums of x per flush segment [2][4] f32, unit weight [2] f32 (+ pad)
constexpr int IM_XIMG_BYTES = IM_WARPS * IM_WP_BYTES;
constexpr int IM_SMEM_LIMIT = 112 * 1024;             // two CTAs per SM

struct ImmaParams {
    const uint16_t* x;       // [K] f16 / bf16 bits
    uint16_t* y;             // [N]
    int K, N, R;             // R = K / 8 packed rows
    int strips;              // strips of the layer: the first n28 are 28 columns wide, the others 24 (the last one may
    int n28;                 // be cut by N): 4096 -> 136 x 28 + 12 x 24 = 148 strips, one per SM (TMA boxes stay 28 wide)
    int strips_q, strips_r;  // strips / grid, strips % grid (CTA b walks strips b, b + grid, ...: strips_q + (b < strips_r) of them)
    int tiles;               // tiles per strip = ceil(R / 256)
    int S;                   // ring slots (<= IM_MAX_STAGES)
    int rpg_shift;           // log2(packed rows per group): groups are powers of two >= 32 values
    int sz_bytes;            // bytes reserved per slot for the scale tile (same again for the zero tile), 128-aligned
    int s_tile_bytes, z_tile_bytes;   // bytes the two TMA boxes deliver
    int early;               // see PipeParams::early
    unsigned long long* trace;
};

__device__ __forceinline__ void im_mma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void im_mma_z(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}

__device__ __forceinline__ uint4 im_lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ uint2 im_lds64(uint32_t a) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
// B fragment: loaded by the lanes with p != 0 only, the others keep their (don't-care) registers
__device__ __forceinline__ void im_lds64_if(uint2& r, uint32_t a, uint32_t p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.shared.v2.u32 {%0,%1}, [%2];\n\t}" : "+r"(r.x), "+r"(r.y) : "r"(a), "r"(p));
}
// ld.volatile: ptxas must not fuse two of these into one LDS.128 (the fused load forces register moves to build the
// {raw, raw, masked, masked} A-fragment quads)
__device__ __forceinline__ uint2 im_lds64v(uint32_t a) {
    uint2 r;
    asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t im_lds32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t im_lds16(uint32_t a) {
    uint32_t r;
    asm volatile("{\n\t.reg .b16 h;\n\tld.shared.u16 h, [%1];\n\tcvt.u32.u16 %0, h;\n\t}" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void im_mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "IM_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra IM_WAIT_DONE;\n\t"
        "bra IM_WAIT_LOOP;\n\t"
        "IM_WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned im_atoms_add(uint32_t a, unsigned v) {
    unsigned r;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ void im_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

#define IM_TRACE(slot_) do { if constexpr (TRACE) { if (p.trace && tid == 0) p.trace[size_t(blockIdx.x) * 8 + (slot_)] = st_gtime(); } } while (0)

// F = k-steps (4 packed rows each) between flushes through the group's affine parameters: min(rpg, 16) / 4.
template <int F, bool ASYM, bool BF16, bool TRACE>
__global__ void __launch_bounds__(IM_THREADS, 2) mpq_imma_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                                 const __grid_constant__ CUtensorMap tm_s,
                                                                 const __grid_constant__ CUtensorMap tm_z,
                                                                 const ImmaParams p) {
    // packed row (inside an 8-row block) that lane c reads in k-step j: 2c + j keeps the LDS.128 conflict-free; groups of
    // 4 packed rows (F == 1) need a k-step to be 4 consecutive rows instead (2-way conflicts, rare configuration)
    constexpr bool SEQ = (F == 1);
    extern __shared__ __align__(1024) unsigned char im_smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.S;
    // carve-up: W ring S x 28672 | scale / zero tiles S x 2 x sz_bytes | per warp {x digit image [2 tiles][512 B],
    //           xsum [2][4] f32, unit weight [2] f32} x 16 | red [2][16][32] f32 | mbarriers full[3] | released[3] u32
    unsigned char* wst = im_smem;
    unsigned char* szst = wst + size_t(S) * IM_TILE_BYTES;
    unsigned char* ximg = szst + size_t(S) * 2 * p.sz_bytes;
    float* red = reinterpret_cast<float*>(ximg + IM_XIMG_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * IM_WARPS * 32);
    unsigned* released = reinterpret_cast<unsigned*>(full + IM_MAX_STAGES);   // per slot: warps that are done with it (mod 16)

    // strips are dealt out cyclically (CTA b: b, b + grid, ...): at any moment the CTAs of a layer read ADJACENT strips,
    // i.e. together whole contiguous rows of the packed matrix.  Contiguous strip ranges per CTA leave two thirds of
    // every DRAM page untouched per visit: 2.9 TB/s instead of 6.4 TB/s on 4096x11008 (tools/microbench5.cu).
    const int s_lo = blockIdx.x, s_step = gridDim.x;
    const int tiles = p.tiles;
    const int T = (p.strips_q + (int(blockIdx.x) < p.strips_r ? 1 : 0)) * tiles;     // tiles of this CTA

    IM_TRACE(0);
    if constexpr (TRACE) {
        if (p.trace && tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[size_t(blockIdx.x) * 8 + 6] = smid + 1;
        }
    }
    if (tid < S) mbar_init(&full[tid], 1);
    else if (tid >= 32 && tid < 32 + S) released[tid - 32] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.early) pdl_launch_dependents();             // see PipeParams::early
    __syncthreads();

    const unsigned tile_tx = unsigned(IM_TILE_BYTES) + unsigned(p.s_tile_bytes) + unsigned(p.z_tile_bytes);
    const uint32_t leader = um_elect();
    // tile (strip, kt) -> ring slot: weight tile + the matching scale / zero rows.  Issued from converged code of warp 0
    // (the elected lane issues).  TMA boxes start on 16-byte boundaries: fp16 rows at column n0 & ~7 (the strip then
    // sits at column offset n0 & 7 in {0, 4} of the 32-column box), packed zero words at word (n0 >> 3) & ~3 (8-word box).
    auto strip_col = [&](int strip) { return strip < p.n28 ? strip * IM_COLS : p.n28 * IM_COLS + (strip - p.n28) * 24; };
    auto issue_tile = [&](int strip, int kt, int slot) {
        const int n0 = strip_col(strip);
        const int row = kt * IM_TILE_ROWS;
        const int g0 = row >> p.rpg_shift;
        unsigned char* sz = szst + size_t(slot) * 2 * p.sz_bytes;
        um_expect_tx(&full[slot], tile_tx, leader);
        um_tma_2d(wst + size_t(slot) * IM_TILE_BYTES, &tm_w, n0, row, &full[slot], leader);
        um_tma_2d(sz, &tm_s, n0 & ~7, g0, &full[slot], leader);
        um_tma_2d(sz + p.sz_bytes, &tm_z, ASYM ? ((n0 >> 3) & ~3) : (n0 & ~7), g0, &full[slot], leader);
    };
    // (rstrip, rkt) = tile that will be requested into the slot of the tile being consumed (S tiles ahead); every warp
    // keeps the pair: the refill is issued by whichever warp releases a slot LAST, so no warp ever waits for the others
    // (a fixed producer warp that first waited for all sixteen, then did its own unit, put its math on the critical path
    // of every refill: 0.95 us per tile instead of 0.6, profiles/r57_timeline_imma.txt)
    int rstrip = s_lo, rkt = 0;
    for (int t = 0; t < S && t < T; ++t) {
        // ---- the first S tiles are requested BEFORE griddepcontrol.wait: weights do not depend on the previous kernel ----
        if (warp == 0) issue_tile(rstrip, rkt, t);
        if (++rkt == tiles) { rkt = 0; rstrip += s_step; }
    }

    const int g = lane >> 2, c = lane & 3;
    if (!p.early) {
        pdl_wait_primary();      // x is produced by the previous kernel; y may still be read by it
        pdl_launch_dependents(); // only now: everything in front of this kernel is complete when its dependents start
    }
    IM_TRACE(1);

    // ---- activations: warp-private.  A staging pass covers the warp's unit (16 packed rows = 128 values) of two
    //      consecutive tiles: lanes 0-15 take tile 2*pass, lanes 16-31 tile 2*pass + 1, one packed row (16 bytes) per
    //      lane.  The raw row is loaded one pass ahead (register xv), so no global load sits in front of the math. ----
    const int hl = lane >> 4, r16 = lane & 15;
    unsigned char* ximg_w = ximg + warp * IM_WP_BYTES;
    float* xsum_w = reinterpret_cast<float*>(ximg_w + 1024);
    float* wt_w = xsum_w + 8;
    const int passes = (tiles + 1) >> 1;
    uint4 xv;
    auto load_x = [&](int pass) {
        const int row = (2 * pass + hl) * IM_TILE_ROWS + warp * IM_UNIT_ROWS + r16;
        xv = make_uint4(0u, 0u, 0u, 0u);
        if (row < p.R) xv = ld_global_v4(p.x + size_t(row) * 8);
    };
    auto stage_x = [&]() {
        const uint32_t w4[4] = {xv.x, xv.y, xv.z, xv.w};
        float v[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[2 * q] = cvt16_lo<BF16>(w4[q]); v[2 * q + 1] = cvt16_hi<BF16>(w4[q]); }
        float amax = 0.f, sum = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) { amax = fmaxf(amax, fabsf(v[q])); sum += v[q]; }
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
#pragma unroll
        for (int off = 1; off < 4 * F; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        // unit exponent: amax < 2^(e - 126); X = x * 2^(30 - (e - 126)) fits 31 bits.  e is clamped from below so that
        // the scale stays finite for all-zero / tiny units (their values then simply use fewer bits).
        int e = int(__float_as_uint(amax) >> 23);
        e = e < 67 ? 67 : e;
        const float scale = __uint_as_float(unsigned(283 - e) << 23);
        uint32_t P[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int Xo = __float2int_rn(v[2 * q + 1] * (scale * 0.0625f));
            const int Xe = __float2int_rn(v[2 * q] * scale) - Xo;        // pairs with the masked low nibble (see header)
            P[2 * q] = (unsigned(Xe) + 0x00808080u) ^ 0x00808080u;       // bytes = balanced base-256 digits, least significant first
            P[2 * q + 1] = (unsigned(Xo) + 0x00808080u) ^ 0x00808080u;
        }
        // digit slot d (= byte d): lo word = digit of Xe - Xo for k = 0,2,4,6, hi word = digit of Xo for k = 1,3,5,7
        uint32_t lo[4], hi[4];
        {
            const uint32_t e01a = __byte_perm(P[0], P[2], 0x5140), e01b = __byte_perm(P[0], P[2], 0x7362);   // bytes (0:P0,0:P2,1:P0,1:P2), (2.., 3..)
            const uint32_t e23a = __byte_perm(P[4], P[6], 0x5140), e23b = __byte_perm(P[4], P[6], 0x7362);
            lo[0] = __byte_perm(e01a, e23a, 0x5410); lo[1] = __byte_perm(e01a, e23a, 0x7632);
            lo[2] = __byte_perm(e01b, e23b, 0x5410); lo[3] = __byte_perm(e01b, e23b, 0x7632);
            const uint32_t o01a = __byte_perm(P[1], P[3], 0x5140), o01b = __byte_perm(P[1], P[3], 0x7362);
            const uint32_t o23a = __byte_perm(P[5], P[7], 0x5140), o23b = __byte_perm(P[5], P[7], 0x7362);
            hi[0] = __byte_perm(o01a, o23a, 0x5410); hi[1] = __byte_perm(o01a, o23a, 0x7632);
            hi[2] = __byte_perm(o01b, o23b, 0x5410); hi[3] = __byte_perm(o01b, o23b, 0x7632);
        }
        __syncwarp();            // every lane is done reading the previous pass's image
        // image of a unit: [8-row block b][k-step j][c][digit][8 B], packed row 8b + 2c + j  (conflict-free LDS.64)
        const int b = r16 >> 3;
        const int cc = SEQ ? (r16 & 3) : ((r16 & 7) >> 1), j = SEQ ? ((r16 >> 2) & 1) : (r16 & 1);
        unsigned char* dst = ximg_w + hl * 512 + b * 256 + j * 128 + cc * 32;
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], lo[0], hi[1], lo[1]);          // per digit: (Xo word, Xe - Xo word)
        *reinterpret_cast<uint4*>(dst + 16) = make_uint4(hi[2], lo[2], hi[3], lo[3]);
        if ((r16 & (4 * F - 1)) == 0) xsum_w[hl * 4 + r16 / (4 * F)] = sum;
        if (r16 == 0) wt_w[hl] = __uint_as_float(unsigned(e - 29) << 23);     // 2^((e - 126) - 30) * 2^... see flush
        __syncwarp();
    };
    load_x(0);
    int next_pass = 1 % passes;
    bool staged_once = false;

    // lane constants of the flush: digit-pair weight (c = 0: 1, c = 1: 65536, c >= 2: 0: duplicate digit columns)
    const float lane_w = (c == 0) ? 1.0f : (c == 1 ? 65536.0f : 0.0f);
    // shared-memory addresses of the hot loop as 32-bit offsets (no 64-bit pointer arithmetic in the loop)
    const uint32_t w_base = smem_u32(wst) + uint32_t((warp * IM_UNIT_ROWS + (SEQ ? c : 2 * c)) * IM_PITCH + g * 8);
    const uint32_t wp_base = smem_u32(ximg_w);                                      // warp-private region
    const uint32_t x_base = wp_base + uint32_t(c * 32 + (g & 3) * 8);
    const uint32_t sz_base = smem_u32(szst) + uint32_t(g * 4);
    const uint32_t zoff_lane = uint32_t(c < 2 ? 4 * g + 2 * c : 32 + 4 * g + 2 * (c - 2));   // column of this lane's zero-point term
    const uint32_t x_loader = lane < 16 ? 1u : 0u;
    uint2 xb = make_uint2(0u, 0u);
    const uint32_t full_base = smem_u32(full), rel_base = smem_u32(released);
    unsigned rel = 0u;
    const uint32_t slot_sz = 2u * uint32_t(p.sz_bytes);
    const int unit_row = warp * IM_UNIT_ROWS;
    // keep the lane's base addresses in registers: ptxas otherwise rebuilds them from %tid in every iteration
    // (~25 of the ~140 instructions per unit)
    uint32_t w_base_r = w_base, x_base_r = x_base, sz_base_r = sz_base, wp_base_r = wp_base;
    asm volatile("" : "+r"(w_base_r), "+r"(x_base_r), "+r"(sz_base_r), "+r"(wp_base_r));

    float yacc[4] = {0.f, 0.f, 0.f, 0.f};     // columns 2g, 2g+1, 16+2g, 17+2g: sum of s * (x . q), this lane's digit pair
    float yz = 0.f;                           // column number c of those four: sum of z * sum(x)
    int strip = s_lo, kt = 0, slot = 0;
    unsigned ph = 0;
    bool out_waited = false;
    int rd_par = 0;
    int n0 = strip_col(strip);
    for (int t = 0; t < T; ++t) {
        if ((kt & 1) == 0 && (passes > 1 || !staged_once)) {
            stage_x();                                 // tiles kt, kt + 1 of the strip
            staged_once = true;
            if (passes > 1) { load_x(next_pass); next_pass = (next_pass + 1 == passes) ? 0 : next_pass + 1; }
            if (t == 0) IM_TRACE(2);
        }
        if (kt * IM_TILE_ROWS + unit_row < p.R) {          // this warp's unit exists in the tile
            im_mbar_wait(full_base + slot * 8, ph);
            if (t == 0) IM_TRACE(3);
            const uint32_t wt = w_base_r + uint32_t(slot) * IM_TILE_BYTES;
            const uint32_t xt = x_base_r + uint32_t(kt & 1) * 512u;
            const uint32_t sz = sz_base_r + uint32_t(slot) * slot_sz + uint32_t(n0 & 7) * 2u;   // strip starts at column n0 & 7 of the fp16 tile rows
            const float wunit = __uint_as_float(im_lds32(wp_base_r + 1056u + uint32_t(kt & 1) * 4u)) * lane_w;
            int A[4], B[4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int b = ks >> 1, j = ks & 1;
                const uint32_t wa_ = wt + uint32_t((b * 8 + (SEQ ? 4 * j : j)) * IM_PITCH);
                const uint2 wa = im_lds64v(wa_), wb = im_lds64v(wa_ + 64u);
                im_lds64_if(xb, xt + uint32_t(b * 256 + j * 128), x_loader);
                if (ks % F == 0) {
                    im_mma_z(A, wa.x, wa.y, wa.x & 0x0f0f0f0fu, wa.y & 0x0f0f0f0fu, xb.x, xb.y);
                    im_mma_z(B, wb.x, wb.y, wb.x & 0x0f0f0f0fu, wb.y & 0x0f0f0f0fu, xb.x, xb.y);
                } else {
                    im_mma(A, wa.x, wa.y, wa.x & 0x0f0f0f0fu, wa.y & 0x0f0f0f0fu, xb.x, xb.y);
                    im_mma(B, wb.x, wb.y, wb.x & 0x0f0f0f0fu, wb.y & 0x0f0f0f0fu, xb.x, xb.y);
                }
                if ((ks + 1) % F == 0) {
                    // ---- flush F k-steps (rows of one group) through the group's affine parameters ----
                    const int seg = ks / F;
                    const uint32_t gl = uint32_t((unit_row + seg * 4 * F) >> p.rpg_shift);    // group row inside the tile
                    // every shared-memory read of the segment first ...
                    const float xs = __uint_as_float(im_lds32(wp_base_r + 1024u + uint32_t((kt & 1) * 16 + seg * 4)));
                    const uint32_t s2[2] = {im_lds32(sz + gl * 64u), im_lds32(sz + gl * 64u + 32u)};   // columns 2g, 2g+1 | 16+2g, 17+2g
                    uint32_t zraw;      // zero point of this lane's column q = c (one column per lane, combined with yacc at the end of the strip)
                    int cz = 0;
                    if constexpr (ASYM) {
                        // zero tile row = 8 packed words from word (n0 >> 3) & ~3; the strip starts at nibble (n0 & 7) in {0, 4}
                        // of word (n0 >> 3) & 3 of the box
                        cz = int(zoff_lane >> 1) + (n0 & 7);                    // nibble index inside the 64-nibble box row
                        zraw = im_lds32(smem_u32(szst) + uint32_t(slot) * slot_sz + uint32_t(p.sz_bytes) + gl * 32u +
                                        uint32_t((((n0 >> 3) & 3) + (cz >> 3)) * 4));
                    } else {
                        zraw = im_lds16(sz - uint32_t(g * 4) + uint32_t(p.sz_bytes) + gl * 64u + zoff_lane);
                    }
                    if (ks == 3) {
                        // ... then, after the unit's last read of the slot, release it: the atomic's round trip overlaps the
                        // flush arithmetic below instead of sitting at the end of the unit
                        __syncwarp();
                        if (lane == 0) rel = im_atoms_add(rel_base + uint32_t(slot) * 4u, 1u);
                    }
                    float sq[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) sq[q] = (q & 1) ? cvt16_hi<BF16>(s2[q >> 1]) : cvt16_lo<BF16>(s2[q >> 1]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {           // q = 0,1: columns 2g, 2g+1 (IMMA alpha rows g, g+8); q = 2,3: 16+2g, 17+2g (beta)
                        const int dlo = (q < 2) ? A[(q & 1) * 2] : B[(q & 1) * 2];
                        const int dhi = (q < 2) ? A[(q & 1) * 2 + 1] : B[(q & 1) * 2 + 1];
                        yacc[q] = fmaf(sq[q] * wunit, float(dhi * 256 + dlo), yacc[q]);
                    }
                    float zf;
                    if constexpr (ASYM) {
                        const float sc = c == 0 ? sq[0] : (c == 1 ? sq[1] : (c == 2 ? sq[2] : sq[3]));
                        zf = sc * float(((zraw >> ((cz & 7) * 4)) & 15u) + 1u);
                    } else {
                        zf = cvt16_lo<BF16>(zraw);
                    }
                    yz = fmaf(zf, xs, yz);
                }
            }
        } else {
            __syncwarp();
            if (lane == 0) rel = im_atoms_add(rel_base + uint32_t(slot) * 4u, 1u);
        }
        // ---- the warp that released the slot last requests tile t + S into it ----
        const unsigned last = ((__shfl_sync(0xffffffffu, rel, 0) & unsigned(IM_WARPS - 1)) == unsigned(IM_WARPS - 1)) ? 1u : 0u;
        if (last != 0u && t + S < T) issue_tile(rstrip, rkt, slot);
        if (++rkt == tiles) { rkt = 0; rstrip += s_step; }
        if (++slot == S) { slot = 0; ph ^= 1u; }
        if (++kt == tiles) {
            if (t + 1 == T) IM_TRACE(4);
            // ---- strip finished: sum the digit-pair lanes, subtract the zero-point terms, then the sixteen warps in
            //      fixed order; y written exactly once ----
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                yacc[q] += __shfl_xor_sync(0xffffffffu, yacc[q], 1);
                yacc[q] += __shfl_xor_sync(0xffffffffu, yacc[q], 2);
                yacc[q] -= __shfl_sync(0xffffffffu, yz, (lane & ~3) + q);
            }
            float* rd = red + rd_par * (IM_WARPS * 32);
            rd_par ^= 1;
            if (c == 0) {
                *reinterpret_cast<float2*>(rd + warp * 32 + 2 * g) = make_float2(yacc[0], yacc[1]);
                *reinterpret_cast<float2*>(rd + warp * 32 + 16 + 2 * g) = make_float2(yacc[2], yacc[3]);
            }
            yacc[0] = yacc[1] = yacc[2] = yacc[3] = 0.f;
            yz = 0.f;
            __syncthreads();
            if (p.early && !out_waited) {             // y may still be in use by the previous kernel
                pdl_wait_primary();
                out_waited = true;
                IM_TRACE(7);
            }
            const int width = strip < p.n28 ? IM_COLS : 24;
            if (tid < width && n0 + tid < p.N) {
                float total = 0.f;
#pragma unroll
                for (int w = 0; w < IM_WARPS; ++w) total += rd[w * 32 + tid];
                p.y[n0 + tid] = f32_to_16<BF16>(total);
            }
            kt = 0;
            strip += s_step;
            n0 = strip_col(strip);
        }
    }
    IM_TRACE(5);
}

struct ImmaLaunch {
    int F, grid;
    bool asym, bf16;
    size_t smem;
    unsigned flags;
    cudaStream_t stream;
};
int launch_imma(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const ImmaParams& p, const ImmaLaunch& l);

}  // namespace b200bit
