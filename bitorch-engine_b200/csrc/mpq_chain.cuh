// mpq_chain.cuh -- decode CHAIN: one persistent launch runs a whole list of batch-1 4-bit Linear layers ("nodes") with
// their dataflow resolved on the device.  Same arithmetic as mpq_imma.cuh (integer tensor pipe, base-256 digit
// activations; results are bit-identical to the per-layer kernel), different scheduling:
//
//   * grid = one CTA per SM, the whole shared memory: a ring of up to 7 weight tiles (28 KB + scale / zero rows each)
//     kept full by a dedicated TMA producer warp ACROSS node boundaries; sixteen compute warps; one epilogue /
//     dependency warp.  The ring is refilled across layers -- the CTA's tile sequence (node, strip, k-tile) is known up
//     front, so the HBM stream never stops for a layer boundary: while a CTA waits for the input of node j its ring
//     already holds the first ~200 KB of node j's (and j+1's) weights.  The per-layer kernels (one launch per layer,
//     programmatic dependent launch) could only prefetch 84 KB per SM and paid 0.55 - 1.4 us between the completion of
//     a layer and the release of its dependent (profiles/r60_imma_timeline.txt): 30 us per decoder block against 16.4 us
//     of HBM time.
//   * dependencies travel WITH the data: a node whose output feeds a later node also writes it as 8-byte words
//     {two 16-bit values, launch epoch} into a shadow buffer ("LL" protocol: a naturally aligned 64-bit store is
//     single-copy atomic, so a reader that sees the epoch sees the values -- no fence).  A relaxed per-node counter is
//     the HINT that pulling is worth it: the dependency warp polls it and releases the compute warps once half of the
//     producer's strips are counted (early_pct); they read their own rows of x from the shadow and simply retry the rows
//     whose words do not carry this launch's epoch yet.  A fence.gpu on an SM that has 150 KB of TMA loads in flight
//     costs ~1.5 us (measured, profiles/r2_02_*), which is why the counter protocol below is only the fallback.
//   * fallback for everything that is not "x is exactly an earlier node's y": counters in global memory.  The CTA that
//     has written a strip of y does red.release.gpu on the node's counter; a dependent polls it (one lane per CTA,
//     ld.acquire.gpu, bounded spin).  The host derives the hazards from the pointer ranges: read-after-write before x
//     is read, write-after-read / write-after-write before y is written (buffer reuse).  Sibling nodes (q, k, v of one
//     hidden state; gate, up) carry no wait at all.
//   * counters reset themselves: the last CTA to leave the kernel zeroes them (launches of one plan must be stream-ordered).
//
// A chain is legitimate only for consecutive Linear layers with nothing else between them (the "linear-layer tokens/s"
// metric, fused q/k/v and gate/up segments of a real decoder); foreign kernels between two layers end a chain.
// Replaces N x (quant_mm_kernel + torch::zeros memset) launches (mpq_linear_cuda_kernel.cu:67-451, :618).
#pragma once
#include "mpq_imma.cuh"
#include <type_traits>

namespace b200bit {

constexpr int CH_MAX_STAGES = 7;
constexpr int CH_SMEM_LIMIT = 227 * 1024;
constexpr int CH_TRACE_NODES = 32;          // diagnostics: stamps for the first 32 nodes
constexpr unsigned CH_SPIN_LIMIT = 1u << 22;   // polls before a wait gives up and raises the plan's error flag (~1 s)
constexpr int CH_RED_PLANE = 33;                 // floats per (warp, value) plane of the hand-over buffer: 32 lanes + 1 pad (conflict-free reads)
constexpr int CH_RED_FLOATS = IM_WARPS * 5 * CH_RED_PLANE;   // per warp: yacc[0..3] and yz of every lane
constexpr int CH_THREADS = IM_THREADS + 128;   // sixteen compute warps + one helper warpgroup: the TMA producer warp, the epilogue /
                                               // dependency warp and two idle warps (register reallocation works on whole warpgroups)

struct __align__(16) ChainNode {     // 64 bytes, device + host
    const uint16_t* x;       // [K] f16 / bf16 bits
    uint16_t* y;             // [N]
    const uint64_t* xll;     // [K / 2] {value pair, epoch} shadow of x written by the producing node, or null (x is plain)
    uint64_t* yll;           // [N / 2] shadow of y for the nodes that consume it, or null
    int R, N;                // packed rows (K / 8), columns
    int strips, n28, tiles;  // strips of the node (the first n28 are 28 columns wide, the others 24), tiles per strip
    int off_sig;             // bit 22: x (pointer and length) is the previous node's x: a CTA that staged it keeps the image;
                             // bits 0..19: off -- CTA b owns the strips s with (s + off) % grid == b (the cyclic deal
                             // continues across nodes); bit 20: a later node orders plain memory behind this node's
                             // counter (release); bit 21: shadow readers use the counter as a hint (relaxed)
    int wx_node;             // before x is read: counter[wx_node] >= strips of wx_node (< 0: no counter wait)
    int wy_node;             // before y is written: likewise (buffer reuse; < 0: none)
};
static_assert(sizeof(ChainNode) == 64, "ChainNode layout");

struct ChainParams {
    const ChainNode* nodes;      // [n_nodes] global
    const CUtensorMap* maps;     // [3 * n_nodes] global: (packed weights, scales, zeros) per node
    unsigned* counters;          // [0, n_nodes) strips finished per node | [n_nodes] CTAs that left | [n_nodes + 1] error flag
                                 // (sticky) | [n_nodes + 2] launch epoch (never 0)
    int n_nodes;
    int S;                       // ring slots
    int rpg_shift;               // log2(packed rows per group) -- one group size per chain
    int sz_bytes, s_tile_bytes, z_tile_bytes;
    int poll_depth;              // dependency polls in flight per CTA (1, 2 or 4)
    int early_pct;               // shadow readers start pulling x once (100 - early_pct) % of the producer's strips are counted
    unsigned long long* trace;   // [grid][CH_TRACE_NODES][8] globaltimer stamps (diagnostics build only)
};

__device__ __forceinline__ unsigned ch_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// x is written by other CTAs of this very launch: never through L1
__device__ __forceinline__ uint4 ch_ld_x(const void* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
// y of a strip is complete: release (covers the lanes' stores through the __syncwarp in front) + count the strip
__device__ __forceinline__ void ch_signal(unsigned* ctr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ ulonglong2 ch_ld_ll(const uint64_t* p) {
    ulonglong2 r;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void ch_st_ll(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// hint only (no ordering): the consumers verify the data itself (shadow words carry the epoch)
__device__ __forceinline__ void ch_hint(unsigned* ctr) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ unsigned ch_ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Wait until *ctr >= want.  A loaded L2 round trip is 0.5 - 1 us while the weight stream is running; `depth` polls can be
// kept in flight (1, 2 or 4; measured equal on the final build, default 1).  `ordered`: the counter orders plain memory
// (acquire fence at the end); otherwise it is only a hint and the data validates itself.
__device__ __forceinline__ void ch_spin(const unsigned* ctr, unsigned want, unsigned* err, bool ordered, int depth) {
    unsigned a = ch_ld_relaxed(ctr);
    if (a < want) {
        unsigned spins = 0;
        if (depth <= 1) {
            while ((a = ch_ld_relaxed(ctr)) < want)
                if (++spins > CH_SPIN_LIMIT) { atomicExch(err, 1u); break; }
        } else if (depth == 2) {
            unsigned b = ch_ld_relaxed(ctr);
            while (a < want) {
                a = b;
                b = ch_ld_relaxed(ctr);
                if (++spins > CH_SPIN_LIMIT) { atomicExch(err, 1u); break; }
            }
        } else {
            unsigned b = ch_ld_relaxed(ctr);
            unsigned c = ch_ld_relaxed(ctr);
            unsigned d = ch_ld_relaxed(ctr);
            while (a < want) {          // `a` is the oldest load in flight; three younger ones are behind it
                a = b; b = c; c = d;
                d = ch_ld_relaxed(ctr);
                if (++spins > CH_SPIN_LIMIT) { atomicExch(err, 1u); break; }
            }
        }
    }
    if (ordered) asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// mbarrier wait with a bound: a protocol bug must end in a trap (the launch fails), never in a hung GPU
__device__ __forceinline__ void ch_mbar_wait(uint32_t bar, unsigned parity) {
    unsigned tries = 0;
    while (true) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++tries > (1u << 22)) asm volatile("trap;");
    }
}

#define CH_TRACE(node_, slot_) do { if constexpr (TRACE) { if (p.trace && tid == 0 && (node_) < CH_TRACE_NODES) \
    p.trace[(size_t(blockIdx.x) * CH_TRACE_NODES + (node_)) * 8 + (slot_)] = st_gtime(); } } while (0)

template <int F, bool ASYM, bool BF16, bool TRACE>
__global__ void __launch_bounds__(CH_THREADS, 1) mpq_chain_kernel(const ChainParams p) {
    constexpr bool SEQ = (F == 1);
    extern __shared__ __align__(1024) unsigned char ch_smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.S;
    const int grid = int(gridDim.x), bid = int(blockIdx.x);
    const int n_nodes = p.n_nodes;
    // carve-up: W ring S x 28672 | scale / zero tiles S x 2 x sz_bytes | per-warp x digit images | red [16][5][33] f32 |
    //           node table n x 64 | mbarriers full[7], empty[7], part[2], free[2], ready
    unsigned char* wst = ch_smem;
    unsigned char* szst = wst + size_t(S) * IM_TILE_BYTES;
    unsigned char* ximg = szst + size_t(S) * 2 * p.sz_bytes;
    float* red = reinterpret_cast<float*>(ximg + IM_XIMG_BYTES);
    ChainNode* nodes_s = reinterpret_cast<ChainNode*>(red + CH_RED_FLOATS);
    uint64_t* full = reinterpret_cast<uint64_t*>(nodes_s + n_nodes);   // [7] tile landed (TMA transaction bytes)
    uint64_t* empty = full + CH_MAX_STAGES;      // [7] the sixteen compute warps are done with the slot
    uint64_t* part = empty + CH_MAX_STAGES;      // [2] partial sums of a strip are in red[par]: 16 compute warps arrive
    uint64_t* freeb = part + 2;                  // [2] the epilogue warp has read red[par]
    uint64_t* ready = freeb + 2;                 // the input of the CTA's next dependent node is complete
    unsigned* err_flag = p.counters + n_nodes + 1;

    if (tid < S) { mbar_init(&full[tid], 1); mbar_init(&empty[tid], IM_WARPS); }
    else if (tid == 64) { mbar_init(&part[0], IM_WARPS); mbar_init(&part[1], IM_WARPS); }
    else if (tid == 65) { mbar_init(&freeb[0], 1); mbar_init(&freeb[1], 1); mbar_init(ready, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.nodes);
        uint4* dst = reinterpret_cast<uint4*>(nodes_s);
        for (int i = tid; i < n_nodes * 4; i += CH_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    auto strip_col = [&](const ChainNode& nd, int strip) {
        return strip < nd.n28 ? strip * IM_COLS : nd.n28 * IM_COLS + (strip - nd.n28) * 24;
    };
    auto first_strip = [&](const ChainNode& nd) { int s0 = bid - (nd.off_sig & 0xfffff); return s0 < 0 ? s0 + grid : s0; };
    auto next_node = [&](int node) {      // first node >= `node` in which this CTA owns a strip (n_nodes: none)
        while (node < n_nodes) {
            const ChainNode& nd = nodes_s[node];
            if (first_strip(nd) < nd.strips) break;
            ++node;
        }
        return node;
    };
    unsigned epoch;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(epoch) : "l"(p.counters + n_nodes + 2) : "memory");

    // 20 warps x 96 registers at launch = the CTA's register pool; the helper warpgroup goes down to 56 (frees 5120), the
    // compute warps up to 104 (take 4096): setmaxnreg.inc blocks until the pool can serve it, so the sums must work out
    if (warp >= IM_WARPS + 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        return;
    }

    if (warp == IM_WARPS) {
        // ===== TMA producer warp: walks the CTA's tile sequence (node, own strip, k-tile) and keeps the ring full; the
        //       only thing it ever waits for is a free slot, never a layer boundary =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const unsigned tile_tx = unsigned(IM_TILE_BYTES) + unsigned(p.s_tile_bytes) + unsigned(p.z_tile_bytes);
        const uint32_t leader = um_elect();
        int slot = 0;
        unsigned eph = 1u;                   // parity of the phase in front of the slot's first release: passes at once
        for (int node = next_node(0); node < n_nodes; node = next_node(node + 1)) {
            const ChainNode nd = nodes_s[node];
            const CUtensorMap* tm = p.maps + 3 * node;
            for (int strip = first_strip(nd); strip < nd.strips; strip += grid) {
                const int n0 = strip_col(nd, strip);
                for (int kt = 0; kt < nd.tiles; ++kt) {
                    ch_mbar_wait(smem_u32(&empty[slot]), eph);
                    const int row = kt * IM_TILE_ROWS;
                    const int g0 = row >> p.rpg_shift;
                    unsigned char* sz = szst + size_t(slot) * 2 * p.sz_bytes;
                    um_expect_tx(&full[slot], tile_tx, leader);
                    um_tma_2d(wst + size_t(slot) * IM_TILE_BYTES, tm, n0, row, &full[slot], leader);
                    um_tma_2d(sz, tm + 1, n0 & ~7, g0, &full[slot], leader);
                    um_tma_2d(sz + p.sz_bytes, tm + 2, ASYM ? ((n0 >> 3) & ~3) : (n0 & ~7), g0, &full[slot], leader);
                    if (++slot == S) { slot = 0; eph ^= 1u; }
                }
            }
        }
        return;
    }

    if (warp == IM_WARPS + 1) {
        // ===== epilogue / dependency warp: everything of the path that talks to other CTAs.  The compute warps never
        //       wait for each other or for a global-memory round trip: they hand their partial sums over through
        //       red[par] + an mbarrier and go on with the next strip. =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        unsigned use = 0u;                       // how often red has been consumed
        int waited = -1;
        // this lane's column of a strip: columns 2g, 2g + 1 are yacc[0], yacc[1] of the lane quad g, columns 16 + 2g, 17 + 2g
        // are yacc[2], yacc[3]
        const int eg = (lane & 15) >> 1, eq = (lane & 1) + ((lane >> 4) << 1);
        for (int node = next_node(0); node < n_nodes; node = next_node(node + 1)) {
            const ChainNode nd = nodes_s[node];
            if (nd.wx_node > waited) {       // read-after-write: the producer's counter (the ordering itself when x is plain,
                                             // a hint that the shadow words are there when x is read through xll); nodes
                                             // finish in order, so a wait covers every earlier node (siblings skip)
                waited = nd.wx_node;
                if (lane == 0) {
                    // shadow words validate themselves, so their readers may start before the last strip is counted: the
                    // pull (one loaded L2 round trip) then overlaps the stragglers instead of following them
                    unsigned want = unsigned(nodes_s[nd.wx_node].strips);
                    if (nd.xll != nullptr) { want -= want * unsigned(p.early_pct) / 100u; want = want ? want : 1u; }
                    ch_spin(p.counters + nd.wx_node, want, err_flag, nd.xll == nullptr, p.poll_depth);
                    mbar_arrive(ready);
                    if constexpr (TRACE) { if (p.trace && node < CH_TRACE_NODES) p.trace[(size_t(bid) * CH_TRACE_NODES + node) * 8 + 4] = st_gtime(); }
                }
                __syncwarp();
            }
            for (int strip = first_strip(nd); strip < nd.strips; strip += grid) {
                const int n0 = strip_col(nd, strip);
                ch_mbar_wait(smem_u32(&part[0]), use & 1u);
                if constexpr (TRACE) { if (p.trace && lane == 0 && node < CH_TRACE_NODES) p.trace[(size_t(bid) * CH_TRACE_NODES + node) * 8 + 0] = st_gtime(); }
                // the compute warps hand over their lanes' raw partial sums; the digit-pair lanes, the zero-point term and
                // the sixteen warps are combined here, in the order of the per-layer kernel (bit-identical results)
                float total = 0.f;
#pragma unroll
                for (int w = 0; w < IM_WARPS; ++w) {
                    const float* rw = red + w * 5 * CH_RED_PLANE;
                    const float* rq = rw + eq * CH_RED_PLANE + eg * 4;
                    const float t = (rq[0] + rq[1]) + (rq[2] + rq[3]);
                    total += t - rw[4 * CH_RED_PLANE + eg * 4 + eq];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&freeb[0]);
                ++use;
                if (nd.wy_node >= 0) {       // write-after-read / write-after-write: the buffer behind y was used by node wy_node
                    if (lane == 0) ch_spin(p.counters + nd.wy_node, unsigned(nodes_s[nd.wy_node].strips), err_flag, true, p.poll_depth);
                    __syncwarp();
                }
                const int width = strip < nd.n28 ? IM_COLS : 24;
                const bool mine = lane < width && n0 + lane < nd.N;
                const unsigned h = f32_to_16<BF16>(total);
                const unsigned hn = __shfl_down_sync(0xffffffffu, h, 1);
                if (mine) nd.y[n0 + lane] = (unsigned short)h;
                // strips start on even columns and are even wide: lane pairs (2i, 2i + 1) form one shadow word
                if (nd.yll != nullptr && mine && (lane & 1) == 0)
                    ch_st_ll(nd.yll + ((n0 + lane) >> 1), (uint64_t(epoch) << 32) | uint64_t(h | (hn << 16)));
                if (nd.off_sig & (1 << 20)) {            // someone orders plain memory behind this node: release
                    __syncwarp();
                    if (lane == 0) ch_signal(p.counters + node);
                } else if (nd.off_sig & (1 << 21)) {     // shadow readers only: the count is a hint that polling the data
                    __syncwarp();                        // is worth it now
                    if (lane == 0) ch_hint(p.counters + node);
                }
                if constexpr (TRACE) { if (p.trace && lane == 0 && node < CH_TRACE_NODES) p.trace[(size_t(bid) * CH_TRACE_NODES + node) * 8 + 3] = st_gtime(); }
            }
        }
        // ---- the last CTA to leave resets the plan's counters for the next launch (this warp wrote the CTA's last y) ----
        if (lane == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            const unsigned prev = atomicAdd(p.counters + n_nodes, 1u);
            if (prev == unsigned(grid) - 1u) {
                for (int i = 0; i <= n_nodes; ++i) p.counters[i] = 0u;
                p.counters[n_nodes + 2] = epoch + 1u == 0u ? 1u : epoch + 1u;
            }
        }
        return;
    }

    // ===== sixteen compute warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int g = lane >> 2, c = lane & 3;
    const int hl = lane >> 4, r16 = lane & 15;
    unsigned char* ximg_w = ximg + warp * IM_WP_BYTES;
    float* xsum_w = reinterpret_cast<float*>(ximg_w + 1024);
    float* wt_w = xsum_w + 8;

    const float lane_w = (c == 0) ? 1.0f : (c == 1 ? 65536.0f : 0.0f);
    const uint32_t w_base = smem_u32(wst) + uint32_t((warp * IM_UNIT_ROWS + (SEQ ? c : 2 * c)) * IM_PITCH + g * 8);
    const uint32_t wp_base = smem_u32(ximg_w);
    const uint32_t x_base = wp_base + uint32_t(c * 32 + (g & 3) * 8);
    const uint32_t sz_base = smem_u32(szst) + uint32_t(g * 4);
    const uint32_t zoff_lane = uint32_t(c < 2 ? 4 * g + 2 * c : 32 + 4 * g + 2 * (c - 2));
    const uint32_t x_loader = lane < 16 ? 1u : 0u;
    uint2 xb = make_uint2(0u, 0u);
    const uint32_t full_base = smem_u32(full), empty_base = smem_u32(empty);
    const uint32_t slot_sz = 2u * uint32_t(p.sz_bytes);
    const int unit_row = warp * IM_UNIT_ROWS;
    uint32_t w_base_r = w_base, x_base_r = x_base, sz_base_r = sz_base, wp_base_r = wp_base;
    asm volatile("" : "+r"(w_base_r), "+r"(x_base_r), "+r"(sz_base_r), "+r"(wp_base_r));

    float yacc[4] = {0.f, 0.f, 0.f, 0.f};
    float yz = 0.f;
    int slot = 0;
    unsigned ph = 0;
    unsigned dep_phase = 0u, use = 0u;
    int waited = -1, prev_node = -2;

    for (int node = next_node(0); node < n_nodes; node = next_node(node + 1)) {
        const ChainNode nd = nodes_s[node];
        const int tiles = nd.tiles;
        const int passes = (tiles + 1) >> 1;
        CH_TRACE(node, 6);
        // ---- read-after-write: the dependency warp has seen the producer's counter ----
        if (nd.wx_node > waited) {
            waited = nd.wx_node;
            ch_mbar_wait(smem_u32(ready), dep_phase);
            dep_phase ^= 1u;
        }
        // the digit image of x is still in shared memory when the previous node read the very same x (k, v after q; up
        // after gate), this CTA took part in it, and the whole of x fits one staging pass
        const bool keep_image = (nd.off_sig & (1 << 22)) != 0 && prev_node == node - 1 && passes == 1;
        prev_node = node;
        uint4 xv, xw;             // raw rows of the next two staging passes (a loaded L2 round trip is longer than a pass)
        auto load_x = [&](int pass, uint4& xv) {
            const int row = (2 * pass + hl) * IM_TILE_ROWS + warp * IM_UNIT_ROWS + r16;
            xv = make_uint4(0u, 0u, 0u, 0u);
            if (nd.xll == nullptr) {
                if (row < nd.R) xv = ch_ld_x(nd.x + size_t(row) * 8);
                return;
            }
            // LL shadow: the four words of the row are valid once each carries this launch's epoch
            bool have = row >= nd.R;
            const uint64_t* src = nd.xll + size_t(row) * 4;
            unsigned spins = 0;
            while (true) {
                if (!have) {
                    const ulonglong2 a = ch_ld_ll(src), b = ch_ld_ll(src + 2);
                    if (unsigned(a.x >> 32) == epoch && unsigned(a.y >> 32) == epoch && unsigned(b.x >> 32) == epoch &&
                        unsigned(b.y >> 32) == epoch) {
                        xv = make_uint4(unsigned(a.x), unsigned(a.y), unsigned(b.x), unsigned(b.y));
                        have = true;
                    }
                }
                if (__all_sync(0xffffffffu, have)) break;
                if (++spins > (CH_SPIN_LIMIT >> 2)) { if (lane == 0) atomicExch(err_flag, 1u); break; }
                if (p.early_pct == 0) __nanosleep(64);        // rare: the hint counter said the words are on their way
            }
        };
        auto stage_x = [&]() {
            const uint32_t w4[4] = {xv.x, xv.y, xv.z, xv.w};
            float v[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) { v[2 * q] = cvt16_lo<BF16>(w4[q]); v[2 * q + 1] = cvt16_hi<BF16>(w4[q]); }
            float amax = 0.f, sum = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) { amax = fmaxf(amax, fabsf(v[q])); sum += v[q]; }
            // maximum over the unit's sixteen lanes: non-negative floats order like their bit patterns -> one REDUX
            // (the 4-round shuffle butterfly sat on the critical path of every dependent node)
            {
                unsigned am = __float_as_uint(amax), r;
                asm volatile("redux.sync.max.u32 %0, %1, %2;" : "=r"(r) : "r"(am), "r"(hl ? 0xffff0000u : 0x0000ffffu));
                amax = __uint_as_float(r);
            }
#pragma unroll
            for (int off = 1; off < 4 * F; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            int e = int(__float_as_uint(amax) >> 23);
            e = e < 67 ? 67 : e;
            const float scale = __uint_as_float(unsigned(283 - e) << 23);
            uint32_t P[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int Xo = __float2int_rn(v[2 * q + 1] * (scale * 0.0625f));
                const int Xe = __float2int_rn(v[2 * q] * scale) - Xo;
                P[2 * q] = (unsigned(Xe) + 0x00808080u) ^ 0x00808080u;
                P[2 * q + 1] = (unsigned(Xo) + 0x00808080u) ^ 0x00808080u;
            }
            uint32_t lo[4], hi[4];
            {
                const uint32_t e01a = __byte_perm(P[0], P[2], 0x5140), e01b = __byte_perm(P[0], P[2], 0x7362);
                const uint32_t e23a = __byte_perm(P[4], P[6], 0x5140), e23b = __byte_perm(P[4], P[6], 0x7362);
                lo[0] = __byte_perm(e01a, e23a, 0x5410); lo[1] = __byte_perm(e01a, e23a, 0x7632);
                lo[2] = __byte_perm(e01b, e23b, 0x5410); lo[3] = __byte_perm(e01b, e23b, 0x7632);
                const uint32_t o01a = __byte_perm(P[1], P[3], 0x5140), o01b = __byte_perm(P[1], P[3], 0x7362);
                const uint32_t o23a = __byte_perm(P[5], P[7], 0x5140), o23b = __byte_perm(P[5], P[7], 0x7362);
                hi[0] = __byte_perm(o01a, o23a, 0x5410); hi[1] = __byte_perm(o01a, o23a, 0x7632);
                hi[2] = __byte_perm(o01b, o23b, 0x5410); hi[3] = __byte_perm(o01b, o23b, 0x7632);
            }
            __syncwarp();
            const int b = r16 >> 3;
            const int cc = SEQ ? (r16 & 3) : ((r16 & 7) >> 1), j = SEQ ? ((r16 >> 2) & 1) : (r16 & 1);
            unsigned char* dst = ximg_w + hl * 512 + b * 256 + j * 128 + cc * 32;
            *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], lo[0], hi[1], lo[1]);
            *reinterpret_cast<uint4*>(dst + 16) = make_uint4(hi[2], lo[2], hi[3], lo[3]);
            if ((r16 & (4 * F - 1)) == 0) xsum_w[hl * 4 + r16 / (4 * F)] = sum;
            if (r16 == 0) wt_w[hl] = __uint_as_float(unsigned(e - 29) << 23);
            __syncwarp();
        };
        if (!keep_image) load_x(0, xv);
        if (passes > 1) load_x(1, xw);
        CH_TRACE(node, 7);
        int next_pass = 2 % passes;
        bool staged_once = keep_image;

        for (int strip = first_strip(nd); strip < nd.strips; strip += grid) {
            const int n0 = strip_col(nd, strip);
            // NT tiles (kt0, kt0 + 1) of the strip at once: two independent unit chains per warp
            auto process = [&](auto nt_tag, int kt0) {
                constexpr int NT = decltype(nt_tag)::value;
                int sl[NT];
                uint32_t wt[NT], xt[NT], sz[NT];
                float wunit[NT];
                int A[NT][4], B[NT][4];
                uint2 xbu[NT];
#pragma unroll
                for (int u = 0; u < NT; ++u) {
                    sl[u] = slot + u >= S ? slot + u - S : slot + u;
                    ch_mbar_wait(full_base + sl[u] * 8, slot + u >= S ? (ph ^ 1u) : ph);
                    wt[u] = w_base_r + uint32_t(sl[u]) * IM_TILE_BYTES;
                    xt[u] = x_base_r + uint32_t((kt0 + u) & 1) * 512u;
                    sz[u] = sz_base_r + uint32_t(sl[u]) * slot_sz + uint32_t(n0 & 7) * 2u;
                    wunit[u] = __uint_as_float(im_lds32(wp_base_r + 1056u + uint32_t((kt0 + u) & 1) * 4u)) * lane_w;
                    xbu[u] = xb;
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int b = ks >> 1, j = ks & 1;
                    uint2 wa[NT], wb[NT];
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
                        const uint32_t wa_ = wt[u] + uint32_t((b * 8 + (SEQ ? 4 * j : j)) * IM_PITCH);
                        wa[u] = im_lds64v(wa_);
                        wb[u] = im_lds64v(wa_ + 64u);
                        im_lds64_if(xbu[u], xt[u] + uint32_t(b * 256 + j * 128), x_loader);
                    }
#pragma unroll
                    for (int u = 0; u < NT; ++u) {
                        if (ks % F == 0) {
                            im_mma_z(A[u], wa[u].x, wa[u].y, wa[u].x & 0x0f0f0f0fu, wa[u].y & 0x0f0f0f0fu, xbu[u].x, xbu[u].y);
                            im_mma_z(B[u], wb[u].x, wb[u].y, wb[u].x & 0x0f0f0f0fu, wb[u].y & 0x0f0f0f0fu, xbu[u].x, xbu[u].y);
                        } else {
                            im_mma(A[u], wa[u].x, wa[u].y, wa[u].x & 0x0f0f0f0fu, wa[u].y & 0x0f0f0f0fu, xbu[u].x, xbu[u].y);
                            im_mma(B[u], wb[u].x, wb[u].y, wb[u].x & 0x0f0f0f0fu, wb[u].y & 0x0f0f0f0fu, xbu[u].x, xbu[u].y);
                        }
                    }
                    if ((ks + 1) % F == 0) {
                        // ---- flush F k-steps (rows of one group) through the group's affine parameters ----
                        const int seg = ks / F;
                        const uint32_t gl = uint32_t((unit_row + seg * 4 * F) >> p.rpg_shift);
                        float xs[NT];
                        uint32_t s2[NT][2], zraw[NT];
                        int cz = 0;
#pragma unroll
                        for (int u = 0; u < NT; ++u) {
                            xs[u] = __uint_as_float(im_lds32(wp_base_r + 1024u + uint32_t(((kt0 + u) & 1) * 16 + seg * 4)));
                            s2[u][0] = im_lds32(sz[u] + gl * 64u);
                            s2[u][1] = im_lds32(sz[u] + gl * 64u + 32u);
                            if constexpr (ASYM) {
                                cz = int(zoff_lane >> 1) + (n0 & 7);
                                zraw[u] = im_lds32(smem_u32(szst) + uint32_t(sl[u]) * slot_sz + uint32_t(p.sz_bytes) + gl * 32u +
                                                   uint32_t((((n0 >> 3) & 3) + (cz >> 3)) * 4));
                            } else {
                                zraw[u] = im_lds16(sz[u] - uint32_t(g * 4) + uint32_t(p.sz_bytes) + gl * 64u + zoff_lane);
                            }
                        }
                        if (ks == 3) {
                            // after the last read of the slots: hand them back to the producer warp
                            __syncwarp();
#pragma unroll
                            for (int u = 0; u < NT; ++u)
                                if (lane == 0) im_mbar_arrive(empty_base + uint32_t(sl[u]) * 8u);
                        }
#pragma unroll
                        for (int u = 0; u < NT; ++u) {
                            float sq[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) sq[q] = (q & 1) ? cvt16_hi<BF16>(s2[u][q >> 1]) : cvt16_lo<BF16>(s2[u][q >> 1]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int dlo = (q < 2) ? A[u][(q & 1) * 2] : B[u][(q & 1) * 2];
                                const int dhi = (q < 2) ? A[u][(q & 1) * 2 + 1] : B[u][(q & 1) * 2 + 1];
                                yacc[q] = fmaf(sq[q] * wunit[u], float(dhi * 256 + dlo), yacc[q]);
                            }
                            float zf;
                            if constexpr (ASYM) {
                                const float sc = c == 0 ? sq[0] : (c == 1 ? sq[1] : (c == 2 ? sq[2] : sq[3]));
                                zf = sc * float(((zraw[u] >> ((cz & 7) * 4)) & 15u) + 1u);
                            } else {
                                zf = cvt16_lo<BF16>(zraw[u]);
                            }
                            yz = fmaf(zf, xs[u], yz);
                        }
                    }
                }
                xb = xbu[NT - 1];
                slot += NT;
                if (slot >= S) { slot -= S; ph ^= 1u; }
                if (kt0 == 0) CH_TRACE(node, 5);
            };
            for (int kt = 0; kt < tiles;) {
                if ((kt & 1) == 0 && (passes > 1 || !staged_once)) {
                    stage_x();
                    staged_once = true;
                    if (passes > 1) { xv = xw; load_x(next_pass, xw); next_pass = (next_pass + 1 == passes) ? 0 : next_pass + 1; }
                    if (kt == 0) CH_TRACE(node, 1);
                }
                const bool have0 = kt * IM_TILE_ROWS + unit_row < nd.R;
                const bool have1 = (kt + 1 < tiles) && ((kt + 1) * IM_TILE_ROWS + unit_row < nd.R);
                // pairs only when a tile flushes once (groups >= 128 values): with more flushes per tile the interleaved
                // order of the fp32 accumulation would differ from the per-layer kernel's
                if (F == 4 && (kt & 1) == 0 && have0 && have1) {
                    process(std::integral_constant<int, 2>{}, kt);
                    kt += 2;
                } else {
                    if (have0) {
                        process(std::integral_constant<int, 1>{}, kt);
                    } else {
                        // the warp's unit lies below the last packed row of the matrix: nothing to compute
                        __syncwarp();
                        if (lane == 0) im_mbar_arrive(empty_base + uint32_t(slot) * 8u);
                        if (++slot == S) { slot = 0; ph ^= 1u; }
                    }
                    kt += 1;
                }
            }
            CH_TRACE(node, 2);
            // ---- strip finished: the lanes' raw partial sums go to the epilogue warp as they are (five conflict-free stores;
            //      the shuffle tree over the digit-pair lanes and the zero-point term moved there: ~60 fewer instructions per
            //      strip in each of the sixteen warps that set the pace of every K = 4096 layer) ----
            {
                // red is free again once the epilogue warp has read the previous strip
                if (use > 0u) ch_mbar_wait(smem_u32(&freeb[0]), (use - 1u) & 1u);
                ++use;
            }
            float* rw = red + warp * 5 * CH_RED_PLANE + lane;
            rw[0] = yacc[0];
            rw[CH_RED_PLANE] = yacc[1];
            rw[2 * CH_RED_PLANE] = yacc[2];
            rw[3 * CH_RED_PLANE] = yacc[3];
            rw[4 * CH_RED_PLANE] = yz;
            yacc[0] = yacc[1] = yacc[2] = yacc[3] = 0.f;
            yz = 0.f;
            __syncwarp();
            if (lane == 0) mbar_arrive(&part[0]);
        }
    }
}

struct ChainLaunch {
    int F, grid;
    bool asym, bf16;
    size_t smem;
    cudaStream_t stream;
};
int launch_chain(const ChainParams& p, const ChainLaunch& l);

}  // namespace b200bit
