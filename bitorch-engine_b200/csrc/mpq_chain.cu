// mpq_chain.cu -- host side of the decode chain (mpq_chain.cuh): hazard analysis, plan image, launch, C ABI.
#include "mpq_chain.cuh"

#include <stdlib.h>
#include <string.h>
#include <vector>

namespace b200bit {

unsigned long long* trace_buffer();     // mpq_forward.cu (b200bit_set_trace_buffer)

template <int F, bool ASYM, bool BF16, bool TRACE>
static int launch_chain_one(const ChainParams& p, const ChainLaunch& l) {
    auto kern = mpq_chain_kernel<F, ASYM, BF16, TRACE>;
    static bool configured_dev[64] = {false};
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_LIMIT));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(l.grid, 1, 1);
    cfg.blockDim = dim3(CH_THREADS, 1, 1);
    cfg.dynamicSmemBytes = l.smem;
    cfg.stream = l.stream;
    // every CTA spins on counters other CTAs advance: all of them must be resident -> cooperative launch
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    return B200BIT_OK;
}

template <int F, bool TRACE>
static int launch_chain_f(const ChainParams& p, const ChainLaunch& l) {
    if (l.asym) return l.bf16 ? launch_chain_one<F, true, true, TRACE>(p, l) : launch_chain_one<F, true, false, TRACE>(p, l);
    return l.bf16 ? launch_chain_one<F, false, true, TRACE>(p, l) : launch_chain_one<F, false, false, TRACE>(p, l);
}

int launch_chain(const ChainParams& p, const ChainLaunch& l) {
    if (p.trace && l.F == 4) return launch_chain_f<4, true>(p, l);
    switch (l.F) {
        case 1: return launch_chain_f<1, false>(p, l);
        case 2: return launch_chain_f<2, false>(p, l);
        case 4: return launch_chain_f<4, false>(p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "decode chain: flush interval %d", l.F);
}

// plan image in device memory: nodes | tensor maps | counters (+ exit count, error flag, epoch) | LL shadows
static size_t chain_maps_offset(int n) { return (size_t(n) * sizeof(ChainNode) + 127) & ~size_t(127); }
static size_t chain_counters_offset(int n) { return chain_maps_offset(n) + size_t(3 * n) * sizeof(CUtensorMap); }
static size_t chain_shadow_offset(int n) { return chain_counters_offset(n) + (((size_t(n) + 3) * sizeof(unsigned) + 127) & ~size_t(127)); }
static size_t shadow_bytes(int N) { return (size_t(N) * 4 + 127) & ~size_t(127); }      // N / 2 words of 8 bytes
// upper bound: every node's output shadowed
static size_t chain_plan_bytes(const b200bit_chain_node* nodes, int n) {
    size_t b = chain_shadow_offset(n);
    for (int i = 0; i < n; ++i) b += shadow_bytes(nodes[i].N > 0 ? nodes[i].N : 0);
    return b;
}

static bool overlap(const void* a, size_t abytes, const void* b, size_t bbytes) {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(a), b0 = reinterpret_cast<uintptr_t>(b);
    return a0 < b0 + bbytes && b0 < a0 + abytes;
}

enum { CI_MAGIC = 0, CI_NODES, CI_GRID, CI_S, CI_F, CI_ASYM, CI_BF16, CI_SMEM, CI_RPG_SHIFT, CI_SZ, CI_STILE, CI_ZTILE,
       CI_MAX_WAIT, CI_COUNT };
constexpr int CHAIN_MAGIC = 0x43483031;

}  // namespace b200bit

using namespace b200bit;

extern "C" {

size_t b200bit_mpq_chain_plan_bytes(const b200bit_chain_node* nodes, int n_nodes) {
    return (nodes && n_nodes > 0) ? chain_plan_bytes(nodes, n_nodes) : 0;
}

// host_only != nullptr: planning only (hazards, strip deal, shadow assignment) into a host array of n 64-byte node records --
// no tensor maps, nothing touches the device; the shadow pointers are offsets from `plan_device` (any non-null base)
static int chain_build(const b200bit_chain_node* nodes, int n, int w_bit, int asym, int dtype, void* plan_device,
                       size_t plan_bytes, int* info16, void* host_only) {
    B200_REQUIRE(nodes && plan_device && info16, B200BIT_ERR_ARG, "mpq_chain_build: null pointer argument");
    B200_REQUIRE(n > 0 && n <= 4096, B200BIT_ERR_SHAPE, "mpq_chain_build: %d nodes (1..4096)", n);
    B200_REQUIRE(w_bit == 4, B200BIT_ERR_UNSUPPORTED, "mpq_chain_build: w_bit=%d (the decode chain is the 4-bit kernel)", w_bit);
    B200_REQUIRE(dtype == B200BIT_F16 || dtype == B200BIT_BF16, B200BIT_ERR_UNSUPPORTED, "mpq_chain_build: dtype code %d (f16 / bf16)", dtype);
    B200_REQUIRE(plan_bytes >= chain_plan_bytes(nodes, n), B200BIT_ERR_WORKSPACE, "mpq_chain_build: plan buffer of %zu bytes needed, %zu given",
                 chain_plan_bytes(nodes, n), plan_bytes);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(plan_device) & 127) == 0, B200BIT_ERR_ARG, "mpq_chain_build: plan buffer must be 128-byte aligned");
    const int sms = sm_count();
    static const bool direct_poll = getenv("B200BIT_CHAIN_POLL") ? atoi(getenv("B200BIT_CHAIN_POLL")) != 0 : false;
    // ---- one group size per chain (it fixes the flush interval and the scale-tile geometry of the ring) ----
    int rpg = 0;
    int max_strips = 0;
    for (int i = 0; i < n; ++i) {
        const b200bit_chain_node& nd = nodes[i];
        B200_REQUIRE(nd.x && nd.y && nd.qweight && nd.scales && nd.zeros, B200BIT_ERR_ARG, "mpq_chain_build: node %d has a null pointer", i);
        B200_REQUIRE(nd.K > 0 && nd.N > 0 && nd.G > 0 && nd.K % nd.G == 0, B200BIT_ERR_SHAPE, "mpq_chain_build: node %d: bad sizes K=%d N=%d G=%d", i, nd.K, nd.N, nd.G);
        B200_REQUIRE(nd.K % (8 * IM_UNIT_ROWS) == 0, B200BIT_ERR_SHAPE, "mpq_chain_build: node %d: K=%d must be a multiple of 128", i, nd.K);
        B200_REQUIRE(nd.N % 8 == 0 && (!asym || nd.N % 32 == 0), B200BIT_ERR_SHAPE, "mpq_chain_build: node %d: N=%d must be a multiple of %d", i, nd.N, asym ? 32 : 8);
        const int gs = nd.K / nd.G;
        B200_REQUIRE(gs % 32 == 0 && ((gs / 8) & (gs / 8 - 1)) == 0, B200BIT_ERR_SHAPE, "mpq_chain_build: node %d: group size %d must be 32 * 2^i", i, gs);
        if (i == 0) rpg = gs / 8;
        B200_REQUIRE(gs / 8 == rpg, B200BIT_ERR_UNSUPPORTED, "mpq_chain_build: node %d: group size %d differs from node 0 (%d): one group size per chain", i, gs, rpg * 8);
        B200_REQUIRE(!overlap(nd.x, size_t(nd.K) * 2, nd.y, size_t(nd.N) * 2), B200BIT_ERR_ARG, "mpq_chain_build: node %d writes its own input", i);
        const int strips = (nd.N + IM_COLS - 1) / IM_COLS;
        if (strips > max_strips) max_strips = strips;
    }
    int rpg_shift = 0;
    while ((1 << rpg_shift) < rpg) ++rpg_shift;
    const int gps = rpg <= IM_TILE_ROWS ? IM_TILE_ROWS / rpg : 1;
    const int F = (rpg < IM_UNIT_ROWS ? rpg : IM_UNIT_ROWS) / 4;
    const int s_tile = gps * 64, z_tile = asym ? gps * 32 : gps * 64;
    const int sz_bytes = (s_tile + 127) & ~127;
    const int grid = max_strips < sms ? max_strips : sms;
    const size_t fixed = size_t(IM_XIMG_BYTES) + size_t(CH_RED_FLOATS) * sizeof(float) + size_t(n) * sizeof(ChainNode) +
                         (2 * CH_MAX_STAGES + 5) * 8 + 64;
    static const int slot_cap = getenv("B200BIT_CHAIN_SLOTS") ? atoi(getenv("B200BIT_CHAIN_SLOTS")) : CH_MAX_STAGES;   // sweep hook
    int S = slot_cap < CH_MAX_STAGES ? (slot_cap < 2 ? 2 : slot_cap) : CH_MAX_STAGES;
    for (; S >= 1; --S)
        if (fixed + size_t(S) * (IM_TILE_BYTES + 2 * sz_bytes) <= size_t(CH_SMEM_LIMIT)) break;
    B200_REQUIRE(S >= 2, B200BIT_ERR_SHAPE, "mpq_chain_build: %d nodes leave no room for the weight ring", n);

    // host image of everything in front of the shadows (those only need their zero fill, which the caller provides)
    std::vector<unsigned char> image(chain_shadow_offset(n), 0);
    size_t shadow_at = chain_shadow_offset(n);
    unsigned char* dev_base = reinterpret_cast<unsigned char*>(plan_device);
    ChainNode* cn = reinterpret_cast<ChainNode*>(image.data());
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(image.data() + chain_maps_offset(n));
    int off = 0, max_wait = -1;
    for (int i = 0; i < n; ++i) {
        const b200bit_chain_node& nd = nodes[i];
        ChainNode& c = cn[i];
        c.x = reinterpret_cast<const uint16_t*>(nd.x);
        c.y = reinterpret_cast<uint16_t*>(nd.y);
        c.R = nd.K / 8;
        c.N = nd.N;
        c.tiles = (c.R + IM_TILE_ROWS - 1) / IM_TILE_ROWS;
        int strips = (nd.N + IM_COLS - 1) / IM_COLS, n28 = strips;
        {   // round the strip count up to a multiple of the grid with some 24-column strips when that costs < 3 % extra
            // traffic (4096 columns -> 136 x 28 + 12 x 24 = 148 strips), as the per-layer kernel does (plan_imma)
            const int want = (strips + grid - 1) / grid * grid;
            const int m28 = (nd.N - 24 * want) / 4;
            if (want != strips && nd.N % 4 == 0 && (nd.N - 24 * want) >= 0 && m28 <= want && (want - m28) * 4 * 100 < 3 * nd.N) {
                strips = want;
                n28 = m28;
            }
        }
        c.strips = strips;
        c.n28 = n28;
        c.off_sig = off;
        off = (off + strips) % grid;
        // hazards against every earlier node (the last one that conflicts is enough: CTAs finish nodes in order)
        c.wx_node = c.wy_node = -1;
        c.xll = nullptr;
        c.yll = nullptr;
        for (int j = i - 1; j >= 0 && c.wx_node < 0; --j)
            if (overlap(nd.x, size_t(nd.K) * 2, nodes[j].y, size_t(nodes[j].N) * 2)) c.wx_node = j;
        for (int j = i - 1; j >= 0 && c.wy_node < 0; --j)
            if (overlap(nd.y, size_t(nd.N) * 2, nodes[j].y, size_t(nodes[j].N) * 2) ||
                overlap(nd.y, size_t(nd.N) * 2, nodes[j].x, size_t(nodes[j].K) * 2)) c.wy_node = j;
        if (c.wx_node >= 0 && nodes[c.wx_node].y == nd.x && nodes[c.wx_node].N == nd.K) {
            // x is exactly the output of node wx_node: read its shadow (values + epoch in one word), no counter wait
            ChainNode& prod = cn[c.wx_node];
            if (prod.yll == nullptr) {
                prod.yll = reinterpret_cast<uint64_t*>(dev_base + shadow_at);
                shadow_at += shadow_bytes(prod.N);
            }
            c.xll = prod.yll;
            // hint mode: the counter wait stays as a hint (relaxed, unordered) -- polling the shadow words only starts
            // once the producer's strips have been counted; direct mode: the consumer's warps poll their own rows of the
            // shadow from the start (one L2 round trip less on the dependent chain, more polling traffic)
            if (direct_poll) c.wx_node = -1;
            else prod.off_sig |= 1 << 21;
        } else {
            if (c.wy_node >= 0 && c.wy_node <= c.wx_node) c.wy_node = -1;     // already implied by the wait in front of x
            if (c.wx_node >= 0) cn[c.wx_node].off_sig |= 1 << 20;
        }
        if (c.wy_node >= 0) cn[c.wy_node].off_sig |= 1 << 20;
        if (i > 0 && nodes[i - 1].x == nd.x && nodes[i - 1].K == nd.K) c.off_sig |= 1 << 22;      // sibling: same x
        if (c.wx_node > max_wait) max_wait = c.wx_node;
    }
    if (host_only) {
        memcpy(host_only, image.data(), size_t(n) * sizeof(ChainNode));
    } else {
        for (int i = 0; i < n; ++i) {
            const b200bit_chain_node& nd = nodes[i];
            int rc = make_map_2d(&maps[3 * i], CU_TENSOR_MAP_DATA_TYPE_UINT32, nd.qweight, uint64_t(nd.N), uint64_t(nd.K / 8),
                                 uint64_t(nd.N) * 4, IM_COLS, IM_TILE_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc != B200BIT_OK) return rc;
            rc = make_map_2d(&maps[3 * i + 1], CU_TENSOR_MAP_DATA_TYPE_UINT16, nd.scales, uint64_t(nd.N), uint64_t(nd.G),
                             uint64_t(nd.N) * 2, 32, uint32_t(gps), CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc != B200BIT_OK) return rc;
            if (asym)
                rc = make_map_2d(&maps[3 * i + 2], CU_TENSOR_MAP_DATA_TYPE_UINT32, nd.zeros, uint64_t(nd.N / 8), uint64_t(nd.G),
                                 uint64_t(nd.N / 8) * 4, 8, uint32_t(gps), CU_TENSOR_MAP_SWIZZLE_NONE);
            else
                rc = make_map_2d(&maps[3 * i + 2], CU_TENSOR_MAP_DATA_TYPE_UINT16, nd.zeros, uint64_t(nd.N), uint64_t(nd.G),
                                 uint64_t(nd.N) * 2, 32, uint32_t(gps), CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc != B200BIT_OK) return rc;
        }
    }
    reinterpret_cast<unsigned*>(image.data() + chain_counters_offset(n))[n + 2] = 1u;      // first launch epoch
    if (!host_only) {
        B200_CUDA_OK(cudaMemcpy(plan_device, image.data(), image.size(), cudaMemcpyHostToDevice));
        B200_CUDA_OK(cudaMemset(dev_base + chain_shadow_offset(n), 0, shadow_at - chain_shadow_offset(n)));
    }
    memset(info16, 0, 16 * sizeof(int));
    info16[CI_MAGIC] = CHAIN_MAGIC; info16[CI_NODES] = n; info16[CI_GRID] = grid; info16[CI_S] = S; info16[CI_F] = F;
    info16[CI_ASYM] = asym ? 1 : 0; info16[CI_BF16] = dtype == B200BIT_BF16 ? 1 : 0;
    info16[CI_SMEM] = int(fixed + size_t(S) * (IM_TILE_BYTES + 2 * sz_bytes));
    info16[CI_RPG_SHIFT] = rpg_shift; info16[CI_SZ] = sz_bytes; info16[CI_STILE] = s_tile; info16[CI_ZTILE] = z_tile;
    info16[CI_MAX_WAIT] = max_wait;
    return B200BIT_OK;
}

int b200bit_mpq_chain_build(const b200bit_chain_node* nodes, int n, int w_bit, int asym, int dtype, void* plan_device,
                            size_t plan_bytes, int* info16) {
    return chain_build(nodes, n, w_bit, asym, dtype, plan_device, plan_bytes, info16, nullptr);
}

int b200bit_mpq_chain_plan_host(const b200bit_chain_node* nodes, int n, int w_bit, int asym, int dtype, void* node_table_host,
                                int* info16) {
    B200_REQUIRE(node_table_host != nullptr, B200BIT_ERR_ARG, "mpq_chain_plan_host: null output");
    // a fake, aligned, never dereferenced base: shadow pointers in the table are base + offset
    return chain_build(nodes, n, w_bit, asym, dtype, reinterpret_cast<void*>(uintptr_t(1) << 20), ~size_t(0), info16, node_table_host);
}

int b200bit_mpq_chain_launch(void* plan_device, const int* info16, unsigned flags, void* stream_) {
    (void)flags;
    B200_REQUIRE(plan_device && info16, B200BIT_ERR_ARG, "mpq_chain_launch: null pointer argument");
    B200_REQUIRE(info16[CI_MAGIC] == CHAIN_MAGIC, B200BIT_ERR_ARG, "mpq_chain_launch: info block was not written by b200bit_mpq_chain_build");
    const int n = info16[CI_NODES];
    unsigned char* base = reinterpret_cast<unsigned char*>(plan_device);
    ChainParams p{};
    p.nodes = reinterpret_cast<const ChainNode*>(base);
    p.maps = reinterpret_cast<const CUtensorMap*>(base + chain_maps_offset(n));
    p.counters = reinterpret_cast<unsigned*>(base + chain_counters_offset(n));
    p.n_nodes = n;
    p.S = info16[CI_S];
    p.rpg_shift = info16[CI_RPG_SHIFT];
    p.sz_bytes = info16[CI_SZ];
    p.s_tile_bytes = info16[CI_STILE];
    p.z_tile_bytes = info16[CI_ZTILE];
    static const int poll_depth = getenv("B200BIT_CHAIN_POLLS") ? atoi(getenv("B200BIT_CHAIN_POLLS")) : 1;     // sweep hook; measured 1 / 2 / 4 in flight: 1164.6 / 1162.4 / 1156.5 tokens/s
    p.poll_depth = poll_depth;
    static const int early_pct = getenv("B200BIT_CHAIN_EARLY") ? atoi(getenv("B200BIT_CHAIN_EARLY")) : 50;     // sweep hook; measured 0 / 25 / 50 / 75 / 95 %: 1153.6 / 1220.0 / 1239.5 / 1234.9 / 1232.2 tokens/s
    p.early_pct = early_pct < 0 ? 0 : (early_pct > 99 ? 99 : early_pct);
    p.trace = trace_buffer();
    ChainLaunch l{};
    l.F = info16[CI_F]; l.grid = info16[CI_GRID]; l.asym = info16[CI_ASYM] != 0; l.bf16 = info16[CI_BF16] != 0;
    l.smem = size_t(info16[CI_SMEM]);
    l.stream = reinterpret_cast<cudaStream_t>(stream_);
    return launch_chain(p, l);
}

int b200bit_mpq_chain_status(const void* plan_device, const int* info16, int* error_flag_host, void* stream_) {
    B200_REQUIRE(plan_device && info16 && error_flag_host, B200BIT_ERR_ARG, "mpq_chain_status: null pointer argument");
    B200_REQUIRE(info16[CI_MAGIC] == CHAIN_MAGIC, B200BIT_ERR_ARG, "mpq_chain_status: bad info block");
    const int n = info16[CI_NODES];
    const unsigned char* base = reinterpret_cast<const unsigned char*>(plan_device);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    unsigned v = 0;
    B200_CUDA_OK(cudaMemcpyAsync(&v, base + chain_counters_offset(n) + (size_t(n) + 1) * sizeof(unsigned), sizeof(unsigned),
                                 cudaMemcpyDeviceToHost, stream));
    B200_CUDA_OK(cudaStreamSynchronize(stream));
    *error_flag_host = int(v);
    return B200BIT_OK;
}

}  // extern "C"
