// mpq_forward.cu -- C-ABI entry b200bit_mpq_forward: validation, path selection (decode GEMV / batched GEMM /
// general fallback) and the general fallback kernel.
//
// Reference path replaced: q_linear_cuda.mpq_forward -> mpq_linear_cuda_forward -> quantmatmul_cuda
// (bitorch_engine/layers/qlinear/nbit/cuda/q_linear_cuda.cpp:258-270, mpq_linear_cuda_kernel.cu:603-626, 482-577).
#include "mpq_gemv.cuh"
#include "mpq_mma.cuh"
#include "mpq_stream.cuh"
#include "mpq_pipe.cuh"
#include "mpq_pipe_mma.cuh"
#include "mpq_imma.cuh"

#include <stdlib.h>
#include <string.h>
#include <mutex>

namespace b200bit {

// ---------------------------------------------------------------------------------------------------------------
// error plumbing / device info
// ---------------------------------------------------------------------------------------------------------------
char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// process-wide tuning override for sweeps (0 = heuristic); set through b200bit_set_gemv_tuning()
static int g_tune_L = 0, g_tune_warps = 0, g_tune_splitk = 0;
// 0 = auto, 1 = CUDA-core FHFMA GEMV, 2 = mma.sync small-batch kernel, 3 = general fallback,
// 4 = TMA-streamed small-batch kernel, 5 = tcgen05 batched kernel (mpq_tc.cu, more than 32 rows), 6 = cross-kernel pipelined decode GEMV (mpq_pipe.cuh),
// 7 = persistent integer-tensor-pipe decode GEMV (mpq_imma.cuh)
static int g_path = 0;
// in auto mode, does M == 1 go to the tensor kernel (1) or stay on the CUDA-core GEMV (0)?
static int g_mma_for_m1 = 0;
static unsigned long long* g_trace = nullptr;   // diagnostics: per-warp globaltimer stamps of the stream kernel
unsigned long long* trace_buffer() { return g_trace; }

// ---------------------------------------------------------------------------------------------------------------
// Programmatic-dependent-launch chains (PipeParams::early).  Per (device, stream): the output ranges of the decode
// kernels this library launched last -- one late-trigger kernel followed by the early ("sibling") kernels admitted
// behind it.  A new launch may run early iff the caller promises that x is ready (B200BIT_FLAG_INPUT_READY), the
// record is valid (the launch in front was a pipe kernel: every other kernel of the library triggers its dependents
// before its own dependency wait) and x overlaps none of the recorded outputs.  Anything else starts a new chain.
// Kernels launched by others between two calls do not trigger early, so they only make the order stricter.
// ---------------------------------------------------------------------------------------------------------------
struct ChainState {
    int dev;
    cudaStream_t stream;
    bool used, valid;
    int n;
    uintptr_t lo[8], hi[8];
    unsigned long long tick;
};
static ChainState g_chain[32];
static std::mutex g_chain_mu;
static unsigned long long g_chain_tick = 0;

static ChainState* chain_entry(cudaStream_t stream) {      // caller holds g_chain_mu
    int dev = 0;
    cudaGetDevice(&dev);
    ChainState* victim = &g_chain[0];
    for (ChainState& e : g_chain) {
        if (e.used && e.dev == dev && e.stream == stream) { e.tick = ++g_chain_tick; return &e; }
        if (!e.used) { if (victim->used) victim = &e; }
        else if (victim->used && e.tick < victim->tick) victim = &e;
    }
    *victim = ChainState{};
    victim->used = true; victim->dev = dev; victim->stream = stream; victim->tick = ++g_chain_tick;
    return victim;
}
// B200BIT_EARLY: 0 = never run early, 1 = whenever the caller's promise and the record allow it (default),
// 2 = only layers whose whole CTA range is resident in the ring before the wait (resident = true)
static bool chain_admit(cudaStream_t stream, const void* x, size_t xbytes, const void* y, size_t ybytes, unsigned flags,
                        bool resident = true) {
    static const int mode = getenv("B200BIT_EARLY") ? atoi(getenv("B200BIT_EARLY")) : 1;
    const bool enabled = mode == 1 || (mode == 2 && resident);
    std::lock_guard<std::mutex> lk(g_chain_mu);
    ChainState* e = chain_entry(stream);
    const uintptr_t xl = reinterpret_cast<uintptr_t>(x), xh = xl + xbytes;
    const uintptr_t yl = reinterpret_cast<uintptr_t>(y), yh = yl + ybytes;
    bool early = enabled && (flags & B200BIT_FLAG_PDL) && (flags & B200BIT_FLAG_INPUT_READY) && e->valid && e->n < 8;
    for (int i = 0; early && i < e->n; ++i)
        if (xl < e->hi[i] && e->lo[i] < xh) early = false;
    if (!early) e->n = 0;
    e->lo[e->n] = yl; e->hi[e->n] = yh; ++e->n;
    e->valid = true;
    return early;
}
static void chain_invalidate(cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_chain_mu);
    chain_entry(stream)->valid = false;
}

// ---------------------------------------------------------------------------------------------------------------
// General fallback: any g_idx (act-order), any dtype incl. f32, any N / group size.  One thread per column,
// fp32 math on the exact model (s*q - z), M tiled by 4 through gridDim.y.  Correctness path, not a fast path.
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ float load_el(const void* p, size_t i) {
    if constexpr (DT == B200BIT_F32) return reinterpret_cast<const float*>(p)[i];
    else if constexpr (DT == B200BIT_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    else return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
template <int DT>
__device__ __forceinline__ void store_el(void* p, size_t i, float v) {
    if constexpr (DT == B200BIT_F32) reinterpret_cast<float*>(p)[i] = v;
    else if constexpr (DT == B200BIT_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

template <int DT>
__global__ void __launch_bounds__(128) mpq_general_kernel(const void* __restrict__ x, const uint32_t* __restrict__ qw,
                                                          const void* __restrict__ scales,
                                                          const void* __restrict__ zeros,
                                                          const int32_t* __restrict__ g_idx, void* __restrict__ y,
                                                          int M, int K, int N, int G, int w_bit, int asym) {
    pdl_launch_dependents();
    pdl_wait_primary();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m0 = blockIdx.y * 4;
    const int mc = min(4, M - m0);
    if (n >= N) return;
    const int nb = 32 / w_bit;
    const uint32_t mask = (w_bit == 32) ? 0xffffffffu : ((1u << w_bit) - 1u);
    const int gs = K / G;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int R = K / nb;
    int g_prev = -1;
    float s = 0.f, z = 0.f;
    for (int r = 0; r < R; ++r) {
        const uint32_t w = qw[size_t(r) * N + n];
        for (int j = 0; j < nb; ++j) {
            const int k = r * nb + j;
            const int g = g_idx ? g_idx[k] : k / gs;
            if (g != g_prev) {
                s = load_el<DT>(scales, size_t(g) * N + n);
                if (asym) {
                    const uint32_t zw = reinterpret_cast<const uint32_t*>(zeros)[size_t(g) * (N / nb) + n / nb];
                    z = s * float(((zw >> ((n % nb) * w_bit)) & mask) + 1u);
                } else {
                    z = load_el<DT>(zeros, size_t(g) * N + n);
                }
                g_prev = g;
            }
            const float wv = fmaf(s, float((w >> (j * w_bit)) & mask), -z);
            for (int m = 0; m < mc; ++m) acc[m] = fmaf(load_el<DT>(x, size_t(m0 + m) * K + k), wv, acc[m]);
        }
    }
    for (int m = 0; m < mc; ++m) store_el<DT>(y, size_t(m0 + m) * N + n, acc[m]);
}

static int launch_general(const void* x, const int32_t* qw, const void* scales, const void* zeros,
                          const int32_t* g_idx, void* y, int M, int K, int N, int G, int w_bit, int asym, int dtype,
                          unsigned flags, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((N + 127) / 128, (M + 3) / 4, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (flags & B200BIT_FLAG_PDL) ? 1 : 0;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(qw);
    switch (dtype) {
        case B200BIT_F32:
            B200_CUDA_OK(cudaLaunchKernelEx(&cfg, mpq_general_kernel<B200BIT_F32>, x, q, scales, zeros, g_idx, y, M, K,
                                            N, G, w_bit, asym));
            break;
        case B200BIT_F16:
            B200_CUDA_OK(cudaLaunchKernelEx(&cfg, mpq_general_kernel<B200BIT_F16>, x, q, scales, zeros, g_idx, y, M, K,
                                            N, G, w_bit, asym));
            break;
        default:
            B200_CUDA_OK(cudaLaunchKernelEx(&cfg, mpq_general_kernel<B200BIT_BF16>, x, q, scales, zeros, g_idx, y, M,
                                            K, N, G, w_bit, asym));
    }
    return B200BIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// decode GEMV launch plan
// ---------------------------------------------------------------------------------------------------------------
struct GemvPlan {
    bool ok;
    int FR, L_log2, warps, splitk, runs_total, runs_per_split, rpr, rpr_shift;
    size_t smem;
};

static size_t gemv_smem_bytes(int M, int nb, int nruns, int FR, int warps, int L) {
    const size_t xs = size_t(M) * nruns * (GEMV_RUN * nb + 8) * 2;
    const size_t xseg = size_t((M * nruns * (GEMV_RUN / FR) + 3) & ~3) * 4;
    return xs + xseg + size_t(warps) * M * 4 * L * 4;
}

static GemvPlan plan_gemv(int M, int K, int N, int G, int w_bit, int asym, int dtype, bool trivial_gidx) {
    GemvPlan pl{};
    pl.ok = false;
    if (!trivial_gidx || dtype == B200BIT_F32) return pl;
    if (dtype == B200BIT_BF16 && w_bit > 4) return pl;
    const int nb = 32 / w_bit;
    if (K % (nb * GEMV_RUN) != 0 || K % G != 0) return pl;      // whole runs only
    if (asym && N % nb != 0) return pl;
    const int gs = K / G;
    if (gs % nb != 0) return pl;
    const int rpg = gs / nb;
    if (rpg >= GEMV_RUN) {
        if (rpg % GEMV_RUN != 0) return pl;                     // runs would straddle groups
        pl.FR = GEMV_RUN;
        pl.rpr = rpg / GEMV_RUN;
        pl.rpr_shift = -1;
        for (int sh = 0; sh < 30; ++sh) if ((1 << sh) == pl.rpr) pl.rpr_shift = sh;
    } else {
        if (GEMV_RUN % rpg != 0) return pl;
        pl.FR = rpg;
        pl.rpr = 1;
        pl.rpr_shift = 0;
    }
    const int R = K / nb;
    pl.runs_total = R / GEMV_RUN;
    // lanes per row segment / warps / split-K: measured on B200 (profiles/r1_gemv_sweep.md): few wide-K strips ->
    // one 16-warp CTA per 32-column strip; many strips (N >= ~2 strips per SM) -> 64-column strips, 4 warps, split-K 4
    const bool many_strips = (N / 32) >= 2 * sm_count();
    int L = g_tune_L ? g_tune_L : (many_strips ? 16 : 8);
    while (L > 8 && N % (4 * L) != 0) L >>= 1;
    if (N % (4 * L) != 0) return pl;
    pl.L_log2 = L == 32 ? 5 : L == 16 ? 4 : 3;
    const int LG = 32 / L;
    pl.warps = g_tune_warps ? g_tune_warps : (many_strips ? 4 : 16);
    const int slots = pl.warps * LG;
    const int strips = N / (4 * L);
    int splitk;
    if (g_tune_splitk) {
        splitk = g_tune_splitk;
    } else {
        splitk = many_strips ? 4 : 1;
        const int most = pl.runs_total / slots > 0 ? pl.runs_total / slots : 1;   // >= one run per slot
        if (splitk > most) splitk = most;
    }
    (void)strips;
    if (splitk < 1) splitk = 1;
    if (splitk > pl.runs_total) splitk = pl.runs_total;
    if (splitk > 64) splitk = 64;
    // bound shared memory (x chunk) by splitting K further
    for (;; ++splitk) {
        pl.runs_per_split = (pl.runs_total + splitk - 1) / splitk;
        pl.smem = gemv_smem_bytes(M, nb, pl.runs_per_split, pl.FR, pl.warps, L);
        if (pl.smem <= 160 * 1024 || splitk >= pl.runs_total || splitk >= 64) break;
    }
    pl.splitk = (pl.runs_total + pl.runs_per_split - 1) / pl.runs_per_split;
    pl.ok = pl.smem <= 200 * 1024;
    return pl;
}

static int launch_gemv(const GemvParams& p, const GemvLaunch& l, int w_bit, bool bf16) {
    switch (w_bit) {
        case 1: return bf16 ? launch_gemv_family<1, true>(p, l) : launch_gemv_family<1, false>(p, l);
        case 2: return bf16 ? launch_gemv_family<2, true>(p, l) : launch_gemv_family<2, false>(p, l);
        case 4: return bf16 ? launch_gemv_family<4, true>(p, l) : launch_gemv_family<4, false>(p, l);
        case 8: if (!bf16) return launch_gemv_family<8, false>(p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "gemv: w_bit=%d bf16=%d", w_bit, int(bf16));
}


// ---------------------------------------------------------------------------------------------------------------
// small-batch mma.sync launch plan
// ---------------------------------------------------------------------------------------------------------------
struct MmaPlan {
    bool ok;
    int FJ, warps, splitk, runs_total, runs_per_split, rpr, rpr_shift;
};

static size_t mma_smem_bytes(int M, int nb, int nruns, int FJ, int warps) {
    const size_t chunk_rows = size_t(nruns) * MMA_RUN_ROWS;
    const size_t xs = size_t(M) * (chunk_rows * nb + 32) * 2;
    const size_t xseg = size_t((M * nruns * (MMA_U / FJ) + 3) & ~3) * 4;
    return xs + xseg + size_t(warps) * M * 32 * 4;
}

static MmaPlan plan_mma(int M, int K, int N, int G, int w_bit, int asym, int dtype, bool trivial_gidx) {
    MmaPlan pl{};
    pl.ok = false;
    if (!trivial_gidx || dtype != B200BIT_F16) return pl;
    if (w_bit != 2 && w_bit != 4 && w_bit != 8) return pl;
    const int nb = 32 / w_bit;
    if (N % 32 != 0 || K % (nb * MMA_RUN_ROWS) != 0 || K % G != 0) return pl;
    if (asym && N % nb != 0) return pl;
    const int gs = K / G;
    if (gs % nb != 0) return pl;
    const int rpg = gs / nb;
    if (rpg % 4 != 0) return pl;
    if (rpg >= MMA_RUN_ROWS) {
        if (rpg % MMA_RUN_ROWS != 0) return pl;
        pl.FJ = MMA_U;
        pl.rpr = rpg / MMA_RUN_ROWS;
        pl.rpr_shift = -1;
        for (int sh = 0; sh < 30; ++sh) if ((1 << sh) == pl.rpr) pl.rpr_shift = sh;
    } else {
        if (MMA_RUN_ROWS % rpg != 0) return pl;
        pl.FJ = rpg / 4;
        pl.rpr = 1;
        pl.rpr_shift = 0;
    }
    pl.runs_total = (K / nb) / MMA_RUN_ROWS;
    pl.warps = g_tune_warps ? (g_tune_warps > 8 ? 8 : g_tune_warps) : 4;
    const int strips = N / 32;
    int splitk;
    if (g_tune_splitk) {
        splitk = g_tune_splitk;
    } else {
        const int want = (4 * sm_count() + strips - 1) / strips;
        const int most = pl.runs_total / pl.warps > 0 ? pl.runs_total / pl.warps : 1;
        splitk = want < most ? want : most;
    }
    if (splitk < 1) splitk = 1;
    if (splitk > pl.runs_total) splitk = pl.runs_total;
    if (splitk > 64) splitk = 64;
    const int mm = M < 32 ? M : 32;
    for (;; ++splitk) {
        pl.runs_per_split = (pl.runs_total + splitk - 1) / splitk;
        if (mma_smem_bytes(mm, nb, pl.runs_per_split, pl.FJ, pl.warps) <= 100 * 1024 || splitk >= pl.runs_total ||
            splitk >= 64)
            break;
    }
    pl.splitk = (pl.runs_total + pl.runs_per_split - 1) / pl.runs_per_split;
    pl.ok = mma_smem_bytes(mm, nb, pl.runs_per_split, pl.FJ, pl.warps) <= 200 * 1024;
    return pl;
}

static int launch_mma(const MmaParams& p, const MmaLaunch& l, int w_bit) {
    switch (w_bit) {
        case 2: return launch_mma_family<2>(p, l);
        case 4: return launch_mma_family<4>(p, l);
        case 8: return launch_mma_family<8>(p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "mma: w_bit=%d", w_bit);
}


// ---------------------------------------------------------------------------------------------------------------
// TMA-streamed small-batch kernel: plan + tensor maps
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}

int make_map_2d(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                       uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
    EncodeTiledFn fn = get_encode_fn();
    B200_REQUIRE(fn != nullptr, B200BIT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    const CUresult rc = fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(rc == CUDA_SUCCESS, B200BIT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(rc));
    return B200BIT_OK;
}

struct StreamPlan {
    bool ok;
    int FJ, warps, grid, S, ngr, rpr, rpr_shift, rps, strips, xs, maxseg;
    size_t smem;
};

static StreamPlan plan_stream(int M, int K, int N, int G, int w_bit, int asym, int dtype, bool trivial_gidx) {
    StreamPlan pl{};
    pl.ok = false;
    if (!trivial_gidx || dtype != B200BIT_F16) return pl;
    if (w_bit != 2 && w_bit != 4 && w_bit != 8) return pl;
    const int nb = 32 / w_bit;
    if (N % 32 != 0 || K % (nb * ST_RUN_ROWS) != 0 || K % G != 0) return pl;
    if (asym && (w_bit == 2 || N % (4 * nb) != 0)) return pl;
    const int gs = K / G;
    if (gs % nb != 0) return pl;
    const int rpg = gs / nb;
    if (rpg == 8 || rpg == 16) {
        pl.FJ = rpg / 4; pl.ngr = 32 / rpg; pl.rpr = 1; pl.rpr_shift = 0;
    } else if (rpg % 32 == 0) {
        pl.FJ = 8; pl.ngr = 1; pl.rpr = rpg / 32; pl.rpr_shift = -1;
        for (int sh = 0; sh < 30; ++sh) if ((1 << sh) == pl.rpr) pl.rpr_shift = sh;
    } else {
        return pl;
    }
    pl.strips = N / 32;
    pl.rps = (K / nb) / ST_RUN_ROWS;
    const int cps = g_tune_splitk > 0 ? g_tune_splitk : 1;       // CTAs per SM (sweep hook)
    int grid = sm_count() * cps;
    if (grid > pl.strips) grid = pl.strips;
    pl.grid = grid;
    const int strips_max = (pl.strips + grid - 1) / grid;          // strips of the busiest CTA
    const long long rc_ll = (long long)strips_max * pl.rps;
    if (rc_ll > (1 << 20)) return pl;
    const int rc_max = int(rc_ll);
    const int mm = M < 32 ? M : 32;
    // decode configuration: x staged in shared memory (M <= 8 and the permuted x image fits), up to 16 consumer warps
    const size_t xs_bytes = size_t(mm) * (size_t(K) * 2 + 64) + ST_RUN_ROWS * nb * 2 +
                            size_t(mm) * ((K / nb) / (4 * pl.FJ)) * 4 + 64;
    pl.xs = (mm <= 8 && xs_bytes <= 72 * 1024 && g_tune_L != 32) ? 1 : 0;
    const int max_warps = pl.xs ? ST_MAX_WARPS : 8;
    int warps;
    if (g_tune_warps) {
        warps = g_tune_warps;
    } else {
        const int rounds = (rc_max + max_warps - 1) / max_warps;
        warps = (rc_max + rounds - 1) / rounds;
        if (warps < 4) warps = 4;
    }
    if (warps > max_warps) warps = max_warps;
    if (warps > rc_max) warps = rc_max;
    for (;; warps = (warps + 1) / 2) {                               // shrink the CTA until it fits in shared memory
        pl.warps = warps;
        const int q = (rc_max + warps - 1) / warps;
        const int segs = (q + pl.rps - 1) / pl.rps + 1;              // strips a warp's run range can touch
        if (segs > ST_MAXSEG) return pl;
        pl.maxseg = segs;
        int depth = 16 / warps;                                     // ring slots per consumer warp
        if (depth < 1) depth = 1;
        if (depth > q) depth = q;
        pl.S = warps * depth;
        pl.smem = size_t(pl.S) * (ST_TILE_BYTES + ST_SZ_BYTES) + size_t(2 * pl.S) * 8 +
                  (ST_MAX_WARPS * ST_MAXSEG + ST_MAX_WARPS + 4) * 4 + size_t(warps) * segs * mm * 32 * 4 +
                  (pl.xs ? xs_bytes : 0) + 1024;
        if (pl.smem <= 200 * 1024 || warps <= 2) break;
    }
    pl.ok = pl.smem <= 200 * 1024;
    return pl;
}

// ---------------------------------------------------------------------------------------------------------------
// cross-kernel pipelined decode GEMV (mpq_pipe.cuh): M == 1, f16 / bf16, contiguous groups
// ---------------------------------------------------------------------------------------------------------------
struct PipePlan {
    bool ok;
    int FS, rpg, rpg_shift, gps, stages_total, stages_per_split, splitk, S, sz_bytes, s_tile_bytes, z_tile_bytes, strips;
    size_t smem;
};

static PipePlan plan_pipe(int M, int K, int N, int G, int w_bit, int asym, int dtype, bool trivial_gidx) {
    PipePlan pl{};
    pl.ok = false;
    if (M != 1 || !trivial_gidx || dtype == B200BIT_F32) return pl;
    if (w_bit != 2 && w_bit != 4 && w_bit != 8) return pl;
    if (dtype == B200BIT_BF16 && w_bit > 4) return pl;
    const int nb = 32 / w_bit;
    if (N % 32 != 0 || K % (nb * PG_UNIT_ROWS) != 0 || K % G != 0) return pl;
    if (asym && (w_bit == 2 || N % (4 * nb) != 0)) return pl;       // packed zero rows: TMA box >= 16 B, stride % 16 == 0
    const int gs = K / G;
    if (gs % nb != 0) return pl;
    const int rpg = gs / nb;
    if (rpg % 4 != 0) return pl;
    if (rpg >= PG_UNIT_ROWS ? (rpg % PG_UNIT_ROWS != 0) : (PG_UNIT_ROWS % rpg != 0)) return pl;
    pl.rpg = rpg;
    pl.rpg_shift = -1;
    if (rpg <= PG_STAGE_ROWS) {
        if (PG_STAGE_ROWS % rpg != 0) return pl;
        for (int sh = 0; sh < 8; ++sh) if ((1 << sh) == rpg) pl.rpg_shift = sh;
        if (pl.rpg_shift < 0) return pl;
        pl.gps = PG_STAGE_ROWS / rpg;
    } else {
        if (rpg % PG_STAGE_ROWS != 0) return pl;
        pl.gps = 1;
    }
    pl.FS = (rpg < PG_UNIT_ROWS ? rpg : PG_UNIT_ROWS) / 4;
    pl.s_tile_bytes = pl.gps * 64;
    pl.z_tile_bytes = asym ? pl.gps * (128 / nb) : pl.gps * 64;
    pl.sz_bytes = (pl.s_tile_bytes + 127) & ~127;
    const int R = K / nb;
    pl.stages_total = (R + PG_STAGE_ROWS - 1) / PG_STAGE_ROWS;
    pl.strips = N / 32;
    // split K so that a CTA's packed weights fit its ring whole (everything is requested before griddepcontrol.wait);
    // sweep hook: g_tune_splitk forces the split, g_tune_warps the ring depth
    int S = g_tune_warps > 0 ? g_tune_warps : PG_MAX_STAGES;
    if (S > PG_MAX_STAGES) S = PG_MAX_STAGES;
    int splitk = g_tune_splitk > 0 ? g_tune_splitk : (pl.stages_total + S - 1) / S;
    if (splitk > pl.stages_total) splitk = pl.stages_total;
    if (splitk > 64) splitk = 64;
    pl.stages_per_split = (pl.stages_total + splitk - 1) / splitk;
    pl.splitk = (pl.stages_total + pl.stages_per_split - 1) / pl.stages_per_split;
    if (S > pl.stages_per_split) S = pl.stages_per_split;
    pl.S = S;
    const size_t xseg = size_t(pl.stages_per_split) * PG_STAGE_ROWS * nb / 16 * 4;
    pl.smem = size_t(S) * (PG_TILE_BYTES + 2 * pl.sz_bytes) + 2 * PG_MAX_STAGES * 8 + PG_WARPS * 32 * 4 + xseg + 16;
    pl.ok = pl.smem <= 76800 && size_t(pl.strips) * sizeof(unsigned) <= B200BIT_WS_ZERO_OFFSET;
    return pl;
}

// ---------------------------------------------------------------------------------------------------------------
// integer-tensor-pipe decode GEMV (mpq_imma.cuh): M == 1, 4-bit, f16 / bf16, contiguous groups
// ---------------------------------------------------------------------------------------------------------------
struct ImmaPlan {
    bool ok;
    int F, rpg, rpg_shift, gps, tiles, strips, n28, grid, S, sz_bytes, s_tile_bytes, z_tile_bytes;
    size_t smem;
};

static ImmaPlan plan_imma(int M, int K, int N, int G, int w_bit, int asym, int dtype, bool trivial_gidx) {
    ImmaPlan pl{};
    pl.ok = false;
    if (M != 1 || !trivial_gidx || w_bit != 4 || dtype == B200BIT_F32) return pl;
    if (K % (8 * IM_UNIT_ROWS) != 0 || K % G != 0) return pl;
    if (N % 8 != 0 || (asym && N % 32 != 0)) return pl;            // TMA row strides must be multiples of 16 bytes
    const int gs = K / G;
    if (gs % 32 != 0) return pl;
    const int rpg = gs / 8;                                        // packed rows per group: 4, 8, 16, 32, ... (a power of two)
    pl.rpg = rpg;
    pl.rpg_shift = -1;
    for (int sh = 2; sh <= 20; ++sh) if ((1 << sh) == rpg) pl.rpg_shift = sh;
    if (pl.rpg_shift < 0) return pl;
    pl.gps = rpg <= IM_TILE_ROWS ? IM_TILE_ROWS / rpg : 1;
    pl.F = (rpg < IM_UNIT_ROWS ? rpg : IM_UNIT_ROWS) / 4;
    pl.s_tile_bytes = pl.gps * 64;
    pl.z_tile_bytes = asym ? pl.gps * 32 : pl.gps * 64;
    pl.sz_bytes = (pl.s_tile_bytes + 127) & ~127;
    const int R = K / 8;
    pl.tiles = (R + IM_TILE_ROWS - 1) / IM_TILE_ROWS;
    pl.strips = (N + IM_COLS - 1) / IM_COLS;
    pl.n28 = pl.strips;
    const int sms = sm_count();
    pl.grid = 2 * pl.strips <= sms ? pl.strips : sms;      // a full wave whenever the layer has at least half a wave of strips
    {
        // even out the CTAs: round the strip count up to a multiple of the grid by making some strips 24 columns wide
        // (the TMA box stays 28 wide, so each such strip fetches 4 columns twice) when that costs < 3 % extra traffic:
        // 4096 columns -> 136 x 28 + 12 x 24 = 148 strips, exactly one per SM
        const int want = (pl.strips + pl.grid - 1) / pl.grid * pl.grid;
        const int n28 = (N - 24 * want) / 4;
        if (want != pl.strips && N % 4 == 0 && n28 >= 0 && n28 <= want && (want - n28) * 4 * 100 < 3 * N) {
            pl.strips = want;
            pl.n28 = n28;
        }
        if (pl.grid > pl.strips) pl.grid = pl.strips;
    }
    const int T = ((pl.strips + pl.grid - 1) / pl.grid) * pl.tiles;   // most tiles any CTA walks
    const size_t fixed = size_t(IM_XIMG_BYTES) + (2 * IM_WARPS * 32) * sizeof(float) + IM_MAX_STAGES * 8 + IM_MAX_STAGES * 4 + 64;
    int S = g_tune_warps > 0 ? g_tune_warps : IM_MAX_STAGES;          // sweep hook: ring depth
    if (S > IM_MAX_STAGES) S = IM_MAX_STAGES;
    if (S > T) S = T;
    for (; S >= 1; --S) {
        pl.smem = fixed + size_t(S) * (IM_TILE_BYTES + 2 * pl.sz_bytes);
        if (pl.smem <= size_t(IM_SMEM_LIMIT)) break;
    }
    if (S < 1) return pl;
    pl.S = S;
    pl.ok = true;
    return pl;
}

static int launch_pipe(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const PipeParams& p,
                       const PipeLaunch& l, int w_bit, bool bf16) {
    switch (w_bit) {
        case 2: return bf16 ? launch_pipe_family<2, true>(tw, ts, tz, p, l) : launch_pipe_family<2, false>(tw, ts, tz, p, l);
        case 4: return bf16 ? launch_pipe_family<4, true>(tw, ts, tz, p, l) : launch_pipe_family<4, false>(tw, ts, tz, p, l);
        case 8: if (!bf16) return launch_pipe_family<8, false>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "pipe gemv: w_bit=%d bf16=%d", w_bit, int(bf16));
}

static int launch_stream(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const StreamParams& p,
                         const StreamLaunch& l, int w_bit) {
    switch (w_bit) {
        case 2: return launch_stream_family<2>(tw, ts, tz, p, l);
        case 4: return launch_stream_family<4>(tw, ts, tz, p, l);
        case 8: return launch_stream_family<8>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "stream: w_bit=%d", w_bit);
}

}  // namespace b200bit

using namespace b200bit;

extern "C" {

int b200bit_version(void) { return B200BIT_VERSION; }
const char* b200bit_last_error(void) { return err_buf(); }

int b200bit_device_info(int* sm, int* major, int* minor) {
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    if (sm) B200_CUDA_OK(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
    if (major) B200_CUDA_OK(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
    if (minor) B200_CUDA_OK(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
    return B200BIT_OK;
}

/* host-side plan of the 4-bit decode kernel for a shape (no launch, no device needed) */
int b200bit_mpq_decode_plan(int K, int N, int G, int w_bit, int asym, int dtype, int* out8) {
    B200_REQUIRE(out8 != nullptr, B200BIT_ERR_ARG, "mpq_decode_plan: null output pointer");
    B200_REQUIRE(K > 0 && N > 0 && G > 0, B200BIT_ERR_SHAPE, "mpq_decode_plan: bad sizes K=%d N=%d G=%d", K, N, G);
    const ImmaPlan pl = plan_imma(1, K, N, G, w_bit, asym, dtype, true);
    const int vals[8] = {pl.ok ? 1 : 0, pl.grid, pl.strips, pl.n28, pl.tiles, pl.S, pl.F, int(pl.smem)};
    for (int i = 0; i < 8; ++i) out8[i] = pl.ok ? vals[i] : (i == 0 ? 0 : 0);
    return B200BIT_OK;
}

/* sweep hook (bench / tests only): lanes per row segment (8/16/32), warps per CTA, split-K; 0 = heuristic */
int b200bit_set_gemv_tuning(int L, int warps, int splitk) {
    B200_REQUIRE(L == 0 || L == 8 || L == 16 || L == 32, B200BIT_ERR_ARG, "L must be 0, 8, 16 or 32");
    B200_REQUIRE(warps >= 0 && warps <= 16, B200BIT_ERR_ARG, "warps must be in [0,16]");
    B200_REQUIRE(splitk >= 0, B200BIT_ERR_ARG, "splitk must be >= 0");
    g_tune_L = L; g_tune_warps = warps; g_tune_splitk = splitk;
    return B200BIT_OK;
}

/* path override for benchmarks/tests: 0 auto, 1 CUDA-core GEMV, 2 small-batch mma kernel, 3 general fallback,
 * 4 TMA-streamed small-batch kernel, 5 tcgen05 batched kernel, 6 cross-kernel pipelined decode GEMV (fp16-subnormal
 * FHFMA / HMMA flavours), 7 persistent integer-tensor-pipe decode GEMV (4-bit);
 * mma_for_m1: in auto mode route M == 1 to the mma kernel (1) or to the CUDA-core GEMV (0) */
int b200bit_set_path(int path, int mma_for_m1) {
    B200_REQUIRE(path >= 0 && path <= 7, B200BIT_ERR_ARG, "path must be in [0,7]");
    g_path = path;
    g_mma_for_m1 = mma_for_m1 ? 1 : 0;
    return B200BIT_OK;
}

/* diagnostics: device buffer of [grid][16][8] u64 receiving globaltimer stamps of the stream kernel (NULL = off) */
int b200bit_set_trace_buffer(void* buf) {
    g_trace = reinterpret_cast<unsigned long long*>(buf);
    return B200BIT_OK;
}

size_t b200bit_mpq_forward_workspace_bytes(int M, int K, int N, int w_bit) {
    (void)w_bit;
    const int mm = M < 32 ? M : 32;
    // tickets | the larger of: split-K / shared-strip partials (<= 64 contributions of [mm, N] f32),
    //                          tcgen05 path: B image (<= 64 KB per 1024 K elements at 32 batch slots) + per-group sums of x
    const size_t partials = size_t(64) * mm * N * sizeof(float);
    const size_t image = (size_t(K) / 1024 + 1) * 65536 + (size_t(K) / 64 + 1) * 32 * 4 * sizeof(float) + 4096;
    return size_t(B200BIT_WS_TICKET_BYTES) + (partials > image ? partials : image);
}

int b200bit_mpq_forward(const void* x, const int32_t* qweight, const void* scales, const void* zeros,
                        const int32_t* g_idx, void* y, int M, int K, int N, int G, int w_bit, int asym, int dtype,
                        void* workspace, size_t workspace_bytes, unsigned flags, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    B200_REQUIRE(x && qweight && scales && zeros && y, B200BIT_ERR_ARG, "mpq_forward: null pointer argument");
    B200_REQUIRE(dtype == B200BIT_F32 || dtype == B200BIT_F16 || dtype == B200BIT_BF16, B200BIT_ERR_ARG,
                 "mpq_forward: bad dtype code %d", dtype);
    B200_REQUIRE(w_bit == 1 || w_bit == 2 || w_bit == 4 || w_bit == 8, B200BIT_ERR_UNSUPPORTED,
                 "mpq_forward: w_bit=%d not supported (1, 2, 4, 8)", w_bit);
    B200_REQUIRE(M >= 0 && K > 0 && N > 0 && G > 0, B200BIT_ERR_SHAPE, "mpq_forward: bad sizes M=%d K=%d N=%d G=%d", M,
                 K, N, G);
    const int nb = 32 / w_bit;
    B200_REQUIRE(K % nb == 0, B200BIT_ERR_SHAPE, "mpq_forward: K=%d must be a multiple of %d", K, nb);
    B200_REQUIRE(g_idx || K % G == 0, B200BIT_ERR_SHAPE, "mpq_forward: K=%d not divisible by G=%d", K, G);
    B200_REQUIRE(!asym || N % nb == 0, B200BIT_ERR_SHAPE, "mpq_forward: asym needs N %% %d == 0 (N=%d)", nb, N);
    if (M == 0) return B200BIT_OK;

    const bool trivial = (g_idx == nullptr);
    // ---- decode (M == 1), 4-bit: persistent integer-tensor-pipe kernel (mpq_imma.cuh) ----
    const ImmaPlan ip = (g_path == 7 || g_path == 0) ? plan_imma(M, K, N, G, w_bit, asym, dtype, trivial) : ImmaPlan{};
    if (ip.ok && !(g_path == 0 && g_mma_for_m1)) {
        CUtensorMap tw, ts, tz;
        int rc = make_map_2d(&tw, CU_TENSOR_MAP_DATA_TYPE_UINT32, qweight, uint64_t(N), uint64_t(K / 8), uint64_t(N) * 4,
                             IM_COLS, IM_TILE_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
        rc = make_map_2d(&ts, CU_TENSOR_MAP_DATA_TYPE_UINT16, scales, uint64_t(N), uint64_t(G), uint64_t(N) * 2, 32,
                         uint32_t(ip.gps), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
        if (asym)       // packed zero rows: 8 words (64 nibbles) from word ((28 * strip) >> 3) & ~3 cover the strip
            rc = make_map_2d(&tz, CU_TENSOR_MAP_DATA_TYPE_UINT32, zeros, uint64_t(N / 8), uint64_t(G), uint64_t(N / 8) * 4,
                             8, uint32_t(ip.gps), CU_TENSOR_MAP_SWIZZLE_NONE);
        else
            rc = make_map_2d(&tz, CU_TENSOR_MAP_DATA_TYPE_UINT16, zeros, uint64_t(N), uint64_t(G), uint64_t(N) * 2, 32,
                             uint32_t(ip.gps), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
        ImmaParams p{};
        p.x = reinterpret_cast<const uint16_t*>(x);
        p.y = reinterpret_cast<uint16_t*>(y);
        p.K = K; p.N = N; p.R = K / 8; p.strips = ip.strips; p.n28 = ip.n28; p.tiles = ip.tiles; p.S = ip.S;
        p.strips_q = ip.strips / ip.grid; p.strips_r = ip.strips % ip.grid;
        p.rpg_shift = ip.rpg_shift; p.sz_bytes = ip.sz_bytes;
        p.s_tile_bytes = ip.s_tile_bytes; p.z_tile_bytes = ip.z_tile_bytes; p.trace = g_trace;
        ImmaLaunch l{};
        l.F = ip.F; l.grid = ip.grid; l.asym = asym != 0; l.bf16 = dtype == B200BIT_BF16; l.smem = ip.smem;
        l.flags = flags; l.stream = stream;
        p.early = chain_admit(stream, x, size_t(K) * 2, y, size_t(N) * 2, flags,
                              ((ip.strips + ip.grid - 1) / ip.grid) * ip.tiles <= ip.S) ? 1 : 0;
        rc = launch_imma(tw, ts, tz, p, l);
        if (rc != B200BIT_OK) chain_invalidate(stream);
        return rc;
    }
    // ---- decode (M == 1): cross-kernel pipelined CUDA-core GEMV ----
    const PipePlan pp = (g_path == 6 || g_path == 0) ? plan_pipe(M, K, N, G, w_bit, asym, dtype, trivial) : PipePlan{};
    if (pp.ok && !(g_path == 0 && g_mma_for_m1)) {
        float* part = nullptr;
        unsigned* tick = nullptr;
        if (pp.splitk > 1) {
            const size_t need = size_t(B200BIT_WS_TICKET_BYTES) + size_t(pp.splitk) * N * sizeof(float);
            B200_REQUIRE(workspace && workspace_bytes >= need, B200BIT_ERR_WORKSPACE,
                         "mpq_forward: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
            tick = reinterpret_cast<unsigned*>(workspace);
            part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + B200BIT_WS_TICKET_BYTES);
        }
        const bool bf16 = dtype == B200BIT_BF16;
        CUtensorMap tw, ts, tz;
        int rc = make_map_2d(&tw, CU_TENSOR_MAP_DATA_TYPE_UINT32, qweight, uint64_t(N), uint64_t(K / nb), uint64_t(N) * 4,
                             32, PG_STAGE_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
        rc = make_map_2d(&ts, CU_TENSOR_MAP_DATA_TYPE_UINT16, scales, uint64_t(N), uint64_t(G), uint64_t(N) * 2, 32,
                         uint32_t(pp.gps), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
        if (asym)
            rc = make_map_2d(&tz, CU_TENSOR_MAP_DATA_TYPE_UINT32, zeros, uint64_t(N / nb), uint64_t(G),
                             uint64_t(N / nb) * 4, uint32_t(32 / nb), uint32_t(pp.gps), CU_TENSOR_MAP_SWIZZLE_NONE);
        else
            rc = make_map_2d(&tz, CU_TENSOR_MAP_DATA_TYPE_UINT16, zeros, uint64_t(N), uint64_t(G), uint64_t(N) * 2, 32,
                             uint32_t(pp.gps), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
        PipeParams p{};
        p.x = reinterpret_cast<const uint16_t*>(x);
        p.y = reinterpret_cast<uint16_t*>(y);
        p.ws_part = part; p.tickets = tick;
        p.K = K; p.N = N; p.R = K / nb;
        p.stages_total = pp.stages_total; p.stages_per_split = pp.stages_per_split; p.S = pp.S;
        p.rpg = pp.rpg; p.rpg_shift = pp.rpg_shift; p.sz_bytes = pp.sz_bytes;
        p.s_tile_bytes = pp.s_tile_bytes; p.z_tile_bytes = pp.z_tile_bytes; p.asym = asym; p.trace = g_trace;
        PipeLaunch l{};
        l.FS = pp.FS; l.splitk = pp.splitk; l.strips = pp.strips; l.smem = pp.smem; l.flags = flags; l.stream = stream;
        // 4-bit fp16 with >= 8 packed rows per group and the CTA's K range resident in the ring: 28-column strips,
        // consumer math on the legacy tensor pipe, activations staged per warp (mpq_pipe_mma.cuh).  g_tune_L == 32
        // keeps the CUDA-core FHFMA flavour for A/B measurements.
        if (w_bit == 4 && !bf16 && pp.rpg % 8 == 0 && pp.stages_per_split <= pp.S && g_tune_L != 32) {
            const int strips = (N + PGM_COLS - 1) / PGM_COLS;
            B200_REQUIRE(size_t(strips) * sizeof(unsigned) <= B200BIT_WS_ZERO_OFFSET, B200BIT_ERR_SHAPE,
                         "mpq_forward: N=%d too large for the ticket area", N);
            rc = make_map_2d(&tw, CU_TENSOR_MAP_DATA_TYPE_UINT32, qweight, uint64_t(N), uint64_t(K / nb), uint64_t(N) * 4,
                             PGM_COLS, PG_STAGE_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc != B200BIT_OK) return rc;
            if (asym) {     // packed zero rows: 8 words (64 nibbles) from word (28 * strip) >> 3 cover the strip
                rc = make_map_2d(&tz, CU_TENSOR_MAP_DATA_TYPE_UINT32, zeros, uint64_t(N / nb), uint64_t(G),
                                 uint64_t(N / nb) * 4, 8, uint32_t(pp.gps), CU_TENSOR_MAP_SWIZZLE_NONE);
                if (rc != B200BIT_OK) return rc;
                p.z_tile_bytes = pp.gps * 32;
            }
            l.FS = pp.rpg == 8 ? 1 : 2;
            l.strips = strips;
            l.smem = size_t(PGM_X_BYTES) + size_t(pp.S) * (PGM_TILE_BYTES + 2 * pp.sz_bytes) + PG_MAX_STAGES * 8 +
                     PG_WARPS * 32 * 4 + 16;
            p.early = chain_admit(stream, x, size_t(K) * 2, y, size_t(N) * 2, flags) ? 1 : 0;
            rc = launch_pipe_mma(tw, ts, tz, p, l);
        } else {
            p.early = chain_admit(stream, x, size_t(K) * 2, y, size_t(N) * 2, flags) ? 1 : 0;
            rc = launch_pipe(tw, ts, tz, p, l, w_bit, bf16);
        }
        if (rc != B200BIT_OK) chain_invalidate(stream);
        return rc;
    }
    chain_invalidate(stream);      // every other kernel triggers its dependents before its own dependency wait
    // ---- more than 16 rows: tcgen05 batched kernel (mpq_tc.cu), one pass over the packed matrix ----
    if (M > (g_path == 5 ? 0 : (w_bit == 2 ? 8 : 16)) && (g_path == 0 || g_path == 5) && trivial && (w_bit == 4 || w_bit == 2) && dtype == B200BIT_F16 && K % 64 == 0 && N % 8 == 0 &&
        K % G == 0 && (K / G) % 32 == 0 && ((K / G) & (K / G - 1)) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        b200bit_mpq_forward_tc_supported(M, K, N, G, w_bit, asym, dtype, workspace ? workspace_bytes : 0))
        return b200bit_mpq_forward_tc(x, qweight, scales, zeros, y, M, K, N, G, w_bit, asym, dtype, workspace, workspace_bytes, stream_);
    // ---- path selection: TMA-streamed tensor kernel (f16, M <= 32) > mma.sync kernel > CUDA-core GEMV > general ----
    // auto: M == 1 -> CUDA-core FHFMA GEMV (fastest measured at batch 1, profiles/r1_*); 2 <= M: TMA-streamed tensor kernel
    const StreamPlan sp = (g_path == 4 || (g_path == 0 && M >= 2)) ? plan_stream(M, K, N, G, w_bit, asym, dtype, trivial)
                                                                 : StreamPlan{};
    if (sp.ok) {
        B200_REQUIRE(workspace && workspace_bytes >= B200BIT_WS_TICKET_BYTES, B200BIT_ERR_WORKSPACE,
                     "mpq_forward: workspace (>= %d bytes, zero-initialised head) required", B200BIT_WS_TICKET_BYTES);
        CUtensorMap tw, ts, tz;
        int rc = make_map_2d(&tw, CU_TENSOR_MAP_DATA_TYPE_UINT32, qweight, uint64_t(N), uint64_t(K / nb), uint64_t(N) * 4,
                             32, ST_RUN_ROWS, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != B200BIT_OK) return rc;
        rc = make_map_2d(&ts, CU_TENSOR_MAP_DATA_TYPE_UINT16, scales, uint64_t(N), uint64_t(G), uint64_t(N) * 2, 32,
                         uint32_t(sp.ngr), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
        if (asym)
            rc = make_map_2d(&tz, CU_TENSOR_MAP_DATA_TYPE_UINT32, zeros, uint64_t(N / nb), uint64_t(G),
                             uint64_t(N / nb) * 4, uint32_t(32 / nb), uint32_t(sp.ngr), CU_TENSOR_MAP_SWIZZLE_NONE);
        else
            rc = make_map_2d(&tz, CU_TENSOR_MAP_DATA_TYPE_UINT16, zeros, uint64_t(N), uint64_t(G), uint64_t(N) * 2, 32,
                             uint32_t(sp.ngr), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != B200BIT_OK) return rc;
        for (int m0 = 0; m0 < M; m0 += 32) {
            const int mc = (M - m0) < 32 ? (M - m0) : 32;
            StreamParams p{};
            p.x = reinterpret_cast<const uint16_t*>(x) + size_t(m0) * K;
            p.y = reinterpret_cast<uint16_t*>(y) + size_t(m0) * N;
            p.zero_page = reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(workspace) + B200BIT_WS_ZERO_OFFSET);
            p.M = mc; p.K = K; p.N = N; p.strips = sp.strips; p.rps = sp.rps; p.ngr = sp.ngr; p.rpr = sp.rpr;
            p.rpr_shift = sp.rpr_shift; p.asym = asym; p.S = sp.S; p.maxseg = sp.maxseg; p.trace = g_trace;
            StreamLaunch l{};
            l.MT = (mc + 7) / 8; l.FJ = sp.FJ; l.warps = sp.warps; l.grid = sp.grid; l.xs = sp.xs;
            l.smem = sp.smem; l.flags = flags; l.stream = stream;
            rc = launch_stream(tw, ts, tz, p, l, w_bit);
            if (rc != B200BIT_OK) return rc;
        }
        return B200BIT_OK;
    }
    const MmaPlan mp = (g_path == 0 || g_path == 2) ? plan_mma(M, K, N, G, w_bit, asym, dtype, trivial) : MmaPlan{};
    if (mp.ok && !(g_path == 0 && M == 1 && !g_mma_for_m1)) {
        float* part = nullptr;
        unsigned* tick = nullptr;
        const int mm = M < 32 ? M : 32;
        if (mp.splitk > 1) {
            const size_t need = size_t(B200BIT_WS_TICKET_BYTES) + size_t(mp.splitk) * mm * N * sizeof(float);
            B200_REQUIRE(workspace && workspace_bytes >= need, B200BIT_ERR_WORKSPACE,
                         "mpq_forward: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
            B200_REQUIRE(size_t(N / 32) * sizeof(unsigned) <= B200BIT_WS_ZERO_OFFSET, B200BIT_ERR_SHAPE,
                         "mpq_forward: N=%d too large for the ticket area", N);
            tick = reinterpret_cast<unsigned*>(workspace);
            part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + B200BIT_WS_TICKET_BYTES);
        }
        for (int m0 = 0; m0 < M; m0 += 32) {
            const int mc = (M - m0) < 32 ? (M - m0) : 32;
            MmaParams p{};
            p.x = reinterpret_cast<const uint16_t*>(x) + size_t(m0) * K;
            p.qw = reinterpret_cast<const uint32_t*>(qweight);
            p.scales = reinterpret_cast<const uint16_t*>(scales);
            p.zeros = zeros;
            p.y = reinterpret_cast<uint16_t*>(y) + size_t(m0) * N;
            p.ws_part = part; p.tickets = tick;
            p.M = mc; p.K = K; p.N = N; p.R = K / nb; p.G = G;
            p.runs_total = mp.runs_total; p.runs_per_split = mp.runs_per_split;
            p.rpr = mp.rpr; p.rpr_shift = mp.rpr_shift; p.asym = asym;
            MmaLaunch l{};
            l.MT = (mc + 7) / 8; l.FJ = mp.FJ; l.warps = mp.warps; l.splitk = mp.splitk;
            l.smem = mma_smem_bytes(mc, nb, mp.runs_per_split, mp.FJ, mp.warps);
            l.flags = flags; l.stream = stream;
            const int rc = launch_mma(p, l, w_bit);
            if (rc != B200BIT_OK) return rc;
        }
        return B200BIT_OK;
    }

    const GemvPlan pl = (g_path == 3 || g_path == 2) ? GemvPlan{} :
        plan_gemv(M < GEMV_MAX_M ? M : GEMV_MAX_M, K, N, G, w_bit, asym, dtype, trivial);
    const int max_m = (w_bit == 2 || w_bit == 4) ? GEMV_MAX_M : 1;
    if (!pl.ok) return launch_general(x, qweight, scales, zeros, g_idx, y, M, K, N, G, w_bit, asym, dtype, flags, stream);

    float* ws_part = nullptr;
    unsigned* tickets = nullptr;
    if (pl.splitk > 1) {
        const int L = 1 << pl.L_log2;
        const size_t strips = N / (4 * L);
        const int mm = M < max_m ? M : max_m;
        const size_t need = size_t(B200BIT_WS_TICKET_BYTES) + size_t(pl.splitk) * mm * N * sizeof(float);
        B200_REQUIRE(workspace && workspace_bytes >= need, B200BIT_ERR_WORKSPACE,
                     "mpq_forward: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        B200_REQUIRE(strips * sizeof(unsigned) <= B200BIT_WS_ZERO_OFFSET, B200BIT_ERR_SHAPE,
                     "mpq_forward: N=%d too large for the ticket area", N);
        tickets = reinterpret_cast<unsigned*>(workspace);
        ws_part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + B200BIT_WS_TICKET_BYTES);
    }
    for (int m0 = 0; m0 < M; m0 += max_m) {
        const int mc = (M - m0) < max_m ? (M - m0) : max_m;
        GemvParams p{};
        p.x = reinterpret_cast<const uint16_t*>(x) + size_t(m0) * K;
        p.qw = reinterpret_cast<const uint32_t*>(qweight);
        p.scales = reinterpret_cast<const uint16_t*>(scales);
        p.zeros = zeros;
        p.y = reinterpret_cast<uint16_t*>(y) + size_t(m0) * N;
        p.ws_part = ws_part;
        p.tickets = tickets;
        p.K = K; p.N = N; p.R = K / nb; p.G = G;
        p.rpr = pl.rpr;
        p.rpr_shift = pl.rpr_shift;
        p.runs_total = pl.runs_total;
        p.runs_per_split = pl.runs_per_split;
        p.L_log2 = pl.L_log2;
        p.asym = asym;
        GemvLaunch l{};
        l.M = mc; l.FR = pl.FR; l.warps = pl.warps; l.splitk = pl.splitk;
        l.smem = gemv_smem_bytes(mc, nb, pl.runs_per_split, pl.FR, pl.warps, 1 << pl.L_log2);
        l.flags = flags; l.stream = stream;
        const int rc = launch_gemv(p, l, w_bit, dtype == B200BIT_BF16);
        if (rc != B200BIT_OK) return rc;
    }
    return B200BIT_OK;
}

}  // extern "C"
