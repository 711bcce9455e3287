// microbench5.cu -- how fast can persistent CTAs stream a packed [R][N] int32 matrix through TMA rings, as a function
// of the box geometry?  (diagnostic, not part of the library)   Each CTA owns column strips of `bw` words and walks
// them in tiles of `br` rows through an S-slot ring; nothing is computed: a slot is released as soon as it lands.
// Matrices are cycled through a pool larger than L2, as in bench.py.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) stream_kernel(const __grid_constant__ CUtensorMap tm, int strips, int tiles, int S, int bw, int br,
                                                     int cyclic, unsigned* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned char* ring = smem + 1024;
    const int tile_bytes = bw * 4 * br;
    // cyclic: CTA b walks strips b, b + grid, ... (neighbouring CTAs read adjacent strips at the same time);
    // otherwise a contiguous range of strips per CTA
    const int s_lo = cyclic ? blockIdx.x : int((long long)blockIdx.x * strips / gridDim.x);
    const int s_hi = int((long long)(blockIdx.x + 1) * strips / gridDim.x);
    const int s_step = cyclic ? gridDim.x : 1;
    const int T = (cyclic ? (strips - s_lo + s_step - 1) / s_step : (s_hi - s_lo)) * tiles;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned acc = 0;
        int issued = 0;
        auto issue = [&](int t) {
            const int strip = s_lo + (t / tiles) * s_step, kt = t % tiles, slot = t % S;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(ring + size_t(slot) * tile_bytes)), "l"(&tm), "r"(strip * bw), "r"(kt * br), "r"(smem_u32(&full[slot])) : "memory");
        };
        for (; issued < S && issued < T; ++issued) issue(issued);
        for (int t = 0; t < T; ++t) {
            const int slot = t % S;
            const unsigned ph = (t / S) & 1;
            asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&full[slot])), "r"(ph) : "memory");
            acc += *reinterpret_cast<volatile unsigned*>(ring + size_t(slot) * tile_bytes);
            if (issued < T) { issue(issued); ++issued; }
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
    const int POOL = 10;
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    struct Shape { int R, N; } shapes[] = {{512, 4096}, {512, 11008}, {1376, 4096}};
    struct Cfg { int bw, br, S, cps, cyc; } cfgs[] = {{28, 256, 3, 1, 1}, {28, 256, 2, 1, 1}, {28, 256, 3, 2, 1}, {28, 256, 3, 2}, {28, 128, 6, 2}, {28, 128, 4, 3}, {28, 64, 12, 2}, {28, 256, 6, 1}, {28, 128, 12, 1},
                                                  {32, 128, 6, 2}, {64, 64, 6, 2}, {64, 128, 3, 2}, {128, 32, 6, 2}, {128, 64, 3, 2}, {256, 16, 6, 2}, {256, 32, 3, 2},
                                                  {256, 32, 6, 1}, {64, 128, 6, 1}};
    for (auto sh : shapes) {
        const size_t bytes = size_t(sh.R) * sh.N * 4;
        unsigned char* pool; CK(cudaMalloc(&pool, bytes * POOL)); CK(cudaMemset(pool, 1, bytes * POOL));
        for (auto c : cfgs) {
            if (c.bw > sh.N) continue;
            // the pool is ONE tall matrix [R * POOL][N]: a launch streams all of it (launch gaps do not matter)
            CUtensorMap tm;
            cuuint64_t dims[2] = {cuuint64_t(sh.N), cuuint64_t(sh.R) * POOL}; cuuint64_t strides[1] = {cuuint64_t(sh.N) * 4};
            cuuint32_t box[2] = {cuuint32_t(c.bw), cuuint32_t(c.br)}; cuuint32_t es[2] = {1, 1};
            if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, pool, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("{\"error\":\"encode\",\"bw\":%d}\n", c.bw); continue;
            }
            const int strips = (sh.N + c.bw - 1) / c.bw, tiles = (sh.R * POOL + c.br - 1) / c.br;
            const int grid = strips < 148 * c.cps ? strips : 148 * c.cps;
            const size_t smem = 1024 + size_t(c.S) * c.bw * 4 * c.br;
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            const int reps = 5;
            for (int w = 0; w < 2; ++w) {
                if (w == 1) CK(cudaEventRecord(e0));
                for (int r = 0; r < (w ? reps : 1); ++r) stream_kernel<<<grid, 128, smem>>>(tm, strips, tiles, c.S, c.bw, c.br, c.cyc, sink);
            }
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = ms * 1e3 / (reps * POOL);
            printf("{\"R\":%d,\"N\":%d,\"box_words\":%d,\"box_rows\":%d,\"S\":%d,\"ctas_per_sm\":%d,\"cyclic\":%d,\"grid\":%d,\"smem_kb\":%.0f,\"us_per_matrix\":%.2f,\"GBs\":%.0f}\n",
                   sh.R, sh.N, c.bw, c.br, c.S, c.cps, c.cyc, grid, smem / 1024.0, us, bytes / us / 1e3);
        }
        CK(cudaFree(pool));
    }
    return 0;
}
