// microbench2.cu -- timeline of PDL-chained streaming kernels: when does each grid start, pass griddepcontrol.wait,
// receive its prefetched data, and finish?  (diagnostic)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

template <int PER>
__global__ void __launch_bounds__(512) read_kernel(const uint4* __restrict__ src, size_t nvec, float* out, const float* dep,
                                                   unsigned long long* tl, int node, int trigger_late) {
    unsigned long long t0 = gtime();
    if (!trigger_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 v[PER];
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const size_t j = tid + k * stride;
        if (j < nvec) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(src + j));
        else v[k] = make_uint4(0, 0, 0, 0);
    }
    unsigned long long t1 = gtime();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned long long t2 = gtime();
    float d = dep ? dep[threadIdx.x & 31] : 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    unsigned long long t3 = gtime();   // data arrived (acc depends on v)
    if (acc == 0x12345678u || d == 123.f) out[tid & 1023] = (float)acc + (float)t3;
    if (trigger_late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
        unsigned long long* p = tl + ((size_t)node * 2 + (blockIdx.x == 0 ? 0 : 1)) * 4;
        p[0] = t0; p[1] = t1; p[2] = t2; p[3] = (acc == 1 ? t3 + 1 : t3);
    }
}

template <typename... Args>
static void launch(void (*k)(Args...), dim3 g, dim3 b, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = g; cfg.blockDim = b; cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k, args...));
}

int main() {
    cudaStream_t s; CK(cudaStreamCreate(&s));
    float* out; CK(cudaMalloc(&out, 1 << 20)); CK(cudaMemset(out, 0, 1 << 20));
    const int NODES = 64;
    unsigned long long* tl; CK(cudaMalloc(&tl, NODES * 8 * sizeof(unsigned long long)));
    const size_t pool_bytes = 1200ull << 20;
    char* pool; CK(cudaMalloc(&pool, pool_bytes)); CK(cudaMemset(pool, 1, pool_bytes));
    const size_t bytes = 8912896ull, nvec = bytes / 16, slots = pool_bytes / bytes;
    struct Cfg { int pdl, threads, ctas, late; };
    for (Cfg c : {Cfg{0, 512, 296, 0}, Cfg{1, 512, 296, 0}, Cfg{1, 512, 296, 1}, Cfg{1, 256, 592, 0}, Cfg{1, 128, 1184, 0}, Cfg{1, 512, 148, 0}}) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < NODES; ++i) {
            const uint4* src = reinterpret_cast<const uint4*>(pool + (i % slots) * bytes);
            const size_t per_thread = (nvec + (size_t)c.ctas * c.threads - 1) / ((size_t)c.ctas * c.threads);
            if (per_thread <= 4) launch(read_kernel<4>, dim3(c.ctas), dim3(c.threads), s, c.pdl, src, nvec, out, (const float*)out, tl, i, c.late);
            else launch(read_kernel<8>, dim3(c.ctas), dim3(c.threads), s, c.pdl, src, nvec, out, (const float*)out, tl, i, c.late);
        }
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, s));
        CK(cudaStreamSynchronize(s));
        std::vector<unsigned long long> h(NODES * 8);
        CK(cudaMemcpy(h.data(), tl, h.size() * 8, cudaMemcpyDeviceToHost));
        printf("cfg pdl=%d threads=%d ctas=%d late_trigger=%d  (ns relative to node 20 start; cta0: start loads_issued wait_done data_in | last cta: same)\n", c.pdl, c.threads, c.ctas, c.late);
        unsigned long long base = h[20 * 8];
        for (int i = 20; i < 28; ++i) {
            printf("  node %2d:", i);
            for (int q = 0; q < 8; ++q) printf(" %6lld%s", (long long)(h[i * 8 + q] - base), q == 3 ? " |" : "");
            printf("\n");
        }
        printf("  avg node period: %.0f ns\n", (double)(h[60 * 8] - h[20 * 8]) / 40.0);
        CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    }
    return 0;
}
