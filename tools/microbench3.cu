// microbench3.cu -- does feeding fp16 SUBNORMAL operands slow down HMMA (mma.sync.m16n8k16) or FHFMA on sm_100a? (diagnostic)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int CHAINS>
__global__ void hmma_kernel(unsigned a_bits, unsigned b_bits, int iters, float* out, long long* cyc) {
    float d[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int q = 0; q < 4; ++q) d[c][q] = 0.f;
    unsigned a0 = a_bits + (threadIdx.x & 1), a1 = a_bits, a2 = a_bits, a3 = a_bits, b0 = b_bits, b1 = b_bits;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    float s = 0; for (int c = 0; c < CHAINS; ++c) for (int q = 0; q < 4; ++q) s += d[c][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
__global__ void fhfma_kernel(unsigned a_bits, unsigned b_bits, int iters, float* out, long long* cyc) {
    float d[CHAINS];
    for (int c = 0; c < CHAINS; ++c) d[c] = 0.f;
    unsigned short a = (unsigned short)(a_bits + (threadIdx.x & 1)), b = (unsigned short)b_bits;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(d[c]) : "h"(a), "h"(b));
    }
    long long t1 = clock64();
    float s = 0; for (int c = 0; c < CHAINS; ++c) s += d[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; CK(cudaMalloc(&out, 1 << 22)); CK(cudaMalloc(&cyc, 8));
    const int iters = 2000;
    struct V { const char* name; unsigned a; } vals[] = {{"normal(1.0)", 0x3C003C00u}, {"subnormal(q)", 0x00070005u}, {"subnormal(16q)", 0x00700050u}, {"zero", 0u}};
    for (auto v : vals)
        for (int warps : {1, 4, 8, 16}) {
            long long h;
            hmma_kernel<4><<<148, warps * 32>>>(v.a, 0x3C003800u, iters, out, cyc); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("{\"op\":\"hmma.16816.f32\",\"A\":\"%s\",\"warps_per_sm\":%d,\"chains\":4,\"cycles_per_hmma_per_warp\":%.2f,\"hmma_per_clk_per_sm\":%.3f}\n", v.name, warps, (double)h / (iters * 4), warps * iters * 4.0 / h);
            hmma_kernel<1><<<148, warps * 32>>>(v.a, 0x3C003800u, iters, out, cyc); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("{\"op\":\"hmma.16816.f32\",\"A\":\"%s\",\"warps_per_sm\":%d,\"chains\":1,\"latency_cycles\":%.2f}\n", v.name, warps, (double)h / iters);
            fhfma_kernel<8><<<148, warps * 32>>>(v.a, 0x3800u, iters, out, cyc); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("{\"op\":\"fhfma\",\"A\":\"%s\",\"warps_per_sm\":%d,\"chains\":8,\"cycles_per_op_per_warp\":%.2f,\"thread_ops_per_clk_per_sm\":%.1f}\n", v.name, warps, (double)h / (iters * 8), warps * 32.0 * iters * 8 / h);
        }
    return 0;
}
