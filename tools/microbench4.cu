// microbench4.cu -- issue rate / latency of the legacy-pipe integer and fp8 MMAs on sm_100a next to HMMA.16816:
//   IMMA.16832.U8.S8 (mma.sync.m16n8k32.s32.u8.s8.s32)  and  QMMA.16832 e4m3 (mma.sync.m16n8k32.f32.e4m3.e4m3.f32),
// alone and interleaved with the LOP3 work a 4-bit unpack needs (4 LOP3 per IMMA: w & 0x0f0f0f0f, w & 0xf0f0f0f0).
// Question it answers: is the int8 route (2 LOP3 per packed word, 512 weights per MMA) faster per weight than the
// fp16-subnormal route (4 LOP3 + 1 SHF per word, 256 weights per HMMA)?   (diagnostic, not part of the library)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int CHAINS, int KIND, int LOPS>   // KIND 0 = IMMA u8.s8, 1 = QMMA e4m3, 2 = HMMA f16
__global__ void mma_kernel(unsigned a_bits, unsigned b_bits, int iters, int* out, long long* cyc) {
    int d[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int q = 0; q < 4; ++q) d[c][q] = 0;
    unsigned w0 = a_bits + threadIdx.x, w1 = a_bits * 3 + threadIdx.x, b0 = b_bits, b1 = b_bits ^ 0x01010101u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            unsigned a0, a1, a2, a3;
            if (LOPS) {       // the unpack: two packed words -> four byte-spread registers (values change every iteration)
                w0 += 0x01010101u * (c + 1); w1 ^= w0;
                a0 = w0 & 0x0f0f0f0fu; a2 = w0 & 0xf0f0f0f0u; a1 = w1 & 0x0f0f0f0fu; a3 = w1 & 0xf0f0f0f0u;
            } else { a0 = w0; a1 = w1; a2 = w0; a3 = w1; }
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(d[c][0]), "+r"(d[c][1]), "+r"(d[c][2]), "+r"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(d[c][0]), "+r"(d[c][1]), "+r"(d[c][2]), "+r"(d[c][3]) : "r"(a0 & 0x0f0f0f0fu), "r"(a1 & 0x0f0f0f0fu), "r"(a2 & 0x0f0f0f0fu), "r"(a3 & 0x0f0f0f0fu), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(d[c][0]), "+r"(d[c][1]), "+r"(d[c][2]), "+r"(d[c][3]) : "r"(a0 & 0x000f000fu), "r"(a1 & 0x000f000fu), "r"(a2 & 0x000f000fu), "r"(a3 & 0x000f000fu), "r"(b0), "r"(b1));
        }
    }
    long long t1 = clock64();
    int s = 0; for (int c = 0; c < CHAINS; ++c) for (int q = 0; q < 4; ++q) s += d[c][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS, int KIND, int LOPS>
static void run(const char* name, int warps, int* out, long long* cyc) {
    const int iters = 2000;
    long long h;
    mma_kernel<CHAINS, KIND, LOPS><<<148, warps * 32>>>(0x12345678u, 0x01020304u, iters, out, cyc); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const int wpm = KIND == 2 ? 256 : 512;   // useful weights per MMA at M == 1 (16 columns x k)
    printf("{\"op\":\"%s\",\"unpack_lop3\":%d,\"warps_per_sm\":%d,\"chains\":%d,\"cycles_per_mma_per_warp\":%.2f,\"mma_per_clk_per_sm\":%.3f,\"weights_per_clk_per_sm\":%.1f}\n",
           name, LOPS, warps, CHAINS, (double)h / (iters * CHAINS), warps * iters * double(CHAINS) / h, warps * iters * double(CHAINS) / h * wpm);
}

int main() {
    int* out; long long* cyc; CK(cudaMalloc(&out, 1 << 22)); CK(cudaMalloc(&cyc, 8));
    for (int warps : {1, 4, 8, 16, 24}) {
        run<4, 0, 0>("imma.16832.u8.s8", warps, out, cyc);
        run<1, 0, 0>("imma.16832.u8.s8", warps, out, cyc);
        run<4, 0, 1>("imma.16832.u8.s8", warps, out, cyc);
        run<4, 1, 0>("qmma.16832.e4m3", warps, out, cyc);
        run<1, 1, 0>("qmma.16832.e4m3", warps, out, cyc);
        run<4, 2, 0>("hmma.16816.f16", warps, out, cyc);
        run<4, 2, 1>("hmma.16816.f16", warps, out, cyc);
    }
    return 0;
}
