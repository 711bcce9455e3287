// umma_rate.cu -- tcgen05 timing facts used by DESIGN.md: issue rate / latency of tcgen05.mma kind::f16 (M=128, K=16) as a
// function of N and of where A lives (TMEM / shared memory), with normal and subnormal fp16 A; tcgen05.st / ld latency.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra W;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// out[0] = cycles for `reps` MMAs + commit + wait;  out[1] = tcgen05.st.x16 + wait::st;  out[2] = tcgen05.ld.x4 + wait::ld
__global__ void __launch_bounds__(128) rate(long long* out, int N, int reps, int a_in_smem, uint32_t a_bits, int ndist) {
    extern __shared__ __align__(128) unsigned char sm[];       // B: N x 16 halves canonical (N/8 groups x 256 B) | A smem: 128 x 16
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int b_bytes = (N / 8) * 256, a_bytes = 16 * 256;
    for (int i = tid; i < (b_bytes + a_bytes) / 4; i += 128)
        reinterpret_cast<uint32_t*>(sm)[i] = (i * 4 < b_bytes) ? 0x3C003C00u : a_bits;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t tm_d = tmem, tm_a = tmem + 256;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    uint32_t r[16];
    for (int c = 0; c < 16; ++c) r[c] = a_bits;
    long long t0 = clock64();
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(tm_a + lane_addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    if (tid == 0) out[1] = t1 - t0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {      // converged warp; only the tcgen05 instructions are predicated on the elected lane
        uint32_t leader;
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t bdesc = hi | (uint64_t)((smem_u32(sm) & 0x3FFFF) >> 4);
        const uint64_t adesc = hi | (uint64_t)(((smem_u32(sm) + b_bytes) & 0x3FFFF) >> 4);
        const uint32_t idesc = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        t0 = clock64();
        // 8 MMAs per asm block (no instructions in between); accumulator alternates between ndist buffers
        const uint32_t d0 = tm_d, d1 = tm_d + (ndist > 1 ? (uint32_t)N : 0u);
        for (int rep = 0; rep < reps; rep += 8) {
            if (a_in_smem)
                asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %6, 0;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t}"
                             ::"r"(d0), "r"(d1), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1), "r"(leader) : "memory");
            else
                asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %6, 0;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %4, p;\n\t"
                             "@q tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], %3, %4, p;\n\t}"
                             ::"r"(d0), "r"(d1), "r"(tm_a), "l"(bdesc), "r"(idesc), "r"(1), "r"(leader) : "memory");
        }
        long long ti = clock64();
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)), "r"(leader) : "memory");
        wait_bar(&bar, 0);
        t1 = clock64();
        if (tid == 0) { out[0] = t1 - t0; out[3] = ti - t0; }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t0 = clock64();
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(tm_d + lane_addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    t1 = clock64();
    if (tid == 0) { out[2] = t1 - t0; out[4] = a + b + c + d; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* d; CK(cudaMalloc(&d, 64));
    CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int Ns[] = {16, 32, 64, 128, 256};
    for (int a_in_smem = 0; a_in_smem < 2; ++a_in_smem)
        for (int sub = 1; sub < 2; ++sub)
            for (int ni = 0; ni < 5; ++ni)
                for (int reps : {8, 64, 512})
                    for (int ndist : {1, 2}) {
                        const int N = Ns[ni];
                        if (ndist * N > 256) continue;
                        for (int warm = 0; warm < 2; ++warm) rate<<<1, 128, 48 * 1024>>>(d, N, reps, a_in_smem, sub ? 0x00050003u : 0x3C004000u, ndist);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("{\"error\":\"%s\"}\n", cudaGetErrorString(e)); return 0; }
                        long long h[8]; CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
                        printf("{\"a\":\"%s\",\"subnormal\":%d,\"N\":%d,\"reps\":%d,\"ndist\":%d,\"cycles_total\":%lld,\"cycles_issue\":%lld,\"per_mma\":%.1f,\"st16_cycles\":%lld,\"ld4_cycles\":%lld}\n",
                               a_in_smem ? "smem" : "tmem", sub, N, reps, ndist, h[0], h[3], double(h[0]) / reps, h[1], h[2]);
                    }
    return 0;
}
