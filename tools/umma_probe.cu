// umma_probe.cu -- pins down the tcgen05 operand conventions this repo's kernels rely on, on real hardware:
//   A operand in TMEM (written with tcgen05.st.32x32b), B operand in shared memory (K-major, no swizzle, described by a
//   64-bit matrix descriptor), D in TMEM (read with tcgen05.ld.32x32b), kind::f16, M=128, N=16, K=16, fp32 accumulate.
// Tries both assignments of the descriptor's leading/stride byte offsets and reports the max error of each.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// variant 0: LBO = K-direction core-matrix stride, SBO = N-direction 8-row-group stride
// variant 1: swapped
// subnormal = 1: A holds fp16 subnormals (integer codes * 2^-24), result scaled back
__global__ void __launch_bounds__(128) probe(const __half* A /*[128][16]*/, const __half* B /*[16][16] (n,k)*/, float* D /*[128][16]*/,
                                             int variant, int accumulate_twice) {
    __shared__ __align__(128) unsigned char bsm[1024];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B tile into canonical no-swizzle K-major layout: core matrix (8 rows x 16 bytes) = 128 contiguous bytes;
    // core(ng, kc) at (ng*2 + kc)*128 : K-adjacent cores 128 B apart, N-adjacent 8-row groups 256 B apart
    for (int i = tid; i < 16 * 16; i += 128) {
        const int n = i / 16, k = i % 16;
        const int ng = n / 8, nr = n % 8, kc = k / 8, ke = k % 8;
        reinterpret_cast<__half*>(bsm)[((ng * 2 + kc) * 128 + nr * 16 + ke * 2) / 2] = B[n * 16 + k];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t tm_a = tmem + 32;        // columns [32, 40): A (128 lanes x 8 columns = K 16)
    const uint32_t tm_d = tmem;             // columns [0, 16): D
    // A: thread <-> lane (row m); column c holds k = 2c (low half), 2c+1 (high half)
    {
        const int m = warp * 32 + lane;
        uint32_t r[8];
        for (int c = 0; c < 8; ++c) {
            const uint32_t lo = __half_as_ushort(A[m * 16 + 2 * c]), hi = __half_as_ushort(A[m * 16 + 2 * c + 1]);
            r[c] = lo | (hi << 16);
        }
        const uint32_t addr = tm_a + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // B was written with generic-proxy stores
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t lbo = variant == 0 ? 128 : 256, sbo = variant == 0 ? 256 : 128;
        uint64_t desc = 0;
        desc |= (uint64_t)((smem_u32(bsm) & 0x3FFFF) >> 4);
        desc |= (uint64_t)(lbo >> 4) << 16;
        desc |= (uint64_t)(sbo >> 4) << 32;
        desc |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
        const uint32_t idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);   // D=f32, A=B=f16, K-major, N=16, M=128
        for (int rep = 0; rep < 1 + accumulate_twice; ++rep) {
            const uint32_t en = rep;                      // first MMA overwrites D, second accumulates
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm_d), "r"(tm_a), "l"(desc), "r"(idesc), "r"(en) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMA
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE;\n\tbra W;\n\tDONE:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        uint32_t v[16];
        const uint32_t addr = tm_d + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int m = warp * 32 + lane;
        for (int n = 0; n < 16; ++n) D[m * 16 + n] = __uint_as_float(v[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    __half hA[128 * 16], hB[16 * 16];
    float ref[128 * 16];
    for (int sub = 0; sub < 2; ++sub) {
        for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) {
            const int q = (m * 3 + k * 5 + m / 16) % 16;
            hA[m * 16 + k] = sub ? __ushort_as_half((unsigned short)q) : __float2half((float)q);   // subnormal: bits = q -> q * 2^-24
        }
        for (int n = 0; n < 16; ++n) for (int k = 0; k < 16; ++k) hB[n * 16 + k] = __float2half((float)(((n * 7 + k * 3) % 9) - 4) * 0.5f);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
            float s = 0;
            for (int k = 0; k < 16; ++k) s += (float)((m * 3 + k * 5 + m / 16) % 16) * __half2float(hB[n * 16 + k]);
            ref[m * 16 + n] = s;
        }
        __half *dA, *dB; float* dD;
        CK(cudaMalloc(&dA, sizeof(hA))); CK(cudaMalloc(&dB, sizeof(hB))); CK(cudaMalloc(&dD, sizeof(ref)));
        CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
        for (int variant = 0; variant < 2; ++variant)
            for (int twice = 0; twice < 2; ++twice) {
                CK(cudaMemset(dD, 0, sizeof(ref)));
                probe<<<1, 128>>>(dA, dB, dD, variant, twice);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("{\"subnormal\":%d,\"variant\":%d,\"twice\":%d,\"error\":\"%s\"}\n", sub, variant, twice, cudaGetErrorString(e)); return 0; }
                float out[128 * 16];
                CK(cudaMemcpy(out, dD, sizeof(out), cudaMemcpyDeviceToHost));
                double maxerr = 0; const double scale = sub ? 16777216.0 : 1.0;
                for (int i = 0; i < 128 * 16; ++i) { double d = fabs(out[i] * scale - ref[i] * (1 + twice)); if (d > maxerr) maxerr = d; }
                printf("{\"subnormal\":%d,\"variant\":%d,\"accumulate_twice\":%d,\"max_abs_err\":%.6f,\"d00\":%.4f,\"ref00\":%.4f,\"d_5_3\":%.4f,\"ref_5_3\":%.4f}\n",
                       sub, variant, twice, maxerr, out[0] * scale, ref[0] * (1 + twice), out[5 * 16 + 3] * scale, ref[5 * 16 + 3] * (1 + twice));
            }
    }
    return 0;
}
