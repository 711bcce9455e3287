// mpq_gemv_family.inl -- instantiates the (M, FR) grid of mpq_gemv_kernel for one BITS value; included by
// mpq_gemv_b{1,2,4,8}.cu with B200_GEMV_BITS defined (one translation unit per bit-width so they build in parallel).
#include "mpq_gemv.cuh"

namespace b200bit {

template <int BITS, bool BF16, int M, int FR>
static int launch_one(const GemvParams& p, const GemvLaunch& l) {
    auto kern = mpq_gemv_kernel<BITS, BF16, M, FR>;
    if (l.smem > 48 * 1024) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
    }
    const int L = 1 << p.L_log2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.N / (4 * L), l.splitk, 1);
    cfg.blockDim = dim3(l.warps * 32, 1, 1);
    cfg.dynamicSmemBytes = l.smem;
    cfg.stream = l.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (l.flags & B200BIT_FLAG_PDL) ? 1 : 0;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    return B200BIT_OK;
}

template <int BITS, bool BF16, int M>
static int launch_fr(const GemvParams& p, const GemvLaunch& l) {
    switch (l.FR) {
        case 1: return launch_one<BITS, BF16, M, 1>(p, l);
        case 2: return launch_one<BITS, BF16, M, 2>(p, l);
        case 4: return launch_one<BITS, BF16, M, 4>(p, l);
        case 8: return launch_one<BITS, BF16, M, 8>(p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "gemv: flush interval %d", l.FR);
}

template <int BITS, bool BF16>
int launch_gemv_family(const GemvParams& p, const GemvLaunch& l) {
    constexpr int MAXM = (BITS == 2 || BITS == 4) ? GEMV_MAX_M : 1;   // wide-M variants only where they matter
    switch (l.M) {
        case 1: return launch_fr<BITS, BF16, 1>(p, l);
        case 2: if constexpr (MAXM >= 2) return launch_fr<BITS, BF16, 2>(p, l); break;
        case 3: if constexpr (MAXM >= 3) return launch_fr<BITS, BF16, 3>(p, l); break;
        case 4: if constexpr (MAXM >= 4) return launch_fr<BITS, BF16, 4>(p, l); break;
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "gemv: M=%d not instantiated for %d-bit", l.M, BITS);
}

template int launch_gemv_family<B200_GEMV_BITS, false>(const GemvParams&, const GemvLaunch&);
#if B200_GEMV_BITS <= 4
template int launch_gemv_family<B200_GEMV_BITS, true>(const GemvParams&, const GemvLaunch&);
#endif

}  // namespace b200bit
