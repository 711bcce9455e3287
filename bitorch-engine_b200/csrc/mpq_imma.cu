// mpq_imma.cu -- instantiation + launch of the integer-tensor-pipe decode kernel (mpq_imma.cuh; 4-bit, f16 / bf16).
#include "mpq_imma.cuh"

namespace b200bit {

template <int F, bool ASYM, bool BF16, bool TRACE>
static int launch_imma_one(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const ImmaParams& p,
                           const ImmaLaunch& l) {
    auto kern = mpq_imma_kernel<F, ASYM, BF16, TRACE>;
    static bool configured_dev[64] = {false};     // function attributes are per device: set once per device
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, IM_SMEM_LIMIT));
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(l.grid, 1, 1);
    cfg.blockDim = dim3(IM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = l.smem;
    cfg.stream = l.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (l.flags & B200BIT_FLAG_PDL) ? 1 : 0;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tw, ts, tz, p));
    return B200BIT_OK;
}

template <int F, bool TRACE>
static int launch_imma_f(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const ImmaParams& p,
                         const ImmaLaunch& l) {
    if (l.asym) return l.bf16 ? launch_imma_one<F, true, true, TRACE>(tw, ts, tz, p, l) : launch_imma_one<F, true, false, TRACE>(tw, ts, tz, p, l);
    return l.bf16 ? launch_imma_one<F, false, true, TRACE>(tw, ts, tz, p, l) : launch_imma_one<F, false, false, TRACE>(tw, ts, tz, p, l);
}

int launch_imma(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const ImmaParams& p, const ImmaLaunch& l) {
    if (p.trace) {      // diagnostics build of the same kernel: in-kernel globaltimer stamps (4-bit g >= 128 only)
        if (l.F == 4) return launch_imma_f<4, true>(tw, ts, tz, p, l);
    }
    switch (l.F) {
        case 1: return launch_imma_f<1, false>(tw, ts, tz, p, l);
        case 2: return launch_imma_f<2, false>(tw, ts, tz, p, l);
        case 4: return launch_imma_f<4, false>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "imma gemv: flush interval %d", l.F);
}

}  // namespace b200bit
