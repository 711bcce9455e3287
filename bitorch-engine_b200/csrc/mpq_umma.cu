// mpq_umma.cu -- instantiations + launcher of the tcgen05 small-batch kernel (mpq_umma.cuh).
#include "mpq_umma.cuh"

namespace b200bit {

template <int FJ2>
static int launch_umma_one(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const UmmaParams& p,
                           const UmmaLaunch& l) {
    auto kern = mpq_umma_kernel<FJ2>;
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(l.grid, 1, 1);
    cfg.blockDim = dim3(192, 1, 1);
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

int launch_umma(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const UmmaParams& p, const UmmaLaunch& l) {
    switch (l.FJ2) {
        case 1: return launch_umma_one<1>(tw, ts, tz, p, l);
        case 2: return launch_umma_one<2>(tw, ts, tz, p, l);
        case 4: return launch_umma_one<4>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "umma: flush interval %d", l.FJ2);
}

}  // namespace b200bit
