// mpq_umma.cu -- instantiations + launchers of the tcgen05 batched kernel and its B-image prepare kernel (mpq_umma.cuh).
#include "mpq_umma.cuh"

namespace b200bit {

template <int FJ2, int MB>
static int launch_umma_one(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const UmmaParams& p,
                           const UmmaLaunch& l) {
    auto kern = mpq_umma_kernel<FJ2, MB>;
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(l.grid, 1, 1);
    cfg.blockDim = dim3(UM_THREADS, 1, 1);
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

template <int FJ2>
static int launch_umma_mb(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const UmmaParams& p,
                          const UmmaLaunch& l) {
    switch (l.MB) {
        case 4: return launch_umma_one<FJ2, 4>(tw, ts, tz, p, l);
        case 8: return launch_umma_one<FJ2, 8>(tw, ts, tz, p, l);
        case 16: return launch_umma_one<FJ2, 16>(tw, ts, tz, p, l);
        case 32: return launch_umma_one<FJ2, 32>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "umma: batch slots %d", l.MB);
}

int launch_umma(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const UmmaParams& p, const UmmaLaunch& l) {
    switch (l.FJ2) {
        case 1: return launch_umma_mb<1>(tw, ts, tz, p, l);
        case 2: return launch_umma_mb<2>(tw, ts, tz, p, l);
        case 4: return launch_umma_mb<4>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "umma: flush interval %d", l.FJ2);
}

int launch_umma_prepare(const UmmaPrepParams& p_in, unsigned flags, cudaStream_t stream) {
    UmmaPrepParams p = p_in;
    const int cells = p.steps * (4 * 2 * 2 * (4 * p.MB) * 4);
    p.cell_blocks = (cells + 255) / 256;
    const int sum_blocks = (p.MB * 4 * p.steps + 7) / 8;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.cell_blocks + sum_blocks, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (flags & B200BIT_FLAG_PDL) ? 1 : 0;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, umma_prepare_kernel<0>, p));
    return B200BIT_OK;
}

}  // namespace b200bit
