// mpq_mma_family.inl -- instantiates the (MT, FJ) grid of mpq_mma_kernel for one BITS value.
#include "mpq_mma.cuh"

namespace b200bit {

template <int BITS, int MT, int FJ>
static int launch_mma_one(const MmaParams& p, const MmaLaunch& l) {
    auto kern = mpq_mma_kernel<BITS, MT, FJ>;
    if (l.smem > 48 * 1024) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.N / 32, l.splitk, 1);
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

template <int BITS, int MT>
static int launch_mma_fj(const MmaParams& p, const MmaLaunch& l) {
    switch (l.FJ) {
        case 1: return launch_mma_one<BITS, MT, 1>(p, l);
        case 2: return launch_mma_one<BITS, MT, 2>(p, l);
        case 4: return launch_mma_one<BITS, MT, 4>(p, l);
        case 8: return launch_mma_one<BITS, MT, 8>(p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "mma: flush interval %d", l.FJ);
}

template <int BITS>
int launch_mma_family(const MmaParams& p, const MmaLaunch& l) {
    switch (l.MT) {
        case 1: return launch_mma_fj<BITS, 1>(p, l);
        case 2: return launch_mma_fj<BITS, 2>(p, l);
        case 3: return launch_mma_fj<BITS, 3>(p, l);
        case 4: return launch_mma_fj<BITS, 4>(p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "mma: MT=%d", l.MT);
}

template int launch_mma_family<B200_MMA_BITS>(const MmaParams&, const MmaLaunch&);

}  // namespace b200bit
