// mpq_pipe_family.inl -- instantiates mpq_pipe_kernel for one BITS value; included by mpq_pipe_b{2,4,8}.cu with
// B200_PIPE_BITS defined (one translation unit per bit-width so they build in parallel).
#include "mpq_pipe.cuh"

namespace b200bit {

template <int BITS, bool BF16, int FS>
static int launch_pipe_one(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const PipeParams& p,
                           const PipeLaunch& l) {
    auto kern = mpq_pipe_kernel<BITS, BF16, FS>;
    static bool configured_dev[64] = {false};     // function attributes are per device: set once per device
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 76800));
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(l.strips, l.splitk, 1);
    cfg.blockDim = dim3(PG_THREADS, 1, 1);
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

template <int BITS, bool BF16>
int launch_pipe_family(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const PipeParams& p,
                       const PipeLaunch& l) {
    switch (l.FS) {
        case 1: return launch_pipe_one<BITS, BF16, 1>(tw, ts, tz, p, l);
        case 2: return launch_pipe_one<BITS, BF16, 2>(tw, ts, tz, p, l);
        case 4: return launch_pipe_one<BITS, BF16, 4>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "pipe gemv: flush interval %d", l.FS);
}

template int launch_pipe_family<B200_PIPE_BITS, false>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                                       const PipeParams&, const PipeLaunch&);
#if B200_PIPE_BITS <= 4
template int launch_pipe_family<B200_PIPE_BITS, true>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                                      const PipeParams&, const PipeLaunch&);
#endif

}  // namespace b200bit
