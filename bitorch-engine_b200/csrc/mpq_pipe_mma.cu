// mpq_pipe_mma.cu -- instantiation + launch of the mma.sync flavour of the pipelined decode kernel (4-bit, fp16).
#include "mpq_pipe_mma.cuh"

namespace b200bit {

template <int FS2, bool ASYM, bool TRACE>
static int launch_pipe_mma_one(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const PipeParams& p,
                               const PipeLaunch& l) {
    auto kern = mpq_pipe_mma_kernel<FS2, ASYM, TRACE>;
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
    cfg.blockDim = dim3(PGM_THREADS, 1, 1);
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

int launch_pipe_mma(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const PipeParams& p,
                    const PipeLaunch& l) {
    if (p.trace) {      // diagnostics build of the same kernel: in-kernel globaltimer stamps
        if (l.FS == 2) return p.asym ? launch_pipe_mma_one<2, true, true>(tw, ts, tz, p, l) : launch_pipe_mma_one<2, false, true>(tw, ts, tz, p, l);
        if (l.FS == 1) return p.asym ? launch_pipe_mma_one<1, true, true>(tw, ts, tz, p, l) : launch_pipe_mma_one<1, false, true>(tw, ts, tz, p, l);
    } else {
        if (l.FS == 2) return p.asym ? launch_pipe_mma_one<2, true, false>(tw, ts, tz, p, l) : launch_pipe_mma_one<2, false, false>(tw, ts, tz, p, l);
        if (l.FS == 1) return p.asym ? launch_pipe_mma_one<1, true, false>(tw, ts, tz, p, l) : launch_pipe_mma_one<1, false, false>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "pipe mma: flush interval %d", l.FS);
}

}  // namespace b200bit
