// mpq_stream_family.inl -- instantiates the (MT, FJ) grid of mpq_stream_kernel for one BITS value.
#include "mpq_mma.cuh"      // mma_kperm
#include "mpq_stream.cuh"

namespace b200bit {

template <int BITS, int MT, int FJ, bool XS>
static int launch_stream_one(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const StreamParams& p,
                             const StreamLaunch& l) {
    auto kern = mpq_stream_kernel<BITS, MT, FJ, XS>;
    if (l.smem > 48 * 1024) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(l.grid, 1, 1);
    cfg.blockDim = dim3((l.warps + 1) * 32, 1, 1);
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

template <int BITS, int MT>
static int launch_stream_fj(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const StreamParams& p,
                            const StreamLaunch& l) {
    switch (l.FJ) {
        case 2: return l.xs ? launch_stream_one<BITS, MT, 2, MT == 1>(tw, ts, tz, p, l) : launch_stream_one<BITS, MT, 2, false>(tw, ts, tz, p, l);
        case 4: return l.xs ? launch_stream_one<BITS, MT, 4, MT == 1>(tw, ts, tz, p, l) : launch_stream_one<BITS, MT, 4, false>(tw, ts, tz, p, l);
        case 8: return l.xs ? launch_stream_one<BITS, MT, 8, MT == 1>(tw, ts, tz, p, l) : launch_stream_one<BITS, MT, 8, false>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "stream: flush interval %d", l.FJ);
}

template <int BITS>
int launch_stream_family(const CUtensorMap& tw, const CUtensorMap& ts, const CUtensorMap& tz, const StreamParams& p,
                         const StreamLaunch& l) {
    switch (l.MT) {
        case 1: return launch_stream_fj<BITS, 1>(tw, ts, tz, p, l);
        case 2: return launch_stream_fj<BITS, 2>(tw, ts, tz, p, l);
        case 3: return launch_stream_fj<BITS, 3>(tw, ts, tz, p, l);
        case 4: return launch_stream_fj<BITS, 4>(tw, ts, tz, p, l);
    }
    return set_error(B200BIT_ERR_UNSUPPORTED, "stream: MT=%d", l.MT);
}

template int launch_stream_family<B200_STREAM_BITS>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                                    const StreamParams&, const StreamLaunch&);

}  // namespace b200bit
