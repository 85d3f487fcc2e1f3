// Instantiations and launchers of psis_stream_kernel (b2l_split.cuh).
#include "b2l_split_host.h"

namespace b2l {

#define B2L_STREAM_CASES(X)                                                                        \
    X(128, 8) X(128, 16) X(256, 8) X(256, 16) X(512, 8) X(512, 16) X(1024, 8) X(1024, 16)

template <int NT, int EPT, int MODE>
static cudaError_t setup1(size_t smem, int* occ) {
    cudaError_t e = cudaFuncSetAttribute(psis_stream_kernel<NT, EPT, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // same (maximum) shared-memory carve-out for every kernel of the path: no SM reconfiguration between launches
    e = cudaFuncSetAttribute(psis_stream_kernel<NT, EPT, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, psis_stream_kernel<NT, EPT, MODE>, stream_block(NT, MODE), smem);
}

cudaError_t split_stream_setup(int nt, int ept, int mode, size_t smem, int* occ) {
#define X(NT_, EPT_)                                                                               \
    if (nt == NT_ && ept == EPT_)                                                                  \
        return (mode == MODE_PSISLW) ? setup1<NT_, EPT_, MODE_PSISLW>(smem, occ)                   \
                                     : setup1<NT_, EPT_, MODE_LOO>(smem, occ);
    B2L_STREAM_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t split_stream_launch(int nt, int ept, int mode, int grid, size_t smem, cudaStream_t st,
                                const SplitParams& q) {
#define X(NT_, EPT_)                                                                               \
    if (nt == NT_ && ept == EPT_) {                                                                \
        if (mode == MODE_PSISLW) psis_stream_kernel<NT_, EPT_, MODE_PSISLW><<<grid, stream_block(NT_, MODE_PSISLW), smem, st>>>(q); \
        else psis_stream_kernel<NT_, EPT_, MODE_LOO><<<grid, NT_, smem, st>>>(q);                  \
        return cudaGetLastError();                                                                 \
    }
    B2L_STREAM_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t split_stream_reasons(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_handover, sizeof(unsigned long long) * HO_REASONS);
    if (e == cudaSuccess && reset) {
        unsigned long long z[HO_REASONS] = {0};
        e = cudaMemcpyToSymbol(g_handover, z, sizeof(z));
    }
    return e;
}

}  // namespace b2l
