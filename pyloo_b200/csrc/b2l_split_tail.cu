// Instantiations and launchers of psis_tail_kernel (b2l_split.cuh).
#include "b2l_split_host.h"

namespace b2l {

template <int TL, int MODE>
static cudaError_t setup1(size_t smem, int* occ) {
    cudaError_t e = cudaFuncSetAttribute(psis_tail_kernel<TL, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // same (maximum) shared-memory carve-out for every kernel of the path: no SM reconfiguration between launches
    e = cudaFuncSetAttribute(psis_tail_kernel<TL, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, psis_tail_kernel<TL, MODE>, TAIL_WARPS * 32, smem);
}

#define B2L_TAIL_CASES(X) X(4) X(8) X(16)

cudaError_t split_tail_setup(int tl, int mode, size_t smem, int* occ) {
#define X(TL_)                                                                                     \
    if (tl == TL_)                                                                                 \
        return (mode == MODE_PSISLW) ? setup1<TL_, MODE_PSISLW>(smem, occ) : setup1<TL_, MODE_LOO>(smem, occ);
    B2L_TAIL_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t split_tail_launch(int tl, int mode, int grid, size_t smem, cudaStream_t st, const SplitParams& q) {
#define X(TL_)                                                                                     \
    if (tl == TL_) {                                                                               \
        if (mode == MODE_PSISLW) psis_tail_kernel<TL_, MODE_PSISLW><<<grid, TAIL_WARPS * 32, smem, st>>>(q); \
        else psis_tail_kernel<TL_, MODE_LOO><<<grid, TAIL_WARPS * 32, smem, st>>>(q);              \
        return cudaGetLastError();                                                                 \
    }
    B2L_TAIL_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t split_apply_launch(int grid, cudaStream_t st, const SplitParams& q) {
    psis_apply_kernel<<<grid, APPLY_NT, 0, st>>>(q);
    return cudaGetLastError();
}

cudaError_t split_tail_reasons(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_handover, sizeof(unsigned long long) * HO_REASONS);
    if (e == cudaSuccess && reset) {
        unsigned long long z[HO_REASONS] = {0};
        e = cudaMemcpyToSymbol(g_handover, z, sizeof(z));
    }
    return e;
}

}  // namespace b2l
