// Instantiations and launchers of psis_tail_kernel (b2l_split.cuh).
#include "b2l_split_host.h"

namespace b2l {

template <int TL, int MODE, int W>
static cudaError_t setup1(size_t smem, int* occ) {
    cudaError_t e = cudaFuncSetAttribute(psis_tail_kernel<TL, MODE, W>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // same (maximum) shared-memory carve-out for every kernel of the path: no SM reconfiguration between launches
    e = cudaFuncSetAttribute(psis_tail_kernel<TL, MODE, W>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, psis_tail_kernel<TL, MODE, W>, W * 32, smem);
}

// (registers per lane of the sort, warps per CTA): 4 warps = every warp on its own, 8 CTAs / SM; 16 or 32 warps =
// one or two big CTAs per SM whose warps move through the phases of a row together (instruction-cache locality)
#define B2L_TAIL_CASES(X) X(4, 4) X(4, 8) X(4, 16) X(4, 32) X(8, 4) X(8, 8) X(8, 16) X(8, 32) X(16, 4) X(16, 8) X(16, 16) X(32, 4) X(32, 8)

cudaError_t split_tail_setup(int tl, int warps, int mode, size_t smem, int* occ) {
#define X(TL_, W_)                                                                                 \
    if (tl == TL_ && warps == W_)                                                                  \
        return (mode == MODE_PSISLW) ? setup1<TL_, MODE_PSISLW, W_>(smem, occ) : setup1<TL_, MODE_LOO, W_>(smem, occ);
    B2L_TAIL_CASES(X)
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t split_tail_launch(int tl, int warps, int mode, int grid, size_t smem, cudaStream_t st, const SplitParams& q) {
#define X(TL_, W_)                                                                                 \
    if (tl == TL_ && warps == W_) {                                                                \
        if (mode == MODE_PSISLW) psis_tail_kernel<TL_, MODE_PSISLW, W_><<<grid, W_ * 32, smem, st>>>(q); \
        else psis_tail_kernel<TL_, MODE_LOO, W_><<<grid, W_ * 32, smem, st>>>(q);                  \
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
