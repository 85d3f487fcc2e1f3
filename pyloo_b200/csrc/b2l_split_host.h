// Host-side interface between the C-ABI translation unit and the split-path kernel translation
// units (compiled in parallel: the stream kernel alone has 16 instantiations).
#pragma once
#include <cuda_runtime.h>

#include "b2l_split.cuh"

namespace b2l {
cudaError_t split_stream_setup(int nt, int ept, int mode, size_t smem, int* occ);
cudaError_t split_stream_launch(int nt, int ept, int mode, int grid, size_t smem, cudaStream_t st,
                                const SplitParams& q);
cudaError_t split_tail_setup(int tl, int warps, int mode, size_t smem, int* occ);
cudaError_t split_tail_launch(int tl, int warps, int mode, int grid, size_t smem, cudaStream_t st, const SplitParams& q);
cudaError_t split_apply_launch(int grid, cudaStream_t st, const SplitParams& q);
cudaError_t split_stream_reasons(unsigned long long* out, int reset);
cudaError_t split_tail_reasons(unsigned long long* out, int reset);
}  // namespace b2l
