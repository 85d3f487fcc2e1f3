// Host-side interface of the tile path (b2l_tile.cu): pl.loo's stream pass on the ArviZ (chain, draw, obs)
// layout, i.e. the observation-fastest (S, N) matrix read where it lies (pyloo/loo.py:189 makes the same
// matrix a strided view and pyloo/utils.py:171-176 loops over its columns).
#pragma once
#include <cuda_runtime.h>

#include "b2l_split.cuh"

namespace b2l {

constexpr int TILE_W = 16;      // default observations per tile: 128 B of every draw (8 is the other build: 64 B)
constexpr int TILE_MAXC = 8;    // largest (portable) cluster
constexpr int TILE_MAX_R = 512; // draws per CTA: one 32-bit candidate mask per thread (16 draw slots x 32)

// Long posteriors (S > 4096): the draw axis is cut into n_chunks equal chunks of at most 4096 draws and every
// (tile, chunk) pair is one unit of work of the same kernel -- own minimum, own threshold, sums about its own
// minimum, candidates appended to the column's one list as raw r = -ll.  tile_merge_kernel folds the chunk records
// of a column into its SplitHeader (b2l_tile.cu).
struct __align__(16) ChunkHeader {  // 64 B per (observation, chunk)
    double llmin;           // min of ll over the chunk
    double q0, q1, q2, q3;  // about llmin: sum exp(-(ll - llmin)), sum exp(ll - llmin), sum (ll - llmin), sum (ll - llmin)^2
    double um;              // upper bound of max (ll - llmin)
    double tl;              // the chunk's candidate threshold: its candidates are the draws with ll <= tl
    double pad_;
};
static_assert(sizeof(ChunkHeader) == 64, "ChunkHeader is 64 bytes");

struct TilePlan {
    int ok;
    int n_chunks;   // 1, or the equal chunks a long draw axis is cut into
    int chunk_len;  // draws per chunk (S when n_chunks = 1): the S the shape below is planned for
    int tw;        // observations per tile (16 or 8); threads per CTA = 16 * tw
    int csize;     // CTAs per cluster = segments of the draw axis
    int R;         // draws a CTA owns: ceil(S / csize)
    int nbox, box_rows;  // its TMA boxes: nbox * box_rows >= R rows, box_rows a multiple of 8
    int q_t, q_l;  // ranks (1..32) of a CTA's sorted bin minima that give the tight / loose candidate threshold
    int occ;       // resident CTAs per SM
    int max_clusters;
    size_t smem;
};

struct TileParams {
    int S, M, cap;
    int R, nbox, box_rows;
    int q_t, q_l;
    long long n_tiles;   // tiles (tw observations each) of this round
    long long col0;      // matrix column of the round's first observation
    long long n_obs;     // observations of this round (the last tile may be partial)
    SplitHeader* hdr;    // [n_tiles * tw]   round scratch, as the split path's
    double* cx;          // [n_tiles * tw][cap]
    unsigned short* cs;  // [n_tiles * tw][cap]
    unsigned* cnt;       // [n_tiles * tw][2] candidates emitted so far (tight, loose): zeroed before the launch
    int* fb_list;        // observations (row_base + index in the round) for the general kernel
    int* fb_count;
    unsigned long long* counters;  // optional [4]; [3] += observations handed over
    long long row_base;
    int n_chunks;        // > 1: S above is the chunk length, work units are (tile, chunk) pairs, n_tiles still counts tiles
    ChunkHeader* chdr;   // [n_tiles * tw][n_chunks] (chunked only; hdr is then written by tile_merge_kernel)
    int debug;  // measurement aids (B2L_TILE_DEBUG): 2 loads only, 4 no candidate pass, 8 no exp pass (2, 4, 8: no valid results)
};

// shape part (pure arithmetic) and device part (occupancy) of the plan
bool tile_shape(long long S, int M, int csize, int tw, TilePlan* tp);
bool tile_pick(long long S, int M, TilePlan* tp);  // tile_shape with the library's choice of tile width and cluster size
cudaError_t tile_plan(long long S, int M, TilePlan* tp);
// 2-D tensor map of the (S, N) matrix: inner dimension = observations, box = TILE_W x box_rows
cudaError_t tile_tensor_map(const double* ll, long long S, long long N, long long stride_s, int tw, int box_rows,
                            void* tmap_out /* CUtensorMap, 128 B */);
cudaError_t tile_launch(const TilePlan& tp, const void* tmap, const TileParams& p, cudaStream_t st);
// chunked rounds: fold the chunk records into the round's SplitHeaders (S_total = all draws), flag what must be handed over
cudaError_t tile_merge_launch(const TileParams& p, long long S_total, cudaStream_t st);
cudaError_t tile_reasons(unsigned long long* out, int reset);

}  // namespace b2l
