// Host-side interface between the C-ABI translation unit and b2l_is.cu (SIS / TIS importance weights and
// the e_loo weighted expectations: the consumers either side of psislw).
#pragma once
#include <cuda_runtime.h>

namespace b2l {

constexpr int IS_METHOD_SIS = 1;  // pyloo/sis.py:86-106
constexpr int IS_METHOD_TIS = 2;  // pyloo/tis.py:91-120
constexpr int IS_MODE_WEIGHTS = 0;  // write the normalised log weights + ess (compute_importance_weights)
constexpr int IS_MODE_LOO = 1;      // input is the log-likelihood; write elpd_i, ess_i, lppd_i (pyloo/loo.py:286-337)

struct IsParams {
    const double* in;  // rows of S doubles, element stride 1
    long long in_stride;
    double* out;  // weights mode: rows of S doubles
    long long out_stride;
    double* ess;                  // N
    double* elpd;                 // loo mode: N
    double* lppd;                 // loo mode: N
    unsigned long long* counters; // loo mode, nullable: NaN / +inf / -inf inputs
    long long n_rows;
    int S;
    int bulk;      // rows may be staged with 1-D bulk TMA (S even, 16 B aligned base and stride)
    double log_S;  // np.log(n_samples), computed by the host like the reference does
};

constexpr int ELOO_MEAN = 0, ELOO_VARIANCE = 1, ELOO_SD = 2, ELOO_NONE = 3;
constexpr int ELOO_MAX_TAIL = 128;

struct ElooParams {
    const double* x;  // nullable (type NONE): h(theta) rows
    long long x_stride;
    const double* lw;  // log weights rows (any normalisation)
    long long lw_stride;
    const double* lr;  // raw log ratios rows; == lw when the caller has none
    long long lr_stride;
    double* value;    // N (unused for type NONE)
    double* khat;     // N
    double* scratch;  // unstaged mode: one row of S doubles per CTA for h * r
    long long n_rows;
    int S;
    int type;
    int tail_len;
    int bulk;
    int grid_cap;  // unstaged mode: number of scratch rows the workspace holds (0 = no limit)
};

constexpr int ELOO_MAX_PROBS = 32;
constexpr int ELOO_QUANT_MAX_S = 16384;  // sort buffer: 12 B per (padded) draw in shared memory

struct QuantParams {
    const double* x;  // draws, rows of S doubles
    long long x_stride;
    const double* lw;  // log weights rows (any normalisation)
    long long lw_stride;
    double* out;  // N x n_probs
    long long n_rows;
    int S;
    int P2;  // S padded to a power of two (>= 256)
    int n_probs;
    double probs[ELOO_MAX_PROBS];
};

// Leave-one-group-out: out[g][s] = sum over the members i of group g (ascending i) of ll(s, i), NaN -> -1e10
// (pyloo/loo_group.py:188-222).  members / offsets: CSR of the groups (device).  out: G rows of S doubles.
struct GroupSumParams {
    const double* ll;
    long long stride_s, stride_n;  // element (s, i) at ll[s * stride_s + i * stride_n]
    const int* members;            // [N] observation indices grouped by group, ascending inside a group
    const int* offsets;            // [G + 1]
    double* out;
    long long out_stride;          // row stride of out (>= S)
    unsigned long long* counters;  // nullable: [0] += NaN inputs
    int S, G;
};
cudaError_t group_sum_launch(const GroupSumParams& p, cudaStream_t st);

// WAIC pointwise pass straight on the observation-fastest (S, N) matrix (pyloo/waic.py:122-145): no
// transposed panels, one read of the matrix.  Outputs as b2l_loo_dev_f64 with B2L_FLAG_WAIC_ONLY.
struct WaicColsParams {
    const double* ll;      // element (s, i) at ll[s * stride_s + i]
    long long stride_s;
    double *elpd_i, *k_i, *lppd_i, *var_i, *lppdw_i;  // N each
    unsigned long long* counters;                     // nullable: NaN / +inf / -inf inputs
    long long N;
    int S;
    double log_S;
};
cudaError_t waic_cols_launch(const WaicColsParams& p, cudaStream_t st);

// loo(method = "sis" | "tis") pointwise pass straight on the observation-fastest (S, N) matrix: two (SIS) or
// three (TIS) reads of the matrix, no transposed panels.  Same outputs as is_launch in IS_MODE_LOO.
struct IsColsParams {
    const double* ll;  // element (s, i) at ll[s * stride_s + i]
    long long stride_s;
    double *elpd, *ess, *lppd;     // N each
    unsigned long long* counters;  // nullable: NaN / +inf / -inf inputs
    long long N;
    int S;
    double log_S;
};
cudaError_t is_cols_launch(int method, const IsColsParams& p, cudaStream_t st);

// Launch planners + launchers.  `*_info`: [0] staged in shared memory, [1] grid, [2] dynamic smem bytes,
// [3] CTAs per SM.  The e_loo launcher needs `scratch` only when info[0] == 0.
cudaError_t is_plan(int method, int mode, int S, long long n_rows, int* info);
cudaError_t is_launch(int method, int mode, const IsParams& p, cudaStream_t st);
cudaError_t eloo_plan(int S, long long n_rows, bool has_x, int tail_len, int* info);
cudaError_t eloo_launch(const ElooParams& p, cudaStream_t st);
cudaError_t eloo_quantile_launch(QuantParams p, cudaStream_t st);  // fills p.P2

}  // namespace b2l
