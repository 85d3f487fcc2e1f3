/* psisloo_b200.h -- C ABI of libpsisloo_b200.so (B200 / sm_100a PSIS-LOO engine).
 *
 * The reference (jordandeklerk/pyloo) has no FFI: its one extension point for this path is the
 * 1-D callable handed to the per-observation Python loop (pyloo/base.py:138-166 -> pyloo/utils.py
 * :171-176).  Each entry point below replaces that loop *for a whole batch*; the reference
 * interface it stands in for is cited per function.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - "dev" entry points take caller-owned DEVICE memory and are asynchronous on `stream`;
 *     "host" entry points take HOST memory, stage it through the GPU in observation chunks
 *     (H2D / kernel / D2H overlapped on internal streams) and return when results are on the host.
 *   - return 0 on success, a negative B2L_E_* for argument errors, a positive cudaError_t for CUDA
 *     failures; b2l_last_error() returns a thread-local message.  Nothing throws across the ABI.
 *   - M ("tail length") and cutoffmin are computed by the caller exactly as pyloo/psis.py:89-90
 *     does (Python floats): M = ceil(min(S/5, 3*sqrt(S/reff))), cutoffmin = log(DBL_MIN).
 *   - all arithmetic is IEEE float64.
 */
#ifndef PSISLOO_B200_H
#define PSISLOO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2L_VERSION 100 /* 0.1.0 */

#define B2L_E_INVALID (-1)     /* bad pointer / size / stride */
#define B2L_E_UNSUPPORTED (-2) /* shape not supported by this build (e.g. S too large for smem) */
#define B2L_E_WORKSPACE (-3)   /* workspace too small */
#define B2L_E_NODEVICE (-4)    /* no CUDA device: there is NO CPU fallback */

/* flags for the loo entry points */
#define B2L_FLAG_WAIC_ONLY 1u /* skip PSIS: only lppd_i / var_i / lppdw_i (pyloo/waic.py path) */
#define B2L_FLAG_NO_TILE 2u   /* (S, N) layout: take the transposed-panel route even where the cluster kernel could read the
                                 matrix in place.  Same results; for inputs whose columns the cluster kernel hands to the
                                 general kernel wholesale (log-likelihoods with Student-t(3)-like tails: the normaliser is
                                 carried by a few draws).  The host entry points set it themselves once more than a fifth
                                 of a chunk's observations were handed over. */

/* fixed per-shard statistics record (doubles); combined across GPUs by b2l_stats_merge */
#define B2L_STATS_LEN 32
enum {
    B2L_ST_N = 0,         /* observations in the shard */
    B2L_ST_ELPD_MEAN = 1, /* mean of elpd_i (log scale) */
    B2L_ST_ELPD_M2 = 2,   /* sum (elpd_i - mean)^2 */
    B2L_ST_ELPD_SUM = 3,
    B2L_ST_LPPD_SUM = 4,  /* sum of loo-policy lppd_i (pyloo/loo.py:329) */
    B2L_ST_PWAIC_SUM = 5, /* sum var_i (pyloo/waic.py:160) */
    B2L_ST_WAIC_MEAN = 6, /* mean of (lppdw_i - var_i) */
    B2L_ST_WAIC_M2 = 7,
    B2L_ST_WAIC_SUM = 8,
    B2L_ST_K_GT_GOOD = 9, /* #k > good_k (inf counts, NaN does not; pyloo/loo.py:292-293) */
    B2L_ST_K_GT_1 = 10,
    B2L_ST_K_INF = 11,
    B2L_ST_K_NAN = 12,
    B2L_ST_VAR_GT_04 = 13, /* #var_i > 0.4 (pyloo/waic.py:147) */
    B2L_ST_ELPD_MIN = 14,
    B2L_ST_ELPD_MAX = 15,
    B2L_ST_WAIC_MIN = 16,
    B2L_ST_WAIC_MAX = 17,
    B2L_ST_N_NAN_IN = 18, /* NaN entries seen in the input (pyloo/loo.py:218) */
    B2L_ST_N_PINF_IN = 19,
    B2L_ST_N_NINF_IN = 20,
    B2L_ST_N_FALLBACK = 21, /* rows that took the exact bit-wise selection fallback */
    B2L_ST_ELPD_NAN = 22    /* #elpd_i that are NaN (excluded from nothing; diagnostic) */
};

int b2l_version(void);
const char* b2l_last_error(void);

/* Number of CUDA devices visible, or B2L_E_NODEVICE. */
int b2l_device_count(void);

/* Bytes of device workspace the dev entry points need for this problem.
 * layout_obs_fastest: 1 if the input has stride_n == 1 (ArviZ (chain, draw, obs) layout). */
int b2l_workspace_bytes(int64_t S, int64_t N, int32_t M, int32_t layout_obs_fastest,
                        size_t* out_bytes);

/* psislw for a batch: replaces pyloo/psis.py:100-106 (wrap_xarray_ufunc(_psislw, ...)) and the
 * PSIS branch of pyloo/base.py:160-166.
 *   lw      : N x S log weights, element (i, s) at lw[i*stride_n + s*stride_s]; never modified
 *   lw_out  : same logical shape, element (i, s) at lw_out[i*ostride_n + s*ostride_s]
 *   k_out   : N Pareto shape estimates (+inf where the tail has <= 4 draws, psis.py:142-144)
 * Fast path: stride_s == 1 and ostride_s == 1 (contiguous rows).  stride_n == 1 is supported
 * through an on-device tile transpose.                                                          */
int b2l_psislw_dev_f64(const double* lw, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                       int32_t M, double cutoffmin, double* lw_out, int64_t ostride_s,
                       int64_t ostride_n, double* k_out, double* diag /* nullable, N x 8 */,
                       void* ws, size_t ws_bytes, void* stream);

/* Fused pointwise PSIS-LOO + WAIC for a batch: replaces pyloo/loo.py:286-289 (weights, lw += ll),
 * :319-337 (loo_i and lppd_i loops) and pyloo/waic.py:137-145 (lppd_i, var_i).
 *   ll      : log-likelihood, element (s, i) at ll[s*stride_s + i*stride_n]; NaN -> -1e10
 *             (loo.py:227); for the WAIC outputs +-inf -> +-1e10 (waic.py:129-132)
 *   elpd_i  : N  log-scale elpd_loo_i      k_i     : N  Pareto k
 *   lppd_i  : N  log mean exp ll (loo)     var_i   : N  var_s(ll), ddof 0 (waic policy)
 *   lppdw_i : N  log mean exp ll (waic policy; equals lppd_i unless the row holds +-inf)
 *   counters: nullable, 4 x uint64 device counters (NaN, +inf, -inf inputs; fallback rows), +=
 * Routes (same results on every one): stride_n == 1 (the ArviZ (chain, draw, obs) layout) with a 16-byte aligned
 * base, an even stride_s, S even and 512 <= S <= 16384 (above 4096: divisible into 2, 3 or 4 chunks of <= 4096
 * draws), M + 1 <= 510 is read where it lies by the cluster kernel (one pass over HBM, 2-D TMA tiles); otherwise
 * the matrix goes through transposed row panels; stride_s == 1 (rows contiguous) takes the row kernels directly.
 * B2L_FLAG_NO_TILE forces the panel route.                                                          */
int b2l_loo_dev_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                    int32_t M, double cutoffmin, uint32_t flags, double* elpd_i, double* k_i,
                    double* lppd_i, double* var_i, double* lppdw_i, unsigned long long* counters,
                    double* diag /* nullable, N x 8 */, void* ws, size_t ws_bytes, void* stream);

/* The same call with one more, nullable output: tail_idx, N x M int32 (row-major, device).  Row i receives the draw
 * indices of observation i's tail -- the draws with x > cutoff, pyloo/psis.py:139-141 -- ordered by decreasing x
 * (fast path) or increasing x (general kernel), padded with -1 where the tail is shorter than M.  It is the direct
 * evidence for "the selected tail indices are bit-exact" (compare as a set with np.where(x > x_cutoff)).            */
int b2l_loo_dev_ex_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                       int32_t M, double cutoffmin, uint32_t flags, double* elpd_i, double* k_i,
                       double* lppd_i, double* var_i, double* lppdw_i, unsigned long long* counters,
                       double* diag /* nullable, N x 8 */, int32_t* tail_idx /* nullable, N x M */, void* ws,
                       size_t ws_bytes, void* stream);

/* Per-shard statistics of the pointwise outputs (device pointers): replaces the NumPy reductions
 * pyloo/loo.py:326-342 and pyloo/waic.py:147-160.  Writes B2L_STATS_LEN doubles to stats_out
 * (device).  Deterministic (fixed reduction tree).  counters may be NULL.                       */
int b2l_stats_dev_f64(const double* elpd_i, const double* k_i, const double* lppd_i,
                      const double* var_i, const double* lppdw_i, int64_t N, double good_k,
                      const unsigned long long* counters, double* stats_out, void* ws,
                      size_t ws_bytes, void* stream);

/* Merge n_shards host-side records (rank order) into one: Chan's parallel (n, mean, M2) update,
 * so se = sqrt(N * M2 / N) is independent of the GPU count.  Pure host arithmetic.             */
int b2l_stats_merge(const double* shards, int32_t n_shards, double* merged);

/* Host-buffer variants (what a NumPy caller binds): same semantics, HOST pointers, internal
 * chunking.  `device` selects the GPU; chunk_obs <= 0 picks a default.                          */
int b2l_psislw_host_f64(const double* lw, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                        int32_t M, double cutoffmin, double* lw_out, int64_t ostride_s,
                        int64_t ostride_n, double* k_out, int32_t device, int64_t chunk_obs);

int b2l_loo_host_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                     int32_t M, double cutoffmin, uint32_t flags, double good_k, double* elpd_i,
                     double* k_i, double* lppd_i, double* var_i, double* lppdw_i, double* stats_out,
                     int32_t device, int64_t chunk_obs);

/* The same two calls over SEVERAL GPUs of one box from ONE process: the observation axis is cut into
 * contiguous shards (whole 16-observation tiles), one host thread and one chunk pipeline per device run
 * concurrently, and the shard records are merged in shard order (b2l_stats_merge).  This is the reference's
 * per-observation independence (pyloo/utils.py:171-176) used across devices: no data-path exchange.
 * `devices`: n_devices CUDA device indices.  Pointwise outputs are bit-identical to the one-device call.
 * B2L_HOST_REGISTER=1 pins a pageable input (and psislw output) for the duration of the call.          */
int b2l_loo_host_mgpu_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                          int32_t M, double cutoffmin, uint32_t flags, double good_k, double* elpd_i,
                          double* k_i, double* lppd_i, double* var_i, double* lppdw_i, double* stats_out,
                          const int32_t* devices, int32_t n_devices, int64_t chunk_obs);
int b2l_psislw_host_mgpu_f64(const double* lw, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                             int32_t M, double cutoffmin, double* lw_out, int64_t ostride_s,
                             int64_t ostride_n, double* k_out, const int32_t* devices, int32_t n_devices,
                             int64_t chunk_obs);

/* ---------------------------------------------------------------------------------------------------
 * The callers either side of psislw ("next" rows 3 and 1).  Device pointers, asynchronous
 * on `stream`, rows of S doubles with element stride 1 unless stated otherwise.
 * ------------------------------------------------------------------------------------------------- */
#define B2L_IS_SIS 1 /* standard importance sampling, pyloo/sis.py:86-106 */
#define B2L_IS_TIS 2 /* truncated importance sampling, pyloo/tis.py:91-120 */

/* SIS / TIS branch of compute_importance_weights (pyloo/base.py:146-152,160-166): replaces the
 * per-observation loop over _sislw / _tislw.  lw: N rows (row i at lw + i*stride_n), never modified;
 * lw_out: N rows at lw_out + i*ostride_n; ess_out: N effective sample sizes 1 / sum(w^2).          */
int b2l_islw_dev_f64(const double* lw, int64_t S, int64_t N, int64_t stride_n, int32_t method,
                     double* lw_out, int64_t ostride_n, double* ess_out, void* stream);

/* loo(method="sis"|"tis") pointwise pass: replaces pyloo/loo.py:286-289 (weights of -ll, lw += ll),
 * :319-324 (elpd_i = logsumexp) and :329-337 (lppd_i).  ll element (s, i) at ll[s*stride_s + i*stride_n]
 * with one of the strides equal to 1 (stride_n == 1 is the ArviZ layout: column-form kernel, one sweep of the
 * matrix per logsumexp, no transposed panels); NaN -> -1e10 (loo.py:227).  counters: nullable, 3 x uint64 (+=) NaN / +inf /
 * -inf inputs.  Workspace: b2l_is_workspace_bytes.                                                   */
int b2l_is_workspace_bytes(int64_t S, int64_t N, int32_t layout_obs_fastest, size_t* out_bytes);
int b2l_loo_is_dev_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                       int32_t method, double* elpd_i, double* ess_i, double* lppd_i,
                       unsigned long long* counters, void* ws, size_t ws_bytes, void* stream);

/* e_loo weighted expectations with their Pareto-k diagnostic: replaces pyloo/e_loo.py:429-463,518-531
 * (_compute_weighted_mean / _variance / _sd over normalised weights) and :266-390 (compute_pareto_k ->
 * k_hat, three top-`tail_len` tails + _gpdfit each) for a batch of N observations.
 *   x   : the draws, N rows; moments are taken of x, and k_hat's h is x (mean) or x*x (variance, sd;
 *         e_loo.py:234-239); NULL with B2L_ELOO_NONE (quantiles: k_hat of the ratios only)
 *   lw  : log weights rows, any normalisation (normalised inside, e_loo.py:557-559)
 *   lr  : raw log ratios rows for k_hat, or NULL to use lw (e_loo.py:229-230)
 *   value_out, khat_out : N doubles each.                                                            */
#define B2L_ELOO_MEAN 0
#define B2L_ELOO_VARIANCE 1
#define B2L_ELOO_SD 2
#define B2L_ELOO_NONE 3
int b2l_eloo_workspace_bytes(int64_t S, int64_t N, int32_t has_lr, int32_t type, size_t* out_bytes);
int b2l_eloo_dev_f64(const double* x, int64_t x_stride_n, const double* lw, int64_t lw_stride_n,
                     const double* lr, int64_t lr_stride_n, int64_t S, int64_t N, int32_t type,
                     int32_t tail_len, double* value_out, double* khat_out, void* ws, size_t ws_bytes,
                     void* stream);

/* e_loo(type="quantile"): replaces pyloo/e_loo.py:466-515 (the Python loop over observations x probabilities
 * around _weighted_quantile, :534-554).  probs: HOST array of n_probs (<= 32) values in (0, 1);
 * value_out: N x n_probs (row-major, device).  Each row is sorted in shared memory, so S <= 16384; larger S
 * returns B2L_E_UNSUPPORTED.  Ties in x are ordered by draw index (np.argsort's order among ties is
 * unspecified); the Pareto k of a quantile comes from b2l_eloo_dev_f64 with B2L_ELOO_NONE.             */
int b2l_eloo_quantile_dev_f64(const double* x, int64_t x_stride_n, const double* lw, int64_t lw_stride_n,
                              int64_t S, int64_t N, const double* probs, int32_t n_probs,
                              double* value_out, void* stream);

/* loo_group (leave-one-group-out): replaces the aggregation loop pyloo/loo_group.py:188-222.
 *   out[g * out_stride_g + s] = sum over the observations i of group g of ll(s, i), NaN -> -1e10 first
 * with the groups given in CSR form (device arrays): members[offsets[g] .. offsets[g+1]) are the observation
 * indices of group g in ascending order.  The G x S result (rows contiguous) then goes through
 * b2l_loo_dev_f64 / b2l_loo_is_dev_f64 with stride_s = 1 -- one "observation" per group.
 * counters: nullable, counters[0] += NaN inputs.                                                       */
int b2l_group_sum_dev_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                          const int32_t* members, const int32_t* offsets, int32_t G, double* out,
                          int64_t out_stride_g, unsigned long long* counters, void* stream);

/* loo_subsample's PSIS stage: the subsampled observations (pyloo/loo_subsample.py:330,
 * `log_likelihood.isel(...)`, feeding :371-383) gathered on the device into contiguous rows,
 *   out[j * out_stride_n + s] = ll[s * stride_s + idx[j] * stride_n],  j < m,
 * which then go through b2l_loo_dev_f64 with stride_s = 1.  idx: m observation indices (device, int64;
 * repeats allowed -- the reference samples with replacement); the caller guarantees 0 <= idx[j] < N.       */
int b2l_gather_rows_dev_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                            const int64_t* idx, int64_t m, double* out, int64_t out_stride_n, void* stream);

/* Per-kernel device timing for benchmarks (no reference counterpart): with b2l_profile(1) every kernel
 * launch is bracketed by CUDA events on its stream; b2l_profile_read() synchronises them and returns
 * summed milliseconds and launch counts per kernel kind since the last read.  Not thread safe.    */
#define B2L_PROF_KINDS 8
enum {
    B2L_PROF_STREAM = 0,    /* psis_stream_kernel (row pass; psislw: + fused apply of the previous batch) */
    B2L_PROF_TAIL = 1,      /* psis_tail_kernel (sort, GPD fit, smoothing, normaliser) */
    B2L_PROF_APPLY = 2,     /* psis_apply_kernel (only when the fused apply is disabled) */
    B2L_PROF_ROW = 3,       /* psis_row_kernel: general kernel, whole batch or hand-over rows */
    B2L_PROF_TRANSPOSE = 4, /* transpose_f64_kernel */
    B2L_PROF_STATS = 5,     /* stats_partial_kernel + stats_final_kernel */
    B2L_PROF_IS = 6,        /* is_row_kernel (SIS / TIS weights or loo) */
    B2L_PROF_ELOO = 7       /* eloo_row_kernel (weighted expectations + k_hat) */
};
int b2l_profile(int32_t enable);
int b2l_profile_read(double* ms_out /* [B2L_PROF_KINDS] */, int64_t* launches_out /* [B2L_PROF_KINDS] */);

/* Diagnostics: why observations were handed from the split path to the general kernel on the current
 * device since the last reset; out16[reason]: 1 NaN/inf row, 2 range > 1e7, 3 threshold retries exhausted,
 * 4 long run of equal sort keys, 5 order check, 6-9 GPD fit (quantile <= 0, factor overflow, product,
 * non-finite profile), 10 body cancellation (cluster kernel: the all-draw sum minus the raw tail leaves less
 * than 1e-3 of it).  Synchronises the device.                                                    */
int b2l_handover_reasons(uint64_t* out16, int32_t reset);

/* Launch shape of the split path for (S, M): info[16] = ok, stream threads, draws per thread, tail
 * registers per lane, candidate capacity, q0, row buffers, fused apply, stream grid, tail grid,
 * stream CTAs/SM, tail CTAs/SM, stream smem, tail smem, observations per round, stream block size. */
int b2l_split_launch_info(int64_t S, int32_t M, int32_t mode, int64_t n_rows, int32_t* info);

/* Shape of the cluster kernel's plan for pl.loo on the (chain, draw, obs) layout (pure arithmetic, no device call):
 * info[16] = eligible, observations per tile, CTAs per cluster, chunks of the draw axis, draws per chunk, draws per
 * CTA, TMA boxes per CTA, draws per box, tight / loose threshold rank, shared memory per CTA, tail-kernel registers
 * per lane, candidate capacity per observation, 0, 0, 0.  (Alignment of the caller's matrix is checked per call.) */
int b2l_tile_shape_info(int64_t S, int32_t M, int32_t* info);

/* Launch-shape introspection for benchmarks / DESIGN.md (grid, block, smem, occupancy). */
int b2l_row_launch_info(int64_t S, int32_t M, int32_t mode /*0 psislw, 1 loo*/, int32_t* grid,
                        int32_t* block, int32_t* smem_bytes, int32_t* ctas_per_sm, int32_t* nbuf);

#ifdef __cplusplus
}
#endif
#endif /* PSISLOO_B200_H */
