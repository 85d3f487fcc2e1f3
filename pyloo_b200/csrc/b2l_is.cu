// SIS / TIS importance weights, e_loo weighted expectations / quantiles and the LOGO group sums for sm_100a.
//
// These are the callers either side of psislw.  One CTA of 256 threads per
// observation row, persistent grid of SMs x occupancy; the row (e_loo: the ratios and the draws) is staged in
// shared memory by 1-D bulk TMA when it fits and read straight from global memory otherwise, so every S is
// supported (the quantile kernel sorts in shared memory: S <= 16384).  All arithmetic is IEEE FP64; the exps of
// the normalising sums are table-driven (exp_sum below), everything else uses the CUDA math library.  The
// kernels move 16*S (weights) / 8*S (loo) / 16-24*S (e_loo) bytes per observation.
//
// Reference semantics
//   SIS   pyloo/sis.py:101-106    x -= max; x -= logsumexp(x); ess = 1 / sum(exp(x)^2)
//   TIS   pyloo/tis.py:108-120    x -= max; log_Z = lse(x) - log S; cut = log_Z + 0.5 log S;
//                                 x = minimum(x, cut); x -= logsumexp(x); ess = 1 / sum(exp(x)^2)
//   loo   pyloo/loo.py:286-289    lw += ll;  :319-324 elpd_i = lse(lw);  :329-337 lppd_i = lse(ll) - log S
//   e_loo pyloo/e_loo.py:429-463,518-531 (weighted mean / variance / sd), :328-390 (k_hat),
//         :466-554 (weighted quantiles)
//   logo  pyloo/loo_group.py:215-222 (per-group sums of the log-likelihood)
// NaN handling follows NumPy: np.max and np.minimum propagate NaN, so a row holding NaN (or whose
// maximum is not finite) comes out all-NaN through ordinary IEEE arithmetic.
#include <math_constants.h>

#include "b2l_common.cuh"
#include "b2l_is_host.h"

namespace b2l {

constexpr int IS_NT = 256;
constexpr int IS_NW = IS_NT / 32;
constexpr int IS_RED_WORDS = 194;  // 128 reduction words + mbarrier (16 B) + 64 words of exp table

// exp(a) for the normalising sums, a <= 0: table-driven (relative error < 4e-15, b2l_common.cuh) for the
// ordinary range, the library routine for NaN / -inf / arguments below -1e7 (where the table's one-step
// argument reduction no longer holds) so that IEEE special cases propagate like NumPy's.
static __device__ __noinline__ double exp_slow(double a) { return exp(a); }
__device__ __forceinline__ double exp_sum(double a, const ExpTab& tb) {
    return (a >= -1e7) ? exp_tab_drop(a, tb, false) : exp_slow(a);
}
__device__ __forceinline__ ExpTab exp_table_init(double* red) {
    double* tab = red + 130;
    if (threadIdx.x < 32) {
        tab[threadIdx.x] = exp2((double)threadIdx.x / 32.0);
        tab[32 + threadIdx.x] = exp2(-(double)threadIdx.x / 32.0);
    }
    ExpTab tb;
    tb.t = tab;
    tb.tinv = tab + 32;
    return tb;
}

// running maximum / minimum by compare-select (3 instructions; an IEEE fmax costs twice as much).  A NaN
// candidate never wins -- every caller tracks NaN separately, exactly as it had to with fmax.
__device__ __forceinline__ double sel_max(double m, double v) { return (v > m) ? v : m; }
__device__ __forceinline__ double sel_min(double m, double v) { return (v < m) ? v : m; }
__device__ __forceinline__ double np_minimum(double a, double b) {
    return (a != a || b != b) ? nan_f64() : fmin(a, b);
}
// Python's max(a, b): returns b only if b > a (so a NaN first argument wins, a NaN second loses)
__device__ __forceinline__ double py_max(double a, double b) { return (b > a) ? b : a; }
// np.isclose(a, b) with the default rtol = 1e-5, atol = 1e-8 (equal infinities are close)
__device__ __forceinline__ bool np_isclose(double a, double b) {
    return (a == b) || (fabs(a - b) <= 1e-8 + 1e-5 * fabs(b));
}

// bitwise OR over the block (every thread returns the same word); __syncthreads_or only returns a boolean
__device__ __forceinline__ int block_or(int v, double* red_) {
    int* red = reinterpret_cast<int*>(red_ + 64);
    v = __reduce_or_sync(FULL, v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < IS_NW; ++w) t |= red[w];
    __syncthreads();
    return t;
}

// N reductions behind one pair of barriers: warp butterflies, one word per (value, warp), every thread folds
// the 8 warp results in the same fixed order (deterministic, identical in all threads).  N <= 8.
template <int N, bool IS_MAX>
__device__ __forceinline__ void block_reduce_n(double (&v)[N], double* red) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = IS_MAX ? warp_max(v[i]) : warp_sum(v[i]);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) red[i * IS_NW + (threadIdx.x >> 5)] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double t = red[i * IS_NW];
#pragma unroll
        for (int w = 1; w < IS_NW; ++w) t = IS_MAX ? fmax(t, red[i * IS_NW + w]) : t + red[i * IS_NW + w];
        v[i] = t;
    }
    __syncthreads();
}

template <bool STAGED>
__device__ __forceinline__ const double* stage_rows(double* const* dst, const double* const* src, int n_arr,
                                                    int S, int bulk, uint64_t* bar, uint32_t& parity) {
    if (!STAGED) return nullptr;
    const int tid = threadIdx.x;
    if (bulk) {
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)(n_arr * S * 8));
            for (int a = 0; a < n_arr; ++a) bulk_g2s(dst[a], src[a], (uint32_t)(S * 8), bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1;
    } else {
        for (int a = 0; a < n_arr; ++a)
            for (int s = tid; s < S; s += IS_NT) dst[a][s] = src[a][s];
        __syncthreads();
    }
    return dst[0];
}

// ===================================================================================== SIS / TIS
// EPT > 0 (TIS, S <= IS_NT * EPT): the exp of every draw is kept in EPT registers per thread, so the sum over
// the truncated row reuses it (exp(min(x, cut) - m) == min(exp(x) * exp(-m), exp(cut - m)) up to rounding)
// instead of evaluating a second exp per draw.
template <int METHOD, int MODE, bool STAGED, int EPT = 0>
__global__ void __launch_bounds__(IS_NT) is_row_kernel(const IsParams p) {
    extern __shared__ __align__(16) unsigned char is_smem[];
    double* red = reinterpret_cast<double*>(is_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 128);
    double* buf = red + IS_RED_WORDS;
    const int tid = threadIdx.x, S = p.S;
    const ExpTab tb = exp_table_init(red);
    if (STAGED && p.bulk && tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    unsigned c_nan = 0, c_pinf = 0, c_ninf = 0;

    for (long long row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
        const double* src = p.in + row * p.in_stride;
        double* dsts[1] = {buf};
        const double* srcs[1] = {src};
        stage_rows<STAGED>(dsts, srcs, 1, S, p.bulk, bar, parity);
        const double* r = STAGED ? buf : src;

        // raw log weight of draw s: weights mode reads it, loo mode negates the sanitised log-likelihood
        auto lwraw = [&](int s) -> double {
            double v = r[s];
            if (MODE == IS_MODE_LOO) {
                if (v != v) v = -1e10;  // pyloo/loo.py:227
                v = -v;
            }
            return v;
        };

        // ---- pass 1: maximum (NaN-propagating), loo: also min (= -max ll) and input counters
        double mx = -inf_f64(), mn = inf_f64();
        int bad = 0;
        for (int s = tid; s < S; s += IS_NT) {
            if (MODE == IS_MODE_LOO) {
                const double raw = r[s];
                c_nan += (raw != raw);
                c_pinf += (raw == inf_f64());
                c_ninf += (raw == -inf_f64());
            }
            const double v = lwraw(s);
            bad |= (v != v);
            mx = sel_max(mx, v);
            mn = sel_min(mn, v);
        }
        mx = block_max<IS_NT>(mx, red);
        if (MODE == IS_MODE_LOO) mn = block_min<IS_NT>(mn, red);
        if (__syncthreads_or(bad)) mx = nan_f64();

        // ---- pass 2: first logsumexp (max of x is 0 by construction)
        double s1 = 0.0, s2 = 0.0, sl = 0.0;
        double ecache[EPT > 0 ? EPT : 1];
        if (EPT > 0) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int s = tid + i * IS_NT;
                const double v = (s < S) ? lwraw(s) : -inf_f64();
                const double e = (s < S) ? exp_sum(v - mx, tb) : 0.0;
                ecache[i] = e;
                s1 += e;
                if (MODE == IS_MODE_LOO && s < S) sl += exp_sum(mn - v, tb);
            }
        } else {
#pragma unroll 4
            for (int s = tid; s < S; s += IS_NT) {
                const double v = lwraw(s);
                const double e = exp_sum(v - mx, tb);
                s1 += e;
                s2 += e * e;
                if (MODE == IS_MODE_LOO) sl += exp_sum(mn - v, tb);  // exp(ll - max ll)
            }
        }
        s1 = block_sum<IS_NT>(s1, red);
        if (MODE == IS_MODE_LOO) sl = block_sum<IS_NT>(sl, red);
        double lse, cut = inf_f64(), ess;
        if (METHOD == IS_METHOD_SIS) {
            s2 = block_sum<IS_NT>(s2, red);
            lse = log(s1);
            ess = (s1 * s1) / s2;
        } else {
            const double log_z = log(s1) - p.log_S;  // tis.py:112
            cut = log_z + 0.5 * p.log_S;             // tis.py:114
            const double mx2 = np_minimum(0.0, cut); // max of the truncated row
            double t1 = 0.0, t2 = 0.0;
            if (EPT > 0) {
                const double up = exp(-mx2), ecut = exp(cut - mx2);
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const double a = np_minimum(ecache[i] * up, ecut);  // 0 for the slots past S
                    t1 += a;
                    t2 += a * a;
                }
            } else {
#pragma unroll 4
                for (int s = tid; s < S; s += IS_NT) {
                    const double a = exp_sum(np_minimum(lwraw(s) - mx, cut) - mx2, tb);
                    t1 += a;
                    t2 += a * a;
                }
            }
            t1 = block_sum<IS_NT>(t1, red);
            t2 = block_sum<IS_NT>(t2, red);
            lse = log(t1) + mx2;
            ess = (t1 * t1) / t2;
        }

        // ---- output
        if (MODE == IS_MODE_WEIGHTS) {
            double* dst = p.out + row * p.out_stride;
#pragma unroll 4
            for (int s = tid; s < S; s += IS_NT) {
                double x = lwraw(s) - mx;
                if (METHOD == IS_METHOD_TIS) x = np_minimum(x, cut);
                dst[s] = x - lse;
            }
            if (tid == 0) p.ess[row] = ess;
        } else {
            // elpd_i = logsumexp(lw + ll)  (loo.py:289, :319-324)
            double tmax = -inf_f64();
#pragma unroll 4
            for (int s = tid; s < S; s += IS_NT) {
                const double v = lwraw(s);
                double x = v - mx;
                if (METHOD == IS_METHOD_TIS) x = np_minimum(x, cut);
                tmax = sel_max(tmax, (x - lse) + (-v));
            }
            tmax = block_max<IS_NT>(tmax, red);
            double st = 0.0;
#pragma unroll 4
            for (int s = tid; s < S; s += IS_NT) {
                const double v = lwraw(s);
                double x = v - mx;
                if (METHOD == IS_METHOD_TIS) x = np_minimum(x, cut);
                st += exp_sum(((x - lse) + (-v)) - tmax, tb);
            }
            st = block_sum<IS_NT>(st, red);
            if (tid == 0) {
                p.elpd[row] = log(st) + tmax;
                p.ess[row] = ess;
                p.lppd[row] = log(sl) + ((-mn) - p.log_S);  // utils.py:352-357 with b_inv = S
            }
        }
        __syncthreads();  // every read of the staged row is done before the next bulk load lands
    }
    if (MODE == IS_MODE_LOO && p.counters) {
        if (c_nan) atomicAdd(&p.counters[0], (unsigned long long)c_nan);
        if (c_pinf) atomicAdd(&p.counters[1], (unsigned long long)c_pinf);
        if (c_ninf) atomicAdd(&p.counters[2], (unsigned long long)c_ninf);
    }
}

// ===================================================================================== e_loo
// ---- top-L selection ---------------------------------------------------------------------------
// The caller's lane owns the elements first, first + stride, ... < n of buf.  Extracts the L largest keys
// of the warp's elements in descending order into out[0..L) (padded with -inf when the warp owns fewer).
// Ties are ordered by index, so duplicates are extracted one by one.  NEG selects on -buf[s].
template <bool NEG>
__device__ __forceinline__ void warp_top(const double* buf, int n, int first, int stride, int L,
                                         double* out, int lane) {
    constexpr int NONE = 0x7fffffff;
    double lastv = inf_f64();
    int lasti = -1;
    double bestv;
    int besti;
    auto rescan = [&]() {
        bestv = -inf_f64();
        besti = NONE;
        for (int s = first; s < n; s += stride) {
            const double v = NEG ? -buf[s] : buf[s];
            const bool eligible = (v < lastv) || (v == lastv && s > lasti);
            if (eligible && (besti == NONE || v > bestv)) {
                bestv = v;
                besti = s;
            }
        }
    };
    rescan();
    for (int k = 0; k < L; ++k) {
        double wv = bestv;
        int wi = besti;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(FULL, wv, o);
            const int oi = __shfl_xor_sync(FULL, wi, o);
            const bool take = (oi != NONE) && (wi == NONE || ov > wv || (ov == wv && oi < wi));
            if (take) {
                wv = ov;
                wi = oi;
            }
        }
        if (lane == 0) out[k] = (wi == NONE) ? -inf_f64() : wv;
        if (wi != NONE && wi == besti) {
            lastv = wv;
            lasti = wi;
            rescan();
        }
    }
}

// Merge the per-warp descending lists cb[w * ld + 0..n_tail), w < IS_NW, written by warp_top into the n_tail
// largest overall (descending) in tl.  One warp.
__device__ __forceinline__ void merge_warp_lists(const double* cb, int ld, int n_tail, double* tl, int lane) {
    constexpr int NONE = 0x7fffffff;
    double lastv = inf_f64();
    int lasti = -1;
    double bestv;
    int besti;
    auto rescan = [&]() {
        bestv = -inf_f64();
        besti = NONE;
        for (int c = lane; c < IS_NW * n_tail; c += 32) {
            const int s = (c / n_tail) * ld + (c % n_tail);
            const double v = cb[s];
            const bool eligible = (v < lastv) || (v == lastv && s > lasti);
            if (eligible && (besti == NONE || v > bestv)) {
                bestv = v;
                besti = s;
            }
        }
    };
    rescan();
    for (int k = 0; k < n_tail; ++k) {
        double wv = bestv;
        int wi = besti;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(FULL, wv, o);
            const int oi = __shfl_xor_sync(FULL, wi, o);
            const bool take = (oi != NONE) && (wi == NONE || ov > wv || (ov == wv && oi < wi));
            if (take) {
                wv = ov;
                wi = oi;
            }
        }
        if (lane == 0) tl[k] = wv;
        if (wi != NONE && wi == besti) {
            lastv = wv;
            lasti = wi;
            rescan();
        }
    }
}

// ---- generalised Pareto fit, literal -------------------------------------------------------------
// pyloo/psis.py:181-208 evaluated operation by operation on ary[0..n) IN THE ORDER GIVEN, by one warp.
// k_hat (e_loo.py:357,377,383) hands it descending tails whose last element is 0, so 1/ary[-1] is inf, the
// whole profile is NaN, every grid weight is dropped, b_post = 0 and the result is 5/(n+10); running the
// same arithmetic (IEEE division, NaN-propagating log1p/exp) reproduces that without special cases.
// bs, ls: shared scratch of 64 doubles each (m_est = 30 + floor(sqrt(n)) <= 41 for n <= 128).
__device__ __forceinline__ double gpdfit_literal_warp(const double* ary, int n, double* bs, double* ls,
                                                      int lane) {
    const int m = 30 + (int)sqrt((double)n);
    int q = (int)((double)n / 4.0 + 0.5) - 1;
    if (q < 0) q += n;
    const double aq = ary[q], an = ary[n - 1];
    {
        // Shortcut, exactly equivalent to the literal evaluation below: with ary[n-1] == +-0 and every entry
        // finite, 1/ary[n-1] is +-inf, every b_j is +-inf or NaN, the i = n-1 term of every profile mean is
        // log1p(-(+-inf) * 0) = NaN, so all grid weights are NaN and dropped, b_post = 0 (empty sum),
        // k_post = mean(log1p(-0 * ary)) = -0 and the estimate is (n * -0 + 5) / (n + 10).  This is the only
        // case k_hat produces (the cutoff is an element of the tail), so the loops below never run for it.
        int fin = 1;
        for (int i = lane; i < n; i += 32) fin &= is_finite(ary[i]) ? 1 : 0;
        if (__all_sync(FULL, fin) && an == 0.0) return ((double)n * -0.0 + 5.0) / ((double)n + 10.0);
    }
    for (int j = lane; j < m; j += 32) {
        double b = 1.0 - sqrt((double)m / ((double)(j + 1) - 0.5));
        b /= 3.0 * aq;
        b += 1.0 / an;
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += log1p(-b * ary[i]);
        const double k = acc / (double)n;
        bs[j] = b;
        ls[j] = (double)n * (log(-(b / k)) - k - 1.0);
    }
    __syncwarp();
    double w[2];
    bool keep[2];
    double wsum = 0.0;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        const int j = lane + 32 * jj;
        w[jj] = 0.0;
        keep[jj] = false;
        if (j < m) {
            double d = 0.0;
            for (int l = 0; l < m; ++l) d += exp(ls[l] - ls[j]);
            w[jj] = 1.0 / d;
            keep[jj] = (w[jj] >= 10.0 * 2.220446049250313e-16);
            if (keep[jj]) wsum += w[jj];
        }
    }
    wsum = warp_sum(wsum);
    double bp = 0.0;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
        if (keep[jj]) bp += bs[lane + 32 * jj] * (w[jj] / wsum);
    const double b_post = warp_sum(bp);
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc += log1p(-b_post * ary[i]);
    const double k_post = warp_sum(acc) / (double)n;
    __syncwarp();
    return ((double)n * k_post + 5.0) / ((double)n + 10.0);
}

constexpr int ELOO_FAST_CAP = 128;  // candidates above the sampled threshold handled without the slow path
// scratch area, a union over the phases of a row: the per-warp c-th maxima (3 x 8), the exact-extraction lists
// (8 warps x tail_len) and the literal-fit scratch (3 x 128)
__host__ __device__ inline int eloo_area_words(int tail_len) {
    const int lists = IS_NW * tail_len;
    return lists > 3 * 128 ? lists : 3 * 128;
}

struct ElooSmem {
    __host__ __device__ static size_t bytes(int S, int n_staged, int tail_len) {
        const size_t spad = (size_t)((S + 1) & ~1);
        // red + staged rows + scratch area + tails (3 x tail_len) + fast-path candidates (3 x 128) + results /
        // thresholds / counters
        return sizeof(double) * (IS_RED_WORDS + spad * n_staged + eloo_area_words(tail_len) + 3 * tail_len +
                                 3 * ELOO_FAST_CAP + 16);
    }
};

template <bool STAGED>
__global__ void __launch_bounds__(IS_NT, 3) eloo_row_kernel(const ElooParams p) {
    extern __shared__ __align__(16) unsigned char is_smem[];
    double* red = reinterpret_cast<double*>(is_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, S = p.S;
    const int spad = (S + 1) & ~1;
    const bool has_x = (p.x != nullptr) && (p.type != ELOO_NONE);
    const bool lr_same = (p.lr == p.lw) && (p.lr_stride == p.lw_stride);
    double* sm = red + IS_RED_WORDS;
    double *bx = nullptr, *blw = nullptr, *blr = nullptr;
    if (STAGED) {
        // the ratios are read three times and h * r overwrites x in place: those two rows are staged; log
        // weights that are not also the ratios are read twice straight from global memory (second read: L2)
        blr = sm; sm += spad;
        blw = lr_same ? blr : nullptr;
        if (has_x) { bx = sm; sm += spad; }
    }
    const int LP = p.tail_len;  // row pitch of the tail arrays
    double* area = sm;                        // union, see ElooSmem::bytes
    double* tmax = area;                      // [3][IS_NT] thread-local maxima, sorted per warp
    double* gpd = area;                       // [3][128] literal-fit scratch (after the selection)
    double* tails = area + eloo_area_words(LP);  // [3][L]
    double* fcand = tails + 3 * LP;              // [3][ELOO_FAST_CAP]
    double* res = fcand + 3 * ELOO_FAST_CAP;  // [3] khat per tail
    double* thr = res + 4;                    // [3] thresholds
    int* cnt = reinterpret_cast<int*>(thr + 4);  // [3] candidate counts
    const ExpTab tb = exp_table_init(red);
    if (STAGED && p.bulk && tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    const int L = p.tail_len;
    const int n_tail = (L < S) ? L : S;

    for (long long row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
        const double* gx = has_x ? p.x + row * p.x_stride : nullptr;
        const double* glw = p.lw + row * p.lw_stride;
        const double* glr = p.lr + row * p.lr_stride;
        if (STAGED) {
            double* dsts[3];
            const double* srcs[3];
            int na = 0;
            dsts[na] = blr; srcs[na++] = glr;
            if (has_x) { dsts[na] = bx; srcs[na++] = gx; }
            stage_rows<true>(dsts, srcs, na, S, p.bulk, bar, parity);
        }
        const double* X = STAGED ? bx : gx;
        const double* LW = (STAGED && lr_same) ? blw : glw;
        const double* LR = STAGED ? blr : glr;
        double* HR = STAGED ? bx : p.scratch + (long long)blockIdx.x * spad;  // h * r, in place when staged

        // ---- pass 1: maxima and the degeneracy tests on h = x (mean) or x^2 (variance, sd; e_loo.py:234-239)
        // for k_hat (e_loo.py:359-365) and on x for the weighted variance (e_loo.py:520)
        const bool sq = (p.type != ELOO_MEAN);
        const double x0 = has_x ? X[0] : 0.0;
        const double h0 = sq ? x0 * x0 : x0;
        double lwmax = -inf_f64(), lrmax = -inf_f64(), nemin = inf_f64(), nemax = -inf_f64();
        // 1 NaN in lw, 2 NaN in lr, 4 h not finite, 8 h not all close to h[0], 16 some h != h[0],
        // 32 x not all close to x[0]
        int flags = 0;
        for (int s = tid; s < S; s += IS_NT) {
            const double a = LW[s], b = LR[s];
            flags |= (a != a) ? 1 : 0;
            flags |= (b != b) ? 2 : 0;
            lwmax = sel_max(lwmax, a);
            lrmax = sel_max(lrmax, b);
            if (has_x) {
                const double xv = X[s];
                const double hv = sq ? xv * xv : xv;
                if (!is_finite(hv)) flags |= 4;
                if (!np_isclose(hv, h0)) flags |= 8;
                if (!np_isclose(xv, x0)) flags |= 32;
                if (hv != h0) {
                    flags |= 16;
                    nemin = sel_min(nemin, hv);
                    nemax = sel_max(nemax, hv);
                }
            }
        }
        const double m_r = lrmax;  // this thread's largest log ratio
        {
            double mm[4] = {lwmax, lrmax, -nemin, nemax};
            block_reduce_n<4, true>(mm, red);
            lwmax = mm[0];
            lrmax = mm[1];
            nemin = -mm[2];
            nemax = mm[3];
        }
        flags = block_or(flags, red);
        if (flags & 1) lwmax = nan_f64();
        if (flags & 2) lrmax = nan_f64();
        const bool x_close = has_x && !(flags & 32);
        const bool h_close = has_x && !(flags & 8);
        const bool h_two = has_x && (flags & 16) && (nemin == nemax);  // len(np.unique(h)) == 2
        const bool h_bad = has_x && (flags & 4);
        const bool need_hr = has_x && !h_close && !h_two && !h_bad && is_finite(lrmax);

        // ---- pass 2: weighted sums (e_loo.py:429-463) and h * r (e_loo.py:368)
        double se = 0.0, sex = 0.0, sexx = 0.0, see = 0.0;
        double m_hi = -inf_f64(), m_lo = -inf_f64();  // this thread's largest h*r and largest -(h*r)
        if (has_x) {
            const bool same = (LR == LW);
#pragma unroll 4
            for (int s = tid; s < S; s += IS_NT) {
                const double xv = X[s];
                const double e = exp_sum(LW[s] - lwmax, tb);
                se += e;
                sex += e * xv;
                sexx += e * (xv * xv);
                see += e * e;
                if (need_hr) {
                    // (library exp: these products are ranked, differenced against their cutoff and fitted --
                    // e_loo.py:368-383 -- where the table exponential's 4e-15 would be amplified by a crowded tail)
                    const double hr = (sq ? xv * xv : xv) * exp((same ? LW[s] : LR[s]) - (same ? lwmax : lrmax));
                    HR[s] = hr;
                    m_hi = sel_max(m_hi, hr);
                    m_lo = sel_max(m_lo, -hr);
                }
            }
            double ss[4] = {se, sex, sexx, see};
            block_reduce_n<4, false>(ss, red);
            se = ss[0];
            sex = ss[1];
            sexx = ss[2];
            see = ss[3];
        }
        __syncthreads();  // HR complete

        // ---- k_hat: tails of r and of h * r (e_loo.py:350-390)
        // Top-n_tail selection.  Fast path: every warp finds the c-th largest of its 32 thread-local maxima
        // (c = ceil(n_tail / 8)); the smallest of those 8 values has at least 8c >= n_tail distinct elements
        // at or above it, so it is a lower bound of the n_tail-th largest element and everything >= it is a
        // candidate (a few dozen on ordinary rows); candidates are ranked by counting.  Rows with many ties at
        // the threshold overflow the candidate list and take the exact extraction (warp_top) instead.
        const bool r_ok = is_finite(lrmax);
        const bool want[3] = {r_ok, need_hr, need_hr};
        {
            const int c = (n_tail + IS_NW - 1) / IS_NW;
            const double keys[3] = {m_r, m_lo, m_hi};
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (!want[t]) continue;
                double mine = keys[t], cth = -inf_f64();
                for (int k = 0; k < c; ++k) {  // c rounds: take the warp maximum out, one lane at a time
                    cth = warp_max(mine);
                    const unsigned holders = __ballot_sync(FULL, mine == cth);
                    if (lane == __ffs(holders) - 1) mine = -inf_f64();
                    if (cth == -inf_f64()) break;  // fewer than c lanes own elements (warp-uniform)
                }
                if (lane == 0) tmax[t * IS_NW + warp] = cth;
            }
            if (tid < 3) cnt[tid] = 0;
        }
        __syncthreads();
        if (tid < 3 && want[tid]) {
            double t = tmax[tid * IS_NW];
            for (int w = 1; w < IS_NW; ++w) t = fmin(t, tmax[tid * IS_NW + w]);
            thr[tid] = t;
        }
        __syncthreads();
        {
            const double t0 = thr[0], t1 = thr[1], t2 = thr[2];
            for (int s = tid; s < S; s += IS_NT) {
                if (want[0]) {
                    const double v = LR[s];
                    if (v >= t0) {
                        const int pos = atomicAdd(&cnt[0], 1);
                        if (pos < ELOO_FAST_CAP) fcand[pos] = v;
                    }
                }
                if (want[1]) {
                    const double v = HR[s];
                    if (-v >= t1) {
                        const int pos = atomicAdd(&cnt[1], 1);
                        if (pos < ELOO_FAST_CAP) fcand[ELOO_FAST_CAP + pos] = -v;
                    }
                    if (v >= t2) {
                        const int pos = atomicAdd(&cnt[2], 1);
                        if (pos < ELOO_FAST_CAP) fcand[2 * ELOO_FAST_CAP + pos] = v;
                    }
                }
            }
        }
        __syncthreads();
        const bool slow[3] = {want[0] && cnt[0] > ELOO_FAST_CAP, want[1] && cnt[1] > ELOO_FAST_CAP,
                              want[2] && cnt[2] > ELOO_FAST_CAP};
        if (warp < 3 && want[warp] && !slow[warp]) {
            // rank the candidates by counting (ties by slot): the n_tail largest land in the tail, descending
            double* tl = tails + warp * LP;
            const double* fc = fcand + warp * ELOO_FAST_CAP;
            const int c = cnt[warp];
            for (int i = lane; i < c; i += 32) {
                const double v = fc[i];
                int rank = 0;
                for (int j = 0; j < c; ++j) {
                    const double o = fc[j];
                    rank += (o > v || (o == v && j < i)) ? 1 : 0;
                }
                if (rank < n_tail) tl[rank] = v;
            }
        }
        // exact extraction for the tails that overflowed, one at a time through the shared scratch area
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (!slow[t]) continue;  // block-uniform
            __syncthreads();
            if (t == 0) warp_top<false>(LR, S, tid, IS_NT, n_tail, area + warp * LP, lane);
            else if (t == 1) warp_top<true>(HR, S, tid, IS_NT, n_tail, area + warp * LP, lane);
            else warp_top<false>(HR, S, tid, IS_NT, n_tail, area + warp * LP, lane);
            __syncthreads();
            if (warp == 0) merge_warp_lists(area, LP, n_tail, tails + t * LP, lane);
        }
        __syncthreads();  // tails complete; the scratch area is free for the fit
        if (warp < 3 && want[warp]) {
            double* tl = tails + warp * LP;
            __syncwarp();
            // tail values in the reference's order, then its degeneracy test and the fit argument
            // warp 0: sorted_r descending (e_loo.py:351); warp 1: left tail ascending (:370); warp 2: right
            // tail descending (:371)
            const double first = (warp == 0) ? exp(tl[0] - lrmax) : (warp == 1 ? -tl[0] : tl[0]);
            __syncwarp();
            int far = 0;
            for (int i = lane; i < n_tail; i += 32) {
                double v = tl[i];
                v = (warp == 0) ? exp(v - lrmax) : (warp == 1 ? -v : v);
                if (!np_isclose(v, first)) far = 1;
                tl[i] = v;
            }
            far = __any_sync(FULL, far);
            __syncwarp();
            double kh;
            if (n_tail < 5 || !far) {
                kh = (warp == 0) ? inf_f64() : -inf_f64();  // e_loo.py:353-354, :373-374, :379-380
            } else {
                const double cutoff = tl[n_tail - 1];
                __syncwarp();
                for (int i = lane; i < n_tail; i += 32) {
                    const double d = tl[i] - cutoff;
                    tl[i] = (warp == 1) ? -d : d;  // e_loo.py:357, :377, :383
                }
                __syncwarp();
                kh = gpdfit_literal_warp(tl, n_tail, gpd + warp * 128, gpd + warp * 128 + 64, lane);
            }
            if (lane == 0) res[warp] = kh;
        }
        __syncthreads();

        if (tid == 0) {
            // np.max propagates NaN: r is all-NaN, every fit returns NaN (e_loo.py:387-388).  A +inf maximum
            // leaves r = 0 on every finite draw: the r tail is all-close => +inf.
            double khat_r = r_ok ? res[0] : ((lrmax == inf_f64()) ? inf_f64() : nan_f64());
            double khat = khat_r;
            if (need_hr) {
                const double khat_hr = py_max(res[1], res[2]);
                khat = (khat_hr != khat_hr && khat_r != khat_r) ? nan_f64() : py_max(khat_hr, khat_r);
            }
            p.khat[row] = khat;
            if (has_x) {
                const double mean = sex / se;
                double val = mean;
                if (p.type != ELOO_MEAN) {
                    const double wss = see / (se * se);
                    if (x_close) val = 0.0;                       // e_loo.py:520-521
                    else if (np_isclose(wss, 1.0)) val = 0.0;     // e_loo.py:523-525
                    else {
                        const double var = (sexx / se - mean * mean) / (1.0 - wss);
                        val = (0.0 > var) ? 0.0 : var;            // max(var, 0.0)
                    }
                    if (p.type == ELOO_SD) val = sqrt(val);
                }
                p.value[row] = val;
            }
        }
        __syncthreads();
    }
}

// ===================================================================================== e_loo quantiles
// pyloo/e_loo.py:466-515 + _weighted_quantile :534-554 for a batch: per observation the draws are sorted
// (bitonic network in shared memory on (x, draw index), NaN last like np.argsort), the normalised weights are
// accumulated in sorted order and every requested probability is located by binary search in the cumulative
// weights and interpolated linearly.  Rows whose weights are all close to the first one take np.quantile's
// linear rule instead (:536-537).
// Sort key of a draw: the order-preserving 64-bit image of the double (b2l_common.cuh), NaN above every number
// (np.argsort puts NaN last) and the padding slots above NaN, so one unsigned compare orders a pair.
constexpr uint64_t QKEY_NAN = 0xfffffffffffffffeull, QKEY_PAD = 0xffffffffffffffffull;
__device__ __forceinline__ uint64_t quant_key(double x) { return (x != x) ? QKEY_NAN : key_of(x); }
__device__ __forceinline__ double quant_val(uint64_t k) { return (k >= QKEY_NAN) ? nan_f64() : val_of(k); }

__global__ void __launch_bounds__(IS_NT) eloo_quantile_kernel(const QuantParams p) {
    extern __shared__ __align__(16) unsigned char is_smem[];
    double* red = reinterpret_cast<double*>(is_smem);
    uint64_t* keys = reinterpret_cast<uint64_t*>(red + IS_RED_WORDS);  // [P2] sort keys, later the sorted draws
    double* part = reinterpret_cast<double*>(keys + p.P2);  // [IS_NW] scan partials
    double* ex = part + IS_NW;                              // [IS_NT] cumulative weight before each thread's chunk
    double* ce = ex + IS_NT;                                // [IS_NT] cumulative weight at the end of each chunk
    int* idx = reinterpret_cast<int*>(ce + IS_NT);          // [P2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, S = p.S, P2 = p.P2;

    for (long long row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
        const double* gx = p.x + row * p.x_stride;
        const double* glw = p.lw + row * p.lw_stride;
        // ---- normaliser of the weights (e_loo.py:557-559) and the sort input
        double lwmax = -inf_f64();
        int bad = 0;
        for (int s = tid; s < P2; s += IS_NT) {
            keys[s] = (s < S) ? quant_key(gx[s]) : QKEY_PAD;
            idx[s] = s;
            if (s < S) {
                const double a = glw[s];
                bad |= (a != a);
                lwmax = fmax(lwmax, a);
            }
        }
        lwmax = block_max<IS_NT>(lwmax, red);
        if (block_or(bad, red)) lwmax = nan_f64();
        double se = 0.0;
        for (int s = tid; s < S; s += IS_NT) se += exp(glw[s] - lwmax);
        se = block_sum<IS_NT>(se, red);
        const double lse = log(se) + lwmax;
        const double w0 = exp(glw[0] - lse);
        int far = 0;
        for (int s = tid; s < S; s += IS_NT) far |= np_isclose(exp(glw[s] - lse), w0) ? 0 : 1;
        const bool uniform = !block_or(far, red);  // np.allclose(w, w[0]) (:536)

        // ---- bitonic sort of (x, index), ascending, NaN last.  Exchanges at distance j <= 32 of the pairs a warp
        // owns in one sweep stay inside one block of 64 consecutive elements, so those sub-steps run back to back
        // with warp-level synchronisation only; block barriers are needed for j > 32 and between stages.
        auto exchange = [&](int t, int j, int k) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const int l = i | j;
            const uint64_t ka = keys[i], kb = keys[l];
            const int ia = idx[i], ib = idx[l];
            const bool up = ((i & k) == 0);
            const bool greater = (ka > kb) || (ka == kb && ia > ib);  // ties in x: draw-index order
            if (greater == up) {
                keys[i] = kb; keys[l] = ka;
                idx[i] = ib; idx[l] = ia;
            }
        };
        for (int k = 2; k <= P2; k <<= 1) {
            int j = k >> 1;
            for (; j > 32; j >>= 1) {
                for (int t = tid; t < (P2 >> 1); t += IS_NT) exchange(t, j, k);
                __syncthreads();
            }
            for (int t = tid; t < (P2 >> 1); t += IS_NT) {
                for (int jj = j; jj > 0; jj >>= 1) {
                    exchange(t, jj, k);
                    __syncwarp();
                }
            }
            __syncthreads();
        }

        double* xs = reinterpret_cast<double*>(keys);  // sorted draws, decoded in place
        for (int s = tid; s < P2; s += IS_NT) xs[s] = quant_val(keys[s]);
        __syncthreads();

        const int C = P2 / IS_NT;  // sorted positions per thread
        if (!uniform) {
            // ---- cumulative weights in sorted order.  The cumulative weight of position i is defined as
            // ex[c] + (running sum inside chunk c = i / C); only the per-chunk offsets ex[] and chunk-end values
            // ce[] are kept, the few positions around a crossing are re-accumulated when a quantile is located.
            double run = 0.0;
            for (int i = tid * C; i < (tid + 1) * C; ++i) {
                const int s = idx[i];
                run += (s < S) ? exp(glw[s] - lse) : 0.0;
            }
            double incl = run;  // inclusive scan of the chunk totals across the block
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) part[warp] = incl;
            __syncthreads();
            double base = 0.0;
            for (int w = 0; w < warp; ++w) base += part[w];
            const double excl = base + (incl - run);
            ex[tid] = excl;
            ce[tid] = excl + run;
            __syncthreads();
        }
        if (tid < p.n_probs) {
            const double q = p.probs[tid];
            double val;
            if (uniform) {
                // np.quantile(x, q), method "linear": virtual index n*q + (alpha + q*(1 - alpha - beta)) - 1 with
                // alpha = beta = 1, evaluated in that order; lerp as numpy's _lerp
                const double n = (double)S;
                const double virt = n * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
                double lo = floor(virt);
                const double g = virt - lo;
                int il = (int)lo, ih = il + 1;
                if (il < 0) il = 0;
                if (ih > S - 1) ih = S - 1;
                if (il > S - 1) il = S - 1;
                const double a = xs[il], b = xs[ih];
                const double d = b - a;
                val = a + d * g;
                if (g >= 0.5) val = b - d * (1.0 - g);
                if (d == 0.0) val = a;
                if (xs[S - 1] != xs[S - 1]) val = nan_f64();  // NaN in the data poisons np.quantile
            } else {
                const double total = ce[IS_NT - 1];
                // first sorted position whose normalised cumulative weight reaches q (:542-544): the chunk by
                // binary search over the chunk-end values, the position by re-accumulating that chunk
                int lo = 0, hi = IS_NT;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (ce[mid] / total >= q) hi = mid; else lo = mid + 1;
                }
                int pos = S;
                double c_prev = 0.0, c_cur = 0.0;
                if (lo < IS_NT) {
                    double run = 0.0;
                    c_prev = (lo > 0) ? ce[lo - 1] : 0.0;
                    for (int i = lo * C; i < (lo + 1) * C; ++i) {
                        const int s = idx[i];
                        run += (s < S) ? exp(glw[s] - lse) : 0.0;
                        c_cur = ex[lo] + run;
                        if (c_cur / total >= q) { pos = i; break; }
                        c_prev = c_cur;
                    }
                }
                if (pos >= S) val = xs[S - 1];      // :545-546
                else if (pos == 0) val = xs[0];     // :549-550
                else {
                    const double w1 = c_prev / total, w2 = c_cur / total;
                    const double x1 = xs[pos - 1];
                    val = x1 + (xs[pos] - x1) * (q - w1) / (w2 - w1);  // :552-554
                }
            }
            p.out[row * p.n_probs + tid] = val;
        }
        __syncthreads();
    }
}

// ===================================================================================== group sums (LOGO)
// One warp per (group, draw).  Observation-fastest input (ArviZ layout): lanes stride over the members, which
// is coalesced when a group's observations are adjacent; row input: lanes take adjacent draws of one group and
// loop over the members.  Partial sums are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256) group_sum_kernel(const GroupSumParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned c_nan = 0;
    if (p.stride_n == 1 && p.stride_s != 1) {
        const long long tasks = (long long)p.G * p.S;
        for (long long t = warp_global; t < tasks; t += n_warps) {
            const int g = (int)(t / p.S), s = (int)(t % p.S);
            const int b = p.offsets[g], e = p.offsets[g + 1];
            const double* col = p.ll + (long long)s * p.stride_s;
            double acc = 0.0;
            for (int m = b + lane; m < e; m += 32) {
                double v = col[p.members[m]];
                if (v != v) { v = -1e10; ++c_nan; }
                acc += v;
            }
            acc = warp_sum(acc);
            if (lane == 0) p.out[(long long)g * p.out_stride + s] = acc;
        }
    } else {
        const int tiles = (p.S + 31) / 32;
        const long long tasks = (long long)p.G * tiles;
        for (long long t = warp_global; t < tasks; t += n_warps) {
            const int g = (int)(t / tiles), s = (int)(t % tiles) * 32 + lane;
            if (s >= p.S) continue;
            const int b = p.offsets[g], e = p.offsets[g + 1];
            double acc = 0.0;
            for (int m = b; m < e; ++m) {
                double v = p.ll[(long long)s * p.stride_s + (long long)p.members[m] * p.stride_n];
                if (v != v) { v = -1e10; ++c_nan; }
                acc += v;
            }
            p.out[(long long)g * p.out_stride + s] = acc;
        }
    }
    if (p.counters && c_nan) atomicAdd(&p.counters[0], (unsigned long long)c_nan);
}

// ===================================================================================== WAIC, column form
// lppd_i = logsumexp_s(ll) - log S and var_s(ll) (ddof 0) for every observation of an observation-fastest
// (S, N) matrix in ONE pass over it (pyloo/waic.py:137-145).  A block is 32 observations (lanes, coalesced
// 256 B per draw) x 8 segments of the draw axis (warps).  Each thread runs an online logsumexp (running
// maximum, rescaled sum; chunks of 8 draws so the exps are independent) and a chunked Chan / Welford update of
// (mean, M2) whose shift is the running mean, so there is no E[x^2] - E[x]^2 cancellation; the 8 segments are
// merged in a fixed order.  NaN -> -1e10 and +-inf -> +-1e10 before the sums (waic.py:113-132).  The loo-policy
// lppd_i (infinities kept, pyloo/loo.py:329-337) equals the WAIC one unless the column holds +-inf; those rare
// columns are redone exactly by one lane with two plain passes.
constexpr int WC_SEG = 8, WC_CHUNK = 8;

struct WcAcc {
    double m, s;         // online logsumexp: running maximum, sum of exp(x - m)
    double n, mean, m2;  // Chan / Welford
};
__device__ __forceinline__ void wc_merge(WcAcc& a, const WcAcc& b, const ExpTab& tb) {
    if (b.n == 0.0) return;
    if (a.n == 0.0) { a = b; return; }
    const double mm = fmax(a.m, b.m);
    a.s = a.s * exp_sum(a.m - mm, tb) + b.s * exp_sum(b.m - mm, tb);
    a.m = mm;
    const double nt = a.n + b.n, d = b.mean - a.mean;
    a.mean += d * (b.n / nt);
    a.m2 += b.m2 + d * d * (a.n * b.n / nt);
    a.n = nt;
}

__global__ void __launch_bounds__(32 * WC_SEG) waic_cols_kernel(const WaicColsParams p) {
    __shared__ double tabs[64];
    __shared__ WcAcc part[WC_SEG][32];
    __shared__ int infs[WC_SEG][32];
    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5, S = p.S;
    if (threadIdx.x < 32) {
        tabs[threadIdx.x] = exp2((double)threadIdx.x / 32.0);
        tabs[32 + threadIdx.x] = exp2(-(double)threadIdx.x / 32.0);
    }
    __syncthreads();
    ExpTab tb;
    tb.t = tabs;
    tb.tinv = tabs + 32;
    const long long obs = (long long)blockIdx.x * 32 + lane;
    const bool live = obs < p.N;
    const int per = (S + WC_SEG - 1) / WC_SEG;
    const int s0 = seg * per, s1 = min(S, s0 + per);
    unsigned c_nan = 0, c_pinf = 0, c_ninf = 0;
    WcAcc a = {-inf_f64(), 0.0, 0.0, 0.0, 0.0};
    if (live) {
        const double* col = p.ll + obs;
        // one chunk of nb <= 8 sanitised draws folded into the accumulators
        auto fold = [&](const double (&w)[WC_CHUNK], int nb) {
            double cm = w[0];  // compare-select: the draws are NaN-free after sanitise (an IEEE fmax costs twice as much)
#pragma unroll
            for (int i = 1; i < WC_CHUNK; ++i)
                if (i < nb) cm = (w[i] > cm) ? w[i] : cm;
            if (cm > a.m) {  // new running maximum: rescale the sum (exp_sum(-inf) = 0 on the first chunk)
                a.s *= exp_sum(a.m - cm, tb);
                a.m = cm;
            }
            const double c = (a.n == 0.0) ? w[0] : a.mean;  // shift = running mean
            double es = 0.0, sd = 0.0, sdd = 0.0;
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i) {
                if (i < nb) {
                    es += exp_sum(w[i] - a.m, tb);
                    const double d = w[i] - c;
                    sd += d;
                    sdd += d * d;
                }
            }
            a.s += es;
            const double nbd = (double)nb, inv = 1.0 / nbd;
            const double mean_b = c + sd * inv, m2_b = sdd - sd * sd * inv;
            if (a.n == 0.0) {
                a.n = nbd; a.mean = mean_b; a.m2 = m2_b;
            } else {
                const double nt = a.n + nbd, f = nbd / nt, d = mean_b - a.mean;
                a.mean += d * f;
                a.m2 += m2_b + d * d * (a.n * f);
                a.n = nt;
            }
        };
        auto sanitise = [&](double (&w)[WC_CHUNK], int nb) {
            bool fin = true;
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i) fin = fin && (i >= nb || is_finite(w[i]));
            if (fin) return;
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i) {
                if (i >= nb) continue;
                const double v = w[i];
                if (v != v) { w[i] = -1e10; ++c_nan; }
                else if (v == inf_f64()) { w[i] = 1e10; ++c_pinf; }
                else if (v == -inf_f64()) { w[i] = -1e10; ++c_ninf; }
            }
        };
        int s = s0;
        double nx[WC_CHUNK];
        const double* pn = col + (long long)s0 * p.stride_s;  // next draw to load (pointer stepping, no 64-bit multiplies)
        if (s + WC_CHUNK <= s1) {
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i, pn += p.stride_s) nx[i] = *pn;
        }
        for (; s + WC_CHUNK <= s1; s += WC_CHUNK) {
            double w[WC_CHUNK];
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i) w[i] = nx[i];
            if (s + 2 * WC_CHUNK <= s1) {  // the next chunk's loads are in flight while this one is folded
#pragma unroll
                for (int i = 0; i < WC_CHUNK; ++i, pn += p.stride_s) nx[i] = *pn;
            }
            sanitise(w, WC_CHUNK);
            fold(w, WC_CHUNK);
        }
        if (s < s1) {
            const int nb = s1 - s;
            double w[WC_CHUNK];
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i) w[i] = (i < nb) ? pn[(long long)i * p.stride_s] : 0.0;
            sanitise(w, nb);
            fold(w, nb);
        }
    }
    part[seg][lane] = a;
    infs[seg][lane] = (int)(c_pinf + c_ninf);
    __syncthreads();
    if (seg == 0 && live) {
        WcAcc t = part[0][lane];
        int n_inf = infs[0][lane];
        for (int g = 1; g < WC_SEG; ++g) {
            wc_merge(t, part[g][lane], tb);
            n_inf += infs[g][lane];
        }
        const double lppdw = log(t.s) + (t.m - p.log_S);  // utils.py:352-357 with b_inv = S
        double lppd = lppdw;
        if (n_inf > 0) {
            // loo policy keeps the infinities: plain two-pass logsumexp of the NaN-replaced column
            const double* col = p.ll + obs;
            double mx = -inf_f64();
            for (int s = 0; s < S; ++s) {
                double v = col[(long long)s * p.stride_s];
                if (v != v) v = -1e10;
                mx = fmax(mx, v);
            }
            double sum = 0.0;
            for (int s = 0; s < S; ++s) {
                double v = col[(long long)s * p.stride_s];
                if (v != v) v = -1e10;
                sum += exp(v - mx);
            }
            lppd = log(sum) + (mx - p.log_S);
        }
        p.lppdw_i[obs] = lppdw;
        p.lppd_i[obs] = lppd;
        p.var_i[obs] = t.m2 / (double)S;
        p.k_i[obs] = inf_f64();    // no PSIS stage: same values as the general kernel writes in this mode
        p.elpd_i[obs] = nan_f64();
    }
    if (p.counters) {
        if (c_nan) atomicAdd(&p.counters[0], (unsigned long long)c_nan);
        if (c_pinf) atomicAdd(&p.counters[1], (unsigned long long)c_pinf);
        if (c_ninf) atomicAdd(&p.counters[2], (unsigned long long)c_ninf);
    }
}

// ===================================================================================== SIS / TIS loo, column form
// loo(method = "sis" | "tis") on the observation-fastest matrix without transposed panels: the same block
// shape as waic_cols_kernel (32 observations x 8 draw segments), one sweep of the matrix per logsumexp that
// needs the previous one's result -- SIS: (max, sum, sum of squares of -ll; logsumexp of ll) then
// logsumexp(lw + ll); TIS: one more sweep for the truncated sum.  Every sweep is an online logsumexp in chunks of
// 8 draws with the next chunk's loads in flight.  Columns holding +-inf are redone by one lane with the plain
// multi-pass arithmetic of is_row_kernel, so the IEEE special cases come out as NumPy's.
template <class F>
__device__ __forceinline__ void col_sweep(const double* col, long long stride_s, int s0, int s1, F&& fold) {
    int s = s0;
    double nx[WC_CHUNK];
    const double* pn = col + (long long)s0 * stride_s;
    if (s + WC_CHUNK <= s1) {
#pragma unroll
        for (int i = 0; i < WC_CHUNK; ++i, pn += stride_s) nx[i] = *pn;
    }
    for (; s + WC_CHUNK <= s1; s += WC_CHUNK) {
        double w[WC_CHUNK];
#pragma unroll
        for (int i = 0; i < WC_CHUNK; ++i) w[i] = nx[i];
        if (s + 2 * WC_CHUNK <= s1) {
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i, pn += stride_s) nx[i] = *pn;
        }
        fold(w, WC_CHUNK);
    }
    if (s < s1) {
        const int nb = s1 - s;
        double w[WC_CHUNK];
#pragma unroll
        for (int i = 0; i < WC_CHUNK; ++i) w[i] = (i < nb) ? pn[(long long)i * stride_s] : 0.0;
        fold(w, nb);
    }
}

struct LseAcc {
    double m, s, q;  // running maximum, sum of exp(x - m), sum of exp(x - m)^2
};
__device__ __forceinline__ void lse_fold(LseAcc& a, const double (&x)[WC_CHUNK], int nb, const ExpTab& tb) {
    double cm = x[0];
#pragma unroll
    for (int i = 1; i < WC_CHUNK; ++i)
        if (i < nb) cm = (x[i] > cm) ? x[i] : cm;  // NaN never wins: such columns are redone serially
    if (cm > a.m) {
        const double f = exp_sum(a.m - cm, tb);
        a.s *= f;
        a.q *= f * f;
        a.m = cm;
    }
    double es = 0.0, eq = 0.0;
#pragma unroll
    for (int i = 0; i < WC_CHUNK; ++i) {
        if (i < nb) {
            const double e = exp_sum(x[i] - a.m, tb);
            es += e;
            eq += e * e;
        }
    }
    a.s += es;
    a.q += eq;
}
__device__ __forceinline__ void lse_merge(LseAcc& a, const LseAcc& b, const ExpTab& tb) {
    const double mm = fmax(a.m, b.m);
    const double fa = exp_sum(a.m - mm, tb), fb = exp_sum(b.m - mm, tb);
    a.s = a.s * fa + b.s * fb;
    a.q = a.q * fa * fa + b.q * fb * fb;
    a.m = mm;
}

// plain serial evaluation for one column (rare: +-inf present), same arithmetic as is_row_kernel in loo mode
template <int METHOD>
__device__ __noinline__ void is_loo_serial(const double* col, long long stride_s, int S, double log_S, double& elpd,
                                           double& ess, double& lppd) {
    auto ll = [&](int s) {
        const double v = col[(long long)s * stride_s];
        return (v != v) ? -1e10 : v;
    };
    double mx = -inf_f64(), mn = inf_f64();
    for (int s = 0; s < S; ++s) {
        const double v = -ll(s);
        mx = fmax(mx, v);
        mn = fmin(mn, v);
    }
    double s1 = 0.0, s2 = 0.0, sl = 0.0;
    for (int s = 0; s < S; ++s) {
        const double v = -ll(s), e = exp(v - mx);
        s1 += e;
        s2 += e * e;
        sl += exp(mn - v);
    }
    double lse, cut = inf_f64();
    if (METHOD == IS_METHOD_SIS) {
        lse = log(s1);
        ess = (s1 * s1) / s2;
    } else {
        cut = (log(s1) - log_S) + 0.5 * log_S;
        const double mx2 = np_minimum(0.0, cut);
        double t1 = 0.0, t2 = 0.0;
        for (int s = 0; s < S; ++s) {
            const double a = exp(np_minimum(-ll(s) - mx, cut) - mx2);
            t1 += a;
            t2 += a * a;
        }
        lse = log(t1) + mx2;
        ess = (t1 * t1) / t2;
    }
    double tmax = -inf_f64();
    for (int s = 0; s < S; ++s) {
        const double v = -ll(s);
        double x = v - mx;
        if (METHOD == IS_METHOD_TIS) x = np_minimum(x, cut);
        tmax = fmax(tmax, (x - lse) + (-v));
    }
    double st = 0.0;
    for (int s = 0; s < S; ++s) {
        const double v = -ll(s);
        double x = v - mx;
        if (METHOD == IS_METHOD_TIS) x = np_minimum(x, cut);
        st += exp(((x - lse) + (-v)) - tmax);
    }
    elpd = log(st) + tmax;
    lppd = log(sl) + ((-mn) - log_S);
}

template <int METHOD>
__global__ void __launch_bounds__(32 * WC_SEG) is_cols_kernel(const IsColsParams p) {
    __shared__ double tabs[64];
    __shared__ LseAcc part[2][WC_SEG][32];
    __shared__ int infs[WC_SEG][32];
    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5, S = p.S;
    if (threadIdx.x < 32) {
        tabs[threadIdx.x] = exp2((double)threadIdx.x / 32.0);
        tabs[32 + threadIdx.x] = exp2(-(double)threadIdx.x / 32.0);
    }
    __syncthreads();
    ExpTab tb;
    tb.t = tabs;
    tb.tinv = tabs + 32;
    const long long obs = (long long)blockIdx.x * 32 + lane;
    const bool live = obs < p.N;
    const double* col = p.ll + (live ? obs : 0);
    const int per = (S + WC_SEG - 1) / WC_SEG;
    const int s0 = min(S, seg * per), s1 = min(S, s0 + per);
    const LseAcc zero = {-inf_f64(), 0.0, 0.0};
    unsigned c_nan = 0, c_pinf = 0, c_ninf = 0;

    // ---- sweep 1: logsumexp (+ squares) of -ll and logsumexp of ll
    LseAcc av = zero, al = zero;
    if (live) {
        col_sweep(col, p.stride_s, s0, s1, [&](double (&w)[WC_CHUNK], int nb) {
            double v[WC_CHUNK];
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i) {
                if (i < nb) {
                    if (!is_finite(w[i])) {
                        if (w[i] != w[i]) { w[i] = -1e10; ++c_nan; }  // pyloo/loo.py:227
                        else if (w[i] > 0) ++c_pinf;
                        else ++c_ninf;
                    }
                }
                v[i] = -w[i];
            }
            lse_fold(av, v, nb, tb);
            lse_fold(al, w, nb, tb);
        });
    }
    part[0][seg][lane] = av;
    part[1][seg][lane] = al;
    infs[seg][lane] = (int)(c_pinf + c_ninf);
    __syncthreads();
    bool special = false;
    double ess = 0.0, lppd = 0.0, lse = 0.0, cut = inf_f64(), mx = 0.0, mx2 = 0.0;
    {
        // every segment folds the 8 partials in the same order: identical values in all of them
        LseAcc tv = part[0][0][lane], tl = part[1][0][lane];
        int n_inf = infs[0][lane];
        for (int g = 1; g < WC_SEG; ++g) {
            lse_merge(tv, part[0][g][lane], tb);
            lse_merge(tl, part[1][g][lane], tb);
            n_inf += infs[g][lane];
        }
        special = n_inf > 0;
        mx = tv.m;
        lppd = log(tl.s) + (tl.m - p.log_S);
        if (METHOD == IS_METHOD_SIS) {
            lse = log(tv.s);
            ess = (tv.s * tv.s) / tv.q;
        } else {
            cut = (log(tv.s) - p.log_S) + 0.5 * p.log_S;  // tis.py:112-114
            mx2 = fmin(0.0, cut);
        }
    }
    __syncthreads();

    // ---- TIS sweep 2: sums over the truncated row
    if (METHOD == IS_METHOD_TIS) {
        double t1 = 0.0, t2 = 0.0;
        if (live && !special) {
            col_sweep(col, p.stride_s, s0, s1, [&](double (&w)[WC_CHUNK], int nb) {
#pragma unroll
                for (int i = 0; i < WC_CHUNK; ++i) {
                    if (i < nb) {
                        const double llv = (w[i] != w[i]) ? -1e10 : w[i];
                        const double a = exp_sum(fmin(-llv - mx, cut) - mx2, tb);
                        t1 += a;
                        t2 += a * a;
                    }
                }
            });
        }
        part[0][seg][lane].s = t1;
        part[0][seg][lane].q = t2;
        __syncthreads();
        double a1 = part[0][0][lane].s, a2 = part[0][0][lane].q;
        for (int g = 1; g < WC_SEG; ++g) {
            a1 += part[0][g][lane].s;
            a2 += part[0][g][lane].q;
        }
        lse = log(a1) + mx2;
        ess = (a1 * a1) / a2;
        __syncthreads();
    }

    // ---- last sweep: elpd_i = logsumexp(lw + ll)  (loo.py:289, :319-324)
    LseAcc at = zero;
    if (live && !special) {
        col_sweep(col, p.stride_s, s0, s1, [&](double (&w)[WC_CHUNK], int nb) {
            double t[WC_CHUNK];
#pragma unroll
            for (int i = 0; i < WC_CHUNK; ++i) {
                const double llv = (w[i] != w[i]) ? -1e10 : w[i];
                const double v = -llv;
                double x = v - mx;
                if (METHOD == IS_METHOD_TIS) x = fmin(x, cut);
                t[i] = (x - lse) + (-v);
            }
            lse_fold(at, t, nb, tb);
        });
    }
    part[0][seg][lane] = at;
    __syncthreads();
    if (seg == 0 && live) {
        double elpd;
        if (special) {
            is_loo_serial<METHOD>(col, p.stride_s, S, p.log_S, elpd, ess, lppd);
        } else {
            LseAcc tt = part[0][0][lane];
            for (int g = 1; g < WC_SEG; ++g) lse_merge(tt, part[0][g][lane], tb);
            elpd = log(tt.s) + tt.m;
        }
        p.elpd[obs] = elpd;
        p.ess[obs] = ess;
        p.lppd[obs] = lppd;
    }
    if (p.counters) {
        if (c_nan) atomicAdd(&p.counters[0], (unsigned long long)c_nan);
        if (c_pinf) atomicAdd(&p.counters[1], (unsigned long long)c_pinf);
        if (c_ninf) atomicAdd(&p.counters[2], (unsigned long long)c_ninf);
    }
}

// ===================================================================================== host side
static size_t is_smem_bytes(int S) { return sizeof(double) * (IS_RED_WORDS + (size_t)((S + 1) & ~1)); }
constexpr size_t SMEM_LIMIT = 227 * 1024;

template <typename K>
static cudaError_t plan_kernel(K kern, size_t smem, long long n_rows, int* info) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0, dev = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, IS_NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidConfiguration;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long cap = (long long)sms * occ;
    info[1] = (int)(n_rows < cap ? (n_rows > 0 ? n_rows : 1) : cap);
    info[2] = (int)smem;
    info[3] = occ;
    return cudaSuccess;
}

constexpr int TIS_EPT = 16;
template <int METHOD, int MODE>
static cudaError_t is_plan_t(int S, long long n_rows, int* info) {
    const size_t staged = is_smem_bytes(S);
    if (METHOD == IS_METHOD_TIS && S <= IS_NT * TIS_EPT) {
        info[0] = 2;
        return plan_kernel(is_row_kernel<METHOD, MODE, true, (METHOD == IS_METHOD_TIS ? TIS_EPT : 0)>, staged, n_rows,
                           info);
    }
    if (staged <= SMEM_LIMIT) {
        info[0] = 1;
        return plan_kernel(is_row_kernel<METHOD, MODE, true>, staged, n_rows, info);
    }
    info[0] = 0;
    return plan_kernel(is_row_kernel<METHOD, MODE, false>, sizeof(double) * IS_RED_WORDS, n_rows, info);
}

cudaError_t is_plan(int method, int mode, int S, long long n_rows, int* info) {
    if (method == IS_METHOD_SIS)
        return mode == IS_MODE_WEIGHTS ? is_plan_t<IS_METHOD_SIS, IS_MODE_WEIGHTS>(S, n_rows, info)
                                       : is_plan_t<IS_METHOD_SIS, IS_MODE_LOO>(S, n_rows, info);
    return mode == IS_MODE_WEIGHTS ? is_plan_t<IS_METHOD_TIS, IS_MODE_WEIGHTS>(S, n_rows, info)
                                   : is_plan_t<IS_METHOD_TIS, IS_MODE_LOO>(S, n_rows, info);
}

template <int METHOD, int MODE>
static cudaError_t is_launch_t(const IsParams& p, cudaStream_t st) {
    int info[4];
    cudaError_t e = is_plan_t<METHOD, MODE>(p.S, p.n_rows, info);
    if (e != cudaSuccess) return e;
    if (info[0] == 2)
        is_row_kernel<METHOD, MODE, true, (METHOD == IS_METHOD_TIS ? TIS_EPT : 0)><<<info[1], IS_NT, info[2], st>>>(p);
    else if (info[0]) is_row_kernel<METHOD, MODE, true><<<info[1], IS_NT, info[2], st>>>(p);
    else is_row_kernel<METHOD, MODE, false><<<info[1], IS_NT, info[2], st>>>(p);
    return cudaGetLastError();
}

cudaError_t is_launch(int method, int mode, const IsParams& p, cudaStream_t st) {
    if (method == IS_METHOD_SIS)
        return mode == IS_MODE_WEIGHTS ? is_launch_t<IS_METHOD_SIS, IS_MODE_WEIGHTS>(p, st)
                                       : is_launch_t<IS_METHOD_SIS, IS_MODE_LOO>(p, st);
    return mode == IS_MODE_WEIGHTS ? is_launch_t<IS_METHOD_TIS, IS_MODE_WEIGHTS>(p, st)
                                   : is_launch_t<IS_METHOD_TIS, IS_MODE_LOO>(p, st);
}

cudaError_t eloo_plan(int S, long long n_rows, bool has_x, int tail_len, int* info) {
    const int n_staged = 1 + (has_x ? 1 : 0);
    const size_t staged = ElooSmem::bytes(S, n_staged, tail_len);
    if (staged <= SMEM_LIMIT) {
        info[0] = 1;
        return plan_kernel(eloo_row_kernel<true>, staged, n_rows, info);
    }
    info[0] = 0;
    return plan_kernel(eloo_row_kernel<false>, ElooSmem::bytes(0, 0, tail_len), n_rows, info);
}

cudaError_t eloo_launch(const ElooParams& p, cudaStream_t st) {
    int info[4];
    const bool has_x = (p.x != nullptr) && (p.type != ELOO_NONE);
    cudaError_t e = eloo_plan(p.S, p.n_rows, has_x, p.tail_len, info);
    if (e != cudaSuccess) return e;
    if (info[0]) eloo_row_kernel<true><<<info[1], IS_NT, info[2], st>>>(p);
    else {
        const int grid = (p.grid_cap > 0 && p.grid_cap < info[1]) ? p.grid_cap : info[1];
        eloo_row_kernel<false><<<grid, IS_NT, info[2], st>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t eloo_quantile_launch(QuantParams p, cudaStream_t st) {
    int p2 = IS_NT;
    while (p2 < p.S) p2 <<= 1;
    p.P2 = p2;
    const size_t smem = sizeof(double) * (IS_RED_WORDS + (size_t)p2 + IS_NW + 2 * IS_NT) + sizeof(int) * (size_t)p2;
    int info[4];
    cudaError_t e = plan_kernel(eloo_quantile_kernel, smem, p.n_rows, info);
    if (e != cudaSuccess) return e;
    eloo_quantile_kernel<<<info[1], IS_NT, info[2], st>>>(p);
    return cudaGetLastError();
}

cudaError_t group_sum_launch(const GroupSumParams& p, cudaStream_t st) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    group_sum_kernel<<<sms * 8, 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t waic_cols_launch(const WaicColsParams& p, cudaStream_t st) {
    const long long blocks = (p.N + 31) / 32;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    waic_cols_kernel<<<(unsigned)blocks, 32 * WC_SEG, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t is_cols_launch(int method, const IsColsParams& p, cudaStream_t st) {
    const long long blocks = (p.N + 31) / 32;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    if (method == IS_METHOD_SIS) is_cols_kernel<IS_METHOD_SIS><<<(unsigned)blocks, 32 * WC_SEG, 0, st>>>(p);
    else is_cols_kernel<IS_METHOD_TIS><<<(unsigned)blocks, 32 * WC_SEG, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace b2l
