// Split PSIS path (sm_100a): a streaming kernel and a tail kernel per observation batch.
//
//   psis_stream_kernel  one CTA per observation.  The S draws are prefetched by 1-D bulk TMA into
//                       shared memory and then held in REGISTERS (EPT per thread), so the row is
//                       read once: max (pyloo/psis.py:134), a threshold below the (M+1)-th largest
//                       draw estimated from the per-thread maxima, then one pass that forms
//                       x = fl(r - max r), sums exp(x) over the body (pyloo/utils.py:349-351) and
//                       emits the ~1.2 (M+1) candidate draws above the threshold as packed 64-bit
//                       keys.  LOO mode adds the lppd / variance sums (pyloo/loo.py:329-337,
//                       pyloo/waic.py:137-145).
//   psis_tail_kernel    one WARP per observation, no block barriers: register bitonic sort of the
//                       candidate keys, exact cutoff / tail (psis.py:135-141), Zhang-Stephens GPD
//                       fit (psis.py:181-208), _gpinv smoothing (psis.py:149-157,211-222),
//                       normaliser (psis.py:158), then either the normalised row with the smoothed
//                       tail patched in (psislw) or elpd_i / lppd_i / var_i (loo.py:319-337).
//
// Rows the fast path cannot decide exactly (NaN / inf, ties or key collisions at the cutoff,
// candidate overflow, non-finite GPD profiles, cutoff below log(DBL_MIN)) are appended to a list
// and re-done by the general row kernel (b2l_row_kernel.cuh), which handles every case.
#pragma once

#include "b2l_row_kernel.cuh"

namespace b2l {

constexpr int KEY_IDX_BITS = 14;               // draw index lives in the low bits of a candidate key
constexpr int SPLIT_MAX_S = 1 << KEY_IDX_BITS;
constexpr unsigned KEY_IDX_MASK = (1u << KEY_IDX_BITS) - 1u;

struct __align__(16) SplitHeader {  // 64 B per observation: stream kernel -> tail kernel
    double mx;      // max_s r_s
    double body;    // sum of exp(x_s) over the draws that are NOT candidates
    double lsum;    // LOO: sum_s exp(ll_s - lshift)
    double vsum;    // LOO: sum_s (ll_s - mean ll)^2
    double lshift;  // LOO: shift used for lsum (min ll, or max ll for very wide rows)
    double ll_max;
    int C;          // candidates emitted (M + 1 <= C <= cap)
    int flags;      // != 0: the row was handed to the general kernel
    int attempts;
    int pad;
};

struct SplitParams {
    const double* in;      // row i = in + i * in_stride (PSISLW: r = log weight; LOO: ll, r = -ll)
    long long in_stride;
    double* out;           // PSISLW rows
    long long out_stride;
    double* k_out;
    double* elpd_i;
    double* lppd_i;
    double* var_i;
    double* lppdw_i;
    double* diag;
    long long n_rows;
    int S, M, cap;
    int q0;                // per-warp rank (1..32) of the thread maxima used for the threshold guess
    int m_full;            // 30 + floor(sqrt(M))
    double cutoffmin;
    SplitHeader* hdr;          // [n_rows]
    unsigned long long* ckey;  // [n_rows][cap]
    int* fb_list;              // [n_rows] rows for the general kernel
    int* fb_count;
    unsigned long long* counters;  // optional [4]; [3] += rows handed to the general kernel
};

// ------------------------------------------------------------------ table-driven exp for the sums
// exp(x) = 2^n * 2^(j/32) * e^r, |r| <= ln2/64, degree-5 polynomial: relative error < 4e-15 for
// x >= -36 (one-step reduction; the error grows to 8e-14 at x = -700, where the term is negligible
// next to the row maximum's exp(0) = 1).  Used ONLY for the normalising sums; tail values use exp().
struct ExpTab {
    const double* t;     // 2^(j/32)
    const double* tinv;  // 2^(-j/32)
};
constexpr double EXP_L = 46.166241308446828;       // 32 / ln 2
constexpr double EXP_C1 = -0.021660849392498290;   // -ln 2 / 32
constexpr double EXP_MAGIC = 6755399441055744.0;   // 1.5 * 2^52

__device__ __forceinline__ double exp_poly5(double r) {
    double p = 8.3333333333333333e-03;
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p;
}
__device__ __forceinline__ double scale2(double y, int n) {  // y * 2^n, result normal
    return __hiloint2double(__double2hiint(y) + (n << 20), __double2loint(y));
}
// x in [-700, 0]
__device__ __forceinline__ double exp_tab(double x, const ExpTab& tb) {
    const double t = fma(x, EXP_L, EXP_MAGIC);
    const int ni = __double2loint(t);
    const double nf = t - EXP_MAGIC;
    const double r = fma(nf, EXP_C1, x);
    return scale2(tb.t[ni & 31] * exp_poly5(r), ni >> 5);
}
// both exp(x) and exp(-x) from one reduction, x in [-600, 0]
__device__ __forceinline__ void exp_tab_pm(double x, const ExpTab& tb, double& ep, double& em) {
    const double t = fma(x, EXP_L, EXP_MAGIC);
    const int ni = __double2loint(t);
    const double nf = t - EXP_MAGIC;
    const double r = fma(nf, EXP_C1, x);
    const int j = ni & 31, n = ni >> 5;
    ep = scale2(tb.t[j] * exp_poly5(r), n);
    em = scale2(tb.tinv[j] * exp_poly5(-r), -n);
}

__device__ __forceinline__ float warp_sort32_f(float v, int lane) {  // ascending across lanes
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(FULL, v, j);
            const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
            v = keep_min ? fminf(v, o) : fmaxf(v, o);
        }
    }
    return v;
}

struct StreamSmem {
    size_t row_bytes, off_tab, off_red, off_ctl, off_bar, total;
};
__host__ __device__ inline StreamSmem stream_smem(int S) {
    StreamSmem L;
    L.row_bytes = align_up((size_t)S * 8, 128);
    size_t o = L.row_bytes;
    L.off_tab = o;   // 64 doubles
    o += 64 * 8;
    L.off_red = o;   // 2 x (5 x 32 doubles + 32 floats): sets alternate between consecutive rows
    o += 2 * (5 * 32 * 8 + 32 * 4);
    L.off_ctl = o;
    o += 16 * 4;
    L.off_bar = o;
    o += 16;
    L.total = align_up(o, 128);
    return L;
}

template <int NW>
__device__ __forceinline__ double slots_max(const double* s, int lane) {
    if (NW <= 8) {
        double r = s[0];
#pragma unroll
        for (int i = 1; i < NW; ++i) r = fmax(r, s[i]);
        return r;
    }
    return warp_max(lane < NW ? s[lane] : -inf_f64());
}
template <int NW>
__device__ __forceinline__ double slots_sum(const double* s, int lane) {
    if (NW <= 8) {
        double r = s[0];
#pragma unroll
        for (int i = 1; i < NW; ++i) r += s[i];
        return r;
    }
    return warp_sum(lane < NW ? s[lane] : 0.0);
}

template <int NT>
constexpr int stream_min_blocks() { return NT == 128 ? 6 : (NT == 256 ? 3 : 1); }

// ------------------------------------------------------------------ stream kernel
template <int NT, int EPT, int MODE>
__global__ void __launch_bounds__(NT, stream_min_blocks<NT>()) psis_stream_kernel(const SplitParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    constexpr int EP2 = EPT / 2;
    const int S = p.S, M = p.M, cap = p.cap;
    const int S2 = S >> 1;
    const StreamSmem L = stream_smem(S);
    const double2* rowbuf = reinterpret_cast<const double2*>(smem_raw);
    double* tab = reinterpret_cast<double*>(smem_raw + L.off_tab);
    int* ctl_all = reinterpret_cast<int*>(smem_raw + L.off_ctl);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t row_tx = (uint32_t)S * 8u;
    const double NEG_INF = -inf_f64();
    ExpTab tb;
    tb.t = tab;
    tb.tinv = tab + 32;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (tid < 32) {
        tab[tid] = exp2((double)tid / 32.0);
        tab[32 + tid] = exp2(-(double)tid / 32.0);
    }
    __syncthreads();
    long long row = blockIdx.x;
    if (tid == 0 && row < p.n_rows) {
        mbar_expect_tx(bar, row_tx);
        bulk_g2s(smem_raw, p.in + row * p.in_stride, row_tx, bar);
    }
    int nv = 0;  // valid double2 slots of this thread
#pragma unroll
    for (int j = 0; j < EP2; ++j) nv += (j * NT + tid < S2) ? 1 : 0;

    for (int it = 0; row < p.n_rows; row += gridDim.x, ++it) {
        // reduction slots / candidate counter alternate between consecutive rows, so a warp that
        // runs ahead into the next row never overwrites what a slower warp still has to read
        double* red = reinterpret_cast<double*>(smem_raw + L.off_red + (size_t)(it & 1) * (5 * 32 * 8 + 32 * 4));
        float* redf = reinterpret_cast<float*>(red + 5 * 32);
        int* ctl = ctl_all + (it & 1);
        // ---------------- row -> registers
        mbar_wait(bar, (uint32_t)(it & 1));
        double2 v[EP2];
#pragma unroll
        for (int j = 0; j < EP2; ++j) {
            if (j < nv) {
                v[j] = rowbuf[j * NT + tid];
                if (MODE == MODE_LOO) {  // r = -ll (pyloo/loo.py:286-288)
                    v[j].x = -v[j].x;
                    v[j].y = -v[j].y;
                }
            } else {
                v[j].x = NEG_INF;
                v[j].y = NEG_INF;
            }
        }
        // ---------------- pass A: max r (+ min r, sum r in LOO mode), NaN / inf flag
        double m0 = NEG_INF, n0 = inf_f64(), s0 = 0.0;
        int spec = 0;
#pragma unroll
        for (int j = 0; j < EP2; ++j) {
            m0 = fmax(m0, fmax(v[j].x, v[j].y));  // NaN-free rows only matter; flagged rows leave
            if (j < nv) {
                spec = max(spec, max(__double2hiint(v[j].x) & 0x7fffffff, __double2hiint(v[j].y) & 0x7fffffff));
                if (MODE == MODE_LOO) {
                    n0 = fmin(n0, fmin(v[j].x, v[j].y));
                    s0 += v[j].x + v[j].y;
                }
            }
        }
        const double tmax = m0;  // this thread's maximum: one "bin" of the threshold estimate
        m0 = warp_max(m0);
        if (MODE == MODE_LOO) {
            n0 = warp_min(n0);
            s0 = warp_sum(s0);
        }
        if (lane == 0) {
            red[wid] = m0;
            if (MODE == MODE_LOO) {
                red[32 + wid] = -n0;
                red[64 + wid] = s0;
            }
        }
        if (tid == 0) ctl[0] = 0;
        // (1) the row is in registers everywhere: the buffer can be refilled
        const bool special = __syncthreads_or(spec >= 0x7ff00000) != 0;
        if (tid == 0 && row + gridDim.x < p.n_rows) {
            fence_proxy_async();
            mbar_expect_tx(bar, row_tx);
            bulk_g2s(smem_raw, p.in + (row + gridDim.x) * p.in_stride, row_tx, bar);
        }
        const double mx = slots_max<NW>(red, lane);
        double r_min = 0.0, r_sum = 0.0;
        if (MODE == MODE_LOO) {
            r_min = -slots_max<NW>(red + 32, lane);
            r_sum = slots_sum<NW>(red + 64, lane);
        }
        // LOO quantities in ll = -r terms
        const double ll_max = -r_min, ll_min = -mx, ll_mean = -r_sum / (double)S;
        const bool wide = (MODE == MODE_LOO) && !((ll_max - ll_min) <= 600.0);

        // ---------------- threshold guess: per-warp sorted thread maxima (as float distances to the max)
        const float dsorted = warp_sort32_f((float)(mx - tmax), lane);
        int q = p.q0, attempts = 0, C = 0;
        double body = 0.0, lsum = 0.0, vsum = 0.0;
        bool ok = !special;
        while (ok) {
            if (lane == q - 1) redf[wid] = dsorted;
            __syncthreads();  // (2)
            double taux;
            {
                const float mine = (lane < NW) ? redf[lane] : __int_as_float(0x7f800000);
                int rank = 0;
#pragma unroll
                for (int j = 0; j < NW; ++j) {
                    const float o = redf[j];
                    rank += (o < mine || (o == mine && j < lane)) ? 1 : 0;
                }
                const unsigned b = __ballot_sync(FULL, lane < NW && rank == NW / 2 - 1 + (NW == 1));
                taux = fmax(-(double)__shfl_sync(FULL, mine, __ffs(b) - 1), -1e300);  // padding (-inf) never qualifies
            }
            // -------- pass B: body exp-sum, candidate marks.  `mxl` is laundered through an empty asm so
            // the compiler cannot hoist the (threshold-independent) exps out of the retry loop and
            // then spill all EPT results
            double mxl = mx;
            asm volatile("" : "+d"(mxl));
            double ll_mean_l = ll_mean, ll_max_l = ll_max;
            if (MODE == MODE_LOO) asm volatile("" : "+d"(ll_mean_l), "+d"(ll_max_l));
            double bs = 0.0, ls = 0.0, vs = 0.0;
            unsigned cmask = 0;
#pragma unroll
            for (int j = 0; j < EP2; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double r = h ? v[j].y : v[j].x;
                    const double x = r - mxl;  // psis.py:134
                    const bool cand = x >= taux;
                    cmask |= cand ? (1u << (2 * j + h)) : 0u;
                    const double xc = fmax(x, -700.0);
                    if (MODE == MODE_LOO && !wide) {
                        double ep, em;
                        exp_tab_pm(xc, tb, ep, em);
                        if (!cand && x >= -700.0) bs += ep;
                        if (j < nv) {
                            ls += em;  // exp(ll - ll_min)
                            const double d = -r - ll_mean_l;
                            vs = fma(d, d, vs);
                        }
                    } else {
                        const double e = exp_tab(xc, tb);
                        if (!cand && x >= -700.0) bs += e;
                        if (MODE == MODE_LOO && j < nv) {
                            ls += exp(-r - ll_max_l);  // wide rows: literal (utils.py:349-351)
                            const double d = -r - ll_mean_l;
                            vs = fma(d, d, vs);
                        }
                    }
                }
                // keep the scheduler from interleaving all EPT exp chains at once (register pressure)
                asm volatile("" ::: "memory");
            }
            // -------- emit candidates: one shared atomic per warp, packed keys to global scratch
            {
                const int mine = __popc(cmask);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += u;
                }
                int base = 0;
                if (lane == 31 && incl > 0) base = atomicAdd(&ctl[0], incl);
                base = __shfl_sync(FULL, base, 31);
                int pos = base + incl - mine;
                unsigned long long* dst = p.ckey + (size_t)row * (size_t)cap;
#pragma unroll
                for (int j = 0; j < EP2; ++j) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (cmask & (1u << (2 * j + h))) {
                            if (pos < cap) {
                                const double x = (h ? v[j].y : v[j].x) - mxl;
                                const int s = 2 * (j * NT + tid) + h;
                                const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
                                const unsigned lo = ((unsigned)__double2loint(x) & ~KEY_IDX_MASK) | (KEY_IDX_MASK - (unsigned)s);
                                dst[pos] = ((unsigned long long)hi << 32) | lo;
                            }
                            ++pos;
                        }
                    }
                }
            }
            bs = warp_sum(bs);
            if (MODE == MODE_LOO) {
                ls = warp_sum(ls);
                vs = warp_sum(vs);
            }
            if (lane == 0) {
                red[96 + wid] = bs;
                if (MODE == MODE_LOO) {
                    red[128 + wid] = ls;
                    red[wid] = vs;  // slot 0 is free again (mx was read before barrier 2)
                }
            }
            __syncthreads();  // (3)
            C = ctl[0];
            body = slots_sum<NW>(red + 96, lane);
            if (MODE == MODE_LOO) {
                lsum = slots_sum<NW>(red + 128, lane);
                vsum = slots_sum<NW>(red, lane);
            }
            if (C >= M + 1 && C <= cap) break;
            // -------- retry with a moved rank (rare), then give the row to the general kernel
            ++attempts;
            int qn;
            if (C < M + 1) qn = min(32, q + 2 * attempts);
            else qn = max(1, min(q - 1, (int)((double)q * 1.3 * (double)(M + 1) / (double)C)));
            if (attempts >= 3 || qn == q) ok = false;
            q = qn;
            __syncthreads();  // everyone has read ctl[0] / red before they are reused
            if (tid == 0) ctl[0] = 0;
        }
        if (tid == 0) {
            SplitHeader h;
            h.mx = mx; h.body = body; h.lsum = lsum; h.vsum = vsum;
            h.lshift = wide ? ll_max : ll_min; h.ll_max = ll_max;
            h.C = C; h.flags = ok ? 0 : 1; h.attempts = attempts; h.pad = 0;
            p.hdr[row] = h;
            if (!ok) {
                p.fb_list[atomicAdd(p.fb_count, 1)] = (int)row;
                if (p.counters) atomicAdd(&p.counters[3], 1ull);
            }
        }
    }
}

// ------------------------------------------------------------------ tail kernel helpers
__device__ __forceinline__ double shfl_f64(double v, int src) { return __shfl_sync(FULL, v, src); }

// Bitonic sort of 32 * CAPL doubles (all >= 0, no NaN) held as k[i] <-> element e = 32 i + lane,
// ascending in e.  Partner distances below 32 are lane shuffles, the rest are register pairs.
template <int CAPL>
__device__ __forceinline__ void warp_bitonic_sort(double (&k)[CAPL], int lane) {
#pragma unroll
    for (int kk = 2; kk <= 32 * CAPL; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int dj = j >> 5;
#pragma unroll
                for (int i = 0; i < CAPL; ++i) {
                    if ((i & dj) == 0) {
                        const bool up = (((i << 5) & kk) == 0);
                        const double a = k[i], b = k[i | dj];
                        const bool sw = (a > b) == up;
                        k[i] = sw ? b : a;
                        k[i | dj] = sw ? a : b;
                    }
                }
            } else {
                const bool lower = ((lane & j) == 0);
#pragma unroll
                for (int i = 0; i < CAPL; ++i) {
                    const bool up = (kk >= 32) ? (((i << 5) & kk) == 0) : ((lane & kk) == 0);
                    const double o = __shfl_xor_sync(FULL, k[i], j);
                    const bool take_o = ((k[i] < o) != (up == lower));
                    k[i] = take_o ? o : k[i];
                }
            }
        }
    }
}

__device__ __forceinline__ bool rescale_pos_w(double& P, int& E) {
    int hi = __double2hiint(P), lo = __double2loint(P);
    const int e = (hi >> 20) & 0x7ff;
    if (hi < 0 || e == 0 || e == 0x7ff) return false;
    E += e - 1023;
    hi = (hi & 0x800fffff) | 0x3ff00000;
    P = __hiloint2double(hi, lo);
    return true;
}

// sum_i log1p(nb * t_i) over the warp's tail (t in shared memory, n values), literal form
__device__ __forceinline__ double warp_log1p_sum(const double* t, int n, double nb, int lane) {
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc += log1p(nb * t[i]);
    return warp_sum(acc);
}

// log prod_i (1 + nb t_i) for NJ grid points per lane (t broadcast from shared memory).
template <int NJ>
__device__ __forceinline__ bool gpd_products(const double* t, int n, const double (&nb)[4], int every,
                                             double (&out)[4]) {
    double P[NJ];
    int E[NJ];
    bool ok = true;
#pragma unroll
    for (int r = 0; r < NJ; ++r) {
        P[r] = 1.0;
        E[r] = 0;
    }
    int i = 0;
    while (i < n) {
        const int stop = min(n, i + every);
        for (; i + 2 <= stop; i += 2) {  // t is 16 B aligned and `every` is even
            const double2 tt = *reinterpret_cast<const double2*>(t + i);
#pragma unroll
            for (int r = 0; r < NJ; ++r) {
                P[r] *= fma(nb[r], tt.x, 1.0);
                P[r] *= fma(nb[r], tt.y, 1.0);
            }
        }
        if (i < stop) {
            const double t0 = t[i];
#pragma unroll
            for (int r = 0; r < NJ; ++r) P[r] *= fma(nb[r], t0, 1.0);
            ++i;
        }
#pragma unroll
        for (int r = 0; r < NJ; ++r) ok = rescale_pos_w(P[r], E[r]) && ok;
    }
#pragma unroll
    for (int r = 0; r < NJ; ++r) out[r] = log(P[r]) + (double)E[r] * 0.6931471805599453094;
    return ok;
}

// Zhang-Stephens fit for one warp (pyloo/psis.py:181-208).  t: shared memory, DESCENDING (t[0] is
// the largest), n >= 5.  Returns false when the row must go to the general kernel.
__device__ bool gpdfit_warp(const double* t, int n, int m, double tsum, int lane, double& k_out,
                            double& sigma_out) {
    const double tq = t[n - ((int)((double)n / 4.0 + 0.5))];  // ascending index int(n/4+.5)-1 (psis.py:187)
    const double tn = t[0];
    if (!(tq > 0.0) || !is_finite(tn) || m > 128) return false;
    const int NJ = (m + 31) >> 5;
    double b[4], nb[4], ks[4];
    bool flag[4];
    double bmag = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int j = lane + 32 * r;
        double bj = 1.0 - sqrt((double)m / ((double)(j + 1) - 0.5));  // psis.py:186
        bj /= 3.0 * tq;                                                // psis.py:187
        bj += 1.0 / tn;                                                // psis.py:188
        const bool live = j < m;
        b[r] = live ? bj : 0.0;
        nb[r] = -b[r];
        ks[r] = 0.0;
        flag[r] = live && (fabs(bj) * tsum < 0.015625);
        if (live) bmag = fmax(bmag, fabs(bj));
        if (live && !is_finite(bj)) bmag = inf_f64();
    }
    bmag = warp_max(bmag);
    const double fmx = 1.0 + bmag * tn;
    if (!(fmx < 0x1p31)) return false;
    bool ok;
    const int every = 32;  // (2^31)^32 < 2^1023 and (1 - b_max t_n)^32 > 2^-290: no over/underflow
    if (NJ == 1) ok = gpd_products<1>(t, n, nb, every, ks);
    else if (NJ == 2) ok = gpd_products<2>(t, n, nb, every, ks);
    else if (NJ == 3) ok = gpd_products<3>(t, n, nb, every, ks);
    else ok = gpd_products<4>(t, n, nb, every, ks);
    if (!__all_sync(FULL, ok)) return false;
    // grid points where the product form loses relative accuracy: literal log1p sum
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        unsigned fm = __ballot_sync(FULL, flag[r]);
        while (fm) {
            const int src = __ffs(fm) - 1;
            fm &= fm - 1;
            const double nbj = shfl_f64(nb[r], src);
            const double acc = warp_log1p_sum(t, n, nbj, lane);
            if (lane == src) ks[r] = acc;
        }
    }
    // profile log-likelihood (psis.py:190-191) and weights (psis.py:192)
    double Lj[4];
    double lm = -inf_f64();
    bool fin = true;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const bool live = (lane + 32 * r) < m;
        const double kj = ks[r] / (double)n;
        Lj[r] = live ? (double)n * (log(-(b[r] / kj)) - kj - 1.0) : -inf_f64();
        if (live) {
            fin = fin && is_finite(Lj[r]);
            lm = fmax(lm, Lj[r]);
        }
    }
    if (!__all_sync(FULL, fin)) return false;
    lm = warp_max(lm);
    double w[4], es = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        w[r] = ((lane + 32 * r) < m) ? exp(Lj[r] - lm) : 0.0;
        es += w[r];
    }
    es = warp_sum(es);
    const double thr = 10.0 * 2.220446049250313e-16;
    double ws = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        w[r] = w[r] / es;
        if (w[r] < thr) w[r] = 0.0;  // psis.py:194-197 (dead grid points already carry 0)
        ws += w[r];
    }
    ws = warp_sum(ws);
    double bp = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (w[r] != 0.0) bp += b[r] * (w[r] / ws);  // psis.py:198-201
    bp = warp_sum(bp);
    // k_post = mean log1p(-b_post t) (psis.py:203): product form unless it loses accuracy
    double lsum;
    if (fabs(bp) * tsum < 0.015625 || !is_finite(bp)) {
        lsum = warp_log1p_sum(t, n, -bp, lane);
    } else {
        double P = 1.0;
        int E = 0;
        bool okp = true;
        int cnt = 0;
        for (int i = lane; i < n; i += 32) {
            P *= fma(-bp, t[i], 1.0);
            if (++cnt == 32) {
                cnt = 0;
                okp = rescale_pos_w(P, E) && okp;
            }
        }
        okp = rescale_pos_w(P, E) && okp;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            P *= __shfl_xor_sync(FULL, P, o);
            E += __shfl_xor_sync(FULL, E, o);
        }
        okp = rescale_pos_w(P, E) && okp;
        if (!__all_sync(FULL, okp)) lsum = warp_log1p_sum(t, n, -bp, lane);
        else lsum = log(P) + (double)E * 0.6931471805599453094;
    }
    const double k_post = lsum / (double)n;
    sigma_out = -k_post / bp;                                  // psis.py:205
    k_out = ((double)n * k_post + 5.0) / ((double)n + 10.0);   // psis.py:206
    return true;
}

struct TailSmem {
    size_t off_l1p, off_t, t_stride, total;
};
__host__ __device__ inline TailSmem tail_smem(int M, int TL, int warps) {
    TailSmem L;
    L.off_l1p = 0;
    size_t o = align_up((size_t)(M + 1) * 8, 16);
    L.off_t = o;
    L.t_stride = (size_t)32 * TL * 8;
    o += L.t_stride * warps;
    L.total = align_up(o, 128);
    return L;
}

constexpr int TAIL_WARPS = 4;

// One row for one warp.  Returns false -> general kernel.
template <int CAPL, int TL, int MODE>
__device__ __forceinline__ bool tail_row(const SplitParams& p, long long row, const SplitHeader& h,
                                         const double* l1p, double* tsh, int lane) {
    const int S = p.S, M = p.M, C = h.C;
    const double mx = h.mx;
    const double PAD = inf_f64();
    const double* src = p.in + row * p.in_stride;

    // ---- candidate keys -> registers, sort ascending (= descending x, ties by descending draw index)
    double k[CAPL];
    {
        const unsigned long long* ck = p.ckey + (size_t)row * (size_t)p.cap;
#pragma unroll
        for (int i = 0; i < CAPL; ++i) {
            const int e = 32 * i + lane;
            k[i] = (e < C) ? __longlong_as_double((long long)ck[e]) : PAD;
        }
    }
    warp_bitonic_sort<CAPL>(k, lane);

    // ---- exact values of the head of the order (tail + cutoff), gathered from the row
    double xt[TL];
    int st[TL];
#pragma unroll
    for (int i = 0; i < TL; ++i) {
        const int e = 32 * i + lane;
        st[i] = (int)(KEY_IDX_MASK - ((unsigned)__double2loint(k[i]) & KEY_IDX_MASK));
        const double v = (e < C) ? src[st[i]] : 0.0;
        xt[i] = (e < C) ? ((MODE == MODE_LOO) ? -v : v) - mx : -inf_f64();
    }
    // candidates that can never be in the tail (e >= 32 TL > M): their exp goes to the normaliser
    double nont = 0.0;
#pragma unroll
    for (int i = TL; i < CAPL; ++i) {
        const int e = 32 * i + lane;
        if (32 * i < C) {
            const int s = (int)(KEY_IDX_MASK - ((unsigned)__double2loint(k[i]) & KEY_IDX_MASK));
            if (e < C) {
                const double v = src[s];
                nont += exp(((MODE == MODE_LOO) ? -v : v) - mx);
            }
        }
    }

    // ---- cutoff = (M+1)-th largest = element M of the order (psis.py:135-136)
    double kc = 0.0, xc = 0.0;
#pragma unroll
    for (int i = 0; i < TL; ++i)
        if (i == (M >> 5)) {
            kc = k[i];
            xc = xt[i];
        }
    kc = shfl_f64(kc, M & 31);
    xc = shfl_f64(xc, M & 31);
    const unsigned long long tc = (unsigned long long)__double_as_longlong(kc) >> KEY_IDX_BITS;
    int n_lt = 0, n_eq = 0;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < CAPL; ++i) {
        const unsigned long long ti = (unsigned long long)__double_as_longlong(k[i]) >> KEY_IDX_BITS;
        const bool eq = (ti == tc);
        n_lt += __popc(__ballot_sync(FULL, ti < tc));
        n_eq += __popc(__ballot_sync(FULL, eq));
        if (i < TL) {
            if (eq && xt[i] != xc) bad = true;  // distinct values share the truncated key at the cutoff
        } else if (eq) {
            bad = true;  // the run of cutoff ties leaves the gathered range
        }
    }
    (void)n_eq;
    const int n = n_lt;  // draws with x > cutoff value (ties with the cutoff are not in the tail)
    if (xc < p.cutoffmin) bad = true;  // cutoff clamped at log(DBL_MIN): general kernel
    // order inside the tail must be exact (descending x, descending index on ties)
#pragma unroll
    for (int i = 0; i < TL; ++i) {
        const int e = 32 * i + lane;
        double xn = __shfl_down_sync(FULL, xt[i], 1);
        int sn = __shfl_down_sync(FULL, st[i], 1);
        if (i + 1 < TL) {
            const double x0 = shfl_f64(xt[i + 1], 0);
            const int s0 = __shfl_sync(FULL, st[i + 1], 0);
            if (lane == 31) {
                xn = x0;
                sn = s0;
            }
        }
        if (e + 1 < n && (lane < 31 || i + 1 < TL)) {
            if (!(xt[i] > xn || (xt[i] == xn && st[i] > sn))) bad = true;
        }
    }
    if (__any_sync(FULL, bad)) return false;

    const double c = xc;  // >= cutoffmin here
    const double exp_c = exp(c);  // psis.py:138
    // ---- t_i = exp(x_i) - exp(c) (psis.py:146-147), descending, to shared memory; candidates at or
    //      below the cutoff join the normaliser
    double tsum = 0.0, traw = 0.0;
#pragma unroll
    for (int i = 0; i < TL; ++i) {
        const int e = 32 * i + lane;
        if (32 * i < C) {
            if (e < C) {
                const double ex = exp(xt[i]);
                if (e < n) {
                    const double ti = ex - exp_c;
                    tsh[e] = ti;
                    tsum += ti;
                    traw += ex;
                } else {
                    nont += ex;
                }
            }
        }
    }
    tsum = warp_sum(tsum);
    traw = warp_sum(traw);
    nont = warp_sum(nont);
    __syncwarp();

    double kk = inf_f64(), sigma = nan_f64();
    bool smooth = false;
    if (n > 4) {
        int m = p.m_full;
        if (n != M) {
            m = (int)sqrt((double)n);
            while (m * m > n) --m;
            while ((m + 1) * (m + 1) <= n) ++m;
            m += 30;
        }
        if (!gpdfit_warp(tsh, n, m, tsum, lane, kk, sigma)) return false;
        smooth = is_finite(kk);  // psis.py:150
    }
    // ---- smoothed tail (psis.py:153-157, _gpinv :211-222); element e has ascending rank n-1-e
    double tails = traw;
    double sm[TL];
#pragma unroll
    for (int i = 0; i < TL; ++i) sm[i] = 0.0;
    if (smooth) {
        double tsm = 0.0;
#pragma unroll
        for (int i = 0; i < TL; ++i) {
            const int e = 32 * i + lane;
            if (32 * i < n) {
                if (e < n) {
                    const int rk = n - 1 - e;
                    double q;
                    if (sigma <= 0.0) {
                        q = nan_f64();
                    } else {
                        const double l1 = (n == M) ? l1p[rk] : log1p(-(((double)rk + 0.5) / (double)n));
                        q = (fabs(kk) < 2.220446049250313e-16) ? -l1 : expm1(-kk * l1) / kk;
                        q *= sigma;
                    }
                    double y = q + exp_c;
                    double s_ = log(y);
                    if (s_ > 0.0) {  // psis.py:157
                        s_ = 0.0;
                        y = 1.0;
                    }
                    sm[i] = s_;
                    tsm += y;
                }
            }
        }
        tails = warp_sum(tsm);
    }
    const double body = h.body + nont;
    const double lse = log(body + tails);  // psis.py:158

    if (MODE == MODE_PSISLW) {
        // ---- normalised row, then the smoothed tail on top of it
        double* dst = p.out + row * p.out_stride;
        const double2* s2 = reinterpret_cast<const double2*>(src);
        double2* d2 = reinterpret_cast<double2*>(dst);
        const int S2 = S >> 1;
        int i2 = lane;
        for (; i2 + 96 < S2; i2 += 128) {
            double2 a0 = s2[i2], a1 = s2[i2 + 32], a2 = s2[i2 + 64], a3 = s2[i2 + 96];
            a0.x = (a0.x - mx) - lse; a0.y = (a0.y - mx) - lse;
            a1.x = (a1.x - mx) - lse; a1.y = (a1.y - mx) - lse;
            a2.x = (a2.x - mx) - lse; a2.y = (a2.y - mx) - lse;
            a3.x = (a3.x - mx) - lse; a3.y = (a3.y - mx) - lse;
            d2[i2] = a0; d2[i2 + 32] = a1; d2[i2 + 64] = a2; d2[i2 + 96] = a3;
        }
        for (; i2 < S2; i2 += 32) {
            double2 a0 = s2[i2];
            a0.x = (a0.x - mx) - lse; a0.y = (a0.y - mx) - lse;
            d2[i2] = a0;
        }
        __syncwarp();
        if (smooth) {
#pragma unroll
            for (int i = 0; i < TL; ++i) {
                const int e = 32 * i + lane;
                if (e < n) dst[st[i]] = sm[i] - lse;
            }
        }
        if (lane == 0) p.k_out[row] = kk;
    } else {
        // elpd_i = LSE_s(lw_s + ll_s): body terms are the constant -(mx + lse), tail terms differ
        // from it by (smoothed - raw) (loo.py:289,319-324)
        double dmax = 0.0, es = (double)n;
        if (smooth) {
            double dm = 0.0;
#pragma unroll
            for (int i = 0; i < TL; ++i)
                if (32 * i + lane < n) dm = fmax(dm, sm[i] - xt[i]);
            dmax = warp_max(dm);
            double e2 = 0.0;
#pragma unroll
            for (int i = 0; i < TL; ++i)
                if (32 * i < n && 32 * i + lane < n) e2 += exp((sm[i] - xt[i]) - dmax);
            es = warp_sum(e2);
        }
        const double tot = (double)(S - n) * exp(-dmax) + es;
        const double elpd = ((-mx - lse) + dmax) + log(tot);
        const double lppd = log(h.lsum) + (h.lshift - log((double)S));  // utils.py:352-357, b_inv = S
        const double var = h.vsum / (double)S;                          // waic.py:145
        if (lane == 0) {
            p.k_out[row] = kk;
            p.elpd_i[row] = elpd;
            p.lppd_i[row] = lppd;
            p.var_i[row] = var;
            p.lppdw_i[row] = lppd;
        }
    }
    if (p.diag && lane == 0) {
        double* d = p.diag + row * DIAG_STRIDE;
        d[0] = mx; d[1] = c; d[2] = (double)n; d[3] = (double)C;
        d[4] = (double)h.attempts; d[5] = body; d[6] = tails; d[7] = sigma;
    }
    return true;
}

template <int TL>
constexpr int tail_min_blocks() { return TL <= 8 ? 4 : 2; }

template <int TL, int MODE>
__global__ void __launch_bounds__(TAIL_WARPS * 32, tail_min_blocks<TL>()) psis_tail_kernel(const SplitParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TailSmem L = tail_smem(p.M, TL, TAIL_WARPS);
    double* l1p = reinterpret_cast<double*>(smem_raw + L.off_l1p);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* tsh = reinterpret_cast<double*>(smem_raw + L.off_t + L.t_stride * wid);
    // per-CTA table for the smoothing step: depends only on (rank, M), psis.py:153 + :221
    for (int i = threadIdx.x; i < p.M; i += TAIL_WARPS * 32) l1p[i] = log1p(-(((double)i + 0.5) / (double)p.M));
    __syncthreads();
    const long long nwarps = (long long)gridDim.x * TAIL_WARPS;
    for (long long row = (long long)blockIdx.x * TAIL_WARPS + wid; row < p.n_rows; row += nwarps) {
        const SplitHeader h = p.hdr[row];
        if (h.flags) continue;
        bool ok;
        if (h.C <= 32 * TL) ok = tail_row<TL, TL, MODE>(p, row, h, l1p, tsh, lane);
        else ok = tail_row<2 * TL, TL, MODE>(p, row, h, l1p, tsh, lane);
        if (!ok && lane == 0) {
            p.fb_list[atomicAdd(p.fb_count, 1)] = (int)row;
            if (p.counters) atomicAdd(&p.counters[3], 1ull);
        }
        __syncwarp();
    }
}

}  // namespace b2l
