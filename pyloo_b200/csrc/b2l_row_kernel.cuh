// PSIS row kernel (sm_100a): one CTA owns one observation at a time, its S draws staged in shared
// memory by 1-D bulk TMA (cp.async.bulk + mbarrier), persistent grid, optional double buffering.
//
// Replaces, per observation, the NumPy work of pyloo/psis.py:133-160 (`_psislw`), :181-208
// (`_gpdfit`), :211-231 (`_gpinv`) and pyloo/utils.py:344-359 (`_logsumexp`); in LOO mode also
// pyloo/loo.py:286-289,319-337 (`lw += ll`, `loo_i`, `lppd_i`) and pyloo/waic.py:137-145.
//
// Selection is exact: draws are keyed by the order-preserving 64-bit image of x = fl(r - max r)
// (formed exactly as psis.py:134 does), a sampled threshold isolates a candidate set that is
// bitonic-sorted in shared memory, and the (M+1)-th largest is read from the sorted candidates;
// a bit-wise binary search over the key space is the exact fallback (ties, degenerate rows).
#pragma once

#include "b2l_common.cuh"

namespace b2l {

enum : int { MODE_PSISLW = 0, MODE_LOO = 1 };

constexpr int GPD_MAX_GRID = 128;  // m = 30 + floor(sqrt(n)) <= 128  <=>  n <= 9603
constexpr int DIAG_STRIDE = 8;

struct RowParams {
    const double* in;      // row i = in + i * in_stride, S contiguous doubles
    long long in_stride;   // elements
    double* out;           // PSISLW: row i = out + i * out_stride
    long long out_stride;
    double* k_out;         // [n_rows] Pareto k
    double* elpd_i;        // LOO [n_rows]  log-scale elpd_loo_i
    double* lppd_i;        // LOO [n_rows]  loo-policy lppd_i   (pyloo/loo.py:329-337)
    double* var_i;         // LOO [n_rows]  waic-policy var_s(ll) (pyloo/waic.py:145)
    double* lppdw_i;       // LOO [n_rows]  waic-policy lppd_i  (pyloo/waic.py:137-143)
    double* diag;          // optional [n_rows][DIAG_STRIDE]: max, cutoff, n_tail, n_cand, attempts, body, tail, sigma
    unsigned long long* counters;  // optional [4]: n_nan, n_posinf, n_neginf, n_fallback
    long long n_rows;
    int S;
    int M;         // tail length, cutoff_ind = -M-1 (computed by the caller, pyloo/psis.py:89)
    int cap;       // candidate capacity (power of two)
    int ns;        // threshold sample size (power of two)
    int r0;        // initial sample rank
    int nbuf;      // row buffers per CTA (1 or 2)
    int use_bulk;  // 1: 16 B aligned rows -> bulk TMA; 0: cooperative LDG/STG
    int waic_only; // LOO mode: skip PSIS, only lppd / variance outputs
    double cutoffmin;
};

struct RowSmemLayout {
    size_t row_bytes, off_row, off_ckey, off_cidx, off_tbuf, off_gb, off_gk, off_gw, off_gflag,
        off_part, off_red, off_ctl, off_bar, total;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Same carve-up on host (launch size) and device.
__host__ __device__ inline RowSmemLayout row_smem_layout(int S, int M, int cap, int ns, int nbuf,
                                                         int nt) {
    RowSmemLayout L;
    size_t o = 0;
    L.row_bytes = align_up((size_t)S * 8, 128);
    L.off_row = o;
    o += L.row_bytes * (size_t)nbuf;
    L.off_ckey = o;
    o += (size_t)cap * 8;
    L.off_cidx = o;
    o += align_up((size_t)cap * 4, 16);
    L.off_tbuf = o;
    o += align_up((size_t)(M + 1) * 8, 16);
    L.off_gb = o;
    o += GPD_MAX_GRID * 8;
    L.off_gk = o;
    o += GPD_MAX_GRID * 8;
    L.off_gw = o;
    o += GPD_MAX_GRID * 8;
    L.off_gflag = o;
    o += GPD_MAX_GRID * 4;
    L.off_part = o;  // sample keys (ns * 8) alias the GPD partial products (nt * 12)
    size_t a = (size_t)ns * 8, b = align_up((size_t)nt * 12, 16);
    o += (a > b ? a : b);
    L.off_red = o;
    o += 128 * 8;
    L.off_ctl = o;
    o += 16 * 4;
    L.off_bar = o;
    o += 2 * 8;
    L.total = align_up(o, 128);
    return L;
}

// ------------------------------------------------------------------ bitonic sorts (shared memory)
template <int NT>
__device__ void bitonic_sort_keys(uint64_t* key, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += NT) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool up = ((i & k) == 0);
                uint64_t a = key[i], b = key[l];
                if ((a > b) == up) {
                    key[i] = b;
                    key[l] = a;
                }
            }
            __syncthreads();
        }
    }
}
// (key, index) ascending -- the tie order we define (reference order is unspecified)
template <int NT>
__device__ void bitonic_sort_pairs(uint64_t* key, int* idx, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += NT) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool up = ((i & k) == 0);
                uint64_t a = key[i], b = key[l];
                int ia = idx[i], ib = idx[l];
                bool gt = (a > b) || (a == b && ia > ib);
                if (gt == up) {
                    key[i] = b;
                    key[l] = a;
                    idx[i] = ib;
                    idx[l] = ia;
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ GPD fit (block cooperative)
struct GpdScratch {
    double* b;     // [GPD_MAX_GRID] grid b_j                      (psis.py:186-188)
    double* ks;    // [GPD_MAX_GRID] sum_i log1p(-b_j t_i), then L_j (psis.py:190-191)
    double* w;     // [GPD_MAX_GRID] posterior weights              (psis.py:192-198)
    int* flag;     // [GPD_MAX_GRID] grid points that need the literal log1p path
    double* part;  // [NT] partial products
    int* parte;    // [NT] partial exponents
    double* red;   // reduction scratch
};

__device__ __forceinline__ bool rescale_pos(double& P, int& E) {
    // P > 0 finite normal -> mantissa in [1,2), exponent accumulated in E.  Returns false otherwise.
    int hi = __double2hiint(P), lo = __double2loint(P);
    int e = (hi >> 20) & 0x7ff;
    if (hi < 0 || e == 0 || e == 0x7ff) return false;
    E += e - 1023;
    hi = (hi & 0x800fffff) | 0x3ff00000;
    P = __hiloint2double(hi, lo);
    return true;
}

// Zhang-Stephens empirical-Bayes fit, pyloo/psis.py:181-208, for sorted t[0..n), n >= 5.
// The m x n `log1p` matrix of psis.py:190 is evaluated as log prod_i (1 - b_j t_i) with exponent
// renormalisation (one log per grid point); grid points where that loses relative accuracy
// (|b_j| sum t small, non-positive or non-finite factors) take the literal log1p path.
template <int NT>
__device__ void gpdfit_block(const double* t, int n, GpdScratch g, double& k_out,
                             double& sigma_out) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int m = (int)sqrt((double)n);
    while (m * m > n) --m;
    while ((m + 1) * (m + 1) <= n) ++m;
    m += 30;                                               // psis.py:184
    const double tq = t[(int)((double)n / 4.0 + 0.5) - 1];  // psis.py:187
    const double tn = t[n - 1];

    double tsum = 0.0;
    for (int i = tid; i < n; i += NT) tsum += t[i];
    tsum = block_sum<NT>(tsum, g.red);

    if (tid < m) {
        double b = 1.0 - sqrt((double)m / ((double)(tid + 1) - 0.5));  // psis.py:186
        b /= 3.0 * tq;                                                  // psis.py:187
        b += 1.0 / tn;                                                  // psis.py:188
        g.b[tid] = b;
        g.flag[tid] = (!is_finite(b) || fabs(b) * tsum < 0.015625) ? 1 : 0;
    }
    __syncthreads();

    // ---- product-form profile: thread = (grid point j, chunk of t)
    const int JW = (m + 31) >> 5;          // warps per chunk
    const int NCH = (NT / 32) / JW;        // chunks
    {
        const int jw = wid % JW, ch = wid / JW;
        const int j = (jw << 5) + lane;
        double P = 1.0;
        int E = 0;
        bool ok = true;
        if (j < m && ch < NCH) {
            const double bmag = fmax(fabs(g.b[0]), fabs(g.b[m - 1]));
            const double fmx = 1.0 + bmag * tn;
            const int R = (fmx < 0x1p120) ? 8 : ((fmx < 0x1p250) ? 4 : 1);
            const int len = (n + NCH - 1) / NCH;
            const int i0 = ch * len, i1 = min(n, i0 + len);
            const double nb = -g.b[j];
            int cnt = 0;
            for (int i = i0; i < i1; ++i) {
                double f = fma(nb, t[i], 1.0);
                ok = ok && (f > 0.0);
                P *= f;
                if (++cnt == R) {
                    cnt = 0;
                    ok = rescale_pos(P, E) && ok;
                }
            }
            ok = rescale_pos(P, E) && ok;
        }
        g.part[tid] = P;
        g.parte[tid] = E;
        if (!ok && j < m) g.flag[j] = 1;  // benign race: all writers store 1
    }
    __syncthreads();
    if (tid < m && !g.flag[tid]) {
        const int jw = tid >> 5, l = tid & 31;
        double P = 1.0;
        int E = 0;
        for (int ch = 0; ch < NCH; ++ch) {
            const int slot = ((ch * JW + jw) << 5) + l;
            P *= g.part[slot];
            E += g.parte[slot];
            rescale_pos(P, E);
        }
        g.ks[tid] = log(P) + (double)E * 0.6931471805599453094;
    }
    __syncthreads();
    // ---- literal path for flagged grid points (block-uniform loop)
    for (int j = 0; j < m; ++j) {
        if (g.flag[j]) {
            const double nb = -g.b[j];
            double acc = 0.0;
            for (int i = tid; i < n; i += NT) acc += log1p(nb * t[i]);
            acc = block_sum<NT>(acc, g.red);
            if (tid == 0) g.ks[j] = acc;
        }
    }
    __syncthreads();
    // ---- profile log-likelihood L_j (psis.py:191)
    bool fin = true;
    if (tid < m) {
        const double kj = g.ks[tid] / (double)n;
        const double L = (double)n * (log(-(g.b[tid] / kj)) - kj - 1.0);
        g.ks[tid] = L;
        fin = is_finite(L);
    }
    const int allfin = __syncthreads_and(fin ? 1 : 0);
    // ---- weights (psis.py:192): 1 / sum_l exp(L_l - L_j)
    if (allfin) {
        if (wid == 0) {
            double lm = -inf_f64();
            for (int j = lane; j < m; j += 32) lm = fmax(lm, g.ks[j]);
            lm = warp_max(lm);
            double es = 0.0;
            for (int j = lane; j < m; j += 32) {
                double e = exp(g.ks[j] - lm);
                g.w[j] = e;
                es += e;
            }
            es = warp_sum(es);
            for (int j = lane; j < m; j += 32) g.w[j] = g.w[j] / es;
        }
    } else if (tid < m) {  // literal m x m form keeps the reference's inf/NaN semantics
        const double Lj = g.ks[tid];
        double s = 0.0;
        for (int l = 0; l < m; ++l) s += exp(g.ks[l] - Lj);
        g.w[tid] = 1.0 / s;
    }
    __syncthreads();
    // ---- drop negligible weights, renormalise, posterior mean of b (psis.py:194-201)
    if (wid == 0) {
        const double thr = 10.0 * 2.220446049250313e-16;
        double ws = 0.0;
        int nk = 0;
        for (int j = lane; j < m; j += 32) {
            double w = g.w[j];
            if (w >= thr) {
                ws += w;
                ++nk;
            }
        }
        ws = warp_sum(ws);
        nk = warp_isum(nk);
        double bp = 0.0;
        for (int j = lane; j < m; j += 32) {
            double w = g.w[j];
            if (w >= thr) bp += g.b[j] * (w / ws);
        }
        bp = warp_sum(bp);
        if (nk == 0) bp = 0.0;  // np.sum of an empty array
        if (lane == 0) g.w[0] = bp;
    }
    __syncthreads();
    const double b_post = g.w[0];
    double acc = 0.0;
    for (int i = tid; i < n; i += NT) acc += log1p(-b_post * t[i]);  // psis.py:203
    acc = block_sum<NT>(acc, g.red);
    const double k_post = acc / (double)n;
    sigma_out = -k_post / b_post;                                     // psis.py:205
    k_out = ((double)n * k_post + 5.0) / ((double)n + 10.0);          // psis.py:206
}

// ------------------------------------------------------------------ the kernel
template <int NT, int MODE>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 3 : 1) psis_row_kernel(const RowParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const RowSmemLayout L = row_smem_layout(p.S, p.M, p.cap, p.ns, p.nbuf, NT);
    uint64_t* ckey = reinterpret_cast<uint64_t*>(smem_raw + L.off_ckey);
    int* cidx = reinterpret_cast<int*>(smem_raw + L.off_cidx);
    double* tbuf = reinterpret_cast<double*>(smem_raw + L.off_tbuf);
    uint64_t* skey = reinterpret_cast<uint64_t*>(smem_raw + L.off_part);
    double* red = reinterpret_cast<double*>(smem_raw + L.off_red);
    int* ctl = reinterpret_cast<int*>(smem_raw + L.off_ctl);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
    GpdScratch g;
    g.b = reinterpret_cast<double*>(smem_raw + L.off_gb);
    g.ks = reinterpret_cast<double*>(smem_raw + L.off_gk);
    g.w = reinterpret_cast<double*>(smem_raw + L.off_gw);
    g.flag = reinterpret_cast<int*>(smem_raw + L.off_gflag);
    g.part = reinterpret_cast<double*>(smem_raw + L.off_part);
    g.parte = reinterpret_cast<int*>(smem_raw + L.off_part + (size_t)NT * 8);
    g.red = red;

    const int tid = threadIdx.x, lane = tid & 31;
    const int S = p.S, M = p.M, cap = p.cap, ns = p.ns;
    const uint32_t row_tx = (uint32_t)S * 8u;
    const double NEG_INF = -inf_f64();

    if (p.use_bulk && tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    long long row = blockIdx.x;
    if (p.use_bulk && tid == 0 && row < p.n_rows) {
        mbar_expect_tx(&bars[0], row_tx);
        bulk_g2s(smem_raw + L.off_row, p.in + row * p.in_stride, row_tx, &bars[0]);
    }

    for (int it = 0; row < p.n_rows; row += gridDim.x, ++it) {
        const int bsel = (p.nbuf == 2) ? (it & 1) : 0;
        double* rbuf = reinterpret_cast<double*>(smem_raw + L.off_row + (size_t)bsel * L.row_bytes);
        const long long nrow = row + gridDim.x;

        // ---------------- stage the row
        if (p.use_bulk) {
            if (p.nbuf == 2 && nrow < p.n_rows && tid == 0) {
                // the other buffer was stored from in the previous iteration: drain, then refill
                bulk_wait_read0();
                fence_proxy_async();
                mbar_expect_tx(&bars[bsel ^ 1], row_tx);
                bulk_g2s(smem_raw + L.off_row + (size_t)(bsel ^ 1) * L.row_bytes,
                         p.in + nrow * p.in_stride, row_tx, &bars[bsel ^ 1]);
            }
            const uint32_t parity = (p.nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(&bars[bsel], parity);
        } else {
            const double* src = p.in + row * p.in_stride;
            for (int s = tid; s < S; s += NT) rbuf[s] = src[s];
            __syncthreads();
        }

        // ---------------- pass A: extrema, NaN/inf census (and sum for the WAIC variance)
        double a_max = NEG_INF, a_min = inf_f64(), a_sum = 0.0;
        int c_nan = 0, c_pinf = 0, c_ninf = 0;
        for (int s = tid; s < S; s += NT) {
            double v = rbuf[s];
            if (v != v) {
                ++c_nan;
                if (MODE == MODE_LOO) {  // pyloo/loo.py:227: NaN -> -1e10
                    v = -1e10;
                    rbuf[s] = v;
                }
            }
            if (MODE == MODE_LOO || v == v) {
                a_max = fmax(a_max, v);
                a_min = fmin(a_min, v);
                a_sum += v;
                if (!is_finite(v)) {
                    if (v > 0) ++c_pinf; else ++c_ninf;
                }
            }
        }
        {
            a_max = warp_max(a_max);
            a_min = warp_min(a_min);
            a_sum = warp_sum(a_sum);
            int packed_inf = warp_isum(c_pinf), packed_ninf = warp_isum(c_ninf);
            c_nan = warp_isum(c_nan);
            __syncthreads();
            if (lane == 0) {
                double* r = red + (tid >> 5) * 6;
                r[0] = a_max; r[1] = a_min; r[2] = a_sum;
                r[3] = (double)c_nan; r[4] = (double)packed_inf; r[5] = (double)packed_ninf;
            }
            __syncthreads();
            a_max = red[0]; a_min = red[1]; a_sum = red[2];
            double dn = red[3], dp = red[4], dm = red[5];
#pragma unroll
            for (int w = 1; w < NT / 32; ++w) {
                a_max = fmax(a_max, red[w * 6 + 0]);
                a_min = fmin(a_min, red[w * 6 + 1]);
                a_sum += red[w * 6 + 2];
                dn += red[w * 6 + 3]; dp += red[w * 6 + 4]; dm += red[w * 6 + 5];
            }
            c_nan = (int)dn; c_pinf = (int)dp; c_ninf = (int)dm;
            __syncthreads();
        }
        if (tid == 0 && p.counters && (c_nan | c_pinf | c_ninf)) {
            if (c_nan) atomicAdd(&p.counters[0], (unsigned long long)c_nan);
            if (c_pinf) atomicAdd(&p.counters[1], (unsigned long long)c_pinf);
            if (c_ninf) atomicAdd(&p.counters[2], (unsigned long long)c_ninf);
        }

        // r = raw log weight: PSISLW r = v ; LOO r = -ll (pyloo/loo.py:286-288)
        const double mx = (MODE == MODE_LOO) ? -a_min : a_max;  // max_s r_s
        const double ll_max = a_max;
        const double ll_mean = a_sum / (double)S;

        // Rows on which the reference produces no tail at all (NaN / inf arithmetic, App. D)
        bool run_psis;
        if (MODE == MODE_PSISLW) run_psis = (c_nan == 0) && is_finite(mx);
        else run_psis = (c_ninf == 0) && !p.waic_only;  // ll = -inf  =>  r = +inf  =>  x = NaN, k = inf

        double kk = inf_f64(), sigma = nan_f64(), lse = nan_f64();
        double c = nan_f64(), body = 0.0, tails = 0.0, lsum = 0.0, vsum = 0.0;
        int n = 0, C = 0, attempts = 0, tail0 = 0;
        bool smooth = false;

        if (run_psis) {
            // ---------------- threshold guess from a sorted sample
            const bool sampling = S > cap;
            int R = p.r0;
            if (sampling) {
                for (int j = tid; j < ns; j += NT) {
                    const int idx = (int)(((long long)j * S) / ns);
                    const double v = rbuf[idx];
                    skey[j] = key_of(((MODE == MODE_LOO) ? -v : v) - mx);
                }
                __syncthreads();
                bitonic_sort_keys<NT>(skey, ns);
            }
            bool exact = false;
            double tau = NEG_INF, craw_exact = NEG_INF;
            while (true) {
                if (sampling && !exact) tau = val_of(skey[ns - R]);
                if (tid == 0) ctl[0] = 0;
                __syncthreads();
                // -------- pass B: x = fl(r - mx) (psis.py:134); candidates x > tau; body exp-sum
                double bsum = 0.0;
                lsum = 0.0;
                vsum = 0.0;
                for (int base = 0; base < S; base += NT) {
                    const int s = base + tid;
                    const bool valid = s < S;
                    double x = NEG_INF;
                    if (valid) {
                        const double v = rbuf[s];
                        x = ((MODE == MODE_LOO) ? -v : v) - mx;
                        if (MODE == MODE_LOO) {
                            lsum += exp(v - ll_max);         // loo.py:329-337 / utils.py:349-351
                            const double d = v - ll_mean;    // waic.py:145 (two-pass variance)
                            vsum += d * d;
                        }
                    }
                    const bool isc = valid && (x > tau);
                    const unsigned mask = __ballot_sync(FULL, isc);
                    if (mask) {
                        const int leader = __ffs(mask) - 1;
                        int basepos = 0;
                        if (lane == leader) basepos = atomicAdd(&ctl[0], __popc(mask));
                        basepos = __shfl_sync(FULL, basepos, leader);
                        if (isc) {
                            const int pos = basepos + __popc(mask & ((1u << lane) - 1u));
                            if (pos < cap) {
                                ckey[pos] = key_of(x);
                                cidx[pos] = s;
                            }
                        }
                    }
                    if (valid && !isc) bsum += exp(x);
                }
                body = block_sum<NT>(bsum, red);
                C = ctl[0];
                __syncthreads();
                if (!sampling || exact) break;
                if (C >= M + 1 && C <= cap) break;
                ++attempts;
                int Rn = (C < M + 1) ? min(ns, 2 * R + 8) : max(1, R >> 1);
                if (attempts >= 3 || Rn == R) {
                    // -------- exact fallback: (M+1)-th largest key by bit-wise binary search
                    uint64_t K = 0;
                    for (int bit = 63; bit >= 0; --bit) {
                        const uint64_t trial = K | (1ull << bit);
                        int cnt = 0;
                        for (int s = tid; s < S; s += NT) {
                            const double v = rbuf[s];
                            cnt += (key_of(((MODE == MODE_LOO) ? -v : v) - mx) >= trial) ? 1 : 0;
                        }
                        cnt = block_isum<NT>(cnt, red);
                        if (cnt >= M + 1) K = trial;
                    }
                    craw_exact = val_of(K);
                    tau = fmax(craw_exact, p.cutoffmin);
                    exact = true;
                    if (tid == 0 && p.counters) atomicAdd(&p.counters[3], 1ull);
                } else {
                    R = Rn;
                }
            }
            if (MODE == MODE_LOO) {
                lsum = block_sum<NT>(lsum, red);
                vsum = block_sum<NT>(vsum, red);
            }

            // ---------------- sort candidates ascending by (value, index); pads (key 0) first
            int P2 = 1;
            while (P2 < C) P2 <<= 1;
            for (int i = C + tid; i < P2; i += NT) {
                ckey[i] = 0ull;
                cidx[i] = -1;
            }
            __syncthreads();
            if (P2 > 1) bitonic_sort_pairs<NT>(ckey, cidx, P2);

            // ---------------- cutoff (psis.py:135-136) and tail (psis.py:139-141)
            double c_raw;
            if (exact) c_raw = craw_exact;
            else c_raw = (C >= M + 1) ? val_of(ckey[P2 - M - 1]) : NEG_INF;
            c = fmax(c_raw, p.cutoffmin);
            {
                const uint64_t kc = key_of(c);
                int lo = P2 - C, hi = P2;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (ckey[mid] > kc) hi = mid; else lo = mid + 1;
                }
                tail0 = lo;
                n = P2 - lo;
            }
            const double exp_c = exp(c);  // psis.py:138
            // candidates at or below the cutoff belong to the body sum
            {
                double b2 = 0.0;
                for (int i = P2 - C + tid; i < tail0; i += NT) b2 += exp(val_of(ckey[i]));
                body += block_sum<NT>(b2, red);
            }

            // ---------------- GPD fit on the tail (psis.py:146-148)
            if (n > 4) {
                for (int i = tid; i < n; i += NT) tbuf[i] = exp(val_of(ckey[tail0 + i])) - exp_c;
                __syncthreads();
                gpdfit_block<NT>(tbuf, n, g, kk, sigma);
                smooth = is_finite(kk);  // psis.py:150
                __syncthreads();
            }
            // ---------------- smoothed tail (psis.py:153-157, _gpinv :211-222) and its exp-sum
            double ts = 0.0;
            if (smooth) {
                for (int i = tid; i < n; i += NT) {
                    const double pr = ((double)i + 0.5) / (double)n;
                    double q;
                    if (sigma <= 0.0) {
                        q = nan_f64();
                    } else {
                        const double l1 = log1p(-pr);
                        q = (fabs(kk) < 2.220446049250313e-16) ? -l1 : expm1(-kk * l1) / kk;
                        q *= sigma;
                    }
                    double y = q + exp_c;
                    double sm = log(y);
                    if (sm > 0.0) {  // psis.py:157
                        sm = 0.0;
                        y = 1.0;
                    }
                    tbuf[i] = sm;
                    ts += y;  // exp(log y) == y: the tail needs no exp
                }
            } else {
                for (int i = tid; i < n; i += NT) ts += exp(val_of(ckey[tail0 + i]));
            }
            tails = block_sum<NT>(ts, red);
            lse = log(body + tails);  // psis.py:158 / utils.py:349-357 (row max is 0 or the top smoothed value)
        } else if (MODE == MODE_LOO) {
            // rows with ll = -inf: still need lppd / variance sums below
            for (int s = tid; s < S; s += NT) {
                const double v = rbuf[s];
                lsum += exp(v - ll_max);
                const double d = v - ll_mean;
                vsum += d * d;
            }
            lsum = block_sum<NT>(lsum, red);
            vsum = block_sum<NT>(vsum, red);
        }

        // ---------------- outputs
        if (MODE == MODE_PSISLW) {
            if (run_psis) {
                for (int s = tid; s < S; s += NT) rbuf[s] = (rbuf[s] - mx) - lse;
                __syncthreads();
                if (smooth)
                    for (int i = tid; i < n; i += NT) rbuf[cidx[tail0 + i]] = tbuf[i] - lse;
            } else {
                const double qn = nan_f64();
                for (int s = tid; s < S; s += NT) rbuf[s] = qn;
            }
            if (tid == 0) p.k_out[row] = kk;
            double* dst = p.out + row * p.out_stride;
            if (p.use_bulk) {
                fence_proxy_async();
                __syncthreads();
                if (tid == 0) {
                    bulk_s2g(dst, rbuf, row_tx);
                    bulk_commit();
                }
            } else {
                __syncthreads();
                for (int s = tid; s < S; s += NT) dst[s] = rbuf[s];
            }
        } else {
            // elpd_i = LSE_s(lw_s + ll_s) (loo.py:289,319-324).  Body terms are the constant
            // -(mx + lse); tail terms differ from it by (smoothed - raw), so only n exps are needed.
            double elpd = nan_f64();
            if (run_psis) {
                double dmax = 0.0, es = 0.0;
                if (smooth) {
                    double dm = 0.0;
                    for (int i = tid; i < n; i += NT)
                        dm = fmax(dm, tbuf[i] - val_of(ckey[tail0 + i]));
                    dmax = block_max<NT>(dm, red);
                    for (int i = tid; i < n; i += NT)
                        es += exp((tbuf[i] - val_of(ckey[tail0 + i])) - dmax);
                    es = block_sum<NT>(es, red);
                } else {
                    es = (double)n;
                }
                const double tot = (double)(S - n) * exp(-dmax) + es;
                elpd = ((-mx - lse) + dmax) + log(tot);
                if (c_pinf > 0) elpd = nan_f64();  // lw = -inf, ll = +inf  =>  NaN term
            }
            double lppd = log(lsum) + (ll_max - log((double)S));  // utils.py:352-357, b_inv = S
            double var = vsum / (double)S;
            double lppdw = lppd;
            if (c_pinf + c_ninf > 0) {
                // waic.py:122-132: +-inf -> +-1e10 before lppd / variance (block-uniform, rare)
                double wmax = NEG_INF, wsum = 0.0;
                for (int s = tid; s < S; s += NT) {
                    double v = rbuf[s];
                    if (!is_finite(v)) v = (v > 0) ? 1e10 : -1e10;
                    wmax = fmax(wmax, v);
                    wsum += v;
                }
                wmax = block_max<NT>(wmax, red);
                const double wmean = block_sum<NT>(wsum, red) / (double)S;
                double e2 = 0.0, v2 = 0.0;
                for (int s = tid; s < S; s += NT) {
                    double v = rbuf[s];
                    if (!is_finite(v)) v = (v > 0) ? 1e10 : -1e10;
                    e2 += exp(v - wmax);
                    const double d = v - wmean;
                    v2 += d * d;
                }
                e2 = block_sum<NT>(e2, red);
                v2 = block_sum<NT>(v2, red);
                lppdw = log(e2) + (wmax - log((double)S));
                var = v2 / (double)S;
            }
            if (tid == 0) {
                p.k_out[row] = kk;
                p.elpd_i[row] = elpd;
                p.lppd_i[row] = lppd;
                p.var_i[row] = var;
                p.lppdw_i[row] = lppdw;
            }
            __syncthreads();  // all reads of rbuf are done before it is refilled
        }
        if (p.diag && tid == 0) {
            double* d = p.diag + row * DIAG_STRIDE;
            d[0] = mx; d[1] = c; d[2] = (double)n; d[3] = (double)C;
            d[4] = (double)attempts; d[5] = body; d[6] = tails; d[7] = sigma;
        }

        // ---------------- single-buffer mode: refill this buffer for the next row
        if (p.use_bulk && p.nbuf == 1 && nrow < p.n_rows && tid == 0) {
            if (MODE == MODE_PSISLW) bulk_wait_read0();
            fence_proxy_async();
            mbar_expect_tx(&bars[0], row_tx);
            bulk_g2s(smem_raw + L.off_row, p.in + nrow * p.in_stride, row_tx, &bars[0]);
        }
    }
    if (MODE == MODE_PSISLW && p.use_bulk && tid == 0) bulk_wait0();
}

}  // namespace b2l
