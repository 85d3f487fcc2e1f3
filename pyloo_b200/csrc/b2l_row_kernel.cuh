// PSIS row kernel (sm_100a): one CTA owns one observation at a time, its S draws staged in shared
// memory by 1-D bulk TMA (cp.async.bulk + mbarrier), persistent grid, optional double buffering.
//
// Replaces, per observation, the NumPy work of pyloo/psis.py:133-160 (`_psislw`), :181-208
// (`_gpdfit`), :211-231 (`_gpinv`) and pyloo/utils.py:344-359 (`_logsumexp`); in LOO mode also
// pyloo/loo.py:286-289,319-337 (`lw += ll`, `loo_i`, `lppd_i`) and pyloo/waic.py:137-145.
//
// Selection is exact.  x = fl(r - max r) is formed exactly as psis.py:134 does.  A threshold guessed
// from a 1-per-thread sample (warp-shuffle sorts) isolates ~1.9 (M+1) candidates; each candidate is
// packed as (31-bit linear quantisation of x - tau | draw index) into ONE 64-bit word, so a single
// shared-memory bitonic sort orders them by (value, index); quantisation collisions between
// distinct values are detected after the sort and sent to the full-key path, and a bit-wise binary
// search over the 64-bit ordered key space is the exact fallback for degenerate rows (ties,
// constant rows).  The (M+1)-th largest is then read from the sorted candidates.
#pragma once

#include "b2l_common.cuh"

namespace b2l {

enum : int { MODE_PSISLW = 0, MODE_LOO = 1 };

constexpr int GPD_MAX_GRID = 128;  // m = 30 + floor(sqrt(n)) <= 128  <=>  n <= 9603
constexpr int DIAG_STRIDE = 8;
constexpr int POOL_PER_WARP = 8;   // per-warp top sample keys pooled for the threshold guess

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
    int r0;        // initial pooled sample rank for the threshold guess
    int nbuf;      // row buffers per CTA (1 or 2)
    int use_bulk;  // 1: 16 B aligned rows -> bulk TMA; 0: cooperative LDG/STG
    int waic_only; // LOO mode: skip PSIS, only lppd / variance outputs
    int force_legacy;  // testing: always take the full-key (unpacked) candidate path
    double* row_ws; // non-null: rows too long for shared memory live here (gridDim.x rows of row_ld doubles)
    long long row_ld;
    int m_full;    // GPD grid size 30 + floor(sqrt(M)) for the common full tail n == M (host-computed)
    double cutoffmin;
    // non-null: process rows row_list[0 .. *n_list) instead of 0 .. n_rows (rows the split path handed over)
    const int* row_list;
    const int* n_list;
    // element stride inside a row (0 or 1: contiguous).  Larger values serve the rows handed over by the tile
    // path, which reads the observation-fastest (S, N) matrix where it lies: in_stride = 1, in_estride = stride_s,
    // use_bulk = 0
    long long in_estride;
    // optional [n_rows][tail_ld]: draw indices of the tail (psis.py:139-141), the rest -1 (evidence for tests)
    int* tail_idx;
    long long tail_ld;
};

struct RowSmemLayout {
    size_t row_bytes, off_row, off_ckey, off_cidx, off_tbuf, off_tx, off_ts, off_l1p, off_gb, off_gk, off_gw,
        off_gflag, off_part, off_red, off_ctl, off_bar, total;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Same carve-up on host (launch size) and device.
__host__ __device__ inline RowSmemLayout row_smem_layout(int S, int M, int cap, int nbuf, int nt) {
    RowSmemLayout L;
    size_t o = 0;
    L.row_bytes = align_up((size_t)S * 8, 128);
    L.off_row = o;
    o += L.row_bytes * (size_t)nbuf;
    L.off_ckey = o;  // packed candidates (u64) alias the full keys of the legacy path
    o += (size_t)cap * 8;
    L.off_cidx = o;
    o += align_up((size_t)cap * 4, 16);
    L.off_tbuf = o;  // tail t_i, then smoothed values
    o += align_up((size_t)(M + 1) * 8, 16);
    L.off_tx = o;    // tail raw x_i (ascending)
    o += align_up((size_t)(M + 1) * 8, 16);
    L.off_ts = o;    // tail draw indices
    o += align_up((size_t)(M + 1) * 4, 16);
    L.off_l1p = o;   // log1p(-(i + 0.5) / M), i < M: the _gpinv argument for a full tail
    o += align_up((size_t)(M + 1) * 8, 16);
    L.off_gb = o;
    o += GPD_MAX_GRID * 8;
    L.off_gk = o;
    o += GPD_MAX_GRID * 8;
    L.off_gw = o;
    o += GPD_MAX_GRID * 8;
    L.off_gflag = o;
    o += GPD_MAX_GRID * 4 * 2;  // flags + compact list of flagged grid points
    L.off_part = o;  // sample pool ((nt/32) * 8 keys) aliases the GPD partial products (nt * 12)
    o += align_up((size_t)nt * 12, 16);
    L.off_red = o;
    o += 128 * 8;
    L.off_ctl = o;
    o += 16 * 8;
    L.off_bar = o;
    o += 2 * 8;
    L.total = align_up(o, 128);
    return L;
}

// ------------------------------------------------------------------ fast exp for x <= 0
// exp(x), x in [-708, 0]: Cody-Waite reduction, degree-13 Taylor/Horner, exponent insertion.
// Max error ~1 ulp.  Anything below -708 (denormal results, -inf) goes to the library routine.
static __constant__ double c_exp_poly[12] = {
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};  // 1/13! .. 1/2!
static __constant__ double c_exp_red[4] = {1.4426950408889634074, 6755399441055744.0,
                                    -6.93147180369123816490e-01, -1.90821492927058770002e-10};

__device__ __forceinline__ double exp_nonpos(double x) {
    if (x < -708.0) return exp(x);
    // coefficients come from the constant bank (DFMA c[][] operand): no per-use 64-bit immediates
    const double t = fma(x, c_exp_red[0], c_exp_red[1]);
    const int ni = __double2loint(t);
    const double n = t - c_exp_red[1];
    double r = fma(n, c_exp_red[2], x);
    r = fma(n, c_exp_red[3], r);
    double p = c_exp_poly[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) p = fma(p, r, c_exp_poly[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
}

// ------------------------------------------------------------------ warp-level helpers
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    return (uint64_t)__shfl_xor_sync(FULL, (long long)v, m);
}
// ascending across lanes (lane 31 ends with the largest); 15 shuffle stages, no shared memory
__device__ __forceinline__ uint64_t warp_sort32(uint64_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint64_t o = shfl_xor_u64(v, j);
            const bool up = ((lane & k) == 0);
            const bool lower = ((lane & j) == 0);
            const uint64_t lo = (v < o) ? v : o, hi = (v < o) ? o : v;
            v = (up == lower) ? lo : hi;
        }
    }
    return v;
}

// ------------------------------------------------------------------ bitonic sorts (shared memory)
// single 64-bit words, ascending.  Sizes 256..2048 run a fully unrolled network: every stage's
// (k, j) is a compile-time constant, so the index arithmetic folds and the loop control vanishes.
template <int NT, int N, int K, int J>
__device__ __forceinline__ void bitonic_stage_const(uint64_t* key) {
#pragma unroll
    for (int t0 = 0; t0 < N / 2; t0 += NT) {
        const int t = t0 + (int)threadIdx.x;
        if (N / 2 >= NT || t < N / 2) {
            const int i = 2 * t - (t & (J - 1));
            const uint64_t a = key[i], b = key[i + J];
            if ((a > b) != ((i & K) != 0)) {
                key[i] = b;
                key[i + J] = a;
            }
        }
    }
    // a stage whose partners stay inside one warp's 64-element block only needs a warp barrier
    if (J <= 32 && J > 1 && N / 2 >= NT) __syncwarp();
    else __syncthreads();
}
template <int NT, int N, int K, int J>
struct BitonicJ {
    static __device__ __forceinline__ void run(uint64_t* key) {
        bitonic_stage_const<NT, N, K, J>(key);
        BitonicJ<NT, N, K, J / 2>::run(key);
    }
};
template <int NT, int N, int K>
struct BitonicJ<NT, N, K, 0> {
    static __device__ __forceinline__ void run(uint64_t*) {}
};
template <int NT, int N, int K>
struct BitonicK {
    static __device__ __forceinline__ void run(uint64_t* key) {
        BitonicK<NT, N, K / 2>::run(key);
        BitonicJ<NT, N, K, K / 2>::run(key);
    }
};
template <int NT, int N>
struct BitonicK<NT, N, 1> {
    static __device__ __forceinline__ void run(uint64_t*) {}
};
template <int NT>
__device__ void bitonic_sort_u64(uint64_t* key, int n) {
    switch (n) {
        case 256: BitonicK<NT, 256, 256>::run(key); return;
        case 512: BitonicK<NT, 512, 512>::run(key); return;
        case 1024: BitonicK<NT, 1024, 1024>::run(key); return;
        case 2048: BitonicK<NT, 2048, 2048>::run(key); return;
        default: break;
    }
    const int half = n >> 1;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += NT) {
                const int i = 2 * t - (t & (j - 1));
                uint64_t* pa = key + i;
                const uint64_t a = pa[0], b = pa[j];
                if ((a > b) != ((i & k) != 0)) {
                    pa[0] = b;
                    pa[j] = a;
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
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = ((i & k) == 0);
                const uint64_t a = key[i], b = key[l];
                const int ia = idx[i], ib = idx[l];
                const bool gt = (a > b) || (a == b && ia > ib);
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
    int* list;     // [GPD_MAX_GRID] compact list of flagged grid points
    double* part;  // [NT] partial products
    int* parte;    // [NT] partial exponents
    double* red;   // reduction scratch
    int* ctl;      // small control words
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
__device__ void gpdfit_block(const double* t, int n, int m_hint, GpdScratch g, double& k_out,
                             double& sigma_out) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int m = m_hint;                                        // psis.py:184 (host-computed when n == M)
    if (m <= 0) {
        m = (int)sqrt((double)n);
        while (m * m > n) --m;
        while ((m + 1) * (m + 1) <= n) ++m;
        m += 30;
    }
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

    // ---- product-form profile: thread = (chunk of t, grid point j), flattened over the CTA
    const int NCH = NT / m;                      // chunks (>= 2 since m <= 128 <= NT / 2)
    const int len = (n + NCH - 1) / NCH;
    {
        const int ch = tid / m, j = tid - ch * m;
        double P = 1.0;
        int E = 0, sgn = 0;
        bool ok = true;
        if (ch < NCH) {
            const double bmag = fmax(fabs(g.b[0]), fabs(g.b[m - 1]));
            const double fmx = 1.0 + bmag * tn;
            const int i0 = ch * len, i1 = min(n, i0 + len);
            const double nb = -g.b[j];
            if (fmx < 0x1p120) {  // common case: rescale every 8 factors, sign check by OR of the high words
                int i = i0;
                for (; i + 8 <= i1; i += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const double f = fma(nb, t[i + u], 1.0);
                        sgn |= __double2hiint(f);
                        P *= f;
                    }
                    ok = rescale_pos(P, E) && ok;
                }
                for (; i < i1; ++i) {
                    const double f = fma(nb, t[i], 1.0);
                    sgn |= __double2hiint(f);
                    P *= f;
                }
                ok = rescale_pos(P, E) && ok;
            } else {
                const int R = (fmx < 0x1p250) ? 4 : 1;
                int cnt = 0;
                for (int i = i0; i < i1; ++i) {
                    const double f = fma(nb, t[i], 1.0);
                    sgn |= __double2hiint(f);
                    P *= f;
                    if (++cnt == R) {
                        cnt = 0;
                        ok = rescale_pos(P, E) && ok;
                    }
                }
                ok = rescale_pos(P, E) && ok;
            }
            ok = ok && (sgn >= 0);
            if (!ok) g.flag[j] = 1;  // benign race: all writers store 1
        }
        g.part[tid] = P;
        g.parte[tid] = E;
    }
    __syncthreads();
    if (tid < m && !g.flag[tid]) {
        double P = 1.0;
        int E = 0;
        for (int ch = 0; ch < NCH; ++ch) {
            P *= g.part[ch * m + tid];
            E += g.parte[ch * m + tid];
            rescale_pos(P, E);
        }
        g.ks[tid] = log(P) + (double)E * 0.6931471805599453094;
    }
    // ---- compact list of flagged grid points (warp 0), then the literal log1p path for each
    if (wid == 0) {
        int nf = 0;
        for (int base = 0; base < m; base += 32) {
            const int j = base + lane;
            const bool f = (j < m) && g.flag[j];
            const unsigned mask = __ballot_sync(FULL, f);
            if (f) g.list[nf + __popc(mask & ((1u << lane) - 1u))] = j;
            nf += __popc(mask);
        }
        if (lane == 0) g.ctl[1] = nf;
    }
    __syncthreads();
    const int nflag = g.ctl[1];
    for (int q = 0; q < nflag; ++q) {
        const int j = g.list[q];
        const double nb = -g.b[j];
        double acc = 0.0;
        for (int i = tid; i < n; i += NT) acc += log1p(nb * t[i]);
        acc = block_sum<NT>(acc, g.red);
        if (tid == 0) g.ks[j] = acc;
    }
    if (nflag) __syncthreads();
    // ---- profile log-likelihood L_j (psis.py:191)
    bool fin = true;
    if (tid < m) {
        const double kj = g.ks[tid] / (double)n;
        const double Lj = (double)n * (log(-(g.b[tid] / kj)) - kj - 1.0);
        g.ks[tid] = Lj;
        fin = is_finite(Lj);
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
__global__ void __launch_bounds__(NT, (NT == 128) ? 4 : ((NT == 256) ? 3 : 1)) psis_row_kernel(const RowParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    const RowSmemLayout L = row_smem_layout(p.S, p.M, p.cap, p.nbuf, NT);
    uint64_t* ckey = reinterpret_cast<uint64_t*>(smem_raw + L.off_ckey);  // packed or full keys
    int* cidx = reinterpret_cast<int*>(smem_raw + L.off_cidx);
    double* tbuf = reinterpret_cast<double*>(smem_raw + L.off_tbuf);
    double* tx = reinterpret_cast<double*>(smem_raw + L.off_tx);
    int* ts = reinterpret_cast<int*>(smem_raw + L.off_ts);
    double* l1p = reinterpret_cast<double*>(smem_raw + L.off_l1p);
    uint64_t* pool = reinterpret_cast<uint64_t*>(smem_raw + L.off_part);
    double* red = reinterpret_cast<double*>(smem_raw + L.off_red);
    int* ctl = reinterpret_cast<int*>(smem_raw + L.off_ctl);
    uint64_t* ctl64 = reinterpret_cast<uint64_t*>(smem_raw + L.off_ctl + 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
    GpdScratch g;
    g.b = reinterpret_cast<double*>(smem_raw + L.off_gb);
    g.ks = reinterpret_cast<double*>(smem_raw + L.off_gk);
    g.w = reinterpret_cast<double*>(smem_raw + L.off_gw);
    g.flag = reinterpret_cast<int*>(smem_raw + L.off_gflag);
    g.list = g.flag + GPD_MAX_GRID;
    g.part = reinterpret_cast<double*>(smem_raw + L.off_part);
    g.parte = reinterpret_cast<int*>(smem_raw + L.off_part + (size_t)NT * 8);
    g.red = red;
    g.ctl = ctl;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int S = p.S, M = p.M, cap = p.cap;
    const int S2 = S >> 1;
    const uint32_t row_tx = (uint32_t)S * 8u;
    const double NEG_INF = -inf_f64();

    // nothing to do for this CTA (the usual case when it only serves the split path's hand-over list)
    if ((long long)blockIdx.x >= (p.row_list ? (long long)*p.n_list : p.n_rows)) return;
    if (p.use_bulk && tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    // per-CTA table for the smoothing step: depends only on (i, M), psis.py:153 + :221
    if (MODE == MODE_PSISLW || !p.waic_only)
        for (int i = tid; i < M; i += NT) l1p[i] = log1p(-(((double)i + 0.5) / (double)M));
    __syncthreads();

    const long long n_items = p.row_list ? (long long)*p.n_list : p.n_rows;
    auto row_of = [&](long long item) -> long long { return p.row_list ? (long long)p.row_list[item] : item; };
    long long item = blockIdx.x;
    if (p.use_bulk && tid == 0 && item < n_items) {
        mbar_expect_tx(&bars[0], row_tx);
        bulk_g2s(smem_raw + L.off_row, p.in + row_of(item) * p.in_stride, row_tx, &bars[0]);
    }

    for (int it = 0; item < n_items; item += gridDim.x, ++it) {
        const long long row = row_of(item);
        const int bsel = (p.nbuf == 2) ? (it & 1) : 0;
        double* rbuf = p.row_ws ? p.row_ws + (size_t)blockIdx.x * p.row_ld
                                : reinterpret_cast<double*>(smem_raw + L.off_row + (size_t)bsel * L.row_bytes);
        const double2* rbuf2 = reinterpret_cast<const double2*>(rbuf);
        const long long nitem = item + gridDim.x;
        const long long nrow = (nitem < n_items) ? row_of(nitem) : 0;

        // ---------------- stage the row
        if (p.use_bulk) {
            if (p.nbuf == 2 && nitem < n_items && tid == 0) {
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
            if (p.in_estride > 1) {
                for (int s = tid; s < S; s += NT) rbuf[s] = src[(long long)s * p.in_estride];
            } else {
                for (int s = tid; s < S; s += NT) rbuf[s] = src[s];
            }
            __syncthreads();
        }

        // ---------------- pass A (fast): extrema and sum; any NaN/inf only raises a flag
        double a_max = NEG_INF, a_min = inf_f64(), a_sum = 0.0;
        int c_nan = 0, c_pinf = 0, c_ninf = 0;
        {
            double m0 = NEG_INF, m1 = NEG_INF, n0 = inf_f64(), n1 = inf_f64(), s0 = 0.0, s1 = 0.0;
            int spec = 0;  // max over |hi word|: >= 0x7ff00000 <=> some inf / NaN
            for (int i = tid; i < S2; i += NT) {
                const double2 v = rbuf2[i];
                m0 = (v.x > m0) ? v.x : m0;
                m1 = (v.y > m1) ? v.y : m1;
                spec = max(spec, max(__double2hiint(v.x) & 0x7fffffff, __double2hiint(v.y) & 0x7fffffff));
                if (MODE == MODE_LOO) {
                    n0 = (v.x < n0) ? v.x : n0;
                    n1 = (v.y < n1) ? v.y : n1;
                    s0 += v.x;
                    s1 += v.y;
                }
            }
            if ((S & 1) && tid == 0) {
                const double v = rbuf[S - 1];
                m0 = (v > m0) ? v : m0;
                spec = max(spec, __double2hiint(v) & 0x7fffffff);
                if (MODE == MODE_LOO) {
                    n0 = (v < n0) ? v : n0;
                    s0 += v;
                }
            }
            a_max = block_max<NT>(fmax(m0, m1), red);
            if (MODE == MODE_LOO) {
                a_min = block_min<NT>(fmin(n0, n1), red);
                a_sum = block_sum<NT>(s0 + s1, red);
            }
            const int sp = __syncthreads_or(spec >= 0x7ff00000) ? 0x7ff00000 : 0;
            if (sp >= 0x7ff00000) {
                // ---------------- census (slow, rare): NaN / inf counts, NaN -> -1e10 in LOO mode
                a_max = NEG_INF; a_min = inf_f64(); a_sum = 0.0;
                for (int s = tid; s < S; s += NT) {
                    double v = rbuf[s];
                    if (v != v) {
                        ++c_nan;
                        if (MODE == MODE_LOO) {  // pyloo/loo.py:227
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
                a_max = block_max<NT>(a_max, red);
                a_min = block_min<NT>(a_min, red);
                a_sum = block_sum<NT>(a_sum, red);
                c_nan = block_isum<NT>(c_nan, red);
                c_pinf = block_isum<NT>(c_pinf, red);
                c_ninf = block_isum<NT>(c_ninf, red);
                __syncthreads();
                if (tid == 0 && p.counters) {
                    if (c_nan) atomicAdd(&p.counters[0], (unsigned long long)c_nan);
                    if (c_pinf) atomicAdd(&p.counters[1], (unsigned long long)c_pinf);
                    if (c_ninf) atomicAdd(&p.counters[2], (unsigned long long)c_ninf);
                }
            }
        }

        // r = raw log weight: PSISLW r = v ; LOO r = -ll (pyloo/loo.py:286-288)
        const double mx = (MODE == MODE_LOO) ? -a_min : a_max;  // max_s r_s
        const double ll_max = a_max;
        const double ll_mean = a_sum / (double)S;
        auto xval = [&](int s) -> double {  // x_s = fl(r_s - max r), psis.py:134
            const double v = rbuf[s];
            return ((MODE == MODE_LOO) ? -v : v) - mx;
        };

        // Rows on which the reference produces no tail at all (NaN / inf arithmetic, App. D)
        bool run_psis;
        if (MODE == MODE_PSISLW) run_psis = (c_nan == 0) && is_finite(mx);
        else run_psis = (c_ninf == 0) && !p.waic_only;  // ll = -inf  =>  r = +inf  =>  x = NaN, k = inf

        double kk = inf_f64(), sigma = nan_f64(), lse = nan_f64();
        double c = nan_f64(), body = 0.0, tails = 0.0, lsum = 0.0, vsum = 0.0;
        int n = 0, C = 0, attempts = 0;
        bool smooth = false;

        if (run_psis) {
            // ---------------- threshold guess: 1 sample per thread, warp sorts, pooled per-warp top 8
            const bool sampling = S > cap;
            int R = p.r0;
            if (sampling) {
                const int idx = (int)(((long long)tid * S) / NT);
                const uint64_t sk = warp_sort32(key_of(xval(idx)));
                if (lane >= 32 - POOL_PER_WARP) pool[wid * POOL_PER_WARP + (lane - (32 - POOL_PER_WARP))] = sk;
                __syncthreads();
            }
            bool exact = false, packed = false;
            double tau = NEG_INF, craw_exact = NEG_INF, qscale = 0.0;
            while (true) {
                if (sampling && !exact) {
                    // R-th largest of the pooled keys: warp 0 ranks by counting (pool is tiny)
                    if (wid == 0) {
                        constexpr int PN = NW * POOL_PER_WARP;
                        constexpr int PE = (PN + 31) / 32;
                        uint64_t mine[PE];
                        int gt[PE];
#pragma unroll
                        for (int e = 0; e < PE; ++e) {
                            const int i = e * 32 + lane;
                            mine[e] = (i < PN) ? pool[i] : 0ull;
                            gt[e] = 0;
                        }
                        for (int q = 0; q < PN; ++q) {
                            const uint64_t o = pool[q];
#pragma unroll
                            for (int e = 0; e < PE; ++e) {
                                const int i = e * 32 + lane;
                                gt[e] += (o > mine[e] || (o == mine[e] && q < i)) ? 1 : 0;
                            }
                        }
                        const int want = min(R, PN) - 1;
#pragma unroll
                        for (int e = 0; e < PE; ++e)
                            if (e * 32 + lane < PN && gt[e] == want) ctl64[0] = mine[e];
                    }
                    __syncthreads();
                    tau = val_of(ctl64[0]);
                }
                // packed candidates: 31-bit linear quantisation of (x - tau) over (tau, 0] | draw index
                qscale = 0x1p31 / (0.0 - tau);
                packed = sampling && !p.force_legacy && (tau < 0.0) && (qscale < 0x1p900);
                if (tid == 0) ctl[0] = 0;
                __syncthreads();
                // -------- pass B: x = fl(r - mx); body exp-sum (+ LOO lppd / variance sums).  Candidates
                // (x > tau, ~9 % of the draws) are only marked in a per-thread bit mask here and
                // replayed afterwards, so the streaming loop carries no compaction code.
                double bsum0 = 0.0, bsum1 = 0.0, ls0 = 0.0, ls1 = 0.0, vs0 = 0.0, vs1 = 0.0;
                // segments of <= 32 iterations (64 NT draws) so the mask fits 64 bits for any S
                for (int seg = 0; seg < S2; seg += 32 * NT) {
                    unsigned long long cmask = 0ull;  // bit 2*it + e  <=>  element e of iteration it
                    const int seg_end = min(S2, seg + 32 * NT);
                    {
                        int itb = 0;
                        for (int i2 = seg + tid; i2 < seg_end; i2 += NT, ++itb) {
                            const double2 v = rbuf2[i2];
                            const double x0 = ((MODE == MODE_LOO) ? -v.x : v.x) - mx;
                            const double x1 = ((MODE == MODE_LOO) ? -v.y : v.y) - mx;
                            const bool c0 = x0 > tau, c1 = x1 > tau;
                            cmask |= (unsigned long long)((c0 ? 1u : 0u) | (c1 ? 2u : 0u)) << (2 * itb);
                            const double e0 = exp_nonpos(x0), e1 = exp_nonpos(x1);
                            bsum0 += c0 ? 0.0 : e0;
                            bsum1 += c1 ? 0.0 : e1;
                            if (MODE == MODE_LOO) {
                                ls0 += exp_nonpos(v.x - ll_max);  // loo.py:329-337 / utils.py:349-351
                                ls1 += exp_nonpos(v.y - ll_max);
                                const double d0 = v.x - ll_mean, d1 = v.y - ll_mean;  // waic.py:145
                                vs0 = fma(d0, d0, vs0);
                                vs1 = fma(d1, d1, vs1);
                            }
                        }
                    }
                    // replay: one smem atomic per warp reserves a slot range, then each thread emits its own
                    const int mine = __popcll(cmask);
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += u;
                    }
                    int base = 0;
                    if (lane == 31) base = atomicAdd(&ctl[0], incl);
                    base = __shfl_sync(FULL, base, 31);
                    int pos = base + incl - mine;
                    while (cmask) {
                        const int bpos = __ffsll((long long)cmask) - 1;
                        cmask &= cmask - 1;
                        const int s = 2 * (seg + (bpos >> 1) * NT + tid) + (bpos & 1);
                        if (pos < cap) {
                            const double x = xval(s);
                            if (packed) {
                                const uint32_t q = __double2uint_rz((x - tau) * qscale) + 1u;
                                ckey[pos] = ((uint64_t)q << 32) | (uint32_t)s;
                            } else {
                                ckey[pos] = key_of(x);
                                cidx[pos] = s;
                            }
                        }
                        ++pos;
                    }
                }
                if ((S & 1) && tid == 0) {  // odd S: last draw
                    const int s = S - 1;
                    const double v = rbuf[s];
                    const double x = ((MODE == MODE_LOO) ? -v : v) - mx;
                    if (x > tau) {
                        const int pos = atomicAdd(&ctl[0], 1);
                        if (pos < cap) {
                            if (packed) {
                                const uint32_t q = __double2uint_rz((x - tau) * qscale) + 1u;
                                ckey[pos] = ((uint64_t)q << 32) | (uint32_t)s;
                            } else {
                                ckey[pos] = key_of(x);
                                cidx[pos] = s;
                            }
                        }
                    } else {
                        bsum0 += exp_nonpos(x);
                    }
                    if (MODE == MODE_LOO) {
                        ls0 += exp_nonpos(v - ll_max);
                        const double d0 = v - ll_mean;
                        vs0 = fma(d0, d0, vs0);
                    }
                }
                body = block_sum<NT>(bsum0 + bsum1, red);
                lsum = ls0 + ls1;
                vsum = vs0 + vs1;
                C = ctl[0];
                __syncthreads();
                if (!sampling || exact) break;
                if (C >= M + 1 && C <= cap) break;
                ++attempts;
                const int Rn = (C < M + 1) ? min(NW * POOL_PER_WARP, 2 * R + 8) : max(1, R >> 1);
                if (attempts >= 3 || Rn == R) {
                    // -------- exact fallback: (M+1)-th largest key by bit-wise binary search
                    uint64_t K = 0;
                    for (int bit = 63; bit >= 0; --bit) {
                        const uint64_t trial = K | (1ull << bit);
                        int cnt = 0;
                        for (int s = tid; s < S; s += NT) cnt += (key_of(xval(s)) >= trial) ? 1 : 0;
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

            // ---------------- sort candidates ascending by (value, index); pads (0) first
            int P2 = 1;
            while (P2 < C) P2 <<= 1;
            for (int i = C + tid; i < P2; i += NT) {
                ckey[i] = 0ull;
                if (!packed) cidx[i] = -1;
            }
            __syncthreads();
            if (packed) {
                if (P2 > 1) bitonic_sort_u64<NT>(ckey, P2);
                // equal quantised values must be equal doubles, else order is not trustworthy
                int bad = 0;
                for (int i = P2 - C + 1 + tid; i < P2; i += NT) {
                    const uint64_t a = ckey[i - 1], b = ckey[i];
                    if ((a >> 32) == (b >> 32) && xval((int)(uint32_t)a) != xval((int)(uint32_t)b)) bad = 1;
                }
                bad = __syncthreads_or(bad);
                if (bad) {  // rare: rebuild full keys in place and use the (key, index) sort
                    for (int i = tid; i < P2; i += NT) {
                        const uint64_t a = ckey[i];
                        if (i >= P2 - C) {
                            const int s = (int)(uint32_t)a;
                            ckey[i] = key_of(xval(s));
                            cidx[i] = s;
                        } else {
                            cidx[i] = -1;
                        }
                    }
                    __syncthreads();
                    packed = false;
                }
            }
            if (!packed && P2 > 1) bitonic_sort_pairs<NT>(ckey, cidx, P2);
            auto cand_idx = [&](int i) -> int { return packed ? (int)(uint32_t)ckey[i] : cidx[i]; };
            auto cand_x = [&](int i) -> double { return packed ? xval((int)(uint32_t)ckey[i]) : val_of(ckey[i]); };

            // ---------------- cutoff (psis.py:135-136) and tail (psis.py:139-141)
            double c_raw;
            if (exact) c_raw = craw_exact;
            else c_raw = (C >= M + 1) ? cand_x(P2 - M - 1) : NEG_INF;
            c = fmax(c_raw, p.cutoffmin);
            int tail0;
            {
                int lo = P2 - C, hi = P2;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (cand_x(mid) > c) hi = mid; else lo = mid + 1;
                }
                tail0 = lo;
                n = P2 - lo;
            }
            const double exp_c = exp(c);  // psis.py:138
            // tail in ascending order: raw x_i, draw index, t_i = exp(x_i) - exp(c) (psis.py:146-147);
            // candidates at or below the cutoff belong to the body sum
            {
                double b2 = 0.0;
                for (int i = P2 - C + tid; i < P2; i += NT) {
                    const double xi = cand_x(i);
                    if (i < tail0) {
                        b2 += exp_nonpos(xi);
                    } else {
                        tx[i - tail0] = xi;
                        ts[i - tail0] = cand_idx(i);
                        tbuf[i - tail0] = exp(xi) - exp_c;
                    }
                }
                body += block_sum<NT>(b2, red);  // (syncs: tx / ts / tbuf visible)
            }

            // ---------------- GPD fit on the tail (psis.py:146-148)
            if (n > 4) {
                gpdfit_block<NT>(tbuf, n, (n == M) ? p.m_full : 0, g, kk, sigma);
                smooth = is_finite(kk);  // psis.py:150
                __syncthreads();
            }
            // ---------------- smoothed tail (psis.py:153-157, _gpinv :211-222) and its exp-sum
            double tsm = 0.0;
            if (smooth) {
                for (int i = tid; i < n; i += NT) {
                    const double pr = ((double)i + 0.5) / (double)n;
                    double q;
                    if (sigma <= 0.0) {
                        q = nan_f64();
                    } else {
                        const double l1 = (n == M) ? l1p[i] : log1p(-pr);
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
                    tsm += y;  // exp(log y) == y: the tail needs no exp
                }
            } else {
                for (int i = tid; i < n; i += NT) tsm += exp(tx[i]);
            }
            tails = block_sum<NT>(tsm, red);
            lse = log(body + tails);  // psis.py:158 / utils.py:349-357 (row max is 0 or the top smoothed value)
        } else if (MODE == MODE_LOO) {
            // rows without PSIS (ll = -inf, or WAIC-only): still need the lppd / variance sums
            double ls0 = 0.0, ls1 = 0.0, vs0 = 0.0, vs1 = 0.0;
            for (int i = tid; i < S2; i += NT) {
                const double2 v = rbuf2[i];
                ls0 += exp(v.x - ll_max);
                ls1 += exp(v.y - ll_max);
                const double d0 = v.x - ll_mean, d1 = v.y - ll_mean;
                vs0 = fma(d0, d0, vs0);
                vs1 = fma(d1, d1, vs1);
            }
            if ((S & 1) && tid == 0) {
                const double v = rbuf[S - 1];
                ls0 += exp(v - ll_max);
                const double d0 = v - ll_mean;
                vs0 = fma(d0, d0, vs0);
            }
            lsum = block_sum<NT>(ls0 + ls1, red);
            vsum = block_sum<NT>(vs0 + vs1, red);
        }

        // ---------------- outputs
        if (MODE == MODE_PSISLW) {
            if (run_psis) {
                double2* w2 = reinterpret_cast<double2*>(rbuf);
                for (int i = tid; i < S2; i += NT) {
                    double2 v = w2[i];
                    v.x = (v.x - mx) - lse;
                    v.y = (v.y - mx) - lse;
                    w2[i] = v;
                }
                if ((S & 1) && tid == 0) rbuf[S - 1] = (rbuf[S - 1] - mx) - lse;
                __syncthreads();
                if (smooth)
                    for (int i = tid; i < n; i += NT) rbuf[ts[i]] = tbuf[i] - lse;
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
                    for (int i = tid; i < n; i += NT) dm = fmax(dm, tbuf[i] - tx[i]);
                    dmax = block_max<NT>(dm, red);
                    for (int i = tid; i < n; i += NT) es += exp((tbuf[i] - tx[i]) - dmax);
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
        if (p.tail_idx) {  // (ts is written only on the PSIS branch, where n > 0 implies it is filled)
            int* d = p.tail_idx + row * p.tail_ld;
            const int nt = run_psis ? n : 0;
            for (int i = tid; i < (int)p.tail_ld; i += NT) d[i] = (i < nt) ? ts[i] : -1;
            __syncthreads();  // ts is reused by the next row
        }

        // ---------------- single-buffer mode: refill this buffer for the next row
        if (p.use_bulk && p.nbuf == 1 && nitem < n_items && tid == 0) {
            if (MODE == MODE_PSISLW) bulk_wait_read0();
            fence_proxy_async();
            mbar_expect_tx(&bars[0], row_tx);
            bulk_g2s(smem_raw + L.off_row, p.in + nrow * p.in_stride, row_tx, &bars[0]);
        }
    }
    if (MODE == MODE_PSISLW && p.use_bulk && tid == 0) bulk_wait0();
}

}  // namespace b2l
