// Split PSIS path (sm_100a): a streaming kernel and a tail kernel per round of observations.
//
//   psis_stream_kernel  one CTA per observation.  The S draws arrive by 1-D bulk TMA in a padded
//                       shared-memory row: max (pyloo/psis.py:134), a threshold below the (M+1)-th
//                       largest draw estimated from the per-thread maxima, then one pass that forms
//                       x = fl(r - max r), sums exp(x) over the body (pyloo/utils.py:349-351) and
//                       emits the ~1.15 (M+1) candidate draws above the threshold (exact x + draw
//                       index) to a per-round scratch.  LOO mode adds the lppd / variance sums
//                       (pyloo/loo.py:329-337, pyloo/waic.py:137-145).  In psislw mode one extra
//                       warp per CTA applies the PREVIOUS round: out = (r - max r) - lse with the
//                       smoothed tail on top (psis.py:156-158), bulk TMA in and out.
//   psis_tail_kernel    one WARP per observation, no block barriers: register bitonic sort of
//                       32-bit quantised keys + exact re-ranking of equal-key runs, cutoff / tail
//                       (psis.py:135-141), Zhang-Stephens GPD fit (psis.py:181-208), _gpinv
//                       smoothing (psis.py:149-157,211-222), normaliser (psis.py:158); output: lse, k
//                       and the patch list (psislw) or elpd_i / lppd_i / var_i (loo.py:319-337).
//   psis_apply_kernel   the apply stage alone (after the last round; NT = 1024 shapes).
//
// Rows the fast path cannot decide exactly (NaN / inf, range > 1e7, threshold retries exhausted, long
// runs of equal keys at the cutoff, non-finite GPD profiles) are appended to a list and re-done by
// the general row kernel (b2l_row_kernel.cuh), which handles every case.
#pragma once

#include <type_traits>

#include "b2l_row_kernel.cuh"

namespace b2l {

// why rows leave the split path (debug histogram, read by b2l_handover_reasons): one copy per kernel TU
enum : int { HO_SPECIAL = 1, HO_RANGE = 2, HO_RETRY = 3, HO_RUNS = 4, HO_ORDER = 5, HO_GPD_TQ = 6, HO_GPD_FMX = 7,
             HO_GPD_PROD = 8, HO_GPD_PROFILE = 9, HO_CANCEL = 10, HO_REASONS = 16 };
static __device__ unsigned long long g_handover[HO_REASONS];
__device__ __forceinline__ void note_handover(int reason) { atomicAdd(&g_handover[reason], 1ull); }

constexpr int KEY_IDX_BITS = 14;               // draw indices travel as 16-bit values: S <= 2^14
constexpr int SPLIT_MAX_S = 1 << KEY_IDX_BITS;

struct __align__(16) SplitHeader {  // 96 B per observation: stream kernel -> tail kernel -> apply kernel
    double mx;      // max_s r_s
    double body;    // sum of exp(x_s) over the draws that are NOT candidates
    double lsum;    // LOO: sum_s exp(ll_s - lshift)
    double vsum;    // LOO: sum_s (ll_s - mean ll)^2
    double lshift;  // LOO: shift used for lsum (min ll, or max ll for very wide rows)
    double taux;    // candidate threshold: every candidate has x >= taux
    double lse;     // tail kernel -> apply kernel: the normaliser (psis.py:158)
    int C;          // candidates emitted (M + 1 <= C <= cap)
    int flags;      // != 0: the row was handed to the general kernel
    int attempts;
    int n_patch;    // tail kernel -> apply kernel: smoothed draws to patch in (0: none)
    int C2;         // tile path: candidates of the looser second list, stored from the END of the row's scratch
    int pad_;
    double vmax;    // chunked tile path: every draw that is NOT a candidate has x < vmax (the largest of the chunks'
                    // thresholds); the list holds the tail only if its (M + 1)-th largest x reaches vmax
    double pad2_;
};
static_assert(sizeof(SplitHeader) == 96, "SplitHeader is 96 bytes");

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
    int nbuf;              // stream kernel row buffers in shared memory (2: prefetch during the whole row)
    int q0;                // per-warp rank (1..32) of the thread maxima used for the threshold guess
    int m_full;            // 30 + floor(sqrt(M))
    double cutoffmin;
    SplitHeader* hdr;          // [n_rows]
    double* cx;                // [n_rows][cap] candidate x (exact); then the smoothed values to patch in
    unsigned short* cs;        // [n_rows][cap] candidate draw index; then the patch positions
    int* fb_list;              // [n_rows] rows for the general kernel
    int* fb_count;
    unsigned long long* counters;  // optional [4]; [3] += rows handed to the general kernel
    long long row_base;        // index of row 0 of this batch in the caller's arrays (hand-over list entries)
    // apply stage fused into the stream kernel (psislw): rows of the PREVIOUS batch, whose tail
    // kernel has finished -- one extra warp per CTA streams them through shared memory with bulk TMA
    const double* a_in;
    double* a_out;
    const SplitHeader* a_hdr;
    const double* a_cx;
    const unsigned short* a_cs;
    long long a_rows;
    int a_chunk;               // draws per apply transfer (S, or a fraction of it when shared memory is short)
    // tile path (b2l_tile.cu): `body` holds the sum over ALL draws (the tail kernel subtracts the raw tail terms)
    // and the candidates come as two lists: C tight ones from the front of the scratch row, C2 looser ones from
    // its end (only read when the tight list is shorter than M + 1 or longer than one sort)
    int total_body;
    int ab_lists;
    int chunked;  // tile path, long posteriors: ONE list of raw r = -ll (x = fl(r - max r) is formed here, psis.py:134), valid only if
                  // the cutoff reaches SplitHeader.vmax
    // optional [n_rows][tail_ld]: draw indices of the tail (psis.py:139-141), the rest -1 (evidence for tests)
    int* tail_idx;
    long long tail_ld;
    double log_S;  // np.log(n_samples), computed by the host like the reference does (utils.py:353)
};

// (ExpTab, exp_poly5, scale2 and exp_tab_drop live in b2l_common.cuh: the importance-sampling kernels use them too)
__device__ __forceinline__ void exp_tab_pm_drop(double x, const ExpTab& tb, bool drop_p, bool drop_m, double& ep,
                                                double& em) {
    const double t = fma(x, EXP_L, EXP_MAGIC);
    const int ni = max(__double2loint(t), -1022 * 32);
    const double nf = t - EXP_MAGIC;
    const double r = fma(nf, EXP_C1, x);
    const int j = ni & 31, sh = (ni & ~31) << 15;
    const double yp = tb.t[j] * exp_poly5(r);
    const double ym = tb.tinv[j] * exp_poly5(-r);
    ep = __hiloint2double(drop_p ? 0 : __double2hiint(yp) + sh, __double2loint(yp));
    em = __hiloint2double(drop_m ? 0 : __double2hiint(ym) - sh, __double2loint(ym));
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

// log(y) for the smoothed tail: y = 2^e m, m = c_j (1 + r) with c_j the centre of the j-th of 64
// mantissa intervals: log y = e ln2 + log c_j + log1p(r), |r| <= 2^-7, degree-6 series (absolute
// error < 5e-16).  tab: [0..63] 1/c_j (rounded), [64..127] -log(1/c_j).  Zero, denormal, negative,
// inf and NaN arguments take the library routine.
__device__ __forceinline__ void log_tab_init(double* tab, int j) {  // j < 64
    const double ic = 1.0 / (1.0 + ((double)j + 0.5) / 64.0);
    tab[j] = ic;
    tab[64 + j] = -log(ic);
}
static __device__ __noinline__ double log_slow(double y) { return log(y); }
static __device__ __noinline__ double exp_any(double x) { return exp(x); }  // (one out-of-line copy of the library exp)
__device__ __forceinline__ double log_tab(double y, const double* tab) {
    const int hi = __double2hiint(y);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log_slow(y);
    const int e = (hi >> 20) - 1023;
    const int j = (hi >> 14) & 63;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(y));
    const double r = fma(m, tab[j], -1.0);
    double p = -1.0 / 6.0;
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    return fma((double)e, 0.6931471805599453094, fma(p, r, tab[64 + j]));
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
    size_t row_bytes, off_apply, off_tab, off_red, off_ctl, off_bar, total;
};
// `slots` = NT * EPT: the stream row buffers are padded to one slot per (thread, register) so the passes
// carry no per-slot bounds logic; the pad is written once per CTA and never touched by the TMA loads
__host__ __device__ inline StreamSmem stream_smem(int S, int slots, int nbuf, int apply_chunk) {
    StreamSmem L;
    L.row_bytes = align_up((size_t)(slots > S ? slots : S) * 8, 128);
    size_t o = L.row_bytes * (size_t)nbuf;
    L.off_apply = o;  // one row buffer of the apply warp
    o += align_up((size_t)apply_chunk * 8, 128);  // apply_chunk = 0: no apply warp
    L.off_tab = o;   // 64 doubles
    o += 64 * 8;
    L.off_red = o;   // 2 x (5 x 32 doubles + 32 floats): sets alternate between consecutive rows
    o += 2 * (5 * 32 * 8 + 32 * 4);
    L.off_ctl = o;
    o += 16 * 4;
    L.off_bar = o;
    o += 32;
    L.total = align_up(o, 128);
    return L;
}

// named barriers for the stream warps (the apply warp of the same CTA never joins them)
__device__ __forceinline__ void bar_sync_n(int nthreads) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
__device__ __forceinline__ bool bar_or_n(int nthreads, bool pred) {
    int r;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.s32 p, %2, 0;\n\t"
        "bar.red.or.pred q, 1, %1, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t"
        "}"
        : "=r"(r)
        : "r"(nthreads), "r"((int)pred)
        : "memory");
    return r != 0;
}

// compare-select max / min: fmax / fmin on doubles expand to ~8 instructions each (IEEE NaN rules);
// rows that hold a NaN are detected separately and leave the fast path, so plain selects are enough
__device__ __forceinline__ double max_sel(double a, double b) { return (b > a) ? b : a; }
__device__ __forceinline__ double min_sel(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double warp_max_sel(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max_sel(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
template <int NW>
__device__ __forceinline__ double slots_max(const double* s, int lane) {
    if (NW <= 8) {
        double r = s[0];
#pragma unroll
        for (int i = 1; i < NW; ++i) r = max_sel(r, s[i]);
        return r;
    }
    return warp_max_sel(lane < NW ? s[lane] : -inf_f64());
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

#ifndef B2L_STREAM_OCC
#define B2L_STREAM_OCC 3  // resident 256(+32)-thread CTAs per SM the stream kernel is compiled for
#endif
// psislw stream CTAs carry one extra warp for the fused apply stage (not at 1024 threads: block limit)
__host__ __device__ constexpr bool stream_has_apply(int nt, int mode) { return mode == MODE_PSISLW && nt <= 512; }
__host__ __device__ constexpr int stream_block(int nt, int mode) { return nt + (stream_has_apply(nt, mode) ? 32 : 0); }
template <int NT>
constexpr int stream_min_blocks() {
    return NT == 128 ? 2 * B2L_STREAM_OCC : (NT == 256 ? B2L_STREAM_OCC : (NT == 512 ? (B2L_STREAM_OCC + 1) / 2 : 1));
}

// ------------------------------------------------------------------ stream kernel
// The apply warp (psislw only, threads NT .. NT + 31): out = (r - max r) - lse (psis.py:134,158)
// plus the smoothed tail (psis.py:156) for the rows of the previous batch, one row at a time through
// its own shared-memory buffer: bulk TMA load, in-place transform, patches, bulk TMA store.
__device__ __forceinline__ void apply_warp_loop_chunked(const SplitParams& p, unsigned char* abuf_raw, uint64_t* abar,
                                                int lane) {
    const int S = p.S, chunk = p.a_chunk;
    double* abuf = reinterpret_cast<double*>(abuf_raw);
    double2* abuf2 = reinterpret_cast<double2*>(abuf_raw);
    uint32_t phase = 0;
    for (long long row = blockIdx.x; row < p.a_rows; row += gridDim.x) {
        const SplitHeader* h = p.a_hdr + row;
        if (h->flags) continue;  // handed to the general kernel (warp-uniform)
        const double mx = h->mx, lse = h->lse;
        const int np = h->n_patch;
        // the row in pieces of `chunk` draws (one piece unless shared memory is short)
        for (int off = 0; off < S; off += chunk) {
            const int len = min(chunk, S - off);  // even
            const uint32_t bytes = (uint32_t)len * 8u;
            if (lane == 0) {
                mbar_expect_tx(abar, bytes);
                bulk_g2s(abuf_raw, p.a_in + row * p.in_stride + off, bytes, abar);
            }
            mbar_wait(abar, phase);
            phase ^= 1u;
            // what this warp loads next: HBM -> L2 while this piece is transformed and stored
            if (lane == 0) {
                if (off + chunk < S)
                    bulk_prefetch_l2(p.a_in + row * p.in_stride + off + chunk, (uint32_t)min(chunk, S - off - chunk) * 8u);
                else if (row + gridDim.x < p.a_rows)
                    bulk_prefetch_l2(p.a_in + (row + gridDim.x) * p.in_stride, (uint32_t)min(chunk, S) * 8u);
            }
#pragma unroll 4
            for (int i2 = lane; i2 < (len >> 1); i2 += 32) {
                double2 a = abuf2[i2];
                a.x = (a.x - mx) - lse;
                a.y = (a.y - mx) - lse;
                abuf2[i2] = a;
            }
            __syncwarp();
            if (np > 0) {  // smoothed draws land on top of the streamed values
                const double* pv = p.a_cx + (size_t)row * (size_t)p.cap;
                const unsigned short* ps = p.a_cs + (size_t)row * (size_t)p.cap;
                for (int e = lane; e < np; e += 32) {
                    const int s = (int)ps[e] - off;
                    if (s >= 0 && s < len) abuf[s] = pv[e];
                }
                __syncwarp();
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(p.a_out + row * p.out_stride + off, abuf_raw, bytes);
                bulk_commit();
                bulk_wait_read0();  // the buffer may be refilled once the store has read it
            }
            __syncwarp();
        }
    }
    if (lane == 0) bulk_wait0();
}

__device__ __forceinline__ void apply_warp_loop(const SplitParams& p, unsigned char* abuf_raw, uint64_t* abar,
                                                int lane) {
    if (p.a_chunk < p.S) {  // shared memory too short for a whole row: piecewise variant
        apply_warp_loop_chunked(p, abuf_raw, abar, lane);
        return;
    }
    const int S2 = p.S >> 1;
    const uint32_t row_tx = (uint32_t)p.S * 8u;
    double* abuf = reinterpret_cast<double*>(abuf_raw);
    double2* abuf2 = reinterpret_cast<double2*>(abuf_raw);
    uint32_t phase = 0;
    for (long long row = blockIdx.x; row < p.a_rows; row += gridDim.x) {
        const SplitHeader* h = p.a_hdr + row;
        if (h->flags) continue;  // handed to the general kernel (warp-uniform)
        const double mx = h->mx, lse = h->lse;
        const int np = h->n_patch;
        if (lane == 0) {
            mbar_expect_tx(abar, row_tx);
            bulk_g2s(abuf_raw, p.a_in + row * p.in_stride, row_tx, abar);
        }
        mbar_wait(abar, phase);
        phase ^= 1u;
        // next row of this warp: HBM -> L2 while this one is transformed and stored
        if (lane == 0 && row + gridDim.x < p.a_rows) bulk_prefetch_l2(p.a_in + (row + gridDim.x) * p.in_stride, row_tx);
#pragma unroll 4
        for (int i2 = lane; i2 < S2; i2 += 32) {
            double2 a = abuf2[i2];
            a.x = (a.x - mx) - lse;
            a.y = (a.y - mx) - lse;
            abuf2[i2] = a;
        }
        __syncwarp();
        if (np > 0) {  // smoothed draws land on top of the streamed values
            const double* pv = p.a_cx + (size_t)row * (size_t)p.cap;
            const unsigned short* ps = p.a_cs + (size_t)row * (size_t)p.cap;
            for (int e = lane; e < np; e += 32) abuf[ps[e]] = pv[e];
            __syncwarp();
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(p.a_out + row * p.out_stride, abuf_raw, row_tx);
            bulk_commit();
            bulk_wait_read0();  // the buffer may be refilled once the store has read it
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait0();
}

template <int NT, int EPT, int MODE>
__global__ void __launch_bounds__(stream_block(NT, MODE), stream_min_blocks<NT>()) psis_stream_kernel(const SplitParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    constexpr int EP2 = EPT / 2;
    const int S = p.S, M = p.M, cap = p.cap;
    const int S2 = S >> 1;
    const int nbuf = p.nbuf;
    const StreamSmem L = stream_smem(S, NT * EPT, nbuf, stream_has_apply(NT, MODE) ? p.a_chunk : 0);
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
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        fence_mbar_init();
    }
    if (tid < 32) {
        tab[tid] = exp2((double)tid / 32.0);
        tab[32 + tid] = exp2(-(double)tid / 32.0);
    }
    // pad of the row buffers (draws S .. NT*EPT-1): r = -inf, i.e. ll = +inf in LOO mode
    for (int b = 0; b < nbuf; ++b) {
        double* rb = reinterpret_cast<double*>(smem_raw + (size_t)b * L.row_bytes);
        for (int i = S + tid; i < NT * EPT; i += blockDim.x) rb[i] = (MODE == MODE_LOO) ? inf_f64() : NEG_INF;
    }
    __syncthreads();
    if (stream_has_apply(NT, MODE) && tid >= NT) {  // the apply warp goes its own way (no CTA-wide barrier below)
        apply_warp_loop(p, smem_raw + L.off_apply, &bar[2], lane);
        return;
    }
    // slots j < nfull hold draws for every thread, slot nfull for threads below `rem`, the rest are pads
    const int nfull = S2 / NT;
    const bool part_live = tid < S2 - nfull * NT;
    long long row = blockIdx.x;
    if (tid == 0 && row < p.n_rows) {
        mbar_expect_tx(&bar[0], row_tx);
        bulk_g2s(smem_raw, p.in + row * p.in_stride, row_tx, &bar[0]);
    }

    for (int it = 0; row < p.n_rows; row += gridDim.x, ++it) {
        // reduction slots / candidate counter alternate between consecutive rows, so a warp that
        // runs ahead into the next row never overwrites what a slower warp still has to read
        double* red = reinterpret_cast<double*>(smem_raw + L.off_red + (size_t)(it & 1) * (5 * 32 * 8 + 32 * 4));
        float* redf = reinterpret_cast<float*>(red + 5 * 32);
        int* ctl = ctl_all + (it & 1);
        // ---------------- row -> registers (the shared-memory copy stays valid until the candidates are out)
        const int bsel = (nbuf == 2) ? (it & 1) : 0;
        const double2* rowbuf = reinterpret_cast<const double2*>(smem_raw + (size_t)bsel * L.row_bytes);
        mbar_wait(&bar[bsel], (uint32_t)((nbuf == 2) ? ((it >> 1) & 1) : (it & 1)));
        double2 v[EP2];
#pragma unroll
        for (int j = 0; j < EP2; ++j) {
            v[j] = rowbuf[j * NT + tid];
            if (MODE == MODE_LOO) {  // r = -ll (pyloo/loo.py:286-288)
                v[j].x = -v[j].x;
                v[j].y = -v[j].y;
            }
        }
        // ---------------- pass A: max r (+ min r, sum r in LOO mode), NaN / inf flag
        double m0 = NEG_INF, n0 = inf_f64(), s0 = 0.0;
        int spec = 0;
#pragma unroll
        for (int j = 0; j < EP2; ++j) {
            m0 = max_sel(m0, max_sel(v[j].x, v[j].y));  // NaN-free rows only matter; flagged rows leave
            if (j < nfull || (j == nfull && part_live)) {
                spec = max(spec, max(__double2hiint(v[j].x) & 0x7fffffff, __double2hiint(v[j].y) & 0x7fffffff));
                n0 = min_sel(n0, min_sel(v[j].x, v[j].y));
                if (MODE == MODE_LOO) s0 += v[j].x + v[j].y;
            }
        }
        const double tmax = m0;  // this thread's maximum: one "bin" of the threshold estimate
        m0 = warp_max_sel(m0);
        n0 = -warp_max_sel(-n0);
        if (MODE == MODE_LOO) s0 = warp_sum(s0);
        if (lane == 0) {
            red[wid] = m0;
            red[32 + wid] = -n0;
            if (MODE == MODE_LOO) red[64 + wid] = s0;
        }
        if (tid == 0) ctl[0] = 0;
        // (1) every warp is past the previous row: the other buffer can take the next row now
        const bool special = bar_or_n(NT, spec >= 0x7ff00000);
        if (nbuf == 2 && tid == 0 && row + gridDim.x < p.n_rows) {
            fence_proxy_async();
            mbar_expect_tx(&bar[bsel ^ 1], row_tx);
            bulk_g2s(smem_raw + (size_t)(bsel ^ 1) * L.row_bytes, p.in + (row + gridDim.x) * p.in_stride, row_tx,
                     &bar[bsel ^ 1]);
        }
        // single buffer: the refill has to wait for the end of this row, so at least start the HBM -> L2 leg now
        if (nbuf == 1 && tid == 0 && row + gridDim.x < p.n_rows)
            bulk_prefetch_l2(p.in + (row + gridDim.x) * p.in_stride, row_tx);
        const double mx = slots_max<NW>(red, lane);
        const double r_min = -slots_max<NW>(red + 32, lane);
        double r_sum = 0.0;
        if (MODE == MODE_LOO) r_sum = slots_sum<NW>(red + 64, lane);
        // LOO quantities in ll = -r terms
        const double ll_max = -r_min, ll_min = -mx, ll_mean = -r_sum / (double)S;
        const bool wide = (MODE == MODE_LOO) && !((ll_max - ll_min) <= 600.0);

        // ---------------- threshold guess: per-warp sorted thread maxima (as float distances to the max)
        const float dsorted = warp_sort32_f((float)(mx - tmax), lane);
        int q = p.q0, attempts = 0, C = 0;
        double body = 0.0, lsum = 0.0, vsum = 0.0, taux_used = 0.0;
        // rows spanning more than 1e7 log units would overflow the integer part of the table exp: general kernel
        bool ok = !special && (mx - r_min) <= 1e7;
        while (ok) {
            if (lane == q - 1) redf[wid] = dsorted;
            bar_sync_n(NT);  // (2)
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
                taux = max_sel(-(double)__shfl_sync(FULL, mine, __ffs(b) - 1), -1e300);  // padding (-inf) never qualifies
            }
            taux_used = taux;
            // -------- pass B: body exp-sum, candidate marks.  `mxl` is laundered through an empty asm so
            // the compiler cannot hoist the (threshold-independent) exps out of the retry loop and
            // then spill all EPT results
            double mxl = mx;
            asm volatile("" : "+d"(mxl));
            double ll_mean_l = ll_mean, ll_max_l = ll_max;
            if (MODE == MODE_LOO) asm volatile("" : "+d"(ll_mean_l), "+d"(ll_max_l));
            double bs = 0.0, ls = 0.0, vs = 0.0;
            unsigned cmask = 0;
            // WIDE (LOO rows whose ll range exceeds 600): exp(ll - ll_min) could overflow, so the lppd
            // sum takes the literal exp(ll - ll_max) (utils.py:349-351); compiled as its own loop so
            // the common loop carries no library call
            auto pass_b = [&](auto wide_tag) {
                constexpr bool WIDE = decltype(wide_tag)::value;
#pragma unroll
                for (int j = 0; j < EP2; ++j) {
                    double2 vv = rowbuf[j * NT + tid];
                    const bool live = (j < nfull) || (j == nfull && part_live);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double raw = h ? vv.y : vv.x;
                        const double x = ((MODE == MODE_LOO) ? -raw : raw) - mxl;  // psis.py:134
                        const bool cand = x >= taux;
                        cmask |= cand ? (1u << (2 * j + h)) : 0u;
                        if (MODE == MODE_LOO && !WIDE) {
                            double ep, em;
                            exp_tab_pm_drop(x, tb, cand || !live, !live, ep, em);
                            bs += ep;
                            ls += em;  // exp(ll - ll_min)
                            const double d = live ? raw - ll_mean_l : 0.0;
                            vs = fma(d, d, vs);
                        } else {
                            bs += exp_tab_drop(x, tb, cand || !live);
                            if (MODE == MODE_LOO && live) {
                                ls += exp(raw - ll_max_l);
                                const double d = raw - ll_mean_l;
                                vs = fma(d, d, vs);
                            }
                        }
                    }
                    // keep the scheduler from interleaving all EPT exp chains at once (register pressure)
                    asm volatile("" ::: "memory");
                }
            };
            if (MODE == MODE_LOO && wide) pass_b(std::true_type{});
            else pass_b(std::false_type{});
            // -------- emit candidates: one shared atomic per warp, (exact x, draw index) to the round's scratch
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
                double* dst_x = p.cx + (size_t)row * (size_t)cap;
                unsigned short* dst_s = p.cs + (size_t)row * (size_t)cap;
                // per-thread walk over its marked draws, values re-read from the shared-memory row
                // (a register array cannot be indexed by a run-time bit position)
                const double* rb = reinterpret_cast<const double*>(rowbuf);
                while (cmask) {
                    const int bpos = __ffs((int)cmask) - 1;
                    cmask &= cmask - 1;
                    const int s = 2 * ((bpos >> 1) * NT + tid) + (bpos & 1);
                    if (pos < cap) {
                        const double rv = rb[s];
                        dst_x[pos] = ((MODE == MODE_LOO) ? -rv : rv) - mxl;  // exact x = fl(r - max r)
                        dst_s[pos] = (unsigned short)s;
                    }
                    ++pos;
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
            bar_sync_n(NT);  // (3)
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
            bar_sync_n(NT);  // everyone has read ctl[0] / red before they are reused
            if (tid == 0) ctl[0] = 0;
        }
        if (nbuf == 1 && tid == 0 && row + gridDim.x < p.n_rows) {
            // single buffer (rows too long for two): refill once the candidates are out
            fence_proxy_async();
            mbar_expect_tx(&bar[0], row_tx);
            bulk_g2s(smem_raw, p.in + (row + gridDim.x) * p.in_stride, row_tx, &bar[0]);
        }
        if (tid == 0) {
            SplitHeader h;
            h.mx = mx; h.body = body; h.lsum = lsum; h.vsum = vsum;
            h.lshift = wide ? ll_max : ll_min; h.taux = taux_used;
            h.lse = 0.0; h.C = C; h.flags = ok ? 0 : 1; h.attempts = attempts; h.n_patch = 0; h.C2 = 0; h.pad_ = 0;
            p.hdr[row] = h;
            if (!ok) {
                p.fb_list[atomicAdd(p.fb_count, 1)] = (int)(p.row_base + row);
                if (p.counters) atomicAdd(&p.counters[3], 1ull);
                note_handover(special ? HO_SPECIAL : (((mx - r_min) <= 1e7) ? HO_RETRY : HO_RANGE));
            }
        }
    }
}

// ------------------------------------------------------------------ tail kernel helpers
__device__ __forceinline__ double shfl_f64(double v, int src) { return __shfl_sync(FULL, v, src); }

// Bitonic sort of 32 * CAPL unsigned 32-bit keys held as k[i] <-> element e = 32 i + lane, ascending
// in e.  The kk loop is unrolled, so the direction of every register-pair / register-lane exchange
// is a compile-time constant and one exchange is SHFL + predicated min/max; the partner-distance
// loop below 32 stays a run-time loop to keep the code small (instruction cache).
template <int CAPL>
__device__ __forceinline__ void warp_bitonic_sort32(unsigned (&k)[CAPL], int lane) {
#pragma unroll
    for (int kk = 2; kk <= 32 * CAPL; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j >= 32; j >>= 1) {  // partners in the same lane
            const int dj = j >> 5;
#pragma unroll
            for (int i = 0; i < CAPL; ++i) {
                if ((i & dj) == 0) {
                    const bool up = (((i << 5) & kk) == 0);
                    const unsigned a = k[i], b = k[i | dj];
                    const unsigned lo = min(a, b), hi = max(a, b);
                    k[i] = up ? lo : hi;
                    k[i | dj] = up ? hi : lo;
                }
            }
        }
        const bool upl = ((lane & kk) == 0);  // kk < 32: direction depends on the lane
        int jstart = (kk >> 1) < 16 ? (kk >> 1) : 16;
        asm volatile("" : "+r"(jstart));  // opaque trip count: the compiler must not unroll this loop (code size)
#pragma unroll 1
        for (int j = jstart; j > 0; j >>= 1) {
            const bool lower = ((lane & j) == 0);
#pragma unroll
            for (int i = 0; i < CAPL; ++i) {
                const bool up = (kk >= 32) ? (((i << 5) & kk) == 0) : upl;
                const unsigned o = __shfl_xor_sync(FULL, k[i], j);
                k[i] = (up == lower) ? min(k[i], o) : max(k[i], o);
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

// sum_i log1p(nb * t_i) over the warp's tail (t in shared memory, n values), literal form (rare)
static __device__ __noinline__ double warp_log1p_sum(const double* t, int n, double nb, int lane) {
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc += log1p(nb * t[i]);
    return warp_sum(acc);
}

// log prod_i (1 + nb t_i) for two grid points per lane (t broadcast from shared memory)
__device__ __forceinline__ bool gpd_products2(const double* t, int n, double nb0, double nb1, int every,
                                              const double* ltab, double& out0, double& out1) {
    double P0 = 1.0, P1 = 1.0;
    int E0 = 0, E1 = 0;
    bool ok = true;
    int i = 0;
#pragma unroll 1
    while (i < n) {
        const int stop = min(n, i + every);
#pragma unroll 4
        for (; i + 2 <= stop; i += 2) {  // t is 16 B aligned and `every` is even
            const double2 tt = *reinterpret_cast<const double2*>(t + i);
            P0 *= fma(nb0, tt.x, 1.0);
            P1 *= fma(nb1, tt.x, 1.0);
            P0 *= fma(nb0, tt.y, 1.0);
            P1 *= fma(nb1, tt.y, 1.0);
        }
        if (i < stop) {
            const double t0 = t[i];
            P0 *= fma(nb0, t0, 1.0);
            P1 *= fma(nb1, t0, 1.0);
            ++i;
        }
        ok = rescale_pos_w(P0, E0) && ok;
        ok = rescale_pos_w(P1, E1) && ok;
    }
    out0 = log_tab(P0, ltab) + (double)E0 * 0.6931471805599453094;  // P in [1, 2): the table log is exact enough
    out1 = log_tab(P1, ltab) + (double)E1 * 0.6931471805599453094;
    return ok;
}

// Zhang-Stephens fit for one warp (pyloo/psis.py:181-208).  t: shared memory, DESCENDING (t[0] is
// the largest), n >= 5, grid size m <= 64 (two grid points per lane).  Returns false when the row
// must go to the general kernel.
__device__ __forceinline__ int gpdfit_warp(const double* t, int n, int m, double tsum, int lane,
                                           const ExpTab& tab, double& k_out, double& sigma_out) {
    const double* ltab = tab.t + 64;
    const double inv_n = 1.0 / (double)n;
    const double tq = t[n - ((int)((double)n / 4.0 + 0.5))];  // ascending index int(n/4+.5)-1 (psis.py:187)
    const double tn = t[0];
    if (!(tq > 0.0) || !is_finite(tn) || m > 64) return HO_GPD_TQ;
    double b[2], ks[2];
    bool flag[2];
    double bmag = 0.0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int j = lane + 32 * r;
        double bj = 1.0 - sqrt((double)m / ((double)(j + 1) - 0.5));  // psis.py:186
        bj /= 3.0 * tq;                                                // psis.py:187
        bj += 1.0 / tn;                                                // psis.py:188
        const bool live = j < m;
        b[r] = live ? bj : 0.0;
        ks[r] = 0.0;
        flag[r] = live && (fabs(bj) * tsum < 0.015625);
        if (live) bmag = (fabs(bj) > bmag) ? fabs(bj) : bmag;
        if (live && !is_finite(bj)) bmag = inf_f64();
    }
    bmag = warp_max_sel(bmag);  // (+inf marks a non-finite grid point; no NaN can reach this)
    const double fmx = 1.0 + bmag * tn;
    if (!(fmx < 0x1p1020)) return HO_GPD_FMX;
    // factors lie in [(1 - b_max t_n) > 6e-4, fmx]: (2^31)^32 < 2^1023 and (6e-4)^32 > 2^-340, so rescaling
    // every 32 factors cannot over/underflow; heavy-tailed rows (huge |b| t_n) rescale more often
    const int every = (fmx < 0x1p31) ? 32 : ((fmx < 0x1p250) ? 4 : 1);
    const bool ok = gpd_products2(t, n, -b[0], -b[1], every, ltab, ks[0], ks[1]);
    if (!__all_sync(FULL, ok)) return HO_GPD_PROD;
    // grid points where the product form loses relative accuracy: literal log1p sum
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        unsigned fm = __ballot_sync(FULL, flag[r]);
        while (fm) {
            const int src = __ffs(fm) - 1;
            fm &= fm - 1;
            const double nbj = shfl_f64(-b[r], src);
            const double acc = warp_log1p_sum(t, n, nbj, lane);
            if (lane == src) ks[r] = acc;
        }
    }
    // profile log-likelihood (psis.py:190-191) and weights (psis.py:192)
    double Lj[2];
    double lm = -inf_f64();
    bool fin = true;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const bool live = (lane + 32 * r) < m;
        const double kj = live ? ks[r] * inv_n : -1.0;
        Lj[r] = live ? (double)n * (log_tab(-((live ? b[r] : 1.0) / kj), ltab) - kj - 1.0) : -inf_f64();
        if (live) {
            fin = fin && is_finite(Lj[r]);
            lm = (Lj[r] > lm) ? Lj[r] : lm;
        }
    }
    if (!__all_sync(FULL, fin)) return HO_GPD_PROFILE;
    lm = warp_max_sel(lm);
    double w[2], es = 0.0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const double dl = Lj[r] - lm;  // <= 0
        w[r] = ((lane + 32 * r) < m && dl > -700.0) ? exp_tab(dl, tab) : 0.0;
        es += w[r];
    }
    es = warp_sum(es);
    const double thr = 10.0 * 2.220446049250313e-16;
    double ws = 0.0;
    const double inv_es = 1.0 / es;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        w[r] = w[r] * inv_es;
        if (w[r] < thr) w[r] = 0.0;  // psis.py:194-197 (dead grid points already carry 0)
        ws += w[r];
    }
    ws = warp_sum(ws);
    const double inv_ws = 1.0 / ws;
    double bp = 0.0;
#pragma unroll
    for (int r = 0; r < 2; ++r)
        if (w[r] != 0.0) bp += b[r] * (w[r] * inv_ws);  // psis.py:198-201
    bp = warp_sum(bp);
    // k_post = mean log1p(-b_post t) (psis.py:203): product form unless it loses accuracy
    double lsum;
    bool literal = (fabs(bp) * tsum < 0.015625) || !is_finite(bp);
    if (!literal) {
        double P = 1.0;
        int E = 0;
        bool okp = true;
        for (int i = lane; i < n; i += 32) {
            P *= fma(-bp, t[i], 1.0);
            if (every < 32) okp = rescale_pos_w(P, E) && okp;  // (else <= 16 factors < 2^31 each)
        }
        okp = rescale_pos_w(P, E) && okp;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            P *= __shfl_xor_sync(FULL, P, o);
            E += __shfl_xor_sync(FULL, E, o);
        }
        okp = rescale_pos_w(P, E) && okp;
        if (!__all_sync(FULL, okp)) literal = true;
        else lsum = log_tab(P, ltab) + (double)E * 0.6931471805599453094;
    }
    if (literal) lsum = warp_log1p_sum(t, n, -bp, lane);
    const double k_post = lsum * inv_n;
    sigma_out = -k_post / bp;                                  // psis.py:205
    k_out = ((double)n * k_post + 5.0) / ((double)n + 10.0);   // psis.py:206
    return 0;
}

// literal _gpinv + log (psis.py:153-157, :211-222) for rows outside the fast path's range of k / sigma:
// returns this lane's share of sum_i min(q_i + e^c, 1), smoothed values to tb (descending order)
static __device__ __noinline__ double smooth_tail_literal(double* tb, const double* l1p, int n, int M, double kk,
                                                          double sigma, double exp_c, int lane) {
    double tsm = 0.0;
    for (int e = lane; e < n; e += 32) {
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
        tb[e] = s_;
        tsm += y;
    }
    return tsm;
}

// The candidates that end up outside the tail are summed in double-double arithmetic (error-free TwoSum):
// the result is exact to ~1e-32, so after rounding it does not depend on the (atomics-dependent) order in
// which the stream kernel emitted candidates with equal sort keys -- runs are bit-reproducible.
struct DD {
    double hi, lo;
};
__device__ __forceinline__ void dd_add(DD& a, double e) {
    const double s = a.hi + e;
    const double bb = s - a.hi;
    a.lo += (a.hi - (s - bb)) + (e - bb);
    a.hi = s;
}
__device__ __forceinline__ double warp_dd_sum(DD a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ohi = __shfl_xor_sync(FULL, a.hi, o), olo = __shfl_xor_sync(FULL, a.lo, o);
        const double s = a.hi + ohi;
        const double bb = s - a.hi;
        const double err = (a.hi - (s - bb)) + (ohi - bb);
        const double lo = (a.lo + olo) + err;
        a.hi = s + lo;              // renormalise (fast two-sum)
        a.lo = lo - (a.hi - s);
    }
    return a.hi + a.lo;
}

// heavy-tailed rows (cutoff clamped at log(DBL_MIN)): t_i and the body terms with the library exp
static __device__ __noinline__ void tail_t_literal(const double* xs, double* tb, int n, int staged, double exp_c,
                                                   int lane, DD& nont, double& tsum, double& traw) {
    for (int e = n + lane; e < staged; e += 32) dd_add(nont, exp(xs[e]));
    for (int e = lane; e < n; e += 32) {
        const double ex = exp(xs[e]);
        const double ti = ex - exp_c;
        tb[e] = ti;
        tsum += ti;
        traw += ex;
    }
}

struct TailSmem {
    size_t off_l1p, off_tab, off_w, w_stride, off_x, off_t, off_s, total;
};
// per CTA: l1p table, exp / log tables.  Per warp (18 B per staged element, 4.5 KB at TL = 8, so that
// shared memory does not cap the resident warps): xs[32 TL] exact x of the head of the order,
// tb[32 TL] t_i, then the smoothed values; ss[32 TL] draw indices.
__host__ __device__ inline TailSmem tail_smem(int M, int TL, int warps) {
    // per-warp staging first, at offsets that are compile-time constants of the kernel's template parameters (the
    // M-dependent table goes last): the compiler rematerialises these addresses all over the row loop, and with the
    // table in front every copy re-derived the offset from M (~220 instructions per row)
    TailSmem L;
    L.off_w = 0;
    L.off_x = 0;
    L.off_t = (size_t)32 * TL * 8;
    L.off_s = (size_t)32 * TL * 16;
    L.w_stride = (size_t)32 * TL * 18;
    size_t o = L.w_stride * warps;
    L.off_tab = o;   // 64 doubles exp table + 128 doubles log table
    o += (64 + 128) * 8;
    L.off_l1p = o;
    o += align_up((size_t)(M + 1) * 8, 16);
    L.total = align_up(o, 128);
    return L;
}


struct TailStage {
    double* tb;
    double* xs;
    unsigned short* ss;
    double* gx;           // this row's candidate scratch in global memory: x by slot (cap doubles) ...
    unsigned short* gs;   // ... and draw index by slot; both free again once the head is staged
};

// quantised image of x for the sort: the float image of (x - taux) >= 0, top QB bits, inverted so that
// ascending keys mean descending x.  Monotone in x; fine near the threshold where the candidates
// crowd, coarse towards the row maximum where they are sparse.
template <int QB>
__device__ __forceinline__ unsigned quant_key(double x, double taux) {
    const unsigned fb = __float_as_uint((float)(x - taux));
    return ((1u << QB) - 1u) - (fb >> (31 - QB));
}

// Candidates -> 32-bit sort keys (quantised x | candidate slot), register sort, then the exact values
// of the first 32 TL elements of the order (tail + cutoff) staged to shared memory.  Candidates
// further down can never be in the tail: their exp goes straight to the normaliser (returned).
// Equal quantised values leave a short run in unspecified order; fix_runs() orders it exactly.
template <int CAPL, int TL>
__device__ __forceinline__ DD sort_and_stage(int C, int CA, int cap, bool need_rest, double taux, double xoff,
                                             const TailStage& st, const ExpTab& tab, int lane) {
    constexpr int PB = (TL == 4) ? 8 : ((TL == 8) ? 9 : 10);  // bits of a candidate slot (cap = 64 TL)
    constexpr int QB = 32 - PB;
    // candidate e of the (tight ++ loose) order -> slot of the row's scratch: the tight list grows from the
    // front, the loose one (tile path) from the end; CA = C when there is one list only
    auto slot_of = [&](int e) { return (e < CA) ? e : cap - 1 - (e - CA); };
    unsigned k[CAPL];
#pragma unroll
    for (int i = 0; i < CAPL; ++i) {
        const int e = 32 * i + lane;
        // (xoff: 0, or max r when the scratch holds raw r -- x - 0.0 is x, bit for bit)
        k[i] = (e < C) ? ((quant_key<QB>(st.gx[slot_of(e)] - xoff, taux) << PB) | (unsigned)e) : 0xffffffffu;
    }
    warp_bitonic_sort32<CAPL>(k, lane);
    // exact value and draw index of every element of the order: gathered by slot from the row's scratch
    // (4 KB, just read: L1 / L2 hits)
    DD rest = {0.0, 0.0};
#pragma unroll
    for (int i = 0; i < CAPL; ++i) {
        const int e = 32 * i + lane;
        const int pidx = slot_of((int)(k[i] & ((1u << PB) - 1u)));
        if (i < TL) {
            st.xs[e] = (e < C) ? st.gx[pidx] - xoff : -inf_f64();
            st.ss[e] = (e < C) ? st.gs[pidx] : (unsigned short)0;
        } else if (need_rest && 32 * i < C) {
            if (e < C) {
                const double x = st.gx[pidx] - xoff;
                if (x >= -700.0) dd_add(rest, exp_tab(x, tab));
            }
        }
    }
    __syncwarp();
    return rest;
}

// Exact order inside runs of equal quantised keys among the staged elements: every element of a run
// counts the run members that precede it in (x descending, draw index descending) order and moves
// there.  Returns false if a run is too long or may continue past the staged range.
template <int TL>
__device__ __forceinline__ bool fix_runs(const TailStage& st, int C, int M, double taux, int lane) {
    constexpr int NS = 32 * TL;
    constexpr int PB = (TL == 4) ? 8 : ((TL == 8) ? 9 : 10);
    constexpr int QB = 32 - PB;
    // temporaries (only touched when a run exists): moved values in tb, their draw indices and new
    // positions in the upper half of the row's candidate scratch (free once the head is staged)
    double* mvx = st.tb;
    int* mvs = reinterpret_cast<int*>(st.gx + NS);
    int* mvn = mvs + NS;
    bool bad = false;
    unsigned moved = 0;  // bit i: this lane's element of the i-th group of 32 moves
    const int lim = (C < NS) ? C : NS;
    // only the order of elements 0 .. M (tail + cutoff) matters: runs that start above M are left alone,
    // the run that holds element M is fixed as a whole
    int need = (M + 1 < lim) ? M + 1 : lim;
    if (need == M + 1) {
        const unsigned qm = quant_key<QB>(st.xs[M], taux);
        while (need < lim && need - M < 35 && quant_key<QB>(st.xs[need], taux) == qm) ++need;
    }
#pragma unroll 1
    for (int i = 0; 32 * i < need; ++i) {
        const int e = 32 * i + lane;
        if (e < need) {
            const double myx = st.xs[e];
            const unsigned q = quant_key<QB>(myx, taux);
            const bool in_run = (e > 0 && quant_key<QB>(st.xs[e - 1], taux) == q) ||
                                (e + 1 < lim && quant_key<QB>(st.xs[e + 1], taux) == q);
            if (in_run) {
                int lo = e, hi = e;
                while (lo > 0 && e - lo < 33 && quant_key<QB>(st.xs[lo - 1], taux) == q) --lo;
                while (hi + 1 < lim && hi - e < 33 && quant_key<QB>(st.xs[hi + 1], taux) == q) ++hi;
                if (hi - lo > 32 || (hi == NS - 1 && C > NS)) bad = true;
                const int mys = st.ss[e];
                int cnt = 0;
                for (int f = lo; f <= hi; ++f) {
                    const double xf = st.xs[f];
                    cnt += (xf > myx || (xf == myx && (int)st.ss[f] > mys)) ? 1 : 0;
                }
                mvx[e] = myx;
                mvs[e] = mys;
                mvn[e] = lo + cnt;
                moved |= 1u << i;
            }
        }
    }
    __syncwarp();
    if (__any_sync(FULL, moved != 0)) {
#pragma unroll 1
        for (int i = 0; 32 * i < need; ++i) {
            if (moved & (1u << i)) {
                const int e = 32 * i + lane;
                const int npos = mvn[e];
                st.xs[npos] = mvx[e];
                st.ss[npos] = (unsigned short)mvs[e];
            }
        }
        __syncwarp();
    }
    return !__any_sync(FULL, bad);
}

// One row for one warp, in four phases.  Returns 0, or the reason the row goes to the general kernel (-1: the warp
// had no row).  SYNCP: the warps of the CTA meet at a block barrier between the phases (every warp of the CTA
// must call this function the same number of times): warps of one CTA then execute the same few hundred
// instructions at any time, which keeps the kernel's ~100 KB of code from thrashing the 32 KB instruction cache.
template <int TL, int MODE, bool SYNCP>
__device__ __forceinline__ int tail_row(const SplitParams& p, long long row, const SplitHeader& h, bool alive,
                                         const double* l1p, const TailStage& st, const ExpTab& tab, int lane) {
    const int S = p.S, M = p.M;
    // tile path: the tight list alone when it holds the tail and fits one sort, else both lists
    const bool tight_only = !p.ab_lists || (h.C >= M + 1 && h.C <= 32 * TL);
    const int CA = h.C, C = tight_only ? h.C : h.C + h.C2;
    const double mx = h.mx;
    double* xs = st.xs;
    double* tb = st.tb;
    unsigned short* ss = st.ss;
    double* cx = st.gx;
    unsigned short* cs = st.gs;
    const bool total_body = p.total_body != 0;  // body = sum over all draws: nothing to add back for the non-tail candidates
    int why = alive ? 0 : -1;
    DD nont = {0.0, 0.0};  // double-double: order-independent sum

    // ---------------- phase 1: sort the candidates, stage the head of the order, exact order inside key runs
    if (!why) {
        const double xoff = p.chunked ? mx : 0.0;
        if (C <= 32 * TL) {
            nont = sort_and_stage<TL, TL>(C, CA, p.cap, !total_body, h.taux, xoff, st, tab, lane);
        } else {
            // (TL = 32, tails of 511 to ~800 draws: one 1024-key sort is all there is; longer lists are handed over)
            if constexpr (TL < 32) nont = sort_and_stage<2 * TL, TL>(C, CA, p.cap, !total_body, h.taux, xoff, st, tab, lane);
            else why = HO_RETRY;
        }
        if (!why && !fix_runs<TL>(st, C, M, h.taux, lane)) why = HO_RUNS;
    }
    if (SYNCP) __syncthreads();

    // ---------------- phase 2: cutoff, tail, t_i
    double xc = 0.0, c = 0.0, exp_c = 0.0, tsum = 0.0, traw = 0.0, nont_sum = 0.0;
    int n = 0;
    bool deep = false;
    if (!why) {
        // cutoff = (M+1)-th largest = element M of the order (psis.py:135-136); draws equal to it are
        // not in the tail (psis.py:139)
        xc = xs[M];
        // chunked lists are unions of per-chunk threshold sets: they hold the column's M + 1 largest only if none of
        // the draws left out (x <= vmax) can exceed the cutoff found
        if (p.chunked && !(xc >= h.vmax)) why = HO_RETRY;
        n = M;
        while (n > 0 && xs[n - 1] == xc) --n;
        // heavy-tailed rows: the cutoff sits near / below log(DBL_MIN) and is clamped there (psis.py:136);
        // the tail is then whatever lies above the clamp, and the table exp (no denormals) is not used
        deep = !(xc >= -690.0);
        c = xc;
        if (deep) {
            c = (xc > p.cutoffmin) ? xc : p.cutoffmin;
            int cnt = 0;
            for (int e = lane; e < M; e += 32) cnt += (xs[e] > c) ? 1 : 0;
            n = warp_isum(cnt);  // the order is descending: these are the first n elements
        }
        // safety net: the order inside the tail must be exact (descending x, descending index on ties)
        bool bad = false;
        for (int e = lane; e + 1 < n; e += 32) {
            const double a = xs[e], b = xs[e + 1];
            if (!(a > b || (a == b && ss[e] > ss[e + 1]))) bad = true;
        }
        if (__any_sync(FULL, bad)) why = HO_ORDER;
    }
    if (!why) {
        if (p.tail_idx) {  // the tail's draw indices (the first n elements of the order), the rest -1
            int* d = p.tail_idx + row * p.tail_ld;
            for (int e = lane; e < (int)p.tail_ld; e += 32) d[e] = (e < n) ? (int)ss[e] : -1;
        }
        exp_c = exp_any(c);  // psis.py:138
        if (!deep) {
            // staged candidates at or below the cutoff belong to the normaliser's body too
            if (!total_body) {
#pragma unroll 1
                for (int e = n + lane; e < min(C, 32 * TL); e += 32) dd_add(nont, exp_tab(xs[e], tab));
            }
            // t_i = exp(x_i) - exp(c) (psis.py:146-147), descending
#pragma unroll 1
            for (int e = lane; e < n; e += 32) {
                const double ex = exp_tab(xs[e], tab);  // x >= c >= -690
                const double ti = ex - exp_c;
                tb[e] = ti;
                tsum += ti;
                traw += ex;
            }
        } else {
            tail_t_literal(xs, tb, n, total_body ? n : min(C, 32 * TL), exp_c, lane, nont, tsum, traw);
        }
        tsum = warp_sum(tsum);
        traw = warp_sum(traw);
        nont_sum = total_body ? 0.0 : warp_dd_sum(nont);
        __syncwarp();
    }
    if (SYNCP) __syncthreads();

    // ---------------- phase 3: Zhang-Stephens fit of the tail
    double kk = inf_f64(), sigma = nan_f64();
    bool smooth = false;
    if (!why && n > 4) {
        int m = p.m_full;
        if (n != M) {
            m = (int)sqrt((double)n);
            while (m * m > n) --m;
            while ((m + 1) * (m + 1) <= n) ++m;
            m += 30;
        }
        why = gpdfit_warp(tb, n, m, tsum, lane, tab, kk, sigma);
        smooth = !why && is_finite(kk);  // psis.py:150
    }
    if (SYNCP) __syncthreads();

    // ---------------- phase 4: smoothed tail, normaliser, outputs
    if (why) return why;
    // smoothed tail (psis.py:153-157, _gpinv :211-222); element e has ascending rank n-1-e.
    // The smoothed values replace t in shared memory.
    double tails = traw;
    if (smooth) {
        double tsm = 0.0;
        __syncwarp();
        const double* ltab = tab.t + 64;
        if (sigma > 0.0 && fabs(kk) >= 0.01 && fabs(kk) < 50.0 && n == M) {
            // common case: expm1(u) = exp(u) - 1 is accurate enough once |k| is not tiny (|u| >=
            // 2.5e-5; the lowest ranks, where the relative error of the difference peaks, add the
            // least to q + exp(c)); table-driven exp and log
            const double sk = sigma / kk;
#pragma unroll 1
            for (int e = lane; e < n; e += 32) {
                const double u = -kk * l1p[n - 1 - e];
                double y = fma(exp_tab(u, tab) - 1.0, sk, exp_c);
                double s_ = log_tab(y, ltab);
                if (s_ > 0.0) {  // psis.py:157
                    s_ = 0.0;
                    y = 1.0;
                }
                tb[e] = s_;
                tsm += y;
            }
        } else {
            tsm = smooth_tail_literal(tb, l1p, n, M, kk, sigma, exp_c, lane);
        }
        tails = warp_sum(tsm);
        __syncwarp();
    }
    // tile path: h.body is the sum over all S draws, so the body is what remains without the raw tail terms;
    // when the tail carries (almost) the whole sum the difference is rounding noise -- harmless as long as the
    // smoothed tail is of the same order, else the row goes to the general kernel
    // (the total carries the table exponential's error, up to 4e-14 relative; the difference inherits it
    // amplified by total / (body + tails), and 1e-3 keeps that below 4e-11 of the result)
    const double body = total_body ? h.body - traw : h.body + nont_sum;
    if (total_body && !(body + tails > 1e-3 * h.body)) return HO_CANCEL;
    const double lse = log_tab(body + tails, tab.t + 64);  // psis.py:158

    if (MODE == MODE_PSISLW) {
        // hand the normaliser and the smoothed tail to the apply kernel (the candidate scratch of this
        // row is free again: values to patch in, and where)
        if (smooth) {
            for (int e = lane; e < n; e += 32) {
                cx[e] = tb[e] - lse;
                cs[e] = ss[e];
            }
        }
        if (lane == 0) {
            p.hdr[row].lse = lse;
            p.hdr[row].n_patch = smooth ? n : 0;
            p.k_out[row] = kk;
        }
    } else {
        // elpd_i = LSE_s(lw_s + ll_s): body terms are the constant -(mx + lse), tail terms differ
        // from it by (smoothed - raw) (loo.py:289,319-324)
        double dmax = 0.0, es = (double)n;
        if (smooth) {
            double dm = 0.0;
            for (int e = lane; e < n; e += 32) {
                const double d = tb[e] - xs[e];
                dm = (d > dm) ? d : dm;
            }
            dmax = warp_max_sel(dm);
            double e2 = 0.0;
#pragma unroll 1
            // (arguments <= 0; the table exponential is good to 4e-15 down to -36 and never worse than 8e-14, on
            // terms that small next to the exp(0) of the largest: elpd_i moves by < 1e-14)
            for (int e = lane; e < n; e += 32) e2 += exp_tab_drop((tb[e] - xs[e]) - dmax, tab, false);
            es = warp_sum(e2);
        }
        const double* ltab2 = tab.t + 64;
        const double tot = (double)(S - n) * exp_tab_drop(-dmax, tab, false) + es;
        const double elpd = ((-mx - lse) + dmax) + log_tab(tot, ltab2);
        const double lppd = log_tab(h.lsum, ltab2) + (h.lshift - p.log_S);  // utils.py:352-357, b_inv = S
        const double var = h.vsum / (double)S;                        // waic.py:145
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
    return 0;
}

// warps per CTA: 4 (eight CTAs per SM, every warp on its own) or 16 / 32 with the phase barriers of tail_row
__host__ __device__ constexpr int tail_ctas_per_sm(int warps) { return warps >= 32 ? 1 : 32 / warps; }

template <int TL, int MODE, int W>
__global__ void __launch_bounds__(W * 32, (TL <= 8) ? tail_ctas_per_sm(W) : (W <= 4 ? 3 : (W <= 8 ? 2 : 1)))
psis_tail_kernel(const SplitParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool SYNCP = W > 4;
    const TailSmem L = tail_smem(p.M, TL, W);
    double* l1p = reinterpret_cast<double*>(smem_raw + L.off_l1p);
    const int lane = threadIdx.x & 31;
    // broadcast: lets the compiler see that everything derived from the warp index (row, header, branch
    // conditions) is warp-uniform, so the shuffles below need no re-convergence code around them
    const int wid = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
    unsigned char* wbase = smem_raw + L.off_w + L.w_stride * wid;
    TailStage st;
    st.xs = reinterpret_cast<double*>(wbase + L.off_x);
    st.tb = reinterpret_cast<double*>(wbase + L.off_t);
    st.ss = reinterpret_cast<unsigned short*>(wbase + L.off_s);
    st.gx = nullptr;
    st.gs = nullptr;
    double* tabm = reinterpret_cast<double*>(smem_raw + L.off_tab);
    ExpTab tab;
    tab.t = tabm;
    tab.tinv = tabm + 32;
    if (threadIdx.x < 32) {
        tabm[threadIdx.x] = exp2((double)threadIdx.x / 32.0);
        tabm[32 + threadIdx.x] = exp2(-(double)threadIdx.x / 32.0);
    }
    if (threadIdx.x < 64) log_tab_init(tabm + 64, threadIdx.x);
    // per-CTA table for the smoothing step: depends only on (rank, M), psis.py:153 + :221
    for (int i = threadIdx.x; i < p.M; i += W * 32) l1p[i] = log1p(-(((double)i + 0.5) / (double)p.M));
    __syncthreads();
    const long long step = (long long)gridDim.x * W;
#pragma unroll 1
    for (long long base = (long long)blockIdx.x * W; base < p.n_rows; base += step) {  // (CTA-uniform trip count)
        const long long row = base + wid;
        bool alive = row < p.n_rows;
        SplitHeader h;
        if (alive) h = p.hdr[row];
        else { h.C = 0; h.C2 = 0; h.flags = 1; h.mx = 0.0; h.taux = 0.0; h.body = 0.0; h.lsum = 1.0; h.vsum = 0.0; h.lshift = 0.0; h.attempts = 0; }
        if (h.flags) alive = false;
        if (alive) {
            st.gx = p.cx + (size_t)row * (size_t)p.cap;
            st.gs = p.cs + (size_t)row * (size_t)p.cap;
        }
        const int why = tail_row<TL, MODE, SYNCP>(p, row, h, alive, l1p, st, tab, lane);
        if (why > 0 && lane == 0) {
            note_handover(why);
            p.hdr[row].flags = 1;  // the apply kernel skips the row
            p.fb_list[atomicAdd(p.fb_count, 1)] = (int)(p.row_base + row);
            if (p.counters) atomicAdd(&p.counters[3], 1ull);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ apply kernel (psislw only)
// out = (r - max r) - lse (psis.py:134,158) for one observation per CTA, then the smoothed tail
// (psis.py:156) on top.  Pure streaming: 16 S bytes per observation.
constexpr int APPLY_NT = 256;
static __global__ void __launch_bounds__(APPLY_NT) psis_apply_kernel(const SplitParams p) {
    const int S2 = p.S >> 1;
    for (long long row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
        const SplitHeader* h = p.hdr + row;
        if (h->flags) continue;  // block-uniform
        const double mx = h->mx, lse = h->lse;
        const int np = h->n_patch;
        const double2* s2 = reinterpret_cast<const double2*>(p.in + row * p.in_stride);
        double* dst = p.out + row * p.out_stride;
        double2* d2 = reinterpret_cast<double2*>(dst);
        int i2 = threadIdx.x;
        for (; i2 + 3 * APPLY_NT < S2; i2 += 4 * APPLY_NT) {
            double2 a0 = s2[i2], a1 = s2[i2 + APPLY_NT], a2 = s2[i2 + 2 * APPLY_NT], a3 = s2[i2 + 3 * APPLY_NT];
            a0.x = (a0.x - mx) - lse; a0.y = (a0.y - mx) - lse;
            a1.x = (a1.x - mx) - lse; a1.y = (a1.y - mx) - lse;
            a2.x = (a2.x - mx) - lse; a2.y = (a2.y - mx) - lse;
            a3.x = (a3.x - mx) - lse; a3.y = (a3.y - mx) - lse;
            d2[i2] = a0; d2[i2 + APPLY_NT] = a1; d2[i2 + 2 * APPLY_NT] = a2; d2[i2 + 3 * APPLY_NT] = a3;
        }
        for (; i2 < S2; i2 += APPLY_NT) {
            double2 a0 = s2[i2];
            a0.x = (a0.x - mx) - lse; a0.y = (a0.y - mx) - lse;
            d2[i2] = a0;
        }
        if (np > 0) {
            __syncthreads();  // the patches must land after the streamed values of the same draws
            const double* pv = p.cx + (size_t)row * (size_t)p.cap;
            const unsigned short* ps = p.cs + (size_t)row * (size_t)p.cap;
            for (int e = threadIdx.x; e < np; e += APPLY_NT) dst[ps[e]] = pv[e];
        }
    }
}

}  // namespace b2l
