// libpsisloo_b200.so -- C ABI (include/psisloo_b200.h) over the sm_100a kernels.
// Host side: launch planning, obs-fastest panel transposes, shard statistics, host-buffer pipelines.
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/psisloo_b200.h"
#include "b2l_row_kernel.cuh"
#include "b2l_split_host.h"
#include "b2l_tile_host.h"
#include "b2l_is_host.h"

using namespace b2l;

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return fail((int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                             \
    } while (0)

// ------------------------------------------------------------------------------------ per-kernel timing
// Optional (b2l_profile): CUDA events around every kernel launch on the launching stream, summed per
// kernel kind.  Used by bench.py for the live roofline numbers; off in the timed throughput loop.
namespace {
struct ProfRec {
    cudaEvent_t a, b;
    int kind;
};
bool g_prof = false;
std::vector<ProfRec> g_prof_recs;
std::mutex g_prof_mu;  // the host pipelines of different devices may run on different threads
struct ProfScope {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    int kind;
    ProfScope(int kind_, cudaStream_t st_) : st(st_), kind(kind_) {
        if (g_prof && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) cudaEventRecord(a, st);
    }
    ~ProfScope() {
        if (a && b) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lock(g_prof_mu);
            g_prof_recs.push_back({a, b, kind});
        }
    }
};
}  // namespace

static bool aligned16_ptr(const void* p) { return ((uintptr_t)p & 15) == 0; }

// ------------------------------------------------------------------------------------ planning
constexpr int GROWS_MAX_CTAS = 192;  // global-memory row mode: at most this many resident rows
struct RowPlan {
    int nt, cap, r0, nbuf, ctas_per_sm, grid, sms, global_rows;
    size_t smem;
};

static int pow2ceil(long long v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <int NT, int MODE>
static cudaError_t occupancy_of(size_t smem, int* out) {
    cudaError_t e = cudaFuncSetAttribute(psis_row_kernel<NT, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(psis_row_kernel<NT, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, psis_row_kernel<NT, MODE>, NT, smem);
}

static int plan_row(long long S, int M, int mode, long long n_rows, RowPlan* pl) {
    if (S < 1 || M < 1 || (long long)M + 1 > S)
        return fail(B2L_E_INVALID, "need 1 <= M and M + 1 <= S (S=%lld, M=%d): the reference indexes "
                    "x[sorted[-M-1]] (pyloo/psis.py:136)", S, M);
    if (M > 9000) return fail(B2L_E_UNSUPPORTED, "tail length M=%d exceeds the GPD grid capacity", M);
    if (S > (1ll << 30)) return fail(B2L_E_UNSUPPORTED, "S too large");
    int dev = 0;
    CK(cudaGetDevice(&dev));
    int smem_optin = 0, smem_sm = 0, sms = 0;
    CK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    pl->sms = sms;
    pl->cap = std::max(512, pow2ceil(3ll * (M + 1)));
    if (const char* ev = getenv("B2L_CAP")) {
        int c = atoi(ev);
        if (c >= 2 * (M + 1) && (c & (c - 1)) == 0) pl->cap = c;
    }
    // rows in flight per SM: prefer more resident CTAs (more warps), double-buffer on a tie
    int best_ctas = 0, best_nbuf = 0, best_nt = 256;
    size_t best_smem = 0;
    for (int nbuf = 1; nbuf <= 2; ++nbuf) {
        for (int nt : {256, 512}) {
            size_t smem = row_smem_layout((int)S, M, pl->cap, nbuf, nt).total;
            if (smem > (size_t)smem_optin) continue;
            int ctas = (int)((size_t)smem_sm / (smem + 1024));
            if (ctas < 1) continue;
            if (nt == 512 && ctas > 1) continue;  // wide CTAs only when a single CTA owns the SM
            bool better = ctas > best_ctas || (ctas == best_ctas && nbuf > best_nbuf) ||
                          (ctas == best_ctas && nbuf == best_nbuf && ctas == 1 && nt > best_nt);
            if (better) {
                best_ctas = ctas; best_nbuf = nbuf; best_nt = nt; best_smem = smem;
            }
        }
    }
    // tuning overrides (numerics are unaffected): B2L_NBUF = 1|2, B2L_NT = 256|512
    if (const char* ev = getenv("B2L_NBUF")) {
        int nb = atoi(ev), nt = getenv("B2L_NT") ? atoi(getenv("B2L_NT")) : best_nt;
        if ((nb == 1 || nb == 2) && ((nt == 128 && 30 + (int)std::sqrt((double)M) <= 64) || nt == 256 || nt == 512)) {
            size_t smem = row_smem_layout((int)S, M, pl->cap, nb, nt).total;
            if (smem <= (size_t)smem_optin) {
                best_nbuf = nb; best_nt = nt; best_smem = smem;
                best_ctas = std::max(1, (int)((size_t)smem_sm / (smem + 1024)));
            }
        }
    }
    pl->global_rows = 0;
    if (best_ctas == 0) {
        // rows longer than shared memory: same kernel, the row lives in a global-memory workspace
        // (L2-resident: gridDim.x rows) instead of shared memory.  Slow path, any S.
        size_t smem = row_smem_layout((int)S, M, pl->cap, 0, 512).total;
        if (smem > (size_t)smem_optin)
            return fail(B2L_E_UNSUPPORTED, "tail length M=%d needs %zu bytes of shared memory", M, smem);
        best_ctas = 1; best_nbuf = 0; best_nt = 512; best_smem = smem;
        pl->global_rows = 1;
    }
    pl->nt = best_nt; pl->nbuf = best_nbuf; pl->smem = best_smem;
    {   // pooled sample rank aiming at ~1.9 (M+1) candidates (one sample per thread)
        const int pool = (best_nt / 32) * POOL_PER_WARP;
        long long r0 = std::llround(1.9 * (M + 1) * (double)best_nt / (double)S);
        pl->r0 = (int)std::min<long long>(pool, std::max<long long>(1, r0));
    }
    int occ = 0;
    cudaError_t e;
    if (mode == MODE_PSISLW)
        e = (best_nt == 128) ? occupancy_of<128, MODE_PSISLW>(best_smem, &occ)
            : (best_nt == 256) ? occupancy_of<256, MODE_PSISLW>(best_smem, &occ)
                               : occupancy_of<512, MODE_PSISLW>(best_smem, &occ);
    else
        e = (best_nt == 128) ? occupancy_of<128, MODE_LOO>(best_smem, &occ)
            : (best_nt == 256) ? occupancy_of<256, MODE_LOO>(best_smem, &occ)
                               : occupancy_of<512, MODE_LOO>(best_smem, &occ);
    CK(e);
    if (occ < 1) return fail(B2L_E_UNSUPPORTED, "row kernel does not fit on an SM (smem %zu)", best_smem);
    pl->ctas_per_sm = occ;
    long long g = (long long)sms * occ;
    if (pl->global_rows) g = std::min<long long>(g, GROWS_MAX_CTAS);
    pl->grid = (int)std::max<long long>(1, std::min<long long>(g, n_rows));
    return 0;
}

static int launch_rows(int mode, const RowPlan& pl, RowParams rp, cudaStream_t st) {
    rp.cap = pl.cap; rp.r0 = pl.r0; rp.nbuf = pl.nbuf;
    rp.force_legacy = getenv("B2L_FORCE_LEGACY") ? 1 : 0;
    if (pl.global_rows) {
        if (!rp.row_ws) return fail(B2L_E_WORKSPACE, "rows of S=%d draws need the global-row workspace", rp.S);
        rp.use_bulk = 0;
    } else {
        rp.row_ws = nullptr;
    }
    {
        int r = (int)std::sqrt((double)rp.M);
        while (r * r > rp.M) --r;
        while ((r + 1) * (r + 1) <= rp.M) ++r;
        rp.m_full = 30 + r;
    }
    int grid = (int)std::max<long long>(1, std::min<long long>(pl.grid, rp.n_rows));
    if (rp.n_rows == 0) return 0;
    ProfScope prof(B2L_PROF_ROW, st);
    if (mode == MODE_PSISLW) {
        if (pl.nt == 128) psis_row_kernel<128, MODE_PSISLW><<<grid, 128, pl.smem, st>>>(rp);
        else if (pl.nt == 256) psis_row_kernel<256, MODE_PSISLW><<<grid, 256, pl.smem, st>>>(rp);
        else psis_row_kernel<512, MODE_PSISLW><<<grid, 512, pl.smem, st>>>(rp);
    } else {
        if (pl.nt == 128) psis_row_kernel<128, MODE_LOO><<<grid, 128, pl.smem, st>>>(rp);
        else if (pl.nt == 256) psis_row_kernel<256, MODE_LOO><<<grid, 256, pl.smem, st>>>(rp);
        else psis_row_kernel<512, MODE_LOO><<<grid, 512, pl.smem, st>>>(rp);
    }
    CK(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------ split path
// stream kernel (one CTA per observation, row in registers) + tail kernel (one warp per observation)
// + the general row kernel on the rows those two hand over.  See b2l_split.cuh.
struct SplitPlan {
    int ok, nt, ept, tl, tw, cap, q0, nbuf, fused, a_chunk, grid1, grid2, occ1, occ2;  // tw: warps per tail CTA
    size_t smem1, smem2;
    long long batch;  // observations per stream -> tail -> fallback round
};

// shape part of the plan: pure arithmetic (also sizes the workspace without touching the device)
static bool split_shape(long long S, int M, long long n_rows, SplitPlan* sp) {
    memset(sp, 0, sizeof(*sp));
    if (const char* ev = getenv("B2L_SPLIT")) if (atoi(ev) == 0) return false;
    if (getenv("B2L_FORCE_LEGACY")) return false;
    // (tails beyond 510 draws -- r_eff < 0.14 at S = 4000, < 0.55 at S = 16 000 -- sort 1024 keys per lane-set: M + 2 <= 800)
    if (S % 2 != 0 || S < 64 || S > SPLIT_MAX_S || M < 5 || M + 2 > 800 || (long long)M + 1 > S / 2) return false;
    static const int shapes[8][2] = {{128, 8}, {128, 16}, {256, 8}, {256, 16}, {512, 8}, {512, 16}, {1024, 8}, {1024, 16}};
    int nt = 0, ept = 0;
    for (auto& sh : shapes) {
        if ((long long)sh[0] * sh[1] >= S && (double)sh[0] >= 1.25 * (M + 1) &&
            (nt == 0 || sh[0] * sh[1] < nt * ept)) {
            nt = sh[0]; ept = sh[1];
        }
    }
    if (!nt) return false;
    sp->nt = nt; sp->ept = ept;
    sp->tl = (M + 2 <= 128) ? 4 : ((M + 2 <= 256) ? 8 : ((M + 2 <= 512) ? 16 : 32));
    sp->cap = 64 * sp->tl;
    {
        const double ratio = (double)(M + 1) / nt;
        // balls-in-bins estimate of the per-warp rank whose value sits just below the (M+1)-th largest draw;
        // tuned on the GPU: a retry (threshold too high, ~5 % of rows) costs less than sorting 512
        // instead of 256 candidates in the tail kernel
        int q = (int)std::lround(32.0 * (1.0 - std::exp(-1.25 * ratio))) +
                ((ratio > 0.45 && !(ept == 16 && ratio > 0.6)) ? 1 : 0);
        if (const char* ev = getenv("B2L_Q0")) q = atoi(ev);
        sp->q0 = std::min(30, std::max(1, q));
    }
    // observations per stream -> tail -> apply -> hand-over round: a multiple of every resident-CTA /
    // resident-warp count the kernels reach on 148 SMs (no partial last wave), ~230 MB at S = 4000
    long long b = 148ll * 128;  // (S = 4000: 18 944 observations per round measure 2 % faster than 9 472)
    while (b > 148 * 6 && b * S * 8 > (1ll << 30)) b /= 2;
    if (const char* ev = getenv("B2L_BATCH")) b = std::max<long long>(1, atoll(ev));
    sp->batch = std::max<long long>(1, std::min<long long>(b, std::max<long long>(n_rows, 1)));
    sp->nbuf = 2;
    if (const char* ev = getenv("B2L_SNBUF")) sp->nbuf = (atoi(ev) == 1) ? 1 : sp->nbuf;
    sp->smem1 = stream_smem((int)S, sp->nt * sp->ept, 1, 0).total;  // refined per mode in plan_split
    // warps per tail CTA: big CTAs whose warps pass through the phases of a row together (see tail_row)
    sp->tw = (sp->tl == 32) ? 4 : 8;  // (TL = 32: 18 KB of staging per warp; measured at S = 4000, M = 190, per 2 x 75 776 observations: 4 warps 1.61 ms, 8 warps 1.36 ms, 16 warps 1.40 ms, 32 warps 1.45 ms)
    if (const char* ev = getenv("B2L_TAIL_WARPS")) {
        const int w = atoi(ev);
        if (sp->tl == 32) { if (w == 4 || w == 8) sp->tw = w; }
        else if (w == 4 || w == 8 || w == 16 || (w == 32 && sp->tl <= 8)) sp->tw = w;
    }
    sp->smem2 = tail_smem(M, sp->tl, sp->tw).total;
    return true;
}

static int plan_split(long long S, int M, int mode, long long n_rows, SplitPlan* sp) {
    if (!split_shape(S, M, n_rows, sp)) return 0;
    const int nt = sp->nt, ept = sp->ept;
    int dev = 0, sms = 0, smem_optin = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (sp->smem2 > (size_t)smem_optin) return 0;
    // shared-memory shape of the stream kernel: most resident CTAs first, then whole-row apply transfers,
    // then two row buffers (prefetch during the whole row)
    {
        const bool ap = stream_has_apply(nt, mode);
        sp->fused = (ap && !(getenv("B2L_FUSED_APPLY") && atoi(getenv("B2L_FUSED_APPLY")) == 0)) ? 1 : 0;
        const int max_nbuf = sp->nbuf;
        int best_occ = 0;
        for (int pieces : {1, 2, 4}) {
            if (!ap && pieces > 1) break;
            const int chunk = ap ? (int)(((S / pieces) + 1) / 2 * 2) : 0;  // even number of draws
            for (int nb = max_nbuf; nb >= 1; --nb) {
                const size_t sm = stream_smem((int)S, nt * ept, nb, chunk).total;
                if (sm > (size_t)smem_optin) continue;
                int occ = 0;
                CK(split_stream_setup(nt, ept, mode, sm, &occ));
                if (occ > best_occ) {
                    best_occ = occ; sp->nbuf = nb; sp->a_chunk = chunk; sp->smem1 = sm; sp->occ1 = occ;
                }
            }
        }
        if (best_occ < 1) return 0;
        int occ = 0;
        CK(split_stream_setup(nt, ept, mode, sp->smem1, &occ));  // leave the chosen size as the function attribute
    }
    CK(split_tail_setup(sp->tl, sp->tw, mode, sp->smem2, &sp->occ2));
    if (sp->occ1 < 1 || sp->occ2 < 1) return 0;
    sp->grid1 = sms * sp->occ1;
    sp->grid2 = sms * sp->occ2;
    sp->ok = 1;
    return 0;
}

// scratch: two slots of [headers | candidate x | candidate s] (the apply stage of batch b runs inside the
// stream kernel of batch b + 1), the hand-over list over all rows of the call, its counter
static size_t split_slot_bytes(const SplitPlan& sp) {
    return align_up((size_t)sp.batch * sizeof(SplitHeader), 256) + align_up((size_t)sp.batch * sp.cap * 8, 256) +
           align_up((size_t)sp.batch * sp.cap * 2, 256);
}
static size_t split_ws_bytes(const SplitPlan& sp, long long n_rows) {
    if (!sp.nt) return 0;
    return 2 * split_slot_bytes(sp) + align_up((size_t)std::max<long long>(n_rows, 1) * 4, 256) + 256;
}

static int launch_split(int mode, const RowPlan& pl, const SplitPlan& sp, const RowParams& rp, void* sws,
                        cudaStream_t st) {
    char* w = (char*)sws;
    SplitHeader* hdr[2];
    double* cx[2];
    unsigned short* cs[2];
    for (int k = 0; k < 2; ++k) {
        hdr[k] = reinterpret_cast<SplitHeader*>(w);
        w += align_up((size_t)sp.batch * sizeof(SplitHeader), 256);
        cx[k] = reinterpret_cast<double*>(w);
        w += align_up((size_t)sp.batch * sp.cap * 8, 256);
        cs[k] = reinterpret_cast<unsigned short*>(w);
        w += align_up((size_t)sp.batch * sp.cap * 2, 256);
    }
    int* fb_list = reinterpret_cast<int*>(w);
    w += align_up((size_t)std::max<long long>(rp.n_rows, 1) * 4, 256);
    int* fb_count = reinterpret_cast<int*>(w);
    int msq = (int)std::sqrt((double)rp.M);
    while (msq * msq > rp.M) --msq;
    while ((msq + 1) * (msq + 1) <= rp.M) ++msq;
    CK(cudaMemsetAsync(fb_count, 0, sizeof(int), st));
    SplitParams prev;
    memset(&prev, 0, sizeof(prev));
    long long prev_rows = 0;
    int slot = 0;
    // equal rounds (no short last round whose stream kernel would be nothing but the previous apply stage)
    const long long n_rounds = std::max<long long>(1, (rp.n_rows + sp.batch - 1) / sp.batch);
    const long long per_round = std::min<long long>(sp.batch, ((rp.n_rows + n_rounds - 1) / n_rounds + 7) / 8 * 8);
    for (long long i0 = 0; i0 < rp.n_rows || prev_rows > 0; i0 += per_round, slot ^= 1) {
        const long long nb = std::max<long long>(0, std::min<long long>(per_round, rp.n_rows - i0));
        SplitParams q;
        memset(&q, 0, sizeof(q));
        q.in = rp.in + i0 * rp.in_stride; q.in_stride = rp.in_stride;
        q.out = rp.out ? rp.out + i0 * rp.out_stride : nullptr; q.out_stride = rp.out_stride;
        q.k_out = rp.k_out + i0;
        q.elpd_i = rp.elpd_i ? rp.elpd_i + i0 : nullptr; q.lppd_i = rp.lppd_i ? rp.lppd_i + i0 : nullptr;
        q.var_i = rp.var_i ? rp.var_i + i0 : nullptr; q.lppdw_i = rp.lppdw_i ? rp.lppdw_i + i0 : nullptr;
        q.diag = rp.diag ? rp.diag + i0 * DIAG_STRIDE : nullptr;
        q.n_rows = nb; q.S = rp.S; q.M = rp.M; q.cap = sp.cap; q.nbuf = sp.nbuf; q.q0 = sp.q0; q.m_full = 30 + msq;
        q.cutoffmin = rp.cutoffmin; q.counters = rp.counters; q.hdr = hdr[slot]; q.cx = cx[slot]; q.cs = cs[slot];
        q.fb_list = fb_list; q.fb_count = fb_count; q.row_base = i0; q.a_chunk = sp.a_chunk;
        q.log_S = std::log((double)rp.S);
        q.tail_idx = rp.tail_idx ? rp.tail_idx + i0 * rp.tail_ld : nullptr; q.tail_ld = rp.tail_ld;
        if (sp.fused && prev_rows > 0) {  // the previous batch's apply stage rides along
            q.a_in = prev.in; q.a_out = prev.out; q.a_hdr = prev.hdr; q.a_cx = prev.cx; q.a_cs = prev.cs;
            q.a_rows = prev_rows;
        }
        if (nb == 0 && q.a_rows > 0) {
            // after the last round: its apply stage alone, on the dedicated streaming kernel (all warps copy)
            ProfScope prof(B2L_PROF_APPLY, st);
            CK(split_apply_launch((int)std::min<long long>(prev_rows, 8ll * pl.sms), st, prev));
        } else if (nb > 0) {
            const int g1 = (int)std::min<long long>(sp.grid1, std::max<long long>(nb, q.a_rows));
            ProfScope prof(B2L_PROF_STREAM, st);
            CK(split_stream_launch(sp.nt, sp.ept, mode, g1, sp.smem1, st, q));
        }
        if (nb > 0) {
            const int g2 = (int)std::min<long long>(sp.grid2, (nb + sp.tw - 1) / sp.tw);
            {
                ProfScope prof(B2L_PROF_TAIL, st);
                CK(split_tail_launch(sp.tl, sp.tw, mode, g2, sp.smem2, st, q));
            }
            if (mode == MODE_PSISLW && !sp.fused) {
                ProfScope prof(B2L_PROF_APPLY, st);
                CK(split_apply_launch((int)std::min<long long>(nb, 8ll * pl.sms), st, q));
            }
        }
        prev = q;
        prev_rows = (mode == MODE_PSISLW && sp.fused) ? nb : 0;
    }
    // rows handed over: ONE launch of the general kernel driven by the device-side list (usually empty)
    RowParams r = rp;
    r.row_list = fb_list; r.n_list = fb_count;
    RowPlan pf = pl;
    pf.grid = std::min(pl.grid, pl.sms);
    return launch_rows(mode, pf, r, st);
}

// rows contiguous: split path when the shape and alignment allow it, else the general kernel alone
static int process_rows(int mode, const RowPlan& pl, const SplitPlan& sp, const RowParams& rp, void* sws,
                        cudaStream_t st) {
    if (sp.ok && sws && rp.use_bulk && !rp.waic_only && !rp.row_ws && !getenv("B2L_FORCE_LEGACY"))
        return launch_split(mode, pl, sp, rp, sws, st);
    return launch_rows(mode, pl, rp, st);
}

// ------------------------------------------------------------------------------------ tile path
// pl.loo on the observation-fastest (S, N) matrix without a transposed copy: per round of observations the
// cluster kernel of b2l_tile.cu (one read of HBM through 2-D TMA boxes) and the split path's tail kernel, then
// the general kernel on the observations those two hand over (strided reads of their columns).
static bool tile_eligible(const double* ll, long long S, long long N, long long stride_s, int M) {
    if (getenv("B2L_FORCE_LEGACY")) return false;
    if (N < 1 || N > (1ll << 30) || stride_s < N) return false;
    if (!aligned16_ptr(ll) || (stride_s % 2) != 0) return false;          // TMA: 16 B aligned base and row pitch
    if ((unsigned long long)stride_s * 8ull >= (1ull << 40)) return false;  // TMA: pitch < 2^40 bytes
    TilePlan tp;
    return tile_shape(S, M, TILE_MAXC, TILE_W, &tp);
}

// scratch of the tile path inside the caller's workspace: [hand-over count][hand-over list: N][one round: headers |
// candidate x | candidate draw index | candidate counters].  Rounds are as long as the workspace allows (up to
// B2L_TILE_ROUND observations): a launch's fixed costs and its last, partly filled wave of clusters weigh less.
static size_t tile_round_bytes(const SplitPlan& sp, long long P) {
    return align_up((size_t)P * sizeof(SplitHeader), 256) + align_up((size_t)P * sp.cap * 8, 256) +
           align_up((size_t)P * sp.cap * 2, 256) + align_up((size_t)P * 8, 256) +
           align_up((size_t)P * 4 * sizeof(ChunkHeader), 256);  // (chunk records: up to 4 per observation)
}
static size_t tile_fixed_bytes(long long N) { return 256 + align_up((size_t)std::max<long long>(N, 1) * 4, 256); }
// observations per round the library asks workspace for: every round costs ~50 us of launches and drained waves
// (N = 10^6: 14 rounds of 71 429 observations 25.23 ms per step, 8 rounds 24.96), within ~2 GB of scratch
static long long tile_round_want(const SplitPlan& sp) {
    long long want = 262144;
    while (want > 148ll * 64 * 8 && (size_t)want * ((size_t)sp.cap * 10 + 512) > (2ull << 30)) want /= 2;
    return want;
}
static long long tile_round_obs(const SplitPlan& sp, long long N, size_t avail) {
    long long want = tile_round_want(sp);
    if (const char* ev = getenv("B2L_TILE_ROUND")) want = std::max<long long>(TILE_W, atoll(ev));
    want = std::min<long long>(want, (N + TILE_W - 1) / TILE_W * TILE_W);
    want = want / TILE_W * TILE_W;
    const size_t fixed = tile_fixed_bytes(N);
    while (want > TILE_W && fixed + tile_round_bytes(sp, want) > avail) want = (want / 2 + TILE_W - 1) / TILE_W * TILE_W;
    return (fixed + tile_round_bytes(sp, want) <= avail) ? want : 0;
}

static int launch_tiles(const RowPlan& pl, const SplitPlan& sp, const TilePlan& tp, const double* ll, long long S,
                        long long N, long long stride_s, const RowParams& rp, void* sws, long long P, cudaStream_t st) {
    alignas(64) unsigned char tmap[128];
    CK(tile_tensor_map(ll, S, N, stride_s, tp.tw, tp.box_rows, tmap));
    char* w = (char*)sws;
    int* fb_count = reinterpret_cast<int*>(w);
    w += 256;
    int* fb_list = reinterpret_cast<int*>(w);
    w += align_up((size_t)std::max<long long>(N, 1) * 4, 256);
    SplitHeader* hdr = reinterpret_cast<SplitHeader*>(w);
    w += align_up((size_t)P * sizeof(SplitHeader), 256);
    double* cx = reinterpret_cast<double*>(w);
    w += align_up((size_t)P * sp.cap * 8, 256);
    unsigned short* cs = reinterpret_cast<unsigned short*>(w);
    w += align_up((size_t)P * sp.cap * 2, 256);
    unsigned* cnt = reinterpret_cast<unsigned*>(w);
    w += align_up((size_t)P * 8, 256);
    ChunkHeader* chdr = reinterpret_cast<ChunkHeader*>(w);
    const bool chunked = tp.n_chunks > 1;
    int msq = (int)std::sqrt((double)rp.M);
    while (msq * msq > rp.M) --msq;
    while ((msq + 1) * (msq + 1) <= rp.M) ++msq;
    CK(cudaMemsetAsync(fb_count, 0, sizeof(int), st));
    // equal rounds of whole tiles, so that the last one is not a sliver
    const long long n_rounds = std::max<long long>(1, (N + P - 1) / P);
    const long long per_round = std::min<long long>(P, ((N + n_rounds - 1) / n_rounds + TILE_W - 1) / TILE_W * TILE_W);
    for (long long i0 = 0; i0 < N; i0 += per_round) {
        const long long nb = std::min<long long>(per_round, N - i0);
        TileParams tq;
        memset(&tq, 0, sizeof(tq));
        tq.S = tp.chunk_len; tq.n_chunks = tp.n_chunks; tq.chdr = chdr;
        tq.M = rp.M; tq.cap = sp.cap; tq.R = tp.R; tq.nbox = tp.nbox; tq.box_rows = tp.box_rows;
        tq.q_t = tp.q_t; tq.q_l = tp.q_l; tq.n_tiles = (nb + tp.tw - 1) / tp.tw; tq.col0 = i0; tq.n_obs = nb;
        tq.hdr = hdr; tq.cx = cx; tq.cs = cs; tq.cnt = cnt; tq.fb_list = fb_list; tq.fb_count = fb_count;
        tq.counters = rp.counters; tq.row_base = i0;
        tq.debug = getenv("B2L_TILE_DEBUG") ? atoi(getenv("B2L_TILE_DEBUG")) : 0;
        CK(cudaMemsetAsync(cnt, 0, (size_t)tq.n_tiles * tp.tw * 2 * sizeof(unsigned), st));
        {
            ProfScope prof(B2L_PROF_STREAM, st);
            CK(tile_launch(tp, tmap, tq, st));
            if (chunked && !(tq.debug & ~1)) CK(tile_merge_launch(tq, S, st));
        }
        SplitParams q;
        memset(&q, 0, sizeof(q));
        q.k_out = rp.k_out + i0; q.elpd_i = rp.elpd_i + i0; q.lppd_i = rp.lppd_i + i0; q.var_i = rp.var_i + i0;
        q.lppdw_i = rp.lppdw_i + i0; q.diag = rp.diag ? rp.diag + i0 * DIAG_STRIDE : nullptr;
        q.n_rows = nb; q.S = (int)S; q.M = rp.M; q.cap = sp.cap; q.m_full = 30 + msq; q.cutoffmin = rp.cutoffmin;
        q.counters = rp.counters; q.hdr = hdr; q.cx = cx; q.cs = cs; q.fb_list = fb_list; q.fb_count = fb_count;
        q.row_base = i0; q.total_body = 1; q.ab_lists = chunked ? 0 : 1; q.chunked = chunked ? 1 : 0;
        q.log_S = std::log((double)S);
        q.tail_idx = rp.tail_idx ? rp.tail_idx + i0 * rp.tail_ld : nullptr; q.tail_ld = rp.tail_ld;
        if (tq.debug & ~1) continue;  // measurement aids that leave no valid results: the tile kernel alone
        const int g2 = (int)std::min<long long>(sp.grid2, (nb + sp.tw - 1) / sp.tw);
        ProfScope prof(B2L_PROF_TAIL, st);
        CK(split_tail_launch(sp.tl, sp.tw, MODE_LOO, g2, sp.smem2, st, q));
    }
    // observations handed over: the general kernel reads their columns where they lie
    RowParams r = rp;
    r.in = ll; r.in_stride = 1; r.in_estride = stride_s; r.use_bulk = 0; r.n_rows = N;
    r.row_list = fb_list; r.n_list = fb_count;
    RowPlan pf = pl;
    pf.grid = std::min(pl.grid, pl.sms);
    return launch_rows(MODE_LOO, pf, r, st);
}

// ------------------------------------------------------------------------------------ transpose
// dst[c * dst_ld + r] = src[r * src_ld + c]   for r < rows, c < cols   (32 x 32 tiles, padded smem)
__global__ void __launch_bounds__(256) transpose_f64_kernel(const double* __restrict__ src,
                                                            long long src_ld,
                                                            double* __restrict__ dst,
                                                            long long dst_ld, int rows, int cols) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + 8 * k][tx] = src[r * src_ld + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < rows && c < cols) dst[c * dst_ld + r] = tile[tx][ty + 8 * k];
    }
}

static int launch_transpose(const double* src, long long src_ld, double* dst, long long dst_ld,
                            long long rows, long long cols, cudaStream_t st) {
    if (rows == 0 || cols == 0) return 0;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    if (grid.y > 65535u) return fail(B2L_E_UNSUPPORTED, "transpose: too many rows (%lld)", rows);
    ProfScope prof(B2L_PROF_TRANSPOSE, st);
    transpose_f64_kernel<<<grid, 256, 0, st>>>(src, src_ld, dst, dst_ld, (int)rows, (int)cols);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------ statistics
struct Welford {
    double n, mean, m2;
};
__host__ __device__ inline Welford chan_merge(Welford a, Welford b) {
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    Welford r;
    r.n = a.n + b.n;
    const double d = b.mean - a.mean;
    r.mean = a.mean + d * (b.n / r.n);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
    return r;
}
struct StatAcc {
    Welford e, w;
    double esum, lsum, psum, wsum;
    double kgood, k1, kinf, knan, v04, enan;
    double emin, emax, wmin, wmax;
};
__host__ __device__ inline StatAcc stat_zero() {
    StatAcc a;
    a.e = {0, 0, 0}; a.w = {0, 0, 0};
    a.esum = a.lsum = a.psum = a.wsum = 0;
    a.kgood = a.k1 = a.kinf = a.knan = a.v04 = a.enan = 0;
    a.emin = a.wmin = INFINITY; a.emax = a.wmax = -INFINITY;
    return a;
}
__host__ __device__ inline StatAcc stat_merge(const StatAcc& a, const StatAcc& b) {
    StatAcc r;
    r.e = chan_merge(a.e, b.e); r.w = chan_merge(a.w, b.w);
    r.esum = a.esum + b.esum; r.lsum = a.lsum + b.lsum; r.psum = a.psum + b.psum; r.wsum = a.wsum + b.wsum;
    r.kgood = a.kgood + b.kgood; r.k1 = a.k1 + b.k1; r.kinf = a.kinf + b.kinf; r.knan = a.knan + b.knan;
    r.v04 = a.v04 + b.v04; r.enan = a.enan + b.enan;
    r.emin = fmin(a.emin, b.emin); r.emax = fmax(a.emax, b.emax);
    r.wmin = fmin(a.wmin, b.wmin); r.wmax = fmax(a.wmax, b.wmax);
    return r;
}
__host__ __device__ inline void stat_push(StatAcc& a, double e, double k, double l, double v, double lw,
                                          double good_k) {
    const double w = lw - v;
    Welford one = {1.0, e, 0.0};
    a.e = chan_merge(a.e, one);
    Welford onew = {1.0, w, 0.0};
    a.w = chan_merge(a.w, onew);
    a.esum += e; a.lsum += l; a.psum += v; a.wsum += w;
    a.kgood += (k > good_k) ? 1.0 : 0.0;  // inf counts, NaN does not (loo.py:292)
    a.k1 += (k > 1.0) ? 1.0 : 0.0;
    a.kinf += (k == INFINITY) ? 1.0 : 0.0;
    a.knan += (k != k) ? 1.0 : 0.0;
    a.v04 += (v > 0.4) ? 1.0 : 0.0;
    a.enan += (e != e) ? 1.0 : 0.0;
    a.emin = fmin(a.emin, e); a.emax = fmax(a.emax, e);
    a.wmin = fmin(a.wmin, w); a.wmax = fmax(a.wmax, w);
}
constexpr int STAT_WORDS = sizeof(StatAcc) / sizeof(double);

__device__ inline StatAcc stat_shfl_xor(const StatAcc& a, int o) {
    StatAcc r;
    const double* s = reinterpret_cast<const double*>(&a);
    double* d = reinterpret_cast<double*>(&r);
#pragma unroll
    for (int i = 0; i < STAT_WORDS; ++i) d[i] = __shfl_down_sync(FULL, s[i], o);
    return r;
}

// stage 1: each block reduces a contiguous slab in a fixed tree; stage 2 (1 block) merges partials
__global__ void __launch_bounds__(256) stats_partial_kernel(const double* elpd, const double* k,
                                                            const double* lppd, const double* var,
                                                            const double* lppdw, long long N,
                                                            double good_k, StatAcc* partial) {
    __shared__ StatAcc sh[8];
    const long long per = (N + gridDim.x - 1) / gridDim.x;
    const long long i0 = (long long)blockIdx.x * per, i1 = (i0 + per < N) ? i0 + per : N;
    StatAcc a = stat_zero();
    for (long long i = i0 + threadIdx.x; i < i1; i += 256)
        stat_push(a, elpd[i], k[i], lppd[i], var[i], lppdw[i], good_k);
    for (int o = 16; o > 0; o >>= 1) {
        StatAcc b = stat_shfl_xor(a, o);
        a = stat_merge(a, b);  // lanes >= o hold garbage merges; lane 0 holds the tree result
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        StatAcc t = sh[0];
        for (int w = 1; w < 8; ++w) t = stat_merge(t, sh[w]);
        partial[blockIdx.x] = t;
    }
}
// stage 2: one warp merges the block partials in a fixed tree (lane l takes partials l, l + 32, ... in order, then
// a shuffle tree over the lanes): deterministic, and ~5 merge steps deep instead of a 148-step serial chain
__global__ void __launch_bounds__(32) stats_final_kernel(const StatAcc* partial, int nparts,
                                                         const unsigned long long* counters, double* out) {
    const int lane = threadIdx.x;
    StatAcc t = stat_zero();
    for (int i = lane; i < nparts; i += 32) t = stat_merge(t, partial[i]);
    for (int o = 16; o > 0; o >>= 1) {
        StatAcc b = stat_shfl_xor(t, o);
        t = stat_merge(t, b);  // lanes >= o hold garbage merges; lane 0 holds the tree result
    }
    if (lane != 0) return;
    for (int i = 0; i < B2L_STATS_LEN; ++i) out[i] = 0.0;
    out[B2L_ST_N] = t.e.n; out[B2L_ST_ELPD_MEAN] = t.e.mean; out[B2L_ST_ELPD_M2] = t.e.m2;
    out[B2L_ST_ELPD_SUM] = t.esum; out[B2L_ST_LPPD_SUM] = t.lsum; out[B2L_ST_PWAIC_SUM] = t.psum;
    out[B2L_ST_WAIC_MEAN] = t.w.mean; out[B2L_ST_WAIC_M2] = t.w.m2; out[B2L_ST_WAIC_SUM] = t.wsum;
    out[B2L_ST_K_GT_GOOD] = t.kgood; out[B2L_ST_K_GT_1] = t.k1; out[B2L_ST_K_INF] = t.kinf;
    out[B2L_ST_K_NAN] = t.knan; out[B2L_ST_VAR_GT_04] = t.v04; out[B2L_ST_ELPD_NAN] = t.enan;
    out[B2L_ST_ELPD_MIN] = t.emin; out[B2L_ST_ELPD_MAX] = t.emax;
    out[B2L_ST_WAIC_MIN] = t.wmin; out[B2L_ST_WAIC_MAX] = t.wmax;
    if (counters) {
        out[B2L_ST_N_NAN_IN] = (double)counters[0]; out[B2L_ST_N_PINF_IN] = (double)counters[1];
        out[B2L_ST_N_NINF_IN] = (double)counters[2]; out[B2L_ST_N_FALLBACK] = (double)counters[3];
    }
}
constexpr int STATS_BLOCKS = 148;

// ------------------------------------------------------------------------------------ workspace
// layout: [stats partials][panel A (obs-fastest input transposed to rows)][panel B (psislw rows out)]
static long long panel_obs(long long S, long long N) {
    // one panel = one stream -> tail round of the split path (whole waves of both kernels)
    long long p = 148ll * 64;
    while (p > 148 * 6 && p * S * 8 > (1ll << 30)) p /= 2;
    if (const char* ev = getenv("B2L_PANEL")) p = std::max<long long>(32, atoll(ev));
    p = (p + 31) / 32 * 32;
    return std::min(p, (N + 31) / 32 * 32);
}
static size_t stats_ws_bytes() { return align_up(sizeof(StatAcc) * STATS_BLOCKS, 256); }
// rows that do not fit shared memory are staged in global memory: one row per resident CTA
static size_t grows_ws_bytes(long long S) {
    if (S * 8 <= 160 * 1024) return 0;  // certainly fits shared memory
    return align_up((size_t)((S + 1) & ~1ll) * 8 * GROWS_MAX_CTAS, 256);
}

extern "C" int b2l_version(void) { return B2L_VERSION; }
extern "C" const char* b2l_last_error(void) { return g_err; }
extern "C" int b2l_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        cudaGetLastError();
        return fail(B2L_E_NODEVICE, "no CUDA device visible (%s); this engine has no CPU fallback",
                    cudaGetErrorString(e));
    }
    return n;
}

extern "C" int b2l_workspace_bytes(int64_t S, int64_t N, int32_t M, int32_t layout_obs_fastest,
                                   size_t* out_bytes) {
    if (!out_bytes || S < 1 || N < 0) return fail(B2L_E_INVALID, "bad arguments");
    size_t b = stats_ws_bytes() + grows_ws_bytes(S);
    {
        SplitPlan sp;
        const long long rows = layout_obs_fastest ? panel_obs(S, N) : N;
        split_shape(S, M, rows, &sp);
        b += split_ws_bytes(sp, rows);
    }
    if (layout_obs_fastest) b += 2 * align_up((size_t)panel_obs(S, N) * (size_t)S * 8, 256);
    if (layout_obs_fastest) {  // the cluster kernel's rounds (it uses the same workspace when the shape is eligible)
        TilePlan tp;
        SplitPlan sp;
        if (tile_pick(S, M, &tp) && split_shape(S, M, 1ll << 30, &sp)) {
            const long long P = std::min<long long>(tile_round_want(sp), (N + TILE_W - 1) / TILE_W * TILE_W);
            b = std::max(b, stats_ws_bytes() + grows_ws_bytes(S) + tile_fixed_bytes(N) + tile_round_bytes(sp, std::max<long long>(P, TILE_W)));
        }
    }
    *out_bytes = b;
    return 0;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" int b2l_psislw_dev_f64(const double* lw, int64_t S, int64_t N, int64_t stride_s,
                                  int64_t stride_n, int32_t M, double cutoffmin, double* lw_out,
                                  int64_t ostride_s, int64_t ostride_n, double* k_out, double* diag,
                                  void* ws, size_t ws_bytes, void* stream) {
    if (!lw || !lw_out || !k_out || N < 0) return fail(B2L_E_INVALID, "null pointer or negative N");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    RowPlan pl;
    int rc = plan_row(S, M, MODE_PSISLW, N, &pl);
    if (rc) return rc;
    RowParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.S = (int)S; rp.M = M; rp.cutoffmin = cutoffmin; rp.k_out = k_out; rp.diag = diag;
    const size_t gro = grows_ws_bytes(S);
    if (gro) {
        if (!ws || ws_bytes < stats_ws_bytes() + gro) return fail(B2L_E_WORKSPACE, "workspace too small for S=%lld", (long long)S);
        rp.row_ws = reinterpret_cast<double*>((char*)ws + stats_ws_bytes());
        rp.row_ld = (S + 1) & ~1ll;
    }
    const bool rows_in = (stride_s == 1 || S == 1), rows_out = (ostride_s == 1 || S == 1);
    SplitPlan sp;
    const size_t sws_off = stats_ws_bytes() + gro;
    if (rows_in && rows_out) {
        rp.in = lw; rp.in_stride = stride_n; rp.out = lw_out; rp.out_stride = ostride_n; rp.n_rows = N;
        rp.use_bulk = (S % 2 == 0) && aligned16(lw) && aligned16(lw_out) && (stride_n % 2 == 0) &&
                      (ostride_n % 2 == 0);
        rc = plan_split(S, M, MODE_PSISLW, N, &sp);
        if (rc) return rc;
        void* sws = (ws && ws_bytes >= sws_off + split_ws_bytes(sp, N)) ? (char*)ws + sws_off : nullptr;
        return process_rows(MODE_PSISLW, pl, sp, rp, sws, st);
    }
    if (!((stride_n == 1 || N == 1) || rows_in) || !((ostride_n == 1 || N == 1) || rows_out))
        return fail(B2L_E_INVALID, "one of (stride_s, stride_n) must be 1 for input and output");
    // obs-fastest on either side: go through row panels
    const long long P = panel_obs(S, N);
    const size_t panel_bytes = align_up((size_t)P * (size_t)S * 8, 256);
    rc = plan_split(S, M, MODE_PSISLW, P, &sp);
    if (rc) return rc;
    const size_t pan_off = sws_off + split_ws_bytes(sp, P);
    if (!ws || ws_bytes < pan_off + 2 * panel_bytes)
        return fail(B2L_E_WORKSPACE, "workspace too small: need %zu bytes", pan_off + 2 * panel_bytes);
    void* sws = (char*)ws + sws_off;
    double* pa = reinterpret_cast<double*>((char*)ws + pan_off);
    double* pb = reinterpret_cast<double*>((char*)ws + pan_off + panel_bytes);
    for (long long i0 = 0; i0 < N; i0 += P) {
        const long long np = std::min<long long>(P, N - i0);
        RowParams r = rp;
        r.n_rows = np; r.k_out = k_out + i0; r.diag = diag ? diag + i0 * DIAG_STRIDE : nullptr;
        if (rows_in) { r.in = lw + i0 * stride_n; r.in_stride = stride_n; }
        else {
            rc = launch_transpose(lw + i0, stride_s, pa, S, S, np, st);  // (S x np) -> (np x S)
            if (rc) return rc;
            r.in = pa; r.in_stride = S;
        }
        if (rows_out) { r.out = lw_out + i0 * ostride_n; r.out_stride = ostride_n; }
        else { r.out = pb; r.out_stride = S; }
        r.use_bulk = (S % 2 == 0) && aligned16(r.in) && aligned16(r.out) && (r.in_stride % 2 == 0) &&
                     (r.out_stride % 2 == 0);
        rc = process_rows(MODE_PSISLW, pl, sp, r, sws, st);
        if (rc) return rc;
        if (!rows_out) {
            rc = launch_transpose(pb, S, lw_out + i0, ostride_s, np, S, st);  // (np x S) -> (S x np)
            if (rc) return rc;
        }
    }
    return 0;
}

extern "C" int b2l_loo_dev_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s,
                               int64_t stride_n, int32_t M, double cutoffmin, uint32_t flags,
                               double* elpd_i, double* k_i, double* lppd_i, double* var_i,
                               double* lppdw_i, unsigned long long* counters, double* diag, void* ws,
                               size_t ws_bytes, void* stream) {
    return b2l_loo_dev_ex_f64(ll, S, N, stride_s, stride_n, M, cutoffmin, flags, elpd_i, k_i, lppd_i, var_i, lppdw_i,
                              counters, diag, nullptr, ws, ws_bytes, stream);
}

extern "C" int b2l_loo_dev_ex_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s,
                                  int64_t stride_n, int32_t M, double cutoffmin, uint32_t flags,
                                  double* elpd_i, double* k_i, double* lppd_i, double* var_i,
                                  double* lppdw_i, unsigned long long* counters, double* diag,
                                  int32_t* tail_idx, void* ws, size_t ws_bytes, void* stream) {
    if (!ll || !elpd_i || !k_i || !lppd_i || !var_i || !lppdw_i || N < 0)
        return fail(B2L_E_INVALID, "null pointer or negative N");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    RowPlan pl;
    int rc = plan_row(S, M, MODE_LOO, N, &pl);
    if (rc) return rc;
    RowParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.S = (int)S; rp.M = M; rp.cutoffmin = cutoffmin; rp.counters = counters;
    rp.tail_idx = tail_idx; rp.tail_ld = M;
    rp.waic_only = (flags & B2L_FLAG_WAIC_ONLY) ? 1 : 0;
    const size_t gro = grows_ws_bytes(S);
    if (gro) {
        if (!ws || ws_bytes < stats_ws_bytes() + gro) return fail(B2L_E_WORKSPACE, "workspace too small for S=%lld", (long long)S);
        rp.row_ws = reinterpret_cast<double*>((char*)ws + stats_ws_bytes());
        rp.row_ld = (S + 1) & ~1ll;
    }
    SplitPlan sp;
    const size_t sws_off = stats_ws_bytes() + gro;
    if (stride_s == 1 || S == 1) {  // rows contiguous
        rp.in = ll; rp.in_stride = stride_n; rp.n_rows = N;
        rp.k_out = k_i; rp.elpd_i = elpd_i; rp.lppd_i = lppd_i; rp.var_i = var_i; rp.lppdw_i = lppdw_i;
        rp.diag = diag;
        rp.use_bulk = (S % 2 == 0) && aligned16(ll) && (stride_n % 2 == 0);
        rc = plan_split(S, M, MODE_LOO, N, &sp);
        if (rc) return rc;
        void* sws = (ws && ws_bytes >= sws_off + split_ws_bytes(sp, N)) ? (char*)ws + sws_off : nullptr;
        return process_rows(MODE_LOO, pl, sp, rp, sws, st);
    }
    if (!(stride_n == 1 || N == 1))
        return fail(B2L_E_INVALID, "one of (stride_s, stride_n) must be 1");
    if (rp.waic_only && !diag && !getenv("B2L_FORCE_LEGACY")) {
        // WAIC alone needs no rows: one pass over the observation-fastest matrix
        WaicColsParams wp;
        memset(&wp, 0, sizeof(wp));
        wp.ll = ll; wp.stride_s = stride_s; wp.N = N; wp.S = (int)S; wp.log_S = std::log((double)S);
        wp.elpd_i = elpd_i; wp.k_i = k_i; wp.lppd_i = lppd_i; wp.var_i = var_i; wp.lppdw_i = lppdw_i;
        wp.counters = counters;
        ProfScope prof(B2L_PROF_IS, st);
        CK(waic_cols_launch(wp, st));
        return 0;
    }
    const long long P = panel_obs(S, N);
    const size_t panel_bytes = align_up((size_t)P * (size_t)S * 8, 256);
    if (!rp.waic_only && !gro && !(flags & B2L_FLAG_NO_TILE) && tile_eligible(ll, S, N, stride_s, M)) {
        // the matrix is read where it lies (2-D TMA tiles, one pass over HBM): no panels
        rc = plan_split(S, M, MODE_LOO, P, &sp);  // same scratch slots as the panel path sizes (b2l_workspace_bytes)
        if (rc) return rc;
        TilePlan tp;
        CK(tile_plan(S, M, &tp));
        const long long Pt = (sp.ok && tp.ok && ws && ws_bytes > sws_off) ? tile_round_obs(sp, N, ws_bytes - sws_off) : 0;
        if (Pt > 0) {
            rp.k_out = k_i; rp.elpd_i = elpd_i; rp.lppd_i = lppd_i; rp.var_i = var_i; rp.lppdw_i = lppdw_i;
            rp.diag = diag;
            return launch_tiles(pl, sp, tp, ll, S, N, stride_s, rp, (char*)ws + sws_off, Pt, st);
        }
    }
    rc = plan_split(S, M, MODE_LOO, P, &sp);
    if (rc) return rc;
    const size_t pan_off = sws_off + split_ws_bytes(sp, P);
    if (!ws || ws_bytes < pan_off + panel_bytes)
        return fail(B2L_E_WORKSPACE, "workspace too small: need %zu bytes", pan_off + panel_bytes);
    void* sws = (char*)ws + sws_off;
    double* pa = reinterpret_cast<double*>((char*)ws + pan_off);
    for (long long i0 = 0; i0 < N; i0 += P) {
        const long long np = std::min<long long>(P, N - i0);
        rc = launch_transpose(ll + i0, stride_s, pa, S, S, np, st);
        if (rc) return rc;
        RowParams r = rp;
        r.in = pa; r.in_stride = S; r.n_rows = np;
        r.k_out = k_i + i0; r.elpd_i = elpd_i + i0; r.lppd_i = lppd_i + i0; r.var_i = var_i + i0;
        r.lppdw_i = lppdw_i + i0; r.diag = diag ? diag + i0 * DIAG_STRIDE : nullptr;
        r.tail_idx = tail_idx ? tail_idx + i0 * (long long)M : nullptr;
        r.use_bulk = (S % 2 == 0) && aligned16(pa);
        rc = process_rows(MODE_LOO, pl, sp, r, sws, st);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int b2l_stats_dev_f64(const double* elpd_i, const double* k_i, const double* lppd_i,
                                 const double* var_i, const double* lppdw_i, int64_t N, double good_k,
                                 const unsigned long long* counters, double* stats_out, void* ws,
                                 size_t ws_bytes, void* stream) {
    if (!elpd_i || !k_i || !lppd_i || !var_i || !lppdw_i || !stats_out || N < 0)
        return fail(B2L_E_INVALID, "null pointer or negative N");
    if (!ws || ws_bytes < stats_ws_bytes()) return fail(B2L_E_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    StatAcc* partial = reinterpret_cast<StatAcc*>(ws);
    ProfScope prof(B2L_PROF_STATS, st);
    stats_partial_kernel<<<STATS_BLOCKS, 256, 0, st>>>(elpd_i, k_i, lppd_i, var_i, lppdw_i, N, good_k, partial);
    CK(cudaGetLastError());
    stats_final_kernel<<<1, 32, 0, st>>>(partial, STATS_BLOCKS, counters, stats_out);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2l_stats_merge(const double* shards, int32_t n_shards, double* merged) {
    if (!shards || !merged || n_shards < 1) return fail(B2L_E_INVALID, "bad arguments");
    Welford e = {0, 0, 0}, w = {0, 0, 0};
    double out[B2L_STATS_LEN];
    for (int i = 0; i < B2L_STATS_LEN; ++i) out[i] = 0.0;
    out[B2L_ST_ELPD_MIN] = out[B2L_ST_WAIC_MIN] = INFINITY;
    out[B2L_ST_ELPD_MAX] = out[B2L_ST_WAIC_MAX] = -INFINITY;
    static const int sums[] = {B2L_ST_ELPD_SUM, B2L_ST_LPPD_SUM, B2L_ST_PWAIC_SUM, B2L_ST_WAIC_SUM,
                               B2L_ST_K_GT_GOOD, B2L_ST_K_GT_1, B2L_ST_K_INF, B2L_ST_K_NAN,
                               B2L_ST_VAR_GT_04, B2L_ST_N_NAN_IN, B2L_ST_N_PINF_IN, B2L_ST_N_NINF_IN,
                               B2L_ST_N_FALLBACK, B2L_ST_ELPD_NAN};
    for (int r = 0; r < n_shards; ++r) {
        const double* s = shards + (size_t)r * B2L_STATS_LEN;
        e = chan_merge(e, Welford{s[B2L_ST_N], s[B2L_ST_ELPD_MEAN], s[B2L_ST_ELPD_M2]});
        w = chan_merge(w, Welford{s[B2L_ST_N], s[B2L_ST_WAIC_MEAN], s[B2L_ST_WAIC_M2]});
        for (int id : sums) out[id] += s[id];
        if (s[B2L_ST_N] > 0) {
            out[B2L_ST_ELPD_MIN] = std::fmin(out[B2L_ST_ELPD_MIN], s[B2L_ST_ELPD_MIN]);
            out[B2L_ST_ELPD_MAX] = std::fmax(out[B2L_ST_ELPD_MAX], s[B2L_ST_ELPD_MAX]);
            out[B2L_ST_WAIC_MIN] = std::fmin(out[B2L_ST_WAIC_MIN], s[B2L_ST_WAIC_MIN]);
            out[B2L_ST_WAIC_MAX] = std::fmax(out[B2L_ST_WAIC_MAX], s[B2L_ST_WAIC_MAX]);
        }
    }
    out[B2L_ST_N] = e.n; out[B2L_ST_ELPD_MEAN] = e.mean; out[B2L_ST_ELPD_M2] = e.m2;
    out[B2L_ST_WAIC_MEAN] = w.mean; out[B2L_ST_WAIC_M2] = w.m2;
    memcpy(merged, out, sizeof(out));
    return 0;
}

// ------------------------------------------------------------------------------------ SIS / TIS / e_loo
static bool bulk_ok(long long S, const void* base, long long stride) {
    return (S % 2 == 0) && aligned16(base) && (stride % 2 == 0) && (S * 8 < (1ll << 20));
}
static long long is_panel_obs(long long S, long long N) {
    long long p = 148ll * 32;
    while (p > 148 && p * S * 8 > (1ll << 29)) p /= 2;
    return std::min(p, std::max<long long>(N, 1));
}

extern "C" int b2l_islw_dev_f64(const double* lw, int64_t S, int64_t N, int64_t stride_n, int32_t method,
                                double* lw_out, int64_t ostride_n, double* ess_out, void* stream) {
    if (!lw || !lw_out || !ess_out || N < 0 || S < 1) return fail(B2L_E_INVALID, "null pointer or bad size");
    if (method != B2L_IS_SIS && method != B2L_IS_TIS) return fail(B2L_E_INVALID, "method must be B2L_IS_SIS or B2L_IS_TIS");
    if (S > INT32_MAX) return fail(B2L_E_UNSUPPORTED, "S too large");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    IsParams p;
    memset(&p, 0, sizeof(p));
    p.in = lw; p.in_stride = stride_n; p.out = lw_out; p.out_stride = ostride_n; p.ess = ess_out;
    p.n_rows = N; p.S = (int)S; p.bulk = bulk_ok(S, lw, stride_n); p.log_S = std::log((double)S);
    ProfScope prof(B2L_PROF_IS, st);
    CK(is_launch(method, IS_MODE_WEIGHTS, p, st));
    return 0;
}

extern "C" int b2l_is_workspace_bytes(int64_t S, int64_t N, int32_t layout_obs_fastest, size_t* out_bytes) {
    if (!out_bytes || S < 1 || N < 0) return fail(B2L_E_INVALID, "bad arguments");
    *out_bytes = layout_obs_fastest ? align_up((size_t)is_panel_obs(S, N) * (size_t)S * 8, 256) : 0;
    return 0;
}

extern "C" int b2l_loo_is_dev_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                                  int32_t method, double* elpd_i, double* ess_i, double* lppd_i,
                                  unsigned long long* counters, void* ws, size_t ws_bytes, void* stream) {
    if (!ll || !elpd_i || !ess_i || !lppd_i || N < 0 || S < 1) return fail(B2L_E_INVALID, "null pointer or bad size");
    if (method != B2L_IS_SIS && method != B2L_IS_TIS) return fail(B2L_E_INVALID, "method must be B2L_IS_SIS or B2L_IS_TIS");
    if (S > INT32_MAX) return fail(B2L_E_UNSUPPORTED, "S too large");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    IsParams p;
    memset(&p, 0, sizeof(p));
    p.S = (int)S; p.log_S = std::log((double)S); p.counters = counters;
    if (stride_s == 1 || S == 1) {
        p.in = ll; p.in_stride = stride_n; p.n_rows = N; p.bulk = bulk_ok(S, ll, stride_n);
        p.elpd = elpd_i; p.ess = ess_i; p.lppd = lppd_i;
        ProfScope prof(B2L_PROF_IS, st);
        CK(is_launch(method, IS_MODE_LOO, p, st));
        return 0;
    }
    if (stride_n != 1 && N != 1) return fail(B2L_E_INVALID, "one of (stride_s, stride_n) must be 1");
    if (!getenv("B2L_FORCE_LEGACY")) {
        // observation-fastest layout: column form, no transposed panels
        IsColsParams cp;
        memset(&cp, 0, sizeof(cp));
        cp.ll = ll; cp.stride_s = stride_s; cp.N = N; cp.S = (int)S; cp.log_S = std::log((double)S);
        cp.elpd = elpd_i; cp.ess = ess_i; cp.lppd = lppd_i; cp.counters = counters;
        ProfScope prof(B2L_PROF_IS, st);
        CK(is_cols_launch(method, cp, st));
        return 0;
    }
    const long long P = is_panel_obs(S, N);
    const size_t need = align_up((size_t)P * (size_t)S * 8, 256);
    if (!ws || ws_bytes < need) return fail(B2L_E_WORKSPACE, "workspace too small: need %zu bytes", need);
    double* panel = reinterpret_cast<double*>(ws);
    for (long long i0 = 0; i0 < N; i0 += P) {
        const long long np = std::min<long long>(P, N - i0);
        int rc = launch_transpose(ll + i0, stride_s, panel, S, S, np, st);  // (S x np) -> (np x S)
        if (rc) return rc;
        p.in = panel; p.in_stride = S; p.n_rows = np; p.bulk = bulk_ok(S, panel, S);
        p.elpd = elpd_i + i0; p.ess = ess_i + i0; p.lppd = lppd_i + i0;
        ProfScope prof(B2L_PROF_IS, st);
        CK(is_launch(method, IS_MODE_LOO, p, st));
    }
    return 0;
}

extern "C" int b2l_eloo_workspace_bytes(int64_t S, int64_t N, int32_t has_lr, int32_t type, size_t* out_bytes) {
    if (!out_bytes || S < 1 || N < 0 || S > INT32_MAX) return fail(B2L_E_INVALID, "bad arguments");
    int info[4] = {0, 0, 0, 0};
    (void)has_lr;
    CK(eloo_plan((int)S, std::max<long long>(N, 1), type != B2L_ELOO_NONE, ELOO_MAX_TAIL, info));
    *out_bytes = info[0] ? 0 : align_up((size_t)info[1] * (size_t)((S + 1) & ~1ll) * 8, 256);
    return 0;
}

extern "C" int b2l_eloo_dev_f64(const double* x, int64_t x_stride_n, const double* lw, int64_t lw_stride_n,
                                const double* lr, int64_t lr_stride_n, int64_t S, int64_t N, int32_t type,
                                int32_t tail_len, double* value_out, double* khat_out, void* ws,
                                size_t ws_bytes, void* stream) {
    if (!lw || !khat_out || N < 0 || S < 1) return fail(B2L_E_INVALID, "null pointer or bad size");
    if (type < B2L_ELOO_MEAN || type > B2L_ELOO_NONE) return fail(B2L_E_INVALID, "bad expectation type %d", type);
    if (type != B2L_ELOO_NONE && (!x || !value_out)) return fail(B2L_E_INVALID, "x and value_out are required");
    if (tail_len < 5) return fail(B2L_E_INVALID, "tail_len must be at least 5");  // e_loo.py:295-296
    if (tail_len > ELOO_MAX_TAIL) return fail(B2L_E_UNSUPPORTED, "tail_len > %d", ELOO_MAX_TAIL);
    if (S > INT32_MAX) return fail(B2L_E_UNSUPPORTED, "S too large");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    ElooParams p;
    memset(&p, 0, sizeof(p));
    p.x = (type == B2L_ELOO_NONE) ? nullptr : x; p.x_stride = x_stride_n;
    p.lw = lw; p.lw_stride = lw_stride_n;
    p.lr = lr ? lr : lw; p.lr_stride = lr ? lr_stride_n : lw_stride_n;
    p.value = value_out; p.khat = khat_out; p.n_rows = N; p.S = (int)S; p.type = type; p.tail_len = tail_len;
    p.bulk = bulk_ok(S, p.lr, p.lr_stride) && (!p.x || bulk_ok(S, p.x, x_stride_n));  // the two staged rows
    int info[4] = {0, 0, 0, 0};
    CK(eloo_plan((int)S, N, p.x != nullptr, tail_len, info));
    if (!info[0]) {
        // rows that do not fit shared memory: one scratch row per resident CTA, as many as the workspace holds
        const size_t row = (size_t)((S + 1) & ~1ll) * 8;
        const size_t rows_held = ws ? ws_bytes / row : 0;
        if (rows_held < 1) return fail(B2L_E_WORKSPACE, "workspace too small: need at least %zu bytes", row);
        p.scratch = reinterpret_cast<double*>(ws);
        p.grid_cap = (int)std::min<size_t>(rows_held, 1u << 20);
    }
    ProfScope prof(B2L_PROF_ELOO, st);
    CK(eloo_launch(p, st));
    return 0;
}

extern "C" int b2l_eloo_quantile_dev_f64(const double* x, int64_t x_stride_n, const double* lw,
                                         int64_t lw_stride_n, int64_t S, int64_t N, const double* probs,
                                         int32_t n_probs, double* value_out, void* stream) {
    if (!x || !lw || !probs || !value_out || N < 0 || S < 1) return fail(B2L_E_INVALID, "null pointer or bad size");
    if (n_probs < 1 || n_probs > ELOO_MAX_PROBS) return fail(B2L_E_INVALID, "n_probs must be in 1..%d", ELOO_MAX_PROBS);
    if (S > ELOO_QUANT_MAX_S)
        return fail(B2L_E_UNSUPPORTED, "weighted quantiles sort each row in shared memory: S <= %d", ELOO_QUANT_MAX_S);
    for (int i = 0; i < n_probs; ++i)
        if (!(probs[i] > 0.0 && probs[i] < 1.0)) return fail(B2L_E_INVALID, "probs must be between 0 and 1");  // e_loo.py:161-162
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    QuantParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.x_stride = x_stride_n; p.lw = lw; p.lw_stride = lw_stride_n; p.out = value_out;
    p.n_rows = N; p.S = (int)S; p.n_probs = n_probs;
    for (int i = 0; i < n_probs; ++i) p.probs[i] = probs[i];
    ProfScope prof(B2L_PROF_ELOO, st);
    CK(eloo_quantile_launch(p, st));
    return 0;
}

extern "C" int b2l_group_sum_dev_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                                     const int32_t* members, const int32_t* offsets, int32_t G, double* out,
                                     int64_t out_stride_g, unsigned long long* counters, void* stream) {
    if (!ll || !members || !offsets || !out || S < 1 || N < 1 || G < 1) return fail(B2L_E_INVALID, "null pointer or bad size");
    if (S > INT32_MAX || N > INT32_MAX) return fail(B2L_E_UNSUPPORTED, "S or N too large");
    if (out_stride_g < S) return fail(B2L_E_INVALID, "out_stride_g < S");
    cudaStream_t st = (cudaStream_t)stream;
    GroupSumParams p;
    memset(&p, 0, sizeof(p));
    p.ll = ll; p.stride_s = stride_s; p.stride_n = stride_n; p.members = members; p.offsets = offsets;
    p.out = out; p.out_stride = out_stride_g; p.counters = counters; p.S = (int)S; p.G = G;
    ProfScope prof(B2L_PROF_IS, st);
    CK(group_sum_launch(p, st));
    return 0;
}

// ------------------------------------------------------------------------------------ gather (loo_subsample)
// dst[j * dst_ld + s] = src[s * stride_s + idx[j] * stride_n]  for j < m, s < S   (32 x 32 tiles, padded smem):
// the subsampled observations of pyloo/loo_subsample.py:330 (`log_likelihood.isel(...)`) as contiguous rows.
__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ src, long long stride_s,
                                                          long long stride_n, const long long* __restrict__ idx,
                                                          double* __restrict__ dst, long long dst_ld, int S, int m) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const long long j0 = (long long)blockIdx.x * 32, s0 = (long long)blockIdx.y * 32;
    const long long col = (j0 + tx < m) ? idx[j0 + tx] : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long s = s0 + ty + 8 * k;
        if (s < S && j0 + tx < m) tile[ty + 8 * k][tx] = src[s * stride_s + col * stride_n];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long j = j0 + ty + 8 * k, s = s0 + tx;
        if (j < m && s < S) dst[j * dst_ld + s] = tile[tx][ty + 8 * k];
    }
}

extern "C" int b2l_gather_rows_dev_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                                       const int64_t* idx, int64_t m, double* out, int64_t out_stride_n,
                                       void* stream) {
    if (!ll || !idx || !out || S < 1 || N < 1 || m < 0) return fail(B2L_E_INVALID, "null pointer or bad size");
    if (out_stride_n < S) return fail(B2L_E_INVALID, "out_stride_n < S");
    if (S > INT32_MAX || m > INT32_MAX) return fail(B2L_E_UNSUPPORTED, "S or m too large");
    if (m == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((S + 31) / 32));
    if (grid.y > 65535u) return fail(B2L_E_UNSUPPORTED, "gather: too many draws (%lld)", (long long)S);
    ProfScope prof(B2L_PROF_TRANSPOSE, st);
    gather_rows_kernel<<<grid, 256, 0, st>>>(ll, stride_s, stride_n, reinterpret_cast<const long long*>(idx), out,
                                             out_stride_n, (int)S, (int)m);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2l_handover_reasons(uint64_t* out16, int32_t reset) {
    if (!out16) return fail(B2L_E_INVALID, "null pointer");
    unsigned long long a[HO_REASONS], b[HO_REASONS], c[HO_REASONS];
    CK(cudaDeviceSynchronize());
    CK(split_stream_reasons(a, reset));
    CK(split_tail_reasons(b, reset));
    CK(tile_reasons(c, reset));
    for (int i = 0; i < HO_REASONS; ++i) out16[i] = a[i] + b[i] + c[i];
    return 0;
}

extern "C" int b2l_profile(int32_t enable) {
    for (auto& r : g_prof_recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
    g_prof = enable != 0;
    return 0;
}

extern "C" int b2l_profile_read(double* ms_out, int64_t* launches_out) {
    if (!ms_out || !launches_out) return fail(B2L_E_INVALID, "null pointer");
    for (int k = 0; k < B2L_PROF_KINDS; ++k) {
        ms_out[k] = 0.0;
        launches_out[k] = 0;
    }
    for (auto& r : g_prof_recs) {
        CK(cudaEventSynchronize(r.b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_out[r.kind] += ms;
        launches_out[r.kind] += (r.kind == B2L_PROF_STATS) ? 2 : 1;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
    return 0;
}

extern "C" int b2l_split_launch_info(int64_t S, int32_t M, int32_t mode, int64_t n_rows, int32_t* info) {
    if (!info) return fail(B2L_E_INVALID, "null pointer");
    SplitPlan sp;
    int rc = plan_split(S, M, mode ? MODE_LOO : MODE_PSISLW, n_rows, &sp);
    if (rc) return rc;
    info[0] = sp.ok; info[1] = sp.nt; info[2] = sp.ept; info[3] = sp.tl; info[4] = sp.cap; info[5] = sp.q0;
    info[6] = sp.nbuf; info[7] = sp.fused; info[8] = sp.grid1; info[9] = sp.grid2; info[10] = sp.occ1;
    info[11] = sp.occ2; info[12] = (int)sp.smem1; info[13] = (int)sp.smem2; info[14] = (int)sp.batch;
    info[15] = stream_block(sp.nt, mode ? MODE_LOO : MODE_PSISLW);
    info[7] = sp.fused ? (int)((S + sp.a_chunk - 1) / std::max(sp.a_chunk, 1)) : 0;  // apply transfers per row
    return 0;
}

extern "C" int b2l_tile_shape_info(int64_t S, int32_t M, int32_t* info) {
    if (!info) return fail(B2L_E_INVALID, "null pointer");
    TilePlan tp;
    SplitPlan sp;
    const bool ok = tile_pick(S, M, &tp) && split_shape(S, M, 1ll << 30, &sp);  // (the tail kernel must serve M too)
    info[0] = ok ? 1 : 0; info[1] = tp.tw; info[2] = tp.csize; info[3] = tp.n_chunks; info[4] = tp.chunk_len;
    info[5] = tp.R; info[6] = tp.nbox; info[7] = tp.box_rows; info[8] = tp.q_t; info[9] = tp.q_l;
    info[10] = (int)tp.smem; info[11] = ok ? sp.tl : 0; info[12] = ok ? sp.cap : 0; info[13] = 0; info[14] = 0; info[15] = 0;
    return 0;
}

extern "C" int b2l_row_launch_info(int64_t S, int32_t M, int32_t mode, int32_t* grid, int32_t* block,
                                   int32_t* smem_bytes, int32_t* ctas_per_sm, int32_t* nbuf) {
    RowPlan pl;
    int rc = plan_row(S, M, mode ? MODE_LOO : MODE_PSISLW, 1ll << 40, &pl);
    if (rc) return rc;
    if (grid) *grid = pl.grid;
    if (block) *block = pl.nt;
    if (smem_bytes) *smem_bytes = (int)pl.smem;
    if (ctas_per_sm) *ctas_per_sm = pl.ctas_per_sm;
    if (nbuf) *nbuf = pl.nbuf;
    return 0;
}

// ------------------------------------------------------------------------------------ host pipelines
// Per-device cache of staging buffers and streams so repeated calls do not pay cudaMalloc.
namespace {
constexpr int NSLOT = 3;
struct Slot {
    cudaStream_t st = nullptr;
    void* buf = nullptr;
    size_t bytes = 0;
    // pinned bounce buffers for pageable callers (NumPy arrays): host threads copy a chunk in / out of these, the
    // DMA engines see pinned memory only (a cudaMemcpyAsync on pageable memory is staged by the driver, serially)
    void* hin = nullptr;
    size_t hin_bytes = 0;
    void* hout = nullptr;
    size_t hout_bytes = 0;
    // psislw: a chunk of results waiting in `hout` for its copy to the caller's array
    double* pend_dst = nullptr;
    long long pend_rows = 0, pend_width = 0, pend_dpitch = 0;
};
struct DevCtx {
    std::mutex mu;  // one host pipeline at a time per device; different devices run concurrently
    Slot slot[NSLOT];
    bool init = false;
    // pinned staging for the small per-observation outputs: an async copy into the caller's (usually
    // pageable) vectors would block the host until the whole chunk is done and serialise the pipeline
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
};
DevCtx g_ctx[64];

// the calling thread's current device is restored on every exit path of a host entry point
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
// work queued on the chunk streams must not outlive a failed call (its buffers may be reused or freed)
struct SlotDrain {
    DevCtx& cx;
    explicit SlotDrain(DevCtx& c) : cx(c) {}
    ~SlotDrain() {
        for (int s = 0; s < NSLOT; ++s)
            if (cx.slot[s].st) cudaStreamSynchronize(cx.slot[s].st);
    }
};

int pinned_reserve(DevCtx& cx, size_t bytes) {
    if (cx.pinned_bytes < bytes) {
        if (cx.pinned) CK(cudaFreeHost(cx.pinned));
        cx.pinned = nullptr; cx.pinned_bytes = 0;
        CK(cudaHostAlloc(&cx.pinned, bytes, cudaHostAllocDefault));
        cx.pinned_bytes = bytes;
    }
    return 0;
}
int slot_reserve(Slot& s, size_t bytes) {
    if (!s.st) CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    if (s.bytes < bytes) {
        if (s.buf) CK(cudaFree(s.buf));
        s.buf = nullptr; s.bytes = 0;
        CK(cudaMalloc(&s.buf, bytes));
        s.bytes = bytes;
    }
    return 0;
}
int hpin_reserve(void** p, size_t* have, size_t bytes) {
    if (*have < bytes) {
        if (*p) CK(cudaFreeHost(*p));
        *p = nullptr; *have = 0;
        CK(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
        *have = bytes;
    }
    return 0;
}
bool host_pageable(const void* p) {
    if (const char* ev = getenv("B2L_HOST_BOUNCE")) if (atoi(ev) == 0) return false;
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}
std::atomic<int> g_active_pipes{0};  // host pipelines running now (one per device): they share the host cores
struct PipeCount {
    PipeCount() { ++g_active_pipes; }
    ~PipeCount() { --g_active_pipes; }
};
int host_threads() {
    if (const char* ev = getenv("B2L_HOST_THREADS")) return std::max(1, std::min(atoi(ev), 64));
    const int hw = (int)std::thread::hardware_concurrency();
    const int pipes = std::max(1, g_active_pipes.load());
    return std::max(2, std::min(12, (hw > 2 ? hw - 2 : 2) / pipes));
}
// One row of a staging copy.  The destination is not read by the CPU again soon (a pinned bounce buffer the DMA engine
// reads, or a result array of gigabytes), so the stores bypass the cache: no read-for-ownership of the destination
// lines (a third less memory traffic than memcpy's cached stores at these row sizes, 37 - 75 KB) and the source
// stays the only stream through the cache.  B2L_HOST_NT=0 selects plain memcpy.
void copy_row_streaming(char* dst, const char* src, size_t n, bool nt_stores) {
#if defined(__SSE2__)
    if (nt_stores && n >= 256) {
        size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
        if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
        const size_t blocks = n / 64;
        for (size_t i = 0; i < blocks; ++i) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), d);
            src += 64; dst += 64;
        }
        if (n % 64) memcpy(dst, src, n % 64);
        return;
    }
#endif
    (void)nt_stores;
    memcpy(dst, src, n);
}
// rows x width bytes between two pitched host buffers, the rows split over a few threads
void par_copy_2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, long long rows) {
    const int nt = (int)std::min<long long>(host_threads(), std::max<long long>(1, rows));
    const bool nts = !(getenv("B2L_HOST_NT") && atoi(getenv("B2L_HOST_NT")) == 0);
    auto work = [=](int t) {
        const long long r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        if (dpitch == width && spitch == width) {
            copy_row_streaming(dst + (size_t)r0 * width, src + (size_t)r0 * width, (size_t)(r1 - r0) * width, nts);
        } else {
            for (long long r = r0; r < r1; ++r) copy_row_streaming(dst + (size_t)r * dpitch, src + (size_t)r * spitch, width, nts);
        }
#if defined(__SSE2__)
        if (nts) _mm_sfence();  // the streamed lines are globally visible before the thread is joined
#endif
    };
    if (nt == 1) { work(0); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}
long long default_chunk(long long S, long long N) {
    // one chunk = one round of the split path (whole waves of the stream and tail kernels)
    long long c = 148ll * 64;
    while (c > 148 * 6 && c * S * 8 > (1ll << 28)) c /= 2;
    return std::min(c, std::max<long long>(N, 1));
}
struct Carve {
    char* p;
    explicit Carve(void* b) : p((char*)b) {}
    template <class T> T* take(size_t n) {
        T* r = reinterpret_cast<T*>(p);
        p += align_up(n * sizeof(T), 256);
        return r;
    }
};
}  // namespace

extern "C" int b2l_psislw_host_f64(const double* lw, int64_t S, int64_t N, int64_t stride_s,
                                   int64_t stride_n, int32_t M, double cutoffmin, double* lw_out,
                                   int64_t ostride_s, int64_t ostride_n, double* k_out, int32_t device,
                                   int64_t chunk_obs) {
    if (!lw || !lw_out || !k_out || N < 0 || S < 1) return fail(B2L_E_INVALID, "bad arguments");
    if (b2l_device_count() < 1) return B2L_E_NODEVICE;
    if (N == 0) return 0;
    const bool rows_in = (stride_s == 1 || S == 1), rows_out = (ostride_s == 1 || S == 1);
    if (!rows_in && !(stride_n == 1 || N == 1)) return fail(B2L_E_INVALID, "input must have a unit stride");
    if (!rows_out && !(ostride_n == 1 || N == 1)) return fail(B2L_E_INVALID, "output must have a unit stride");
    if (device < 0 || device >= 64) return fail(B2L_E_INVALID, "device index %d out of range", device);
    DevCtx& cx = g_ctx[device];
    std::lock_guard<std::mutex> lock(cx.mu);
    DeviceGuard restore;
    CK(cudaSetDevice(device));
    SlotDrain drain(cx);
    PipeCount pipe;
    const long long chunk = chunk_obs > 0 ? std::min<long long>(chunk_obs, N) : default_chunk(S, N);
    size_t wsb = 0;
    b2l_workspace_bytes(S, chunk, M, (!rows_in || !rows_out) ? 1 : 0, &wsb);
    const size_t mat = align_up((size_t)chunk * (size_t)S * 8, 256);
    const size_t need = 2 * mat + align_up((size_t)chunk * 8, 256) + wsb + 1024;
    const bool bounce_in = host_pageable(lw), bounce_out = host_pageable(lw_out);
    for (int s = 0; s < NSLOT; ++s) {
        int rc = slot_reserve(cx.slot[s], need);
        if (rc) return rc;
        if (bounce_in) rc = hpin_reserve(&cx.slot[s].hin, &cx.slot[s].hin_bytes, mat);
        if (rc) return rc;
        if (bounce_out) rc = hpin_reserve(&cx.slot[s].hout, &cx.slot[s].hout_bytes, mat);
        if (rc) return rc;
        cx.slot[s].pend_dst = nullptr;
    }
    {
        int rc = pinned_reserve(cx, (size_t)N * 8);
        if (rc) return rc;
    }
    // results of a slot's earlier chunk: out of the bounce buffer into the caller's array
    auto flush_out = [&](Slot& sl) -> int {
        if (!sl.pend_dst) return 0;
        CK(cudaStreamSynchronize(sl.st));
        par_copy_2d((char*)sl.pend_dst, (size_t)sl.pend_dpitch, (const char*)sl.hout, (size_t)sl.pend_width,
                    (size_t)sl.pend_width, sl.pend_rows);
        sl.pend_dst = nullptr;
        return 0;
    };
    double* pk = reinterpret_cast<double*>(cx.pinned);
    int ci = 0;
    for (long long i0 = 0; i0 < N; i0 += chunk, ++ci) {
        const long long nc = std::min<long long>(chunk, N - i0);
        Slot& sl = cx.slot[ci % NSLOT];
        Carve cv(sl.buf);
        double* d_in = cv.take<double>((size_t)chunk * S);
        double* d_out = cv.take<double>((size_t)chunk * S);
        double* d_k = cv.take<double>((size_t)chunk);
        void* d_ws = cv.take<char>(wsb);
        long long dss, dsn, oss, osn;
        if (bounce_out) {
            int rc = flush_out(sl);
            if (rc) return rc;
        }
        if (bounce_in) {
            CK(cudaStreamSynchronize(sl.st));  // the slot's previous chunk has left its bounce buffer
            if (rows_in) par_copy_2d((char*)sl.hin, (size_t)S * 8, (const char*)(lw + i0 * stride_n), (size_t)stride_n * 8,
                                     (size_t)S * 8, nc);
            else par_copy_2d((char*)sl.hin, (size_t)nc * 8, (const char*)(lw + i0), (size_t)stride_s * 8, (size_t)nc * 8, S);
            CK(cudaMemcpyAsync(d_in, sl.hin, (size_t)nc * S * 8, cudaMemcpyHostToDevice, sl.st));
            if (rows_in) { dss = 1; dsn = S; } else { dss = nc; dsn = 1; }
        } else if (rows_in) {  // rows [i0, i0+nc) -> dense nc x S
            if (stride_n == S)  // dense on the host too: one linear DMA instead of one descriptor per row
                CK(cudaMemcpyAsync(d_in, lw + i0 * stride_n, (size_t)nc * S * 8, cudaMemcpyHostToDevice, sl.st));
            else
                CK(cudaMemcpy2DAsync(d_in, (size_t)S * 8, lw + i0 * stride_n, (size_t)stride_n * 8,
                                     (size_t)S * 8, (size_t)nc, cudaMemcpyHostToDevice, sl.st));
            dss = 1; dsn = S;
        } else {        // columns [i0, i0+nc) of the S x N matrix -> dense S x nc
            CK(cudaMemcpy2DAsync(d_in, (size_t)nc * 8, lw + i0, (size_t)stride_s * 8, (size_t)nc * 8,
                                 (size_t)S, cudaMemcpyHostToDevice, sl.st));
            dss = nc; dsn = 1;
        }
        if (rows_out) { oss = 1; osn = S; } else { oss = nc; osn = 1; }
        int rc = b2l_psislw_dev_f64(d_in, S, nc, dss, dsn, M, cutoffmin, d_out, oss, osn, d_k, nullptr,
                                    d_ws, wsb, sl.st);
        if (rc) return rc;
        if (bounce_out) {
            CK(cudaMemcpyAsync(sl.hout, d_out, (size_t)nc * S * 8, cudaMemcpyDeviceToHost, sl.st));
            if (rows_out) { sl.pend_dst = lw_out + i0 * ostride_n; sl.pend_rows = nc; sl.pend_width = S * 8; sl.pend_dpitch = ostride_n * 8; }
            else { sl.pend_dst = lw_out + i0; sl.pend_rows = S; sl.pend_width = nc * 8; sl.pend_dpitch = ostride_s * 8; }
        } else if (rows_out && ostride_n == S)
            CK(cudaMemcpyAsync(lw_out + i0 * ostride_n, d_out, (size_t)nc * S * 8, cudaMemcpyDeviceToHost, sl.st));
        else if (rows_out)
            CK(cudaMemcpy2DAsync(lw_out + i0 * ostride_n, (size_t)ostride_n * 8, d_out, (size_t)S * 8,
                                 (size_t)S * 8, (size_t)nc, cudaMemcpyDeviceToHost, sl.st));
        else
            CK(cudaMemcpy2DAsync(lw_out + i0, (size_t)ostride_s * 8, d_out, (size_t)nc * 8,
                                 (size_t)nc * 8, (size_t)S, cudaMemcpyDeviceToHost, sl.st));
        CK(cudaMemcpyAsync(pk + i0, d_k, (size_t)nc * 8, cudaMemcpyDeviceToHost, sl.st));
    }
    for (int s = 0; s < NSLOT; ++s) {
        if (bounce_out) {
            int rc = flush_out(cx.slot[s]);
            if (rc) return rc;
        }
        CK(cudaStreamSynchronize(cx.slot[s].st));
    }
    memcpy(k_out, pk, (size_t)N * 8);
    return 0;
}

extern "C" int b2l_loo_host_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s,
                                int64_t stride_n, int32_t M, double cutoffmin, uint32_t flags,
                                double good_k, double* elpd_i, double* k_i, double* lppd_i, double* var_i,
                                double* lppdw_i, double* stats_out, int32_t device, int64_t chunk_obs) {
    if (!ll || !elpd_i || !k_i || !lppd_i || !var_i || !lppdw_i || N < 0 || S < 1)
        return fail(B2L_E_INVALID, "bad arguments");
    if (b2l_device_count() < 1) return B2L_E_NODEVICE;
    const bool rows_in = (stride_s == 1 || S == 1);
    if (!rows_in && !(stride_n == 1 || N == 1)) return fail(B2L_E_INVALID, "input must have a unit stride");
    if (device < 0 || device >= 64) return fail(B2L_E_INVALID, "device index %d out of range", device);
    DevCtx& cx = g_ctx[device];
    std::lock_guard<std::mutex> lock(cx.mu);
    DeviceGuard restore;
    CK(cudaSetDevice(device));
    SlotDrain drain(cx);
    PipeCount pipe;
    const long long chunk = chunk_obs > 0 ? std::min<long long>(chunk_obs, std::max<long long>(N, 1))
                                          : default_chunk(S, N);
    size_t wsb = 0;
    b2l_workspace_bytes(S, chunk, M, rows_in ? 0 : 1, &wsb);
    // device pitch of a chunk of the (S, N) layout: even, so that an odd last chunk still meets the 16-byte pitch
    // rule of the 2-D TMA tiles (the tile path) -- the pad column is never read
    const size_t mat = align_up((size_t)(chunk + 1) * (size_t)S * 8, 256);
    const size_t vec = align_up((size_t)chunk * 8, 256);
    const size_t need = mat + 5 * vec + 512 + wsb + 1024;
    const bool bounce = N > 0 && host_pageable(ll);
    for (int s = 0; s < NSLOT; ++s) {
        int rc = slot_reserve(cx.slot[s], need);
        if (rc) return rc;
        if (bounce) {
            rc = hpin_reserve(&cx.slot[s].hin, &cx.slot[s].hin_bytes, mat);
            if (rc) return rc;
        }
    }
    const long long nchunks = N > 0 ? (N + chunk - 1) / chunk : 0;
    std::vector<double> recs((size_t)std::max<long long>(nchunks, 1) * B2L_STATS_LEN, 0.0);
    {
        int rc = pinned_reserve(cx, (size_t)(5 * std::max<long long>(N, 1) + std::max<long long>(nchunks, 1) * B2L_STATS_LEN) * 8);
        if (rc) return rc;
    }
    double* pv = reinterpret_cast<double*>(cx.pinned);                 // [5][N]
    double* prec = pv + 5 * std::max<long long>(N, 1);                 // [nchunks][B2L_STATS_LEN]
    int ci = 0;
    for (long long i0 = 0; i0 < N; i0 += chunk, ++ci) {
        const long long nc = std::min<long long>(chunk, N - i0);
        Slot& sl = cx.slot[ci % NSLOT];
        Carve cv(sl.buf);
        double* d_in = cv.take<double>((size_t)(chunk + 1) * S);
        const long long pitch = rows_in ? nc : ((nc > 1) ? ((nc + 1) & ~1ll) : nc);
        double* d_e = cv.take<double>((size_t)chunk);
        double* d_k = cv.take<double>((size_t)chunk);
        double* d_l = cv.take<double>((size_t)chunk);
        double* d_v = cv.take<double>((size_t)chunk);
        double* d_lw = cv.take<double>((size_t)chunk);
        unsigned long long* d_cnt = cv.take<unsigned long long>(4);
        double* d_stats = cv.take<double>(B2L_STATS_LEN);
        void* d_ws = cv.take<char>(wsb);
        long long dss, dsn;
        if (ci >= NSLOT && !rows_in && !(flags & (B2L_FLAG_NO_TILE | B2L_FLAG_WAIC_ONLY))) {
            // the slot's previous chunk (three chunks back): if the cluster kernel handed more than a fifth of it to
            // the general kernel (heavy-tailed columns), the rest of the matrix takes the transposed-panel route
            CK(cudaStreamSynchronize(sl.st));
            const double* rec = prec + (size_t)(ci - NSLOT) * B2L_STATS_LEN;
            if (rec[B2L_ST_N_FALLBACK] > 0.2 * rec[B2L_ST_N]) flags |= B2L_FLAG_NO_TILE;
        }
        if (bounce) {
            // the slot's previous chunk (three chunks back) has left its bounce buffer once its stream is idle
            CK(cudaStreamSynchronize(sl.st));
            if (rows_in) par_copy_2d((char*)sl.hin, (size_t)S * 8, (const char*)(ll + i0 * stride_n), (size_t)stride_n * 8,
                                     (size_t)S * 8, nc);
            else par_copy_2d((char*)sl.hin, (size_t)pitch * 8, (const char*)(ll + i0), (size_t)stride_s * 8, (size_t)nc * 8, S);
            CK(cudaMemcpyAsync(d_in, sl.hin, (size_t)(rows_in ? nc : pitch) * S * 8, cudaMemcpyHostToDevice, sl.st));
            if (rows_in) { dss = 1; dsn = S; } else { dss = pitch; dsn = 1; }
        } else if (rows_in) {
            CK(cudaMemcpy2DAsync(d_in, (size_t)S * 8, ll + i0 * stride_n, (size_t)stride_n * 8,
                                 (size_t)S * 8, (size_t)nc, cudaMemcpyHostToDevice, sl.st));
            dss = 1; dsn = S;
        } else {
            CK(cudaMemcpy2DAsync(d_in, (size_t)pitch * 8, ll + i0, (size_t)stride_s * 8, (size_t)nc * 8,
                                 (size_t)S, cudaMemcpyHostToDevice, sl.st));
            dss = pitch; dsn = 1;
        }
        CK(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned long long), sl.st));
        int rc = b2l_loo_dev_f64(d_in, S, nc, dss, dsn, M, cutoffmin, flags, d_e, d_k, d_l, d_v, d_lw, d_cnt,
                                 nullptr, d_ws, wsb, sl.st);
        if (rc) return rc;
        rc = b2l_stats_dev_f64(d_e, d_k, d_l, d_v, d_lw, nc, good_k, d_cnt, d_stats, d_ws, wsb, sl.st);
        if (rc) return rc;
        CK(cudaMemcpyAsync(pv + 0 * N + i0, d_e, (size_t)nc * 8, cudaMemcpyDeviceToHost, sl.st));
        CK(cudaMemcpyAsync(pv + 1 * N + i0, d_k, (size_t)nc * 8, cudaMemcpyDeviceToHost, sl.st));
        CK(cudaMemcpyAsync(pv + 2 * N + i0, d_l, (size_t)nc * 8, cudaMemcpyDeviceToHost, sl.st));
        CK(cudaMemcpyAsync(pv + 3 * N + i0, d_v, (size_t)nc * 8, cudaMemcpyDeviceToHost, sl.st));
        CK(cudaMemcpyAsync(pv + 4 * N + i0, d_lw, (size_t)nc * 8, cudaMemcpyDeviceToHost, sl.st));
        CK(cudaMemcpyAsync(prec + (size_t)ci * B2L_STATS_LEN, d_stats, B2L_STATS_LEN * 8, cudaMemcpyDeviceToHost,
                           sl.st));
    }
    for (int s = 0; s < NSLOT; ++s) CK(cudaStreamSynchronize(cx.slot[s].st));
    if (N > 0) {
        memcpy(elpd_i, pv + 0 * N, (size_t)N * 8);
        memcpy(k_i, pv + 1 * N, (size_t)N * 8);
        memcpy(lppd_i, pv + 2 * N, (size_t)N * 8);
        memcpy(var_i, pv + 3 * N, (size_t)N * 8);
        memcpy(lppdw_i, pv + 4 * N, (size_t)N * 8);
        memcpy(recs.data(), prec, (size_t)nchunks * B2L_STATS_LEN * 8);
    }
    if (stats_out) {
        if (nchunks == 0) {
            double z[B2L_STATS_LEN] = {0};
            memcpy(stats_out, z, sizeof(z));
        } else {
            int rc = b2l_stats_merge(recs.data(), (int)nchunks, stats_out);
            if (rc) return rc;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------ several GPUs, one process
// Observations are independent (pyloo/utils.py:171-176): contiguous shards of the observation axis, one host
// thread and one b2l_*_host_f64 pipeline per device, no data-path exchange.  Shard records are merged in
// shard order with Chan's update (b2l_stats_merge), so the totals do not depend on the device count.
namespace {
// pins the caller's (pageable) buffer for the duration of a call when B2L_HOST_REGISTER=1: the chunk copies
// then run as true asynchronous DMA instead of through the driver's staging buffers
struct HostPin {
    void* p = nullptr;
    HostPin(const void* base, size_t bytes) {
        const char* ev = getenv("B2L_HOST_REGISTER");
        if (!ev || atoi(ev) == 0 || !base || bytes == 0) return;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, base) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return;
        cudaGetLastError();
        if (cudaHostRegister(const_cast<void*>(base), bytes, cudaHostRegisterPortable) == cudaSuccess) p = const_cast<void*>(base);
        else cudaGetLastError();
    }
    ~HostPin() { if (p) cudaHostUnregister(p); }
};
size_t span_bytes(int64_t S, int64_t N, int64_t stride_s, int64_t stride_n) {
    if (S < 1 || N < 1) return 0;
    return (size_t)((S - 1) * stride_s + (N - 1) * stride_n + 1) * 8;
}
void shard_bounds(int64_t N, int n, std::vector<int64_t>& b) {
    b.assign(n + 1, 0);
    const int64_t per = ((N + n - 1) / n + 15) / 16 * 16;  // whole tiles: shards stay 16 B aligned and tile aligned
    for (int d = 1; d <= n; ++d) b[d] = std::min<int64_t>(N, b[d - 1] + per);
}
}  // namespace

extern "C" int b2l_loo_host_mgpu_f64(const double* ll, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                                     int32_t M, double cutoffmin, uint32_t flags, double good_k, double* elpd_i,
                                     double* k_i, double* lppd_i, double* var_i, double* lppdw_i, double* stats_out,
                                     const int32_t* devices, int32_t n_devices, int64_t chunk_obs) {
    if (!devices || n_devices < 1 || n_devices > 64) return fail(B2L_E_INVALID, "bad device list");
    if (!ll || !elpd_i || !k_i || !lppd_i || !var_i || !lppdw_i || N < 0 || S < 1)
        return fail(B2L_E_INVALID, "bad arguments");
    HostPin pin(ll, span_bytes(S, N, stride_s, stride_n));
    if (n_devices == 1 || N < 32)
        return b2l_loo_host_f64(ll, S, N, stride_s, stride_n, M, cutoffmin, flags, good_k, elpd_i, k_i, lppd_i, var_i,
                                lppdw_i, stats_out, devices[0], chunk_obs);
    std::vector<int64_t> b;
    shard_bounds(N, n_devices, b);
    std::vector<double> recs((size_t)n_devices * B2L_STATS_LEN, 0.0);
    std::vector<int> rcs(n_devices, 0);
    std::vector<std::string> msgs(n_devices);
    std::vector<std::thread> th;
    int used = 0;
    for (int d = 0; d < n_devices; ++d) {
        const int64_t i0 = b[d], n = b[d + 1] - b[d];
        if (n <= 0) continue;
        ++used;
        th.emplace_back([&, d, i0, n]() {
            rcs[d] = b2l_loo_host_f64(ll + i0 * stride_n, S, n, stride_s, stride_n, M, cutoffmin, flags, good_k,
                                      elpd_i + i0, k_i + i0, lppd_i + i0, var_i + i0, lppdw_i + i0,
                                      recs.data() + (size_t)d * B2L_STATS_LEN, devices[d], chunk_obs);
            if (rcs[d]) msgs[d] = b2l_last_error();
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < n_devices; ++d)
        if (rcs[d]) return fail(rcs[d], "device %d: %s", devices[d], msgs[d].c_str());
    if (stats_out) {
        std::vector<double> packed;
        for (int d = 0; d < n_devices; ++d)
            if (b[d + 1] > b[d]) packed.insert(packed.end(), recs.begin() + (size_t)d * B2L_STATS_LEN,
                                               recs.begin() + (size_t)(d + 1) * B2L_STATS_LEN);
        return b2l_stats_merge(packed.data(), used, stats_out);
    }
    return 0;
}

extern "C" int b2l_psislw_host_mgpu_f64(const double* lw, int64_t S, int64_t N, int64_t stride_s, int64_t stride_n,
                                        int32_t M, double cutoffmin, double* lw_out, int64_t ostride_s,
                                        int64_t ostride_n, double* k_out, const int32_t* devices, int32_t n_devices,
                                        int64_t chunk_obs) {
    if (!devices || n_devices < 1 || n_devices > 64) return fail(B2L_E_INVALID, "bad device list");
    if (!lw || !lw_out || !k_out || N < 0 || S < 1) return fail(B2L_E_INVALID, "bad arguments");
    HostPin pin_in(lw, span_bytes(S, N, stride_s, stride_n));
    HostPin pin_out(lw_out, span_bytes(S, N, ostride_s, ostride_n));
    if (n_devices == 1 || N < 32)
        return b2l_psislw_host_f64(lw, S, N, stride_s, stride_n, M, cutoffmin, lw_out, ostride_s, ostride_n, k_out,
                                   devices[0], chunk_obs);
    std::vector<int64_t> b;
    shard_bounds(N, n_devices, b);
    std::vector<int> rcs(n_devices, 0);
    std::vector<std::string> msgs(n_devices);
    std::vector<std::thread> th;
    for (int d = 0; d < n_devices; ++d) {
        const int64_t i0 = b[d], n = b[d + 1] - b[d];
        if (n <= 0) continue;
        th.emplace_back([&, d, i0, n]() {
            rcs[d] = b2l_psislw_host_f64(lw + i0 * stride_n, S, n, stride_s, stride_n, M, cutoffmin,
                                         lw_out + i0 * ostride_n, ostride_s, ostride_n, k_out + i0, devices[d],
                                         chunk_obs);
            if (rcs[d]) msgs[d] = b2l_last_error();
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < n_devices; ++d)
        if (rcs[d]) return fail(rcs[d], "device %d: %s", devices[d], msgs[d].c_str());
    return 0;
}
