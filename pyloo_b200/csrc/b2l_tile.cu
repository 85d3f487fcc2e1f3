// Tile path (sm_100a): the stream pass of pl.loo on the ArviZ (chain, draw, obs) layout, reading the
// observation-fastest (S, N) log-likelihood matrix where it lies -- no transposed copy, ONE read of HBM.
//
//   loo_tile_kernel   a thread-block CLUSTER owns a tile of TILE_W = 16 observations x all S draws (128 B of
//                     every draw: whole DRAM bursts).  CTA r of the cluster stages draws [r R, (r + 1) R) of the
//                     tile in its shared memory with 2-D TMA boxes {16 observations x box_rows draws}
//                     (cp.async.bulk.tensor.2d with the 128-byte swizzle, SASS UTMALDG), so the cluster as a whole
//                     holds the tile on chip and every pass after the load reads shared memory.
//                     WARP TEAMS: warp w of every CTA owns columns 2w and 2w + 1 of the tile (16 draw slots x 2
//                     columns per warp; the swizzle makes that column-wise access bank-conflict free).  All
//                     reductions of a column are warp shuffles, and the eight warps w of the cluster exchange
//                     their column results directly -- st.async into the peers' shared memory, completing a
//                     per-warp mbarrier there -- so there is no CTA-wide or cluster-wide barrier inside the loop:
//                       pass A   per-thread minima of ll: the CTA's column minimum and its 32 bin minima, whose
//                                sorted ranks q_t / q_l are the CTA's tight / loose threshold statistic
//                       send     (minimum, two statistics) to every CTA; with it the PREVIOUS tile's partial
//                                sums to the column's owner CTA
//                       pass B   sums about the CTA's own minimum c (no peer data needed, hides the exchange):
//                                one range reduction gives exp(-(ll - c)) for the PSIS normaliser
//                                (pyloo/utils.py:349-351) and exp(ll - c) for lppd_i (pyloo/loo.py:329-337);
//                                sums of (ll - c), (ll - c)^2 for var_s(ll) (pyloo/waic.py:145)
//                       receive  column minimum over the CTAs (r = -ll, max r = -min ll: pyloo/loo.py:286-288,
//                                pyloo/psis.py:134), lower / upper median of their thresholds; the owner rescales
//                                and adds the previous tile's partial sums and writes its SplitHeader
//                       pass C   draws at or below the loose threshold are the candidates: exact
//                                x = fl(min ll - ll) = fl(r - max r) (psis.py:134) and draw index to the round's
//                                scratch, tight ones from the front, looser ones from the end (positions from
//                                one atomic per warp, column and list)
//   psis_tail_kernel  (b2l_split.cuh) sort, cutoff, GPD fit, smoothing, elpd_i: one warp per observation.
//
// Observations the fast path cannot decide (NaN / inf, ll range > 600, candidate count outside
// [M + 1, cap]) are appended to the hand-over list; the general row kernel re-does them with strided reads.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "b2l_tile_host.h"

namespace b2l {

// ---------------------------------------------------------------- shared-memory carve-up
struct XMsg {  // one per (source CTA, column) and tile
    double mn;     // minimum of ll over the source CTA's draws
    float st, sl;  // its tight / loose threshold statistic
};
struct RSum {  // one per (source CTA, owned column) and tile
    double q[4];  // about the SOURCE CTA's minimum c: sum exp(-(ll - c)), sum exp(ll - c), sum (ll - c), sum (ll - c)^2
    double c;
    int umax;     // max over the high words of ll - c (>= 0: integer order = value order)
    int pad;
};
static_assert(sizeof(XMsg) == 16 && sizeof(RSum) == 48, "exchange records");
struct TileSmemLayout {
    size_t off_tab, off_xch, off_rsum, off_bar, off_tile, total;
};
// The small structures first, at offsets that depend on the template parameter only (compile-time constants inside
// the kernel); the tile follows on the next 1024-byte boundary of the shared window (its swizzle atom).
__host__ __device__ inline TileSmemLayout tile_smem(int rows_load, int tw) {
    TileSmemLayout L;
    size_t o = 0;
    L.off_tab = o;   // 64 x (2^(j/64), 2^(-j/64))
    o += 64 * 16;
    L.off_xch = o;   // [parity][source CTA][column], written by the peers
    o += 2 * TILE_MAXC * tw * sizeof(XMsg);
    L.off_rsum = o;  // [parity][source CTA][owned slot], written by the peers (tw entries per parity)
    o += 2 * tw * sizeof(RSum);
    L.off_bar = o;   // tile full, tile empty, per warp and parity: exchange complete
    o += (2 + tw) * 8;
    L.off_tile = o;  // (+ up to 1023 bytes of alignment slack)
    L.total = align_up(o + 1023 + (size_t)rows_load * tw * 8, 128);
    return L;
}

// ---------------------------------------------------------------- cluster / DSMEM / TMA helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
// 16 bytes into a peer's shared memory; completes 16 bytes of the transaction count of the peer's mbarrier
__device__ __forceinline__ void st_async_16(uint32_t dst, uint64_t a, uint64_t b, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(dst),
                 "l"(a), "l"(b), "r"(bar)
                 : "memory");
}
// wait for a phase completed by the peers' st.async stores: acquire at cluster scope, so that what the other
// CTAs wrote into this CTA's shared memory is visible to the loads that follow
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t f64_bits(double v) { return (uint64_t)__double_as_longlong(v); }

// exp(u) and exp(-u) from one range reduction, 0 <= u <= 600: u = n ln2/64 + r, |r| <= ln2/128,
// e^(+-u) = 2^(+-(n >> 6)) 2^(+-(n & 63)/64) (cosh r +- sinh r), cosh r ~ 1 + r^2/2 + r^4/24, sinh r ~ r (1 + r^2/6):
// relative error < 4e-14 (the dropped r^5/120 term).  These feed the normalising sums only; the exact tail values
// are formed by the tail kernel from the exact x.
constexpr double TE_L = 92.332482616893657;    // 64 / ln 2
constexpr double TE_C = 0.010830424696249145;  // ln 2 / 64
__device__ __forceinline__ void exp_pm64(double u, const double2* tab, double& ep, double& em) {
    const double t = fma(u, TE_L, EXP_MAGIC);
    const int ni = __double2loint(t);
    const double nf = t - EXP_MAGIC;
    const double r = fma(nf, -TE_C, u);
    const double r2 = r * r;
    const double c = fma(r2, fma(r2, 1.0 / 24.0, 0.5), 1.0);
    const double s1 = fma(r2, 1.0 / 6.0, 1.0);
    const double2 T = tab[ni & 63];
    const int sh = (ni << 14) & 0xfff00000;
    const double yp = T.x * fma(r, s1, c), ym = T.y * fma(-r, s1, c);
    ep = __hiloint2double(__double2hiint(yp) + sh, __double2loint(yp));
    em = __hiloint2double(__double2hiint(ym) - sh, __double2loint(ym));
}

// Ascending bitonic sort of the 32 values a half-warp holds as (a, b) per lane: element e = 16 reg + slot.  Both
// half-warps (= both columns of the warp) sort at the same time; partners at distance 16 sit in the same lane.
__device__ __forceinline__ void half_sort32_f(float& a, float& b, int slot) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 16) {  // (only k = 32: ascending everywhere)
                const float lo = fminf(a, b), hi = fmaxf(a, b);
                a = lo;
                b = hi;
            } else {
                const float oa = __shfl_xor_sync(FULL, a, j), ob = __shfl_xor_sync(FULL, b, j);
                const bool lower = (slot & j) == 0;
                // direction of the k-block: bit k of the element index (slot for k < 32... reg for k = 16 -> e & 16)
                const bool up_a = (k == 32) ? true : (k == 16 ? true : ((slot & k) == 0));
                const bool up_b = (k == 32) ? true : (k == 16 ? false : ((slot & k) == 0));
                a = (up_a == lower) ? fminf(a, oa) : fmaxf(a, oa);
                b = (up_b == lower) ? fminf(b, ob) : fmaxf(b, ob);
            }
        }
    }
}

// ---------------------------------------------------------------- the kernel
// TW: observations per tile: 16 (128 B of a draw, 8 warps, 3 CTAs / SM) or 8 (64 B, 4 warps, 6 CTAs / SM)
// CHUNKED: work units are (tile, chunk) pairs (a separate build: the plain one carries none of its bookkeeping)
template <int TW, bool CHUNKED>
__global__ void __launch_bounds__(16 * TW, 48 / TW) loo_tile_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                    const TileParams p) {
    constexpr int NW = TW / 2;       // warps: two columns each
    constexpr int KSTEP = 16 * TW;   // doubles between a thread's consecutive draws (16 rows of TW observations)
    extern __shared__ unsigned char smem_dyn[];
    const TileSmemLayout L = tile_smem(0, TW);  // (the offsets do not depend on the tile's size)
    unsigned char* aux = smem_dyn;
    // the swizzled tile starts on a 1024-byte boundary of the shared window
    unsigned char* smem_raw = smem_dyn + L.off_tile + ((1024u - ((smem_u32(smem_dyn) + (unsigned)L.off_tile) & 1023u)) & 1023u);
    double* tile = reinterpret_cast<double*>(smem_raw);
    double2* etab = reinterpret_cast<double2*>(aux + L.off_tab);
    XMsg* xch = reinterpret_cast<XMsg*>(aux + L.off_xch);       // [2][TILE_MAXC][TW]
    RSum* rsum = reinterpret_cast<RSum*>(aux + L.off_rsum);     // [2][TW]: index src * nslot + slot
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux + L.off_bar);
    uint64_t* bar_full = &bars[0];
    uint64_t* bar_empty = &bars[1];
    uint64_t* bar_x = &bars[2];  // [NW][2]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int slot = lane & 15, hf = lane >> 4, col = 2 * w + hf;  // draw slot 0..15, column of the tile
    const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
    const int nslot = TW / csize;            // columns a CTA owns: col = oslot * csize + rank
    const int orank = col % csize, oslot = col / csize;
    const bool own = orank == crank;
    const long long cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;
    const int S = p.S, M = p.M, cap = p.cap, R = p.R;
    const int row0 = crank * R;
    const int rows_live = max(0, min(R, S - row0));
    const int kmin = rows_live >> 4;   // draws every slot of this CTA has
    const int k4 = kmin & ~3;          // ... in whole groups of four; the rest runs guarded
    const uint32_t tile_tx = (uint32_t)p.nbox * (uint32_t)p.box_rows * TW * 8u;
    const double INF = inf_f64();
    const float FINF = __int_as_float(0x7f800000);

    if (tid == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_empty, NW);
        for (int i = 0; i < 2 * NW; ++i) mbar_init(&bar_x[i], 1);
        fence_mbar_init();
    }
    if (tid < 64) etab[tid] = make_double2(exp2((double)tid / 64.0), exp2(-(double)tid / 64.0));
    __syncthreads();
    cluster_sync_all();  // every CTA of the cluster runs and has its barriers initialised

    // work unit u = tile * n_chunks + chunk (n_chunks = 1 unless the draw axis is longer than one cluster holds):
    // S is then the chunk length and the unit's draws start at row chunk * S of the matrix
    const int nch = CHUNKED ? p.n_chunks : 1;
    const long long n_units = p.n_tiles * nch;
    auto issue = [&](long long u) {  // one thread: the CTA's draws of unit u, nbox boxes on one mbarrier
        const unsigned tl_ = (unsigned)u / (unsigned)nch;  // (a round has far fewer than 2^31 units)
        const int ch = (int)((unsigned)u - tl_ * (unsigned)nch);
        mbar_expect_tx(bar_full, tile_tx);
        for (int b = 0; b < p.nbox; ++b)
            tma_load_2d(tile + (size_t)b * p.box_rows * TW, &tmap, (int)(p.col0 + (long long)tl_ * TW),
                        ch * S + row0 + b * p.box_rows, bar_full);
    };
    // this thread's draws: row 16 k + slot; the warp's 16-byte chunk w of the row sits at chunk w ^ (row & 7)
    // (128-byte rows, 128-byte swizzle) or w ^ ((row >> 1) & 3) (64-byte rows, 64-byte swizzle); half hf of the chunk
    const int swz = (TW == 16) ? (w ^ (slot & 7)) : (w ^ ((slot >> 1) & 3));
    const double* pcol = tile + (slot * (TW * 8) + (swz << 4) + hf * 8) / 8;  // draw k at pcol[k * KSTEP]
    // guarded tail of every pass: draws k4 .. k4 + 3
    bool tail_ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tail_ok[j] = ((k4 + j) << 4) + slot < rows_live;

    // results of the previous tile that leave one tile later
    double pq0 = 0.0, pq1 = 0.0, pq2 = 0.0, pq3 = 0.0, pc = 0.0, p_llmin = 0.0, p_tl = 0.0;
    int pum = 0;
    long long t_prev = -1;

    // the owner's share of the exchange: rescale and add the CTAs' partial sums, write the header
    auto write_header = [&](long long tt, int par, double llmin_, double tl_v) {
        const int r = slot & 7;
        const bool have = own && r < csize;
        double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, um = 0.0;
        if (have) {
            const RSum e = rsum[par * TW + r * nslot + oslot];
            const double n_r = (double)max(0, min(R, S - r * R));
            if (n_r > 0.0) {
                const double d = e.c - llmin_;  // >= 0: the source CTA's sums are about its own minimum
                q0 = e.q[0] * exp(-d);
                q1 = e.q[1] * exp(d);
                q2 = fma(n_r, d, e.q[2]);
                q3 = fma(d, fma(n_r, d, 2.0 * e.q[2]), e.q[3]);
                um = __hiloint2double(e.umax, 0) * 1.000002 + d;  // upper bound of max (ll - min ll)
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {  // over the source CTAs, fixed order
            q0 += __shfl_xor_sync(FULL, q0, o);
            q1 += __shfl_xor_sync(FULL, q1, o);
            q2 += __shfl_xor_sync(FULL, q2, o);
            q3 += __shfl_xor_sync(FULL, q3, o);
            um = fmax(um, __shfl_xor_sync(FULL, um, o));
        }
        const unsigned tl_ = (unsigned)tt / (unsigned)nch;
        const long long oo = (long long)tl_ * TW + col;
        if (nch > 1) {  // chunked: the raw record; tile_merge_kernel folds a column's chunks and decides
            if (own && slot == 0 && oo < p.n_obs) {
                ChunkHeader c;
                c.llmin = llmin_; c.q0 = q0; c.q1 = q1; c.q2 = q2; c.q3 = q3; c.um = um; c.tl = tl_v; c.pad_ = 0.0;
                p.chdr[oo * nch + ((unsigned)tt - tl_ * (unsigned)nch)] = c;
            }
            return;
        }
        if (own && slot == 0 && oo < p.n_obs) {
            const int ca = (int)atomicAdd(&p.cnt[2 * oo], 0u), cb = (int)atomicAdd(&p.cnt[2 * oo + 1], 0u);
            const bool special = !(is_finite(q2) && is_finite(q3) && is_finite(q0) && is_finite(q1));
            const bool wide = !(um <= 600.0);
            const bool count_bad = (ca + cb < M + 1) || (ca + cb > cap);
            const bool ok = !special && !wide && !count_bad;
            SplitHeader h;
            h.mx = -llmin_;
            h.body = q0;
            h.lsum = q1;
            h.vsum = q3 - q2 * q2 / (double)S;  // sum (ll - mean)^2 taken about the minimum
            if (h.vsum < 0.0) h.vsum = 0.0;
            h.lshift = llmin_;
            h.taux = llmin_ - tl_v;  // every candidate has ll <= t_l, i.e. x = fl(min ll - ll) >= taux
            h.lse = 0.0;
            h.C = ca; h.flags = ok ? 0 : 1; h.attempts = 0; h.n_patch = 0; h.C2 = cb; h.pad_ = 0;
            p.hdr[oo] = h;
            if (!ok) {
                p.fb_list[atomicAdd(p.fb_count, 1)] = (int)(p.row_base + oo);
                if (p.counters) atomicAdd(&p.counters[3], 1ull);
                note_handover(special ? HO_SPECIAL : (wide ? HO_RANGE : HO_RETRY));
            }
        }
    };
    // bytes this warp receives per exchange: 16 per (CTA, column) + the partial sums of the columns it owns here
    const int own_cols = ((2 * w) % csize == crank ? 1 : 0) + ((2 * w + 1) % csize == crank ? 1 : 0);
    auto send_partials = [&](int par) {  // lanes slot 0 of each half: the previous tile's sums to the column's owner
        if (slot == 0) {
            const uint32_t dst = dsmem_addr(&rsum[par * TW + crank * nslot + oslot], (uint32_t)orank);
            const uint32_t bar = dsmem_addr(&bar_x[w * 2 + par], (uint32_t)orank);
            st_async_16(dst, f64_bits(pq0), f64_bits(pq1), bar);
            st_async_16(dst + 16, f64_bits(pq2), f64_bits(pq3), bar);
            st_async_16(dst + 32, f64_bits(pc), (uint64_t)(uint32_t)pum, bar);
        }
    };

    long long t = cluster_id;  // (a work unit: a tile, or a (tile, chunk) pair)
    if (tid == 0 && t < n_units) issue(t);
    int it = 0;
    for (; t < n_units; t += n_clusters, ++it) {
        const int par = it & 1;
        const uint32_t xph = (uint32_t)((it >> 1) & 1);
        if (lane == 0 && !(p.debug & 2))
            mbar_expect_tx(&bar_x[w * 2 + par], (uint32_t)(csize * 2 * 16 + (t_prev >= 0 ? own_cols * csize * 48 : 0)));
        mbar_wait(bar_full, (uint32_t)(it & 1));
        if (p.debug & 2) {  // measurement aid: the loads alone
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty);
            if (tid == 0 && t + n_clusters < n_units) {
                mbar_wait(bar_empty, (uint32_t)(it & 1));
                fence_proxy_async();
                issue(t + n_clusters);
            }
            __syncwarp();
            continue;
        }

        // ---------------- pass A: minima.  Two alternating bins per thread (32 bins per CTA and column).
        double mA = INF, mB = INF;
        {
            int k = 0;
            for (; k + 8 <= k4; k += 8) {  // eight draws in flight
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = pcol[(k + j) * KSTEP];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    mA = min_sel(mA, v[j]);
                    mB = min_sel(mB, v[j + 1]);
                }
            }
            if (k < k4) {
                double v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = pcol[(k + j) * KSTEP];
                mA = min_sel(mA, v[0]);
                mB = min_sel(mB, v[1]);
                mA = min_sel(mA, v[2]);
                mB = min_sel(mB, v[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double v = tail_ok[j] ? pcol[(k4 + j) * KSTEP] : INF;
                if (j & 1) mB = min_sel(mB, v);
                else mA = min_sel(mA, v);
            }
        }
        double cl = min_sel(mA, mB);  // -> this CTA's column minimum
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) cl = min_sel(cl, __shfl_xor_sync(FULL, cl, o));
        {
            float a = (float)mA, b = (float)mB;
            half_sort32_f(a, b, slot);
            const int et = p.q_t - 1, el = p.q_l - 1;  // element e = 16 reg + slot of the sorted order
            const float st = __shfl_sync(FULL, (et & 16) ? b : a, (lane & 16) + (et & 15));
            const float sl = __shfl_sync(FULL, (el & 16) ? b : a, (lane & 16) + (el & 15));
            if (slot < csize) {  // -> CTA `slot`
                const uint64_t pk = (uint64_t)__float_as_uint(st) | ((uint64_t)__float_as_uint(sl) << 32);
                st_async_16(dsmem_addr(&xch[(par * TILE_MAXC + crank) * TW + col], (uint32_t)slot), f64_bits(cl), pk,
                            dsmem_addr(&bar_x[w * 2 + par], (uint32_t)slot));
            }
        }
        if (t_prev >= 0) send_partials(par);

        // ---------------- pass B: sums over all draws about this CTA's column minimum (no peer data needed)
        double bs = 0.0, ls = 0.0, su = 0.0, suu = 0.0;
        int umax = 0;
        {
            // Four draws per step in three stages, so that the table look-ups (whose address comes out of the
            // range reduction) and the next step's draws are in flight while the polynomials run:
            //   1. u = ll - c, range reduction, table look-up issued   2. cosh / sinh polynomials   3. scale, add
            auto step = [&](const double (&v)[4], const bool (&ok)[4]) {
                double u[4], r[4];
                int ni[4];
                double2 T[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u[j] = v[j] - cl;  // >= 0
                    const double t_ = fma(u[j], TE_L, EXP_MAGIC);
                    ni[j] = __double2loint(t_);
                    r[j] = fma(t_ - EXP_MAGIC, -TE_C, u[j]);
                    T[j] = etab[ni[j] & 63];
                }
                double yp[4], ym[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double r2 = r[j] * r[j];
                    const double c = fma(r2, fma(r2, 1.0 / 24.0, 0.5), 1.0);
                    const double s1 = fma(r2, 1.0 / 6.0, 1.0);
                    yp[j] = fma(r[j], s1, c);
                    ym[j] = fma(-r[j], s1, c);
                    if (ok[j]) {
                        su += u[j];
                        suu = fma(u[j], u[j], suu);
                        umax = max(umax, __double2hiint(u[j]));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int sh = (ni[j] << 14) & 0xfff00000;
                    const double a = T[j].x * yp[j], b = T[j].y * ym[j];
                    if (ok[j]) {
                        ls += __hiloint2double(__double2hiint(a) + sh, __double2loint(a));  // exp(ll - c)
                        bs += __hiloint2double(__double2hiint(b) - sh, __double2loint(b));  // exp(-(ll - c))
                    }
                }
            };
            const bool all_ok[4] = {true, true, true, true};
            double v[4];
            if (p.debug & 8) {
            } else if (k4 > 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = pcol[j * KSTEP];
                for (int k = 0; k + 4 < k4; k += 4) {
                    double vn[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) vn[j] = pcol[(k + 4 + j) * KSTEP];
                    step(v, all_ok);
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = vn[j];
                }
                double vn[4];  // draws k4 .. k4 + 3, where the CTA's share of the draw axis ends: guarded
#pragma unroll
                for (int j = 0; j < 4; ++j) vn[j] = tail_ok[j] ? pcol[(k4 + j) * KSTEP] : cl;
                step(v, all_ok);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = vn[j];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = tail_ok[j] ? pcol[j * KSTEP] : cl;
            }
            if (!(p.debug & 8)) step(v, tail_ok);  // (a draw that is not there reads as the minimum, u = 0, and is left out of the sums)
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {  // the 16 draw slots of the column
                bs += __shfl_xor_sync(FULL, bs, o);
                ls += __shfl_xor_sync(FULL, ls, o);
                su += __shfl_xor_sync(FULL, su, o);
                suu += __shfl_xor_sync(FULL, suu, o);
                umax = max(umax, __shfl_xor_sync(FULL, umax, o));
            }
        }

        // ---------------- receive: column minimum, lower / upper median of the CTAs' thresholds
        mbar_wait_cluster(&bar_x[w * 2 + par], xph);
        double llmin, t_t, t_l;
        {
            const int r = slot & 7;  // lanes 8..15 of each half mirror lanes 0..7
            const bool have = r < csize;
            const XMsg e = xch[(par * TILE_MAXC + (have ? r : 0)) * TW + col];
            double mn = have ? e.mn : INF;
            const float st = have ? e.st : FINF, sl = have ? e.sl : FINF;
            int rt = 0, rl = 0;
#pragma unroll
            for (int q = 0; q < TILE_MAXC; ++q) {
                const float ot = __shfl_sync(FULL, st, (lane & 16) + q), ol = __shfl_sync(FULL, sl, (lane & 16) + q);
                rt += (ot < st || (ot == st && q < r)) ? 1 : 0;
                rl += (ol < sl || (ol == sl && q < r)) ? 1 : 0;
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) mn = min_sel(mn, __shfl_xor_sync(FULL, mn, o));
            // (the padding lanes' +inf rank last: positions 0 .. csize - 1 belong to the real CTAs)
            const unsigned mine = hf ? 0x00ff0000u : 0x000000ffu;
            const unsigned bt = __ballot_sync(FULL, have && rt == (csize - 1) / 2) & mine;
            const unsigned bl = __ballot_sync(FULL, have && rl == csize / 2) & mine;
            t_t = (double)__shfl_sync(FULL, st, __ffs(bt) - 1);
            t_l = (double)__shfl_sync(FULL, sl, __ffs(bl) - 1);
            if (nch > 1) t_t = t_l;  // chunked: one list (a chunk's tight list alone proves nothing about the column)
            llmin = mn;
        }
        // the previous tile's header (its partial sums came with this exchange)
        if (t_prev >= 0) write_header(t_prev, par, p_llmin, p_tl);

        // ---------------- pass C: candidates = draws at or below the loose threshold; tight ones separately
        if (!(p.debug & 4)) {
            unsigned mask = 0;
            {
                int k = 0;
                for (; k + 8 <= k4; k += 8) {  // eight draws in flight
                    double v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = pcol[(k + j) * KSTEP];
                    unsigned nib = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) nib |= (v[j] <= t_l) ? (1u << j) : 0u;
                    mask |= nib << k;
                }
                if (k < k4) {
                    unsigned nib = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) nib |= (pcol[(k + j) * KSTEP] <= t_l) ? (1u << j) : 0u;
                    mask |= nib << k;
                }
            }
            {
                unsigned nib = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (tail_ok[j]) nib |= (pcol[(k4 + j) * KSTEP] <= t_l) ? (1u << j) : 0u;
                mask |= nib << k4;
            }
            unsigned maskA = 0;
            for (unsigned m = mask; m; m &= m - 1) {
                const int b = __ffs((int)m) - 1;
                if (pcol[b * KSTEP] <= t_t) maskA |= 1u << b;
            }
            // positions: inclusive scan of the packed counts over the 16 draw slots, one atomic per column and list
            const unsigned nA = (unsigned)__popc(maskA), nB = (unsigned)__popc(mask) - nA;
            unsigned inc = nA | (nB << 16);
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const unsigned v = __shfl_up_sync(FULL, inc, o, 16);
                if (slot >= o) inc += v;
            }
            const unsigned tl_ = (unsigned)t / (unsigned)nch;
            const int draw0 = (int)((unsigned)t - tl_ * (unsigned)nch) * S + row0;  // matrix row of this CTA's first draw
            const long long o = (long long)tl_ * TW + col;
            unsigned baseA = 0, baseB = 0;
            if (slot == 15 && o < p.n_obs) {
                if (inc & 0xffffu) baseA = atomicAdd(&p.cnt[2 * o], inc & 0xffffu);
                if (inc >> 16) baseB = atomicAdd(&p.cnt[2 * o + 1], inc >> 16);
            }
            baseA = __shfl_sync(FULL, baseA, 15, 16);
            baseB = __shfl_sync(FULL, baseB, 15, 16);
            unsigned posA = baseA + (inc & 0xffffu) - nA, posB = baseB + (inc >> 16) - nB;
            if (mask && o < p.n_obs) {
                const unsigned ob = (unsigned)o * (unsigned)cap;  // (a round's scratch has far fewer than 2^32 slots)
                double* const dx = p.cx;
                unsigned short* const ds = p.cs;
                for (unsigned m = mask; m; m &= m - 1) {
                    const int b = __ffs((int)m) - 1;
                    const bool isA = (maskA >> b) & 1u;
                    const unsigned pos = isA ? posA++ : posB++;
                    if (pos < (unsigned)cap) {  // (an overflowing column is flagged by its owner and never read)
                        const unsigned sl_ = isA ? pos : (unsigned)cap - 1u - pos;
                        // x = fl(r - max r), exactly (psis.py:134); chunked: r = -ll, the tail kernel subtracts max r
                        const double v = pcol[b * KSTEP];
                        dx[ob + sl_] = (nch > 1) ? -v : llmin - v;
                        ds[ob + sl_] = (unsigned short)(draw0 + 16 * b + slot);
                    }
                }
            }
        }
        // this warp is done with the tile; the last one lets the producer refill it
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty);
        if (tid == 0 && t + n_clusters < n_units) {
            mbar_wait(bar_empty, (uint32_t)(it & 1));
            fence_proxy_async();
            issue(t + n_clusters);
        }
        __syncwarp();
        pq0 = bs; pq1 = ls; pq2 = su; pq3 = suu; pc = cl; pum = umax;
        p_llmin = llmin; p_tl = t_l;
        t_prev = t;
    }
    // the last tile's partial sums and header: one more exchange that carries only those
    if (t_prev >= 0 && !(p.debug & 2)) {
        const int par = it & 1;
        const uint32_t xph = (uint32_t)((it >> 1) & 1);
        if (lane == 0 && own_cols) mbar_expect_tx(&bar_x[w * 2 + par], (uint32_t)(own_cols * csize * 48));
        send_partials(par);
        if (own_cols) mbar_wait_cluster(&bar_x[w * 2 + par], xph);
        write_header(t_prev, par, p_llmin, p_tl);
    }
    // (a CTA's shared memory must outlive the peers' last stores into it: every warp has waited for all it expects)
    cluster_sync_all();
}

// ---------------------------------------------------------------- host side
bool tile_shape(long long S_total, int M, int csize, int tw, TilePlan* tp) {
    memset(tp, 0, sizeof(*tp));
    if (tw != 8 && tw != 16) return false;
    if (csize > tw) return false;
    tp->tw = tw;
    if (csize < 1 || csize > TILE_MAXC || (csize & (csize - 1)) != 0) return false;
    if (S_total < 512 || S_total > SPLIT_MAX_S) return false;
    // a draw axis longer than one cluster holds (8 x 512) is cut into 2, 3 or 4 equal chunks
    int nch = 1;
    if (S_total > (long long)TILE_MAX_R * TILE_MAXC) {
        nch = 0;
        for (int c = 2; c <= 4 && !nch; ++c)
            if (S_total % c == 0 && S_total / c <= (long long)TILE_MAX_R * TILE_MAXC) nch = c;
        if (!nch) return false;
    }
    if (const char* ev = getenv("B2L_TILE_CHUNKS")) {  // (tests: chunked units on short posteriors)
        const int c = atoi(ev);
        if (c >= 1 && c <= 4 && S_total % c == 0 && S_total / c >= 512 && S_total / c <= (long long)TILE_MAX_R * TILE_MAXC) nch = c;
    }
    const long long S = S_total / nch;
    tp->n_chunks = nch; tp->chunk_len = (int)S;
    const long long per = (S + csize - 1) / csize;  // draws a CTA owns
    if (per > TILE_MAX_R) return false;
    // boxes of the CTA's draws: <= 256 rows each, a multiple of 8 rows (every box starts on a 1024-byte swizzle
    // atom of the shared-memory tile); the few rows loaded beyond `per` belong to the next CTA and are not read
    int best_n = 0, best_rows = 0;
    for (int nb = (int)((per + 255) / 256); nb <= (int)((per + 255) / 256) + 3; ++nb) {
        const int rows = (int)(((per + nb - 1) / nb + 7) / 8 * 8);
        if (rows > 256) continue;
        if (!best_n || nb * rows < best_n * best_rows) { best_n = nb; best_rows = rows; }
    }
    if (!best_n) return false;
    tp->csize = csize; tp->nbox = best_n; tp->box_rows = best_rows; tp->R = (int)per;
    // threshold ranks: 32 bins of R / 32 draws per CTA and column, i.e. B = 32 * csize bins per column.  The draws
    // at or below the q-th smallest bin minimum of a CTA (lower / upper median over the CTAs) number about
    // K(q) = -0.9 B ln(1 - q / 32) with a spread of ~8 % (balls in bins; the 0.9 is measured).  Tight rank: K in
    // the middle of [M + 1, one sort of the tail kernel]; loose rank: K ~ 1.65 (M + 1), far from M + 1 and from cap.
    // Long tails (M + 1 > 0.9 B: reff well below 1) need ranks near 32, where the count follows the harmonic form
    // B (H_32 - H_(32 - q)) of the exponential order statistics, times 0.91 (lower median, tight) / 0.93 (upper
    // median, loose) -- simulated: q = 28: 453 / 482, q = 31: 670 / 719 of 4000 draws, spread 8-9 %.  Rank 31 must
    // still cover M + 1 with three spreads to spare.
    const double B = 32.0 * csize;
    auto h_of = [](int q) { double s = 0.0; for (int i = 32 - q + 1; i <= 32; ++i) s += 1.0 / i; return s; };
    const bool long_tail = (double)(M + 1) > 0.9 * B;
    if (nch == 1 && (double)(M + 1) > 0.70 * 0.93 * B * h_of(31)) return false;
    auto K_t = [&](int q) { return long_tail ? 0.91 * B * h_of(q) : -0.9 * B * std::log(1.0 - (double)q / 32.0); };
    auto K_l = [&](int q) { return long_tail ? 0.93 * B * h_of(q) : -0.9 * B * std::log(1.0 - (double)q / 32.0); };
    const int tl = (M + 2 <= 128) ? 4 : ((M + 2 <= 256) ? 8 : ((M + 2 <= 512) ? 16 : 32));  // the tail kernel's registers per lane (split_shape)
    const double two_sorts = (tl == 32) ? 1024.0 : 64.0 * tl;  // the longest list the tail kernel sorts
    double want_t = 0.5 * ((double)(M + 1) + 32.0 * tl), want_l = std::min(1.65 * (M + 1), 0.8 * two_sorts);
    if (nch > 1) {
        // chunked: ONE list; a chunk's threshold must admit every draw of the chunk that belongs to the column's
        // tail.  Its share of the M + 1 extreme draws is binomial (mean (M + 1) / n_chunks, spread ~9 % at M = 380)
        // and the count at a threshold rank spreads ~8 %: 1.6 shares per chunk leave 3.4 joint spreads.
        want_t = want_l = 1.6 * (double)(M + 1) / nch;
        if (want_l > 0.70 * 0.93 * B * h_of(31) || 1.6 * (double)(M + 1) > 0.9 * two_sorts) return false;
    }
    auto K_c = [&](int q) { return 0.93 * B * h_of(q); };
    int qt = 1, ql = 1;
    for (int q = 1; q <= 31; ++q) {
        if (nch > 1) {
            if (std::fabs(K_c(q) - want_l) < std::fabs(K_c(ql) - want_l)) qt = ql = q;
            continue;
        }
        if (std::fabs(K_t(q) - want_t) < std::fabs(K_t(qt) - want_t)) qt = q;
        if (std::fabs(K_l(q) - want_l) < std::fabs(K_l(ql) - want_l)) ql = q;
    }
    if (const char* ev = getenv("B2L_TILE_QT")) qt = atoi(ev);
    if (const char* ev = getenv("B2L_TILE_QL")) ql = atoi(ev);
    tp->q_t = std::min(31, std::max(1, qt));
    tp->q_l = std::min(31, std::max(tp->q_t, ql));
    tp->smem = tile_smem(tp->nbox * tp->box_rows, tw).total;
    return true;
}

template <int TW, bool CHUNKED>
static cudaError_t tile_occupancy(TilePlan* tp) {
    cudaError_t e = cudaFuncSetAttribute(loo_tile_kernel<TW, CHUNKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp->smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(loo_tile_kernel<TW, CHUNKED>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tp->occ, loo_tile_kernel<TW, CHUNKED>, 16 * TW, tp->smem);
    if (e != cudaSuccess) return e;
    if (tp->occ < 1) return cudaSuccess;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(tp->csize * 1024));
    cfg.blockDim = dim3(16 * TW);
    cfg.dynamicSmemBytes = tp->smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)tp->csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nc = 0;
    e = cudaOccupancyMaxActiveClusters(&nc, loo_tile_kernel<TW, CHUNKED>, &cfg);
    if (e != cudaSuccess) return e;
    tp->max_clusters = nc;
    return cudaSuccess;
}

// Shape of the plan without the device part: observations per tile, cluster size, chunks (pure arithmetic)
bool tile_pick(long long S, int M, TilePlan* tp) {
    int csize = TILE_MAXC, tw = 8;
    if (const char* ev = getenv("B2L_TILE_W")) tw = atoi(ev);
    if (const char* ev = getenv("B2L_TILE")) if (atoi(ev) == 0) { memset(tp, 0, sizeof(*tp)); return false; }
    // Cluster size: the fewest CTAs that hold the draws of a tile (<= 512 each) -- short posteriors in clusters of
    // 8 leave a thread 4 to 8 draws per pass and the per-tile exchange dominates (S = 1000, 606 208 observations:
    // 5.08 ms with 8 CTAs, 3.45 ms with 4; S = 512: 4.69 / 2.85 / 1.93 ms with 8 / 4 / 2) -- as long as the
    // thresholds stay reliable: medians over few CTAs of ranks near 32 are noisy (S = 1000 in clusters of 2,
    // 1.5 tail draws per bin: 0.5 % of the columns handed over), so at most 1.2 tail draws per bin.
    if (const char* ev = getenv("B2L_TILE_CSIZE")) {
        csize = atoi(ev);
    } else {
        for (int c = 2; c <= TILE_MAXC; c <<= 1) {
            TilePlan probe;
            if (tile_shape(S, M, c, tw, &probe) && probe.n_chunks == 1 && (double)(M + 1) <= 1.2 * 32.0 * c) { csize = c; break; }
        }
    }
    return tile_shape(S, M, csize, tw, tp);
}

cudaError_t tile_plan(long long S, int M, TilePlan* tp) {
    // 8-observation tiles (64 B of a draw, 4 warps, 6 CTAs / SM) measure ~10 % faster than 16-observation ones
    // (128 B, 8 warps, 3 CTAs / SM): more, smaller CTAs hide each other's waits better
    if (!tile_pick(S, M, tp)) return cudaSuccess;
    const int tw = tp->tw;
    int dev = 0, smem_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    if (tp->smem > (size_t)smem_optin) return cudaSuccess;
    if (tp->n_chunks > 1) e = (tw == 16) ? tile_occupancy<16, true>(tp) : tile_occupancy<8, true>(tp);
    else e = (tw == 16) ? tile_occupancy<16, false>(tp) : tile_occupancy<8, false>(tp);
    if (e != cudaSuccess) return e;
    if (tp->occ < 1 || tp->max_clusters < 1) return cudaSuccess;
    tp->ok = 1;
    return cudaSuccess;
}

cudaError_t tile_tensor_map(const double* ll, long long S, long long N, long long stride_s, int tw, int box_rows,
                            void* tmap_out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !sym) return cudaErrorNotSupported;
        fn = reinterpret_cast<EncodeFn>(sym);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)S};
    const cuuint64_t strides[1] = {(cuuint64_t)stride_s * 8ull};  // bytes between consecutive draws
    const cuuint32_t box[2] = {(cuuint32_t)tw, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(reinterpret_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
                          const_cast<double*>(ll), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          tw == 16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t tile_launch(const TilePlan& tp, const void* tmap, const TileParams& p, cudaStream_t st) {
    if (p.n_tiles <= 0) return cudaSuccess;
    const long long nc = std::min<long long>(tp.max_clusters, p.n_tiles * std::max(1, p.n_chunks));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(nc * tp.csize));
    cfg.blockDim = dim3(16 * tp.tw);
    cfg.dynamicSmemBytes = tp.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)tp.csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(tmap);
    if (p.n_chunks > 1)
        return (tp.tw == 16) ? cudaLaunchKernelEx(&cfg, loo_tile_kernel<16, true>, tm, p)
                             : cudaLaunchKernelEx(&cfg, loo_tile_kernel<8, true>, tm, p);
    return (tp.tw == 16) ? cudaLaunchKernelEx(&cfg, loo_tile_kernel<16, false>, tm, p)
                         : cudaLaunchKernelEx(&cfg, loo_tile_kernel<8, false>, tm, p);
}

// Chunked rounds: one thread per observation folds the chunk records into the SplitHeader the tail kernel reads
// (sums re-centred about the column minimum exactly as the owner CTA does for a tile's CTAs) and decides what is
// handed over.
static __global__ void __launch_bounds__(128) tile_merge_kernel(const TileParams p, const double S_total) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= p.n_obs) return;
    const int nch = p.n_chunks;
    const ChunkHeader* c = p.chdr + o * nch;
    double llmin = c[0].llmin;
    for (int i = 1; i < nch; ++i) llmin = min_sel(llmin, c[i].llmin);
    const double n_r = (double)p.S;
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, um = 0.0, vmax = -inf_f64(), vmin = inf_f64();
    for (int i = 0; i < nch; ++i) {
        const ChunkHeader e = c[i];
        const double d = e.llmin - llmin;  // >= 0
        q0 += e.q0 * exp(-d);
        q1 += e.q1 * exp(d);
        q2 += fma(n_r, d, e.q2);
        q3 += fma(d, fma(n_r, d, 2.0 * e.q2), e.q3);
        um = fmax(um, e.um + d);
        const double v = llmin - e.tl;  // the chunk's candidates have x >= v, its other draws x <= v
        vmax = fmax(vmax, v);
        vmin = fmin(vmin, v);
    }
    const int ca = (int)p.cnt[2 * o];
    const bool special = !(is_finite(q2) && is_finite(q3) && is_finite(q0) && is_finite(q1));
    const bool wide = !(um <= 600.0);
    const bool count_bad = (ca < p.M + 1) || (ca > p.cap);
    const bool ok = !special && !wide && !count_bad;
    SplitHeader h;
    h.mx = -llmin;
    h.body = q0;
    h.lsum = q1;
    h.vsum = q3 - q2 * q2 / S_total;
    if (h.vsum < 0.0) h.vsum = 0.0;
    h.lshift = llmin;
    h.taux = vmin;
    h.lse = 0.0;
    h.C = ca; h.flags = ok ? 0 : 1; h.attempts = 0; h.n_patch = 0; h.C2 = 0; h.pad_ = 0;
    h.vmax = vmax; h.pad2_ = 0.0;
    p.hdr[o] = h;
    if (!ok) {
        p.fb_list[atomicAdd(p.fb_count, 1)] = (int)(p.row_base + o);
        if (p.counters) atomicAdd(&p.counters[3], 1ull);
        note_handover(special ? HO_SPECIAL : (wide ? HO_RANGE : HO_RETRY));
    }
}

cudaError_t tile_merge_launch(const TileParams& p, long long S_total, cudaStream_t st) {
    if (p.n_obs <= 0 || p.n_chunks <= 1) return cudaSuccess;
    tile_merge_kernel<<<(unsigned)((p.n_obs + 127) / 128), 128, 0, st>>>(p, (double)S_total);
    return cudaGetLastError();
}

cudaError_t tile_reasons(unsigned long long* out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_handover, sizeof(unsigned long long) * HO_REASONS);
    if (e == cudaSuccess && reset) {
        unsigned long long z[HO_REASONS] = {0};
        e = cudaMemcpyToSymbol(g_handover, z, sizeof(z));
    }
    return e;
}

}  // namespace b2l
