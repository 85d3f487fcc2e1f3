// Tile path (sm_100a): the stream pass of pl.loo on the ArviZ (chain, draw, obs) layout, reading the
// observation-fastest (S, N) log-likelihood matrix where it lies -- no transposed copy, ONE read of HBM.
//
//   loo_tile_kernel   a thread-block CLUSTER owns a tile of TILE_W = 16 observations x all S draws (128 B of
//                     every draw: whole DRAM bursts).  CTA r of the cluster stages draws [r R, (r+1) R) of the
//                     tile in its shared memory with 2-D TMA boxes {16 observations x box_rows draws}
//                     (cp.async.bulk.tensor.2d, SASS UTMALDG), so the cluster as a whole holds the tile on chip
//                     and every pass after the first reads shared memory.  Threads map 16 draw slots x 16
//                     observations; a column's partial results meet through distributed shared memory
//                     (st.shared::cluster + barrier.cluster), twice per tile:
//                       pass A   per-thread minima of ll: the column minimum (r = -ll, max r = -min ll,
//                                pyloo/loo.py:286-288 + pyloo/psis.py:134) and 32 bin minima per CTA whose
//                                sorted ranks q_t / q_l give a tight and a loose candidate threshold
//                       -------- exchange 1: (min, two thresholds) from every CTA to every CTA
//                       pass B   u = ll - min ll = -x exactly (x = fl(r - max r), psis.py:134); one range
//                                reduction gives exp(-u) for the PSIS normaliser (pyloo/utils.py:349-351) and
//                                exp(u) for lppd_i (pyloo/loo.py:329-337); sums of u and u^2 for var_s(ll)
//                                (pyloo/waic.py:145); draws at or below the loose threshold are marked
//                       -------- exchange 2: partial sums to the column's owner CTA, candidate counts to all
//                       emit     candidates (exact x, draw index) to the round's scratch in a fixed order
//                                (CTA rank, thread, draw): tight ones from the front, looser ones from the end
//                     The owner CTA writes the 80-byte SplitHeader the tail kernel reads (b2l_split.cuh).
//   psis_tail_kernel  unchanged contract: sort, cutoff, GPD fit, smoothing, elpd_i (one warp per observation).
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
struct XchA {  // exchange 1, one per (source CTA, column)
    double mn;     // minimum of ll over the source CTA's draws
    float st, sl;  // its tight / loose threshold statistic
};
struct RSum {  // exchange 2, one per (source CTA, owned column); double-buffered by tile parity
    double q[4];  // about the SOURCE CTA's minimum c: sum exp(-(ll - c)), sum exp(ll - c), sum (ll - c), sum (ll - c)^2
    double c;
    int umax;     // max over the high words of ll - c (>= 0: integer order = value order)
    int pad;
};
struct ColInfo {
    double llmin, t_t, t_l;  // column minimum, tight / loose candidate threshold (ll <= t)
};
constexpr int TILE_RSUM = 16;  // csize * (16 / csize): the cluster size is a power of two
struct TileSmemLayout {
    size_t off_tab, off_scr, off_um, off_xch, off_rsum, off_info, off_cl, off_bar, total;
};
__host__ __device__ inline TileSmemLayout tile_smem(int R) {
    TileSmemLayout L;
    size_t o = (size_t)R * TILE_W * 8;  // the tile: R draws x 16 observations, 128 B per draw
    L.off_tab = o;   // 64 x (2^(j/64), 2^(-j/64))
    o += 64 * 16;
    L.off_scr = o;   // pass A: 32 x 16 float bin minima + 16 x 16 double thread minima; pass B: 8 x 16 x 4 partial sums
    o += 4096;
    L.off_um = o;    // 8 x 16 high-word maxima
    o += 8 * TILE_W * 4;
    L.off_xch = o;   // written by the other CTAs of the cluster: [parity][source CTA][column]
    o += 2 * TILE_MAXC * TILE_W * sizeof(XchA);
    L.off_rsum = o;  // written by the other CTAs of the cluster: [parity][source CTA][owned slot]
    o += 2 * TILE_RSUM * sizeof(RSum);
    L.off_info = o;  // [parity][column]
    o += 2 * TILE_W * sizeof(ColInfo);
    L.off_cl = o;    // this CTA's column minima
    o += TILE_W * 8;
    L.off_bar = o;
    o += 64;
    L.total = align_up(o, 128);
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
// split cluster barrier: what a CTA wrote to its peers' shared memory before arrive is visible to them after wait
__device__ __forceinline__ void cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_st_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st_v2f32(uint32_t addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void dsmem_st_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1)
                 : "memory");
}

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

// two independent ascending 32-lane bitonic sorts in one pass (the direction logic is shared)
__device__ __forceinline__ void warp_sort32_f2(float& a, float& b, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float oa = __shfl_xor_sync(FULL, a, j), ob = __shfl_xor_sync(FULL, b, j);
            const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
            a = keep_min ? fminf(a, oa) : fmaxf(a, oa);
            b = keep_min ? fminf(b, ob) : fmaxf(b, ob);
        }
    }
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(TILE_NT, 3) loo_tile_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const TileParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmemLayout L = tile_smem(p.R);
    double* tile = reinterpret_cast<double*>(smem_raw);
    double2* etab = reinterpret_cast<double2*>(smem_raw + L.off_tab);
    float* binsf = reinterpret_cast<float*>(smem_raw + L.off_scr);          // [32][16]
    double* cmin = reinterpret_cast<double*>(smem_raw + L.off_scr + 2048);   // [16][16]
    double* part = reinterpret_cast<double*>(smem_raw + L.off_scr);         // [8][16][4]
    int* umW = reinterpret_cast<int*>(smem_raw + L.off_um);                 // [8][16]
    XchA* xch = reinterpret_cast<XchA*>(smem_raw + L.off_xch);              // [2][csize][16]
    RSum* rsum = reinterpret_cast<RSum*>(smem_raw + L.off_rsum);            // [2][csize * nslot]
    ColInfo* info = reinterpret_cast<ColInfo*>(smem_raw + L.off_info);      // [2][16]
    double* clocal = reinterpret_cast<double*>(smem_raw + L.off_cl);        // [16]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int col = lane & 15, hf = lane >> 4, rw = 2 * w + hf;  // observation of the tile, draw slot 0..15
    const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
    const int nslot = (TILE_W + csize - 1) / csize;  // columns a CTA owns: col = slot * csize + rank
    const long long cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;
    const int S = p.S, M = p.M, cap = p.cap, R = p.R;
    const int row0 = crank * R;
    const int rows_live = max(0, min(R, S - row0));
    const int kmin = rows_live >> 4;                  // draws every slot of this CTA has
    const bool extra = (kmin << 4) + rw < rows_live;  // one more for the first slots
    const uint32_t tile_tx = (uint32_t)R * TILE_W * 8u;
    const double INF = inf_f64();

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (tid < 64) etab[tid] = make_double2(exp2((double)tid / 64.0), exp2(-(double)tid / 64.0));
    __syncthreads();
    cluster_arrive();  // every CTA of the cluster is running: its shared memory may be written remotely
    cluster_wait();

    auto issue = [&](long long t) {  // one thread: the CTA's R draws of tile t, nbox boxes on one mbarrier
        mbar_expect_tx(bar, tile_tx);
        for (int b = 0; b < p.nbox; ++b)
            tma_load_2d(tile + (size_t)b * p.box_rows * TILE_W, &tmap, (int)(p.col0 + t * TILE_W),
                        row0 + b * p.box_rows, bar);
    };
    // The header of a tile (what the tail kernel reads) is written by the owner CTA of each column one tile
    // LATER, after the next cluster barrier: by then every CTA's partial sums and candidate counts are in.
    auto write_headers = [&](long long tt, int par) {  // warp 7: lane = 8 * slot' + source rank
        const int slot = lane >> 3, r = lane & 7;
        for (int s0 = 0; s0 < nslot; s0 += 4) {
            const int sl_ = s0 + slot;
            const int c = sl_ * csize + crank;
            const long long oo = tt * TILE_W + c;
            const bool live_col = sl_ < nslot && c < TILE_W && oo < p.n_obs;
            const bool have = live_col && r < csize;
            const ColInfo ci = info[par * TILE_W + (live_col ? c : 0)];
            double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, um = 0.0;
            if (have) {
                const RSum e = rsum[par * TILE_RSUM + r * nslot + sl_];
                const double n_r = (double)max(0, min(R, S - r * R));
                if (n_r > 0.0) {
                    const double d = e.c - ci.llmin;  // >= 0: this CTA's sums are about its own minimum
                    q0 = e.q[0] * exp(-d);
                    q1 = e.q[1] * exp(d);
                    q2 = fma(n_r, d, e.q[2]);
                    q3 = fma(d, fma(n_r, d, 2.0 * e.q[2]), e.q[3]);
                    um = __hiloint2double(e.umax, 0) * 1.000002 + d;  // upper bound of max (ll - min ll)
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {  // over the 8 source CTAs, fixed order
                q0 += __shfl_xor_sync(FULL, q0, o);
                q1 += __shfl_xor_sync(FULL, q1, o);
                q2 += __shfl_xor_sync(FULL, q2, o);
                q3 += __shfl_xor_sync(FULL, q3, o);
                um = fmax(um, __shfl_xor_sync(FULL, um, o));
            }
            if (live_col && r == 0) {
                const int ca = (int)atomicAdd(&p.cnt[2 * oo], 0u), cb = (int)atomicAdd(&p.cnt[2 * oo + 1], 0u);
                const bool special = !(is_finite(q2) && is_finite(q3) && is_finite(q0) && is_finite(q1));
                const bool wide = !(um <= 600.0);
                const bool count_bad = (ca + cb < M + 1) || (ca + cb > cap);
                const bool ok = !special && !wide && !count_bad;
                SplitHeader h;
                h.mx = -ci.llmin;
                h.body = q0;
                h.lsum = q1;
                h.vsum = q3 - q2 * q2 / (double)S;  // sum (ll - mean)^2 taken about the minimum
                if (h.vsum < 0.0) h.vsum = 0.0;
                h.lshift = ci.llmin;
                h.taux = ci.llmin - ci.t_l;  // every candidate has ll <= t_l, i.e. x = fl(min ll - ll) >= taux
                h.lse = 0.0;
                h.C = ca; h.flags = ok ? 0 : 1; h.attempts = 0; h.n_patch = 0; h.C2 = cb; h.pad_ = 0;
                p.hdr[oo] = h;
                if (!ok) {
                    p.fb_list[atomicAdd(p.fb_count, 1)] = (int)(p.row_base + oo);
                    if (p.counters) atomicAdd(&p.counters[3], 1ull);
                    note_handover(special ? HO_SPECIAL : (wide ? HO_RANGE : HO_RETRY));
                }
            }
        }
    };

    long long t = cluster_id;
    if (tid == 0 && t < p.n_tiles) issue(t);
    uint32_t phase = 0;
    const double* pcol = tile + rw * TILE_W + col;  // this thread's draws: pcol[k * 256]
    long long t_prev = -1;
    int par = 0;

    for (; t < p.n_tiles; t += n_clusters, par ^= 1) {
        // the next tile of this CTA: HBM -> L2 while this one is worked on
        if (tid == 0 && t + n_clusters < p.n_tiles)
            for (int b = 0; b < p.nbox; ++b)
                tma_prefetch_2d(&tmap, (int)(p.col0 + (t + n_clusters) * TILE_W), row0 + b * p.box_rows);
        mbar_wait(bar, phase);
        phase ^= 1u;

        // ---------------- pass A: minima.  Two alternating bins per thread (32 bins per CTA and column).
        {
            double mA = INF, mB = INF;
            int k = 0;
#pragma unroll 4
            for (; k + 2 <= kmin; k += 2) {
                const double v0 = pcol[k * 256], v1 = pcol[(k + 1) * 256];
                mA = min_sel(mA, v0);
                mB = min_sel(mB, v1);
            }
            if (k < kmin) mA = min_sel(mA, pcol[k * 256]);
            if (extra) mB = min_sel(mB, pcol[kmin * 256]);
            binsf[(2 * rw) * TILE_W + col] = (float)mA;
            binsf[(2 * rw + 1) * TILE_W + col] = (float)mB;
            cmin[rw * TILE_W + col] = min_sel(mA, mB);
        }
        __syncthreads();
        // per column: this CTA's minimum and the q_t-th / q_l-th smallest of its 32 bin minima -> every CTA
        {
            const int c0 = 2 * w;
            float a = binsf[lane * TILE_W + c0], b = binsf[lane * TILE_W + c0 + 1];
            warp_sort32_f2(a, b, lane);
            const float st0 = __shfl_sync(FULL, a, p.q_t - 1), sl0 = __shfl_sync(FULL, a, p.q_l - 1);
            const float st1 = __shfl_sync(FULL, b, p.q_t - 1), sl1 = __shfl_sync(FULL, b, p.q_l - 1);
            double mn = cmin[col * TILE_W + c0 + hf];  // lanes 0..15: column c0, lanes 16..31: column c0 + 1
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mn = min_sel(mn, __shfl_xor_sync(FULL, mn, o));
            if (col == 0) clocal[c0 + hf] = mn;
            if (col < csize) {
                const uint32_t ad = dsmem_addr(&xch[(par * TILE_MAXC + crank) * TILE_W + c0 + hf], (uint32_t)col);
                dsmem_st_f64(ad, mn);
                dsmem_st_v2f32(ad + 8, hf ? st1 : st0, hf ? sl1 : sl0);
            }
        }
        cluster_arrive();  // exchange 1 is on its way; its wait sits behind pass B
        __syncthreads();

        // ---------------- pass B: sums over all draws about this CTA's column minimum (no peer data needed)
        {
            const double cl = clocal[col];
            double bs = 0.0, ls = 0.0, su = 0.0, suu = 0.0;
            int umax = 0;
            auto fold = [&](double v) {
                const double u = v - cl;  // >= 0
                double ep, em;
                exp_pm64(u, etab, ep, em);
                bs += em;  // exp(-(ll - c))
                ls += ep;  // exp(ll - c)
                su += u;
                suu = fma(u, u, suu);
                umax = max(umax, __double2hiint(u));
            };
            int k = 0;
#pragma unroll 4
            for (; k < kmin; ++k) fold(pcol[k * 256]);
            if (extra) fold(pcol[kmin * 256]);
            // the two draw slots of a warp that share a column, then the 8 warps through shared memory
            bs += __shfl_xor_sync(FULL, bs, 16);
            ls += __shfl_xor_sync(FULL, ls, 16);
            su += __shfl_xor_sync(FULL, su, 16);
            suu += __shfl_xor_sync(FULL, suu, 16);
            umax = max(umax, __shfl_xor_sync(FULL, umax, 16));
            if (lane < 16) {
                double* d = part + (w * TILE_W + col) * 4;
                d[0] = bs; d[1] = ls; d[2] = su; d[3] = suu;
                umW[w * TILE_W + col] = umax;
            }
        }
        cluster_wait();  // exchange 1 has arrived from every CTA
        // per column: the global minimum and the lower / upper median of the CTAs' thresholds
        {
            const int c = 2 * w + hf;
            const int r = col & 7;  // lanes 8..15 of each half mirror lanes 0..7
            const bool have = r < csize;
            const XchA e = xch[(par * TILE_MAXC + (have ? r : 0)) * TILE_W + c];
            double mn = have ? e.mn : INF;
            const float st = have ? e.st : __int_as_float(0x7f800000), sl = have ? e.sl : __int_as_float(0x7f800000);
            int rt = 0, rl = 0;
#pragma unroll
            for (int q = 0; q < TILE_MAXC; ++q) {
                const float ot = __shfl_sync(FULL, st, (lane & 16) + q), ol = __shfl_sync(FULL, sl, (lane & 16) + q);
                rt += (ot < st || (ot == st && q < r)) ? 1 : 0;
                rl += (ol < sl || (ol == sl && q < r)) ? 1 : 0;
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) mn = min_sel(mn, __shfl_xor_sync(FULL, mn, o));
            // (ranks count the padding lanes' +inf as larger: positions 0 .. csize - 1 belong to the real CTAs)
            const unsigned bt = __ballot_sync(FULL, have && col < 8 && rt == (csize - 1) / 2);
            const unsigned bl = __ballot_sync(FULL, have && col < 8 && rl == csize / 2);
            const unsigned mine = hf ? 0xffff0000u : 0x0000ffffu;
            const float tt_ = __shfl_sync(FULL, st, __ffs(bt & mine) - 1), tl_ = __shfl_sync(FULL, sl, __ffs(bl & mine) - 1);
            if (col == 0) {
                ColInfo ci;
                ci.llmin = mn; ci.t_t = (double)tt_; ci.t_l = (double)tl_;
                info[par * TILE_W + c] = ci;
            }
        }
        __syncthreads();
        // this CTA's partial sums -> the column's owner (for the header written one tile later)
        if (tid < 64) {
            const int c = tid & 15, q = tid >> 4;
            double a = 0.0;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) a += part[(ww * TILE_W + c) * 4 + q];
            dsmem_st_f64(dsmem_addr(&rsum[par * TILE_RSUM + crank * nslot + c / csize].q[q], (uint32_t)(c % csize)), a);
        } else if (tid < 80) {
            const int c = tid - 64;
            int a = 0;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) a = max(a, umW[ww * TILE_W + c]);
            RSum* dst = &rsum[par * TILE_RSUM + crank * nslot + c / csize];
            dsmem_st_u32(dsmem_addr(&dst->umax, (uint32_t)(c % csize)), (uint32_t)a);
            dsmem_st_f64(dsmem_addr(&dst->c, (uint32_t)(c % csize)), clocal[c]);
        }
        // the previous tile's headers: everything they need arrived before this tile's barrier
        if (w == 7 && t_prev >= 0) write_headers(t_prev, par ^ 1);

        // ---------------- pass C: candidates = draws at or below the loose threshold; tight ones separately
        {
            const ColInfo ci = info[par * TILE_W + col];
            unsigned mask = 0;
            int k = 0;
            for (; k + 4 <= kmin; k += 4) {
                unsigned nib = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) nib |= (pcol[(k + j) * 256] <= ci.t_l) ? (1u << j) : 0u;
                mask |= nib << k;
            }
            for (; k < kmin; ++k) mask |= (pcol[k * 256] <= ci.t_l) ? (1u << k) : 0u;
            if (extra) mask |= (pcol[kmin * 256] <= ci.t_l) ? (1u << kmin) : 0u;
            unsigned maskA = 0;
            for (unsigned m = mask; m; m &= m - 1) {
                const int b = __ffs((int)m) - 1;
                if (pcol[b * 256] <= ci.t_t) maskA |= 1u << b;
            }
            // positions: one atomic per (warp, column, list) on the observation's global counters
            const long long o = t * TILE_W + col;
            const unsigned nA = (unsigned)__popc(maskA), nB = (unsigned)__popc(mask) - nA;
            const unsigned nA2 = __shfl_down_sync(FULL, nA, 16), nB2 = __shfl_down_sync(FULL, nB, 16);
            unsigned baseA = 0, baseB = 0;
            if (lane < 16 && o < p.n_obs) {
                if (nA + nA2) baseA = atomicAdd(&p.cnt[2 * o], nA + nA2);
                if (nB + nB2) baseB = atomicAdd(&p.cnt[2 * o + 1], nB + nB2);
            }
            const unsigned bA = __shfl_sync(FULL, baseA, col), bB = __shfl_sync(FULL, baseB, col);
            const unsigned nA1 = __shfl_sync(FULL, nA, col), nB1 = __shfl_sync(FULL, nB, col);
            unsigned posA = bA + (hf ? nA1 : 0u), posB = bB + (hf ? nB1 : 0u);
            if (mask && o < p.n_obs) {
                double* dx = p.cx + (size_t)o * (size_t)cap;
                unsigned short* ds = p.cs + (size_t)o * (size_t)cap;
                for (unsigned m = mask; m; m &= m - 1) {
                    const int b = __ffs((int)m) - 1;
                    const bool isA = (maskA >> b) & 1u;
                    const unsigned pos = isA ? posA++ : posB++;
                    if (pos < (unsigned)cap) {  // (an overflowing column is flagged by its owner and never read)
                        const unsigned slot = isA ? pos : (unsigned)cap - 1u - pos;
                        dx[slot] = ci.llmin - pcol[b * 256];  // x = fl(r - max r), exactly (psis.py:134)
                        ds[slot] = (unsigned short)(row0 + 16 * b + rw);
                    }
                }
            }
        }
        t_prev = t;
        __syncthreads();  // the tile has been read for the last time
        if (tid == 0 && t + n_clusters < p.n_tiles) {
            fence_proxy_async();
            issue(t + n_clusters);
        }
    }
    // the last tile's headers
    cluster_arrive();
    cluster_wait();
    if (w == 7 && t_prev >= 0) write_headers(t_prev, par ^ 1);
}

// ---------------------------------------------------------------- host side
bool tile_shape(long long S, int M, int csize, TilePlan* tp) {
    memset(tp, 0, sizeof(*tp));
    if (csize < 1 || csize > TILE_MAXC || (csize & (csize - 1)) != 0) return false;
    if (S < 1024 || S > SPLIT_MAX_S) return false;
    const long long per = (S + csize - 1) / csize;
    if (per > TILE_MAX_R) return false;
    const int nbox = (int)((per + 255) / 256);
    const int box_rows = (int)((per + nbox - 1) / nbox);
    tp->csize = csize; tp->nbox = nbox; tp->box_rows = box_rows; tp->R = nbox * box_rows;
    if (tp->R > TILE_MAX_R) return false;
    // threshold ranks: 32 bins of R / 32 draws per CTA and column, i.e. B = 32 * csize bins per column.  The draws
    // at or below the q-th smallest bin minimum of a CTA (lower / upper median over the CTAs) number about
    // K(q) = -0.9 B ln(1 - q / 32) with a spread of ~8 % (balls in bins; the 0.9 is measured).  Tight rank: K in
    // the middle of [M + 1, one sort of the tail kernel]; loose rank: K ~ 1.65 (M + 1), far from M + 1 and from cap.
    const double B = 32.0 * csize;
    if ((double)(M + 1) > 0.9 * B) return false;
    auto K_of = [&](int q) { return -0.9 * B * std::log(1.0 - (double)q / 32.0); };
    const int tl = (M + 2 <= 128) ? 4 : ((M + 2 <= 256) ? 8 : 16);  // the tail kernel's registers per lane (split_shape)
    const double want_t = 0.5 * ((double)(M + 1) + 32.0 * tl), want_l = std::min(1.65 * (M + 1), 0.8 * 64.0 * tl);
    int qt = 1, ql = 1;
    for (int q = 1; q <= 31; ++q) {
        if (std::fabs(K_of(q) - want_t) < std::fabs(K_of(qt) - want_t)) qt = q;
        if (std::fabs(K_of(q) - want_l) < std::fabs(K_of(ql) - want_l)) ql = q;
    }
    if (const char* ev = getenv("B2L_TILE_QT")) qt = atoi(ev);
    if (const char* ev = getenv("B2L_TILE_QL")) ql = atoi(ev);
    tp->q_t = std::min(31, std::max(1, qt));
    tp->q_l = std::min(31, std::max(tp->q_t, ql));
    tp->smem = tile_smem(tp->R).total;
    return true;
}

cudaError_t tile_plan(long long S, int M, TilePlan* tp) {
    int csize = TILE_MAXC;
    if (const char* ev = getenv("B2L_TILE_CSIZE")) csize = atoi(ev);
    if (const char* ev = getenv("B2L_TILE")) if (atoi(ev) == 0) { memset(tp, 0, sizeof(*tp)); return cudaSuccess; }
    if (!tile_shape(S, M, csize, tp)) return cudaSuccess;
    int dev = 0, smem_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    if (tp->smem > (size_t)smem_optin) return cudaSuccess;
    e = cudaFuncSetAttribute(loo_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp->smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(loo_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tp->occ, loo_tile_kernel, TILE_NT, tp->smem);
    if (e != cudaSuccess) return e;
    if (tp->occ < 1) return cudaSuccess;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(tp->csize * 1024));
    cfg.blockDim = dim3(TILE_NT);
    cfg.dynamicSmemBytes = tp->smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)tp->csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nc = 0;
    e = cudaOccupancyMaxActiveClusters(&nc, loo_tile_kernel, &cfg);
    if (e != cudaSuccess) return e;
    if (nc < 1) return cudaSuccess;
    tp->max_clusters = nc;
    tp->ok = 1;
    return cudaSuccess;
}

cudaError_t tile_tensor_map(const double* ll, long long S, long long N, long long stride_s, int box_rows,
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
    const cuuint32_t box[2] = {(cuuint32_t)TILE_W, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(reinterpret_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
                          const_cast<double*>(ll), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t tile_launch(const TilePlan& tp, const void* tmap, const TileParams& p, cudaStream_t st) {
    if (p.n_tiles <= 0) return cudaSuccess;
    const long long nc = std::min<long long>(tp.max_clusters, p.n_tiles);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(nc * tp.csize));
    cfg.blockDim = dim3(TILE_NT);
    cfg.dynamicSmemBytes = tp.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)tp.csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, loo_tile_kernel, *reinterpret_cast<const CUtensorMap*>(tmap), p);
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
