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
    double mn;
    float st, sl;
};
struct RSum {  // exchange 2, one per (source CTA, owned column)
    double q[4];  // sum exp(-u), sum exp(u), sum u, sum u^2
    int umax;     // max over the high words of u (u >= 0: integer order = value order)
    int pad;
};
struct ColInfo {
    double llmin, tu_t, tu_l, pad;
};
struct TileSmemLayout {
    size_t off_tab, off_scr, off_cnt, off_um, off_xch, off_xcnt, off_rsum, off_info, off_bar, total;
};
__host__ __device__ inline TileSmemLayout tile_smem(int R) {
    TileSmemLayout L;
    size_t o = (size_t)R * TILE_W * 8;  // the tile: R draws x 16 observations, 128 B per draw
    L.off_tab = o;   // 128 x (2^(j/128), 2^(-j/128))
    o += 128 * 16;
    L.off_scr = o;   // pass A: 32 x 16 float bin minima + 16 x 16 double thread minima; pass B: 8 x 16 x 4 partial sums
    o += 4096;
    L.off_cnt = o;   // 16 x 16 packed candidate counts (tight | loose << 16), then their exclusive prefixes
    o += 16 * TILE_W * 4;
    L.off_um = o;    // 8 x 16 high-word maxima
    o += 8 * TILE_W * 4;
    L.off_xch = o;   // written by the other CTAs of the cluster
    o += TILE_MAXC * TILE_W * sizeof(XchA);
    L.off_xcnt = o;  // written by the other CTAs of the cluster
    o += TILE_MAXC * TILE_W * 4;
    L.off_rsum = o;  // written by the other CTAs of the cluster: [source CTA][owned slot], <= 24 entries
    o += 24 * sizeof(RSum);
    L.off_info = o;
    o += TILE_W * sizeof(ColInfo);
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
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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

// exp(u) and exp(-u) from one range reduction, 0 <= u <= 600: u = n ln2/128 + r, |r| <= ln2/256,
// e^(+-u) = 2^(+-(n >> 7)) 2^(+-(n & 127)/128) (cosh r +- sinh r) with cosh r ~ 1 + r^2/2, sinh r ~ r (1 + r^2/6):
// relative error < 2.3e-12 (the dropped r^4/24 term) -- these feed normalising sums of S terms whose logarithm is
// compared at 1e-10; the exact tail values are formed by the tail kernel from the exact x.
constexpr double TE_L = 184.66496523378731;     // 128 / ln 2
constexpr double TE_C = 0.0054152123481245727;  // ln 2 / 128
__device__ __forceinline__ void exp_pm128(double u, const double2* tab, double& ep, double& em) {
    const double t = fma(u, TE_L, EXP_MAGIC);
    const int ni = __double2loint(t);
    const double nf = t - EXP_MAGIC;
    const double r = fma(nf, -TE_C, u);
    const double r2 = r * r;
    const double c = fma(r2, 0.5, 1.0);
    const double s1 = fma(r2, 1.0 / 6.0, 1.0);
    const double2 T = tab[ni & 127];
    const int sh = (ni << 13) & 0xfff00000;
    const double yp = T.x * fma(r, s1, c), ym = T.y * fma(-r, s1, c);
    ep = __hiloint2double(__double2hiint(yp) + sh, __double2loint(yp));
    em = __hiloint2double(__double2hiint(ym) - sh, __double2loint(ym));
}

constexpr int WIDE_HI = 0x4082c000;  // high word of 600.0

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(TILE_NT, 3) loo_tile_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const TileParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmemLayout L = tile_smem(p.R);
    double* tile = reinterpret_cast<double*>(smem_raw);
    double2* etab = reinterpret_cast<double2*>(smem_raw + L.off_tab);
    float* binsf = reinterpret_cast<float*>(smem_raw + L.off_scr);         // [32][16]
    double* cmin = reinterpret_cast<double*>(smem_raw + L.off_scr + 2048);  // [16][16]
    double* part = reinterpret_cast<double*>(smem_raw + L.off_scr);        // [8][16][4]
    unsigned* cntT = reinterpret_cast<unsigned*>(smem_raw + L.off_cnt);    // [16][16]
    int* umW = reinterpret_cast<int*>(smem_raw + L.off_um);                // [8][16]
    XchA* xch = reinterpret_cast<XchA*>(smem_raw + L.off_xch);             // [csize][16]
    unsigned* xcnt = reinterpret_cast<unsigned*>(smem_raw + L.off_xcnt);   // [csize][16]
    RSum* rsum = reinterpret_cast<RSum*>(smem_raw + L.off_rsum);           // [csize][nslot]
    ColInfo* info = reinterpret_cast<ColInfo*>(smem_raw + L.off_info);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int col = lane & 15, rw = 2 * w + (lane >> 4);  // observation of the tile, draw slot 0..15
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
    if (tid < 128) etab[tid] = make_double2(exp2((double)tid / 128.0), exp2(-(double)tid / 128.0));
    __syncthreads();
    cluster_sync_all();  // every CTA of the cluster is running: its shared memory may be written remotely

    auto issue = [&](long long t) {  // one thread: the CTA's R draws of tile t, nbox boxes on one mbarrier
        mbar_expect_tx(bar, tile_tx);
        for (int b = 0; b < p.nbox; ++b)
            tma_load_2d(tile + (size_t)b * p.box_rows * TILE_W, &tmap, (int)(p.col0 + t * TILE_W),
                        row0 + b * p.box_rows, bar);
    };
    long long t = cluster_id;
    if (tid == 0 && t < p.n_tiles) issue(t);
    uint32_t phase = 0;
    const double* pcol = tile + rw * TILE_W + col;  // this thread's draws: pcol[k * 256]

    for (; t < p.n_tiles; t += n_clusters) {
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
        // per column: CTA minimum, q_t-th and q_l-th smallest bin minimum -> every CTA of the cluster
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * w + cc;
            const float sorted = warp_sort32_f(binsf[lane * TILE_W + c], lane);
            const float st = __shfl_sync(FULL, sorted, p.q_t - 1), sl = __shfl_sync(FULL, sorted, p.q_l - 1);
            double mn = cmin[(lane & 15) * TILE_W + c];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mn = min_sel(mn, __shfl_xor_sync(FULL, mn, o));
            if (lane < csize) {
                const uint32_t a = dsmem_addr(&xch[crank * TILE_W + c], (uint32_t)lane);
                dsmem_st_f64(a, mn);
                dsmem_st_v2f32(a + 8, st, sl);
            }
        }
        cluster_sync_all();  // exchange 1
        if (tid < TILE_W) {
            double mn = INF;
            for (int r = 0; r < csize; ++r) mn = min_sel(mn, xch[r * TILE_W + tid].mn);
            // lower / upper median of the CTAs' thresholds (ranked with the CTA index as tie-break)
            const int want_t = (csize - 1) / 2, want_l = csize / 2;
            float st = 0.f, sl = 0.f;
            for (int r = 0; r < csize; ++r) {
                const float vt = xch[r * TILE_W + tid].st, vl = xch[r * TILE_W + tid].sl;
                int rt = 0, rl = 0;
                for (int q = 0; q < csize; ++q) {
                    const float ot = xch[q * TILE_W + tid].st, ol = xch[q * TILE_W + tid].sl;
                    rt += (ot < vt || (ot == vt && q < r)) ? 1 : 0;
                    rl += (ol < vl || (ol == vl && q < r)) ? 1 : 0;
                }
                if (rt == want_t) st = vt;
                if (rl == want_l) sl = vl;
            }
            ColInfo ci;
            ci.llmin = mn;
            ci.tu_t = (double)st - mn;
            ci.tu_l = (double)sl - mn;
            ci.pad = 0.0;
            info[tid] = ci;
        }
        __syncthreads();

        // ---------------- pass B: sums over all draws, marks of the draws at or below the loose threshold
        const double llmin = info[col].llmin, tu_t = info[col].tu_t, tu_l = info[col].tu_l;
        double bs = 0.0, ls = 0.0, su = 0.0, suu = 0.0;
        unsigned mask = 0;
        int umax = 0;
        auto fold = [&](double v, unsigned bit, unsigned& msk) {
            const double u = v - llmin;  // = -x with x = fl(r - max r) (psis.py:134): exact, >= 0
            double ep, em;
            exp_pm128(u, etab, ep, em);
            msk |= (u <= tu_l) ? bit : 0u;
            bs += em;  // exp(x)
            ls += ep;  // exp(ll - min ll)
            su += u;
            suu = fma(u, u, suu);
            umax = max(umax, __double2hiint(u));
        };
        {
            int k = 0;
            for (; k + 4 <= kmin; k += 4) {
                unsigned nib = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) fold(pcol[(k + j) * 256], 1u << j, nib);
                mask |= nib << k;
            }
            for (; k < kmin; ++k) fold(pcol[k * 256], 1u << k, mask);
            if (extra) fold(pcol[kmin * 256], 1u << kmin, mask);
        }
        // tight candidates among the marked draws; counts of both kinds
        unsigned maskA = 0;
        for (unsigned m = mask; m; m &= m - 1) {
            const int b = __ffs((int)m) - 1;
            if (pcol[b * 256] - llmin <= tu_t) maskA |= 1u << b;
        }
        const unsigned nA = (unsigned)__popc(maskA), nB = (unsigned)__popc(mask) - nA;
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
        cntT[rw * TILE_W + col] = nA | (nB << 16);
        __syncthreads();
        if (tid < 64) {  // CTA partial of one (column, quantity) -> the column's owner
            const int c = tid & 15, q = tid >> 4;
            double a = 0.0;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) a += part[(ww * TILE_W + c) * 4 + q];
            dsmem_st_f64(dsmem_addr(&rsum[crank * nslot + c / csize].q[q], (uint32_t)(c % csize)), a);
        } else if (tid < 80) {
            const int c = tid - 64;
            int a = 0;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) a = max(a, umW[ww * TILE_W + c]);
            dsmem_st_u32(dsmem_addr(&rsum[crank * nslot + c / csize].umax, (uint32_t)(c % csize)), (uint32_t)a);
        } else if (tid >= 96 && tid < 112) {  // exclusive prefix of the packed counts over the 16 draw slots
            const int c = tid - 96;
            unsigned run = 0;
            for (int s = 0; s < 16; ++s) {
                const unsigned v = cntT[s * TILE_W + c];
                cntT[s * TILE_W + c] = run;
                run += v;
            }
            for (int r = 0; r < csize; ++r) dsmem_st_u32(dsmem_addr(&xcnt[crank * TILE_W + c], (uint32_t)r), run);
        }
        cluster_sync_all();  // exchange 2

        // ---------------- emit: fixed order (CTA rank, draw slot, draw)
        {
            unsigned before = 0, total = 0;
            for (int r = 0; r < csize; ++r) {
                const unsigned v = xcnt[r * TILE_W + col];
                total += v;
                before += (r < crank) ? v : 0u;
            }
            const int CA = (int)(total & 0xffffu), CB = (int)(total >> 16);
            const long long o = t * TILE_W + col;  // observation of the round
            if (mask && CA + CB <= cap && o < p.n_obs) {
                const unsigned pre = cntT[rw * TILE_W + col];
                int posA = (int)((before & 0xffffu) + (pre & 0xffffu));
                int posB = (int)((before >> 16) + (pre >> 16));
                double* dx = p.cx + (size_t)o * (size_t)cap;
                unsigned short* ds = p.cs + (size_t)o * (size_t)cap;
                for (unsigned m = mask; m; m &= m - 1) {
                    const int b = __ffs((int)m) - 1;
                    const int slot = ((maskA >> b) & 1u) ? posA++ : cap - 1 - posB++;
                    dx[slot] = llmin - pcol[b * 256];  // x = fl(r - max r), exactly
                    ds[slot] = (unsigned short)(row0 + 16 * b + rw);
                }
            }
            if (tid < nslot) {  // owner of column tid * csize + crank: the header the tail kernel reads
                const int c = tid * csize + crank;
                const long long oo = t * TILE_W + c;
                if (c < TILE_W && oo < p.n_obs) {
                    double q[4] = {0.0, 0.0, 0.0, 0.0};
                    int um = 0;
                    unsigned tot = 0;
                    for (int r = 0; r < csize; ++r) {
                        const RSum& e = rsum[r * nslot + tid];
                        q[0] += e.q[0]; q[1] += e.q[1]; q[2] += e.q[2]; q[3] += e.q[3];
                        um = max(um, e.umax);
                        tot += xcnt[r * TILE_W + c];
                    }
                    const int ca = (int)(tot & 0xffffu), cb = (int)(tot >> 16);
                    const bool special = !(is_finite(q[2]) && is_finite(q[3]) && is_finite(q[0]) && is_finite(q[1]));
                    const bool wide = um > WIDE_HI;
                    const bool count_bad = (ca + cb < M + 1) || (ca + cb > cap);
                    const bool ok = !special && !wide && !count_bad;
                    SplitHeader h;
                    h.mx = -info[c].llmin;
                    h.body = q[0];
                    h.lsum = q[1];
                    h.vsum = q[3] - q[2] * q[2] / (double)S;  // sum (ll - mean)^2 about the minimum: >= 0 up to rounding
                    h.lshift = info[c].llmin;
                    h.taux = -info[c].tu_l;
                    h.lse = 0.0;
                    h.C = ca; h.flags = ok ? 0 : 1; h.attempts = 0; h.n_patch = 0; h.C2 = cb; h.pad_ = 0;
                    if (h.vsum < 0.0) h.vsum = 0.0;
                    p.hdr[oo] = h;
                    if (!ok) {
                        p.fb_list[atomicAdd(p.fb_count, 1)] = (int)(p.row_base + oo);
                        if (p.counters) atomicAdd(&p.counters[3], 1ull);
                        note_handover(special ? HO_SPECIAL : (wide ? HO_RANGE : HO_RETRY));
                    }
                }
            }
        }
        __syncthreads();  // the tile has been read for the last time
        if (tid == 0 && t + n_clusters < p.n_tiles) {
            fence_proxy_async();
            issue(t + n_clusters);
        }
    }
}

// ---------------------------------------------------------------- host side
bool tile_shape(long long S, int M, int csize, TilePlan* tp) {
    memset(tp, 0, sizeof(*tp));
    if (csize < 1 || csize > TILE_MAXC) return false;
    if (S < 1024 || S > SPLIT_MAX_S) return false;
    const long long per = (S + csize - 1) / csize;
    if (per > TILE_MAX_R) return false;
    const int nbox = (int)((per + 255) / 256);
    const int box_rows = (int)((per + nbox - 1) / nbox);
    tp->csize = csize; tp->nbox = nbox; tp->box_rows = box_rows; tp->R = nbox * box_rows;
    if (tp->R > TILE_MAX_R) return false;
    // threshold ranks: 32 bins of R / 32 draws per CTA and column, i.e. 32 * csize bins per column (the split
    // path's balls-in-bins estimate, b2l_split.cuh); the loose rank sits three bins further down
    const double ratio = (double)(M + 1) / (32.0 * csize);
    if (ratio > 0.95) return false;
    int q = (int)std::lround(32.0 * (1.0 - std::exp(-1.25 * ratio))) + ((ratio > 0.45 && ratio <= 0.6) ? 1 : 0);
    if (const char* ev = getenv("B2L_TILE_QT")) q = atoi(ev);
    tp->q_t = std::min(29, std::max(1, q));
    int ql = tp->q_t + 3;
    if (const char* ev = getenv("B2L_TILE_QL")) ql = atoi(ev);
    tp->q_l = std::min(32, std::max(tp->q_t, ql));
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
