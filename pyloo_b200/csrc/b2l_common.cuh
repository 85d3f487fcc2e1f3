// Common device helpers for the B200 PSIS-LOO kernels (sm_100a).
//   * order-preserving 64-bit image of a double (radix/bitonic selection works on integers)
//   * deterministic block reductions (fixed association order => batch-invariant results)
//   * mbarrier + 1-D bulk-TMA (cp.async.bulk) wrappers for row staging, 2-D TMA tensor loads
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b2l {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------- ordered keys
// Monotone map double -> uint64 (total order, -inf < ... < -0 < +0 < ... < +inf).  NaNs are
// never keyed (rows holding NaN are short-circuited, pyloo/psis.py:134-144 semantics).
__device__ __forceinline__ uint64_t key_of(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double val_of(uint64_t k) {
    uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// ---------------------------------------------------------------- reductions
// Butterfly reductions: both partners of every exchange add the same two numbers, so all lanes
// end bit-identical and the association order is fixed (deterministic, batch-invariant).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_isum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// `red` = shared scratch of 128 doubles.  Every thread returns the same value.
// Two levels: warp butterfly -> one word per warp -> warp 0 butterfly -> broadcast word.
#define B2L_BLOCK_REDUCE(NAME, TYPE, WARPOP, IDENT, OFFS)                             \
    template <int NT>                                                                 \
    __device__ __forceinline__ TYPE NAME(TYPE v, double* red_) {                      \
        TYPE* red = reinterpret_cast<TYPE*>(red_ + (OFFS));                           \
        constexpr int NW = NT / 32;                                                   \
        v = WARPOP(v);                                                                \
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;                       \
        __syncthreads();                                                              \
        if (threadIdx.x < 32) {                                                       \
            TYPE t = (threadIdx.x < NW) ? red[threadIdx.x] : (IDENT);                 \
            t = WARPOP(t);                                                            \
            if (threadIdx.x == 0) red[NW] = t;                                        \
        }                                                                             \
        __syncthreads();                                                              \
        return red[NW];                                                               \
    }
B2L_BLOCK_REDUCE(block_sum, double, warp_sum, 0.0, 0)
B2L_BLOCK_REDUCE(block_max, double, warp_max, -__longlong_as_double(0x7ff0000000000000ll), 0)
B2L_BLOCK_REDUCE(block_min, double, warp_min, __longlong_as_double(0x7ff0000000000000ll), 0)
B2L_BLOCK_REDUCE(block_isum, int, warp_isum, 0, 64)  // ints live in their own half of `red`

// ---------------------------------------------------------------- mbarrier / bulk TMA
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, 1-D, bytes % 16 == 0, both addresses 16 B aligned.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// pull a 1-D range into L2 ahead of the bulk load that will need it (no shared-memory destination)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// shared -> global, 1-D bulk store (bulk_group completion).
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// 2-D tiled TMA load (tensor map in param/const space).  SASS: UTMALDG.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------- misc
__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
__device__ __forceinline__ bool is_finite(double x) {
    return (((uint64_t)__double_as_longlong(x) >> 52) & 0x7ff) != 0x7ff;
}
__device__ __forceinline__ double nan_f64() { return __longlong_as_double(0x7ff8000000000000ll); }
__device__ __forceinline__ double inf_f64() { return __longlong_as_double(0x7ff0000000000000ll); }

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
// exp(x) for finite x <= 0 with the result forced to (almost) zero when `drop`: the exponent
// argument is clamped so that any finite x gives a finite, tiny value (no guard needed), and a
// dropped term keeps only its low word (< 2^-1022): one SEL instead of a predicated FP64 add
__device__ __forceinline__ double exp_tab_drop(double x, const ExpTab& tb, bool drop) {
    const double t = fma(x, EXP_L, EXP_MAGIC);
    const int ni = max(__double2loint(t), -1022 * 32);
    const double nf = t - EXP_MAGIC;
    const double r = fma(nf, EXP_C1, x);
    const double y = tb.t[ni & 31] * exp_poly5(r);
    const int hi = __double2hiint(y) + ((ni & ~31) << 15);
    return __hiloint2double(drop ? 0 : hi, __double2loint(y));
}

}  // namespace b2l
