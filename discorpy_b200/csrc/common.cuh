// common.cuh -- PTX wrappers (mbarrier, TMA) and the fp64 helpers shared by the
// sm_100a kernels of libdiscorpy_b200.  Nothing in this file exists in the
// reference (pure Python); see DESIGN.md for the kernel design.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

namespace dcb {

// ---------------------------------------------------------------------------
// shared-memory / mbarrier / TMA primitives (inline PTX; SASS: SYNCS.*, UTMALDG)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}

// make mbarrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

// one arrival (release) on an mbarrier of this CTA
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Blocking wait on a phase parity.  try_wait suspends in hardware for a
// bounded time; the outer loop is bounded too so that a mis-programmed copy
// traps instead of hanging the GPU.
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)   // suspend-time hint (ns): fewer polls per wait
        : "memory");
    return done != 0;
}
// (round 1 measured no difference from a __nanosleep back-off or the suspend-time hint; with the
// leaner round-2 sampling loop the ~10 polls per wait were 5 % of all issued instructions, so the
// hint is on)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        if (mbar_try(addr, parity)) return;
    }
    __trap();
}

// 3-D tiled TMA load global -> shared, completion signalled on an mbarrier.
// Coordinates are in elements, innermost first; out-of-bounds elements are
// zero-filled (they are never sampled: every tap index is clamped).
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y,
                                            int z, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

// The same load with an L2 eviction-priority hint (createpolicy: 0 evict_first, 1 evict_last).
__device__ __forceinline__ uint64_t l2_policy(int evict_last) {
    uint64_t pol;
    if (evict_last)
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(void *smem_dst, const CUtensorMap *map, int x, int y,
                                                 int z, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// The same box into L2 only (no shared-memory destination, no completion to wait for).
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap *map, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(x), "r"(y), "r"(z)
                 : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---------------------------------------------------------------------------
// fp64 helpers
// ---------------------------------------------------------------------------

// sqrt(s) for s >= 0, correctly rounded for all but ~2^-35 of inputs
// (dcb_selftest_sqrt measures it): MUFU.RSQ64H seed (~2^-22), one coupled
// Goldschmidt step (-> ~2^-43) and one Newton correction on the residual
// (-> ~2^-87 before the final rounding).  7 fp64 ops + 1 MUFU; no slow path.
// s == 0 (the pixel on an integer distortion centre) and subnormal s return 0.
__device__ __forceinline__ double dsqrt_pos(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double g = s * y;
    double h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    double d = fma(-g, g, s);
    g = fma(d, h, g);
    return (__double2hiint(s) < 0x00100000) ? 0.0 : g;
}

// Horner evaluation of N coefficients held in kernel-parameter (constant) space.
template <int N>
__device__ __forceinline__ double horner(const double *a, double r) {
    double f = a[N - 1];
#pragma unroll
    for (int i = N - 2; i >= 0; --i) f = fma(f, r, a[i]);
    return f;
}

// F(r) for NPX independent radii at once; the switch is warp-uniform and
// outside the per-pixel chains so that the NPX chains interleave (ILP).
template <int NPX>
__device__ __forceinline__ void radial_factor(const double *a, int n, const double (&r)[NPX],
                                              double (&f)[NPX]) {
#define DCB_CASE(N)                                                  \
    case N:                                                          \
        _Pragma("unroll") for (int k = 0; k < NPX; ++k) f[k] = horner<N>(a, r[k]); \
        break;
    switch (n) {
        DCB_CASE(1) DCB_CASE(2) DCB_CASE(3) DCB_CASE(4) DCB_CASE(5) DCB_CASE(6) DCB_CASE(7)
        DCB_CASE(8) DCB_CASE(9) DCB_CASE(10) DCB_CASE(11) DCB_CASE(12) DCB_CASE(13) DCB_CASE(14)
        DCB_CASE(15) DCB_CASE(16)
        default:
#pragma unroll
            for (int k = 0; k < NPX; ++k) f[k] = 0.0;
    }
#undef DCB_CASE
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

}  // namespace dcb
