// microbench.cuh -- pipe-rate probes and the sqrt self-test (diagnostics only).
// They size the fp64 budget of the radial kernel (DESIGN.md "fp64 budget").
#pragma once
#include "common.cuh"

namespace dcb {

constexpr int kMbIters = 2048;
constexpr int kMbChains = 8;

// which: 0 DFMA, 1 F2F f32<->f64 pair, 2 MUFU.RSQ64H, 4 FFMA, 5 DFMA + F2F mixed
template <int WHICH>
__global__ void __launch_bounds__(256) microbench_kernel(double *sink, double seed) {
    double d[kMbChains];
    float f[kMbChains];
#pragma unroll
    for (int c = 0; c < kMbChains; ++c) {
        d[c] = seed + c + threadIdx.x * 1e-3;
        f[c] = (float)d[c];
    }
    for (int it = 0; it < kMbIters; ++it) {
#pragma unroll
        for (int c = 0; c < kMbChains; ++c) {
            if (WHICH == 0) {
                d[c] = fma(d[c], 0.999999, 1e-7);
            } else if (WHICH == 1) {
                double w;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w) : "f"(f[c]));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(w));
            } else if (WHICH == 2) {
                asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(d[c]) : "d"(d[c]));
            } else if (WHICH == 4) {
                f[c] = fmaf(f[c], 0.999999f, 1e-7f);
            } else if (WHICH == 10) {
                // DFMA with three register operands (what the kernels issue)
                d[c] = fma(d[c], d[(c + 3) % kMbChains], d[(c + 5) % kMbChains]);
            } else if (WHICH == 20) {   // DFMA, two registers + constant (Horner step)
                d[c] = fma(d[c], d[(c + 3) % kMbChains], 1e-7);
            } else if (WHICH == 21) {   // DMUL, two registers
                d[c] = __dmul_rn(d[c], d[(c + 3) % kMbChains]);
            } else if (WHICH == 22) {   // DADD, two registers
                d[c] = __dadd_rn(d[c], d[(c + 3) % kMbChains]);
            } else if (WHICH == 23) {   // DFMA, two distinct registers, one used twice
                d[c] = fma(d[c], d[(c + 3) % kMbChains], d[c]);
            } else if (WHICH == 30) {   // 2 widenings (of values that change) + 1 DFMA + 2 FFMA
                f[c] = fmaf(f[c], 0.999999f, 1e-7f);
                const float g = fmaf(f[c], 0.5f, 0.25f);
                double w1, w2;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w1) : "f"(f[c]));
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w2) : "f"(g));
                d[c] = fma(w1, w2, d[c]);
            } else if (WHICH == 31) {   // 2 narrowings (of values that change) + 2 DFMA + 1 FADD
                d[c] = fma(d[c], 0.999999, 1e-7);
                const double e = fma(d[c], 0.5, 0.25);
                float n1, n2;
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(n1) : "d"(d[c]));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(n2) : "d"(e));
                f[c] += n1 * n2;
            } else if (WHICH == 32) {   // 1 widening + 1 DFMA + 1 FFMA
                f[c] = fmaf(f[c], 0.999999f, 1e-7f);
                double w;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w) : "f"(f[c]));
                d[c] = fma(d[c], 0.999999, w);
            } else if (WHICH == 33) {   // 1 narrowing + 2 DFMA
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(d[c]));
                d[c] = fma(d[c], 0.999999, 1e-7);
                d[(c + 1) % kMbChains] = fma(d[(c + 1) % kMbChains], 0.999999, 1e-7);
            } else if (WHICH == 34) {   // balanced: 1 widening (8 XU cycles) + 4 DFMA (8 fp64 cycles)
                f[c] = fmaf(f[c], 0.999999f, 1e-7f);
                double w;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w) : "f"(f[c]));
                d[c] = fma(d[c], 0.999999, w);
                d[(c + 1) % kMbChains] = fma(d[(c + 1) % kMbChains], 0.999999, 1e-7);
                d[(c + 2) % kMbChains] = fma(d[(c + 2) % kMbChains], 0.999999, 1e-7);
                d[(c + 3) % kMbChains] = fma(d[(c + 3) % kMbChains], 0.999999, 1e-7);
            } else if (WHICH == 35) {   // balanced: 1 narrowing + 4 DFMA
                float n1;
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(n1) : "d"(d[c]));
                f[c] += n1;
                d[c] = fma(d[c], 0.999999, 1e-7);
                d[(c + 1) % kMbChains] = fma(d[(c + 1) % kMbChains], 0.999999, 1e-7);
                d[(c + 2) % kMbChains] = fma(d[(c + 2) % kMbChains], 0.999999, 1e-7);
                d[(c + 3) % kMbChains] = fma(d[(c + 3) % kMbChains], 0.999999, 1e-7);
            } else if (WHICH == 36) {   // 4 DFMA only, same shape (the fp64 side of 34/35 alone)
                d[c] = fma(d[c], 0.999999, 1e-7);
                d[(c + 1) % kMbChains] = fma(d[(c + 1) % kMbChains], 0.999999, 1e-7);
                d[(c + 2) % kMbChains] = fma(d[(c + 2) % kMbChains], 0.999999, 1e-7);
                d[(c + 3) % kMbChains] = fma(d[(c + 3) % kMbChains], 0.999999, 1e-7);
            } else if (WHICH == 5) {
                d[c] = fma(d[c], 0.999999, 1e-7);
                double w;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w) : "f"(f[c]));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(w));
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int c = 0; c < kMbChains; ++c) acc += d[c] + f[c];
    if (acc == 123.456) sink[0] = acc;
}

// which 24..29: does a slow-pipe instruction block the issue port?  One group per
// iteration, all operations independent across 8 chains; the host reports the
// cycles one warp-level group costs an SM sub-partition.
//   24: 16 FFMA                       25: 2 F2F (f32->f64->f32) + 16 FFMA
//   26: 1 MUFU.RSQ64H + 8 FFMA        27: 4 DFMA + 8 FFMA
//   28: 2 F2F + 8 DFMA                29: 2 F2F alone
template <int WHICH>
__global__ void __launch_bounds__(256) microbench_mix_kernel(double *sink, double seed) {
    float f[16];
    double d[8];
    float g = (float)seed + threadIdx.x * 1e-3f;
#pragma unroll
    for (int c = 0; c < 16; ++c) f[c] = (float)seed + c + threadIdx.x * 1e-3f;
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c] = seed + c + threadIdx.x * 1e-3;
    for (int it = 0; it < kMbIters; ++it) {
        if (WHICH == 24 || WHICH == 25) {
#pragma unroll
            for (int c = 0; c < 16; ++c) f[c] = fmaf(f[c], 0.999999f, 1e-7f);
        }
        if (WHICH == 26 || WHICH == 27) {
#pragma unroll
            for (int c = 0; c < 8; ++c) f[c] = fmaf(f[c], 0.999999f, 1e-7f);
        }
        if (WHICH == 25 || WHICH == 28 || WHICH == 29) {
            double w;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w) : "f"(g));
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(g) : "d"(w));
        }
        if (WHICH == 26) asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(d[0]) : "d"(d[0]));
        if (WHICH == 27) {
#pragma unroll
            for (int c = 0; c < 4; ++c) d[c] = fma(d[c], 0.999999, 1e-7);
        }
        if (WHICH == 28) {
#pragma unroll
            for (int c = 0; c < 8; ++c) d[c] = fma(d[c], 0.999999, 1e-7);
        }
    }
    double acc = g;
#pragma unroll
    for (int c = 0; c < 16; ++c) acc += f[c];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc += d[c];
    if (acc == 123.456) sink[0] = acc;
}

// which 6..9: dependent-issue latency in cycles of DFMA (6), an F2F f32->f64->f32
// round trip (7, two conversions), MUFU.RSQ64H (8) and LDS.64 (9): one warp, one
// dependent chain, clock64 around it.
template <int WHICH>
__global__ void microbench_latency_kernel(double *out) {
    __shared__ double sm[64];
    sm[threadIdx.x] = 1.0 + threadIdx.x * 1e-3;
    sm[threadIdx.x + 32] = 0.0;
    __syncwarp();
    double d = 1.0 + threadIdx.x * 1e-3;
    float f = (float)d;
    int idx = threadIdx.x;
    const int iters = 4096;
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < iters; ++it) {
        if (WHICH == 6) {
            d = fma(d, 0.999999, 1e-7);
        } else if (WHICH == 7) {
            double w;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(w) : "f"(f));
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(w));
        } else if (WHICH == 8) {
            asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(d) : "d"(d));
        } else {
            d = sm[idx];
            idx = (int)__double2hiint(d) & 31;  // stays in range; value-dependent address
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / iters;
    if (d + f + idx == 123.456) out[1] = d;
}

// which == 3: the radial coordinate evaluation alone, 5 terms, 4 px per step
__global__ void __launch_bounds__(256) microbench_coords_kernel(double *sink, double xc, double yc) {
    const double a[5] = {1.0, -2e-5, 6e-8, -1e-10, 5e-14};
    double xu[4], xu2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        xu[k] = (double)(threadIdx.x + 32 * k + blockIdx.x) - xc;
        xu2[k] = xu[k] * xu[k];
    }
    float acc = 0.f;
    for (int it = 0; it < kMbIters; ++it) {
        const double yu = (double)it - yc;
        const double yu2 = yu * yu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double r = dsqrt_pos(xu2[k] + yu2);
            double f = a[4];
            f = fma(f, r, a[3]);
            f = fma(f, r, a[2]);
            f = fma(f, r, a[1]);
            f = fma(f, r, a[0]);
            acc += __double2float_rn(fma(f, xu[k], xc)) + __double2float_rn(fma(f, yu, yc));
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

// custom sqrt vs IEEE sqrt on pseudo-random positive inputs spanning the
// magnitudes r^2 takes on images up to 65536^2 (and a few exact squares)
__global__ void __launch_bounds__(256)
    selftest_sqrt_kernel(size_t n, uint64_t seed, unsigned long long *mismatch) {
    unsigned long long bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t h = splitmix64(seed ^ i);
        double s;
        if ((i & 7) == 0) {
            const double q = (double)(h >> 44);  // exact squares
            s = q * q;
        } else {
            const double m = (double)(h >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
            const int e = (int)((h & 0x3f)) - 20;                             // 2^-20 .. 2^43
            s = ldexp(1.0 + m, e);
        }
        if (dsqrt_pos(s) != sqrt(s)) ++bad;
    }
    if (bad) atomicAdd(mismatch, bad);
}

// the image kernel's 5-operation sqrt (dsqrt_nz, remap_image.cuh) on the same
// inputs: out[0] = results that differ from IEEE sqrt, out[1] = results more
// than one ulp away (must be 0)
__global__ void __launch_bounds__(256)
    selftest_sqrt_fast_kernel(size_t n, uint64_t seed, unsigned long long *out) {
    unsigned long long diff = 0, bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t h = splitmix64(seed ^ i);
        double s;
        if ((i & 7) == 0) {
            const double q = (double)(h >> 44) + 1.0;  // exact squares
            s = q * q;
        } else {
            const double m = (double)(h >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
            const int e = (int)((h & 0x3f)) - 20;                             // 2^-20 .. 2^43
            s = ldexp(1.0 + m, e);
        }
        const long long a = __double_as_longlong(dsqrt_nz(s));
        const long long b = __double_as_longlong(sqrt(s));
        if (a != b) ++diff;
        if (a - b > 1 || b - a > 1) ++bad;
    }
    if (diff) atomicAdd(&out[0], diff);
    if (bad) atomicAdd(&out[1], bad);
}

// TMA probe: load one (bw x bh) box at element coordinates (x0, y0, z0) of a
// (D, H, W) float tensor into shared memory and copy it out.  status[0] = 0 ok,
// 1 = the mbarrier never completed within ~0.2 s (no trap, so the context
// survives and the host can report which geometry failed).
__global__ void __launch_bounds__(256)
    selftest_tma_kernel(const __grid_constant__ CUtensorMap tmap, int bw, int bh, int x0, int y0,
                        int z0, float *out, int *status) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)(bw * bh * 4));
        tma_load_3d(smem, &tmap, x0, y0, z0, &bar);
    }
    const long long t0 = clock64();
    bool ok = false;
    while (clock64() - t0 < 400000000LL) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(&bar)), "r"(0u)
            : "memory");
        if (done) {
            ok = true;
            break;
        }
    }
    if (!ok) {
        if (threadIdx.x == 0) status[0] = 1;
        return;
    }
    const float *tile = reinterpret_cast<const float *>(smem);
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

}  // namespace dcb
