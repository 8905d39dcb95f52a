// remap.cuh -- shared pieces of the backward-remap kernels of libdiscorpy_b200
// (sm_100a): parameter blocks, tap fetchers, the per-pixel order-0/1 arithmetic
// of scipy.ndimage.map_coordinates for pre-clipped coordinates (used by the
// direct-gather paths), the caller-supplied-coordinates kernel and the
// synthetic input generator.  The tiled kernels live in remap_image.cuh (one
// image: a1, a4 of SURVEY.md section 8a) and remap_stack.cuh (Z-stacks: a2, a3).
// Numerics restate discorpy/post/postprocessing.py:138-147 / :214-228 /
// :302-312 / :448-457 -- see DESIGN.md "Numerics".
#pragma once
#include <type_traits>
#include <limits.h>
#include "common.cuh"
#include "../../include/discorpy_b200.h"

namespace dcb {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kCols = 4;            // columns per thread
constexpr int kTileW = 32 * kCols;  // 128 output pixels per tile row

enum { MAP_RADIAL = 0, MAP_PERSP = 1 };

struct RadialDev {
    double xc, yc;
    double a[DCB_MAX_TERMS];
    int n, pad;
};
struct PerspDev {
    double c[8];
};

struct RemapParams {
    const float *src;
    float *dst;
    long long src_pitch, src_slice, dst_pitch, dst_slice;  // in elements
    int H, W, D;
    int row0, nrows;
    int yorg, ylast;  // image rows held by src: yorg .. ylast (window); src points at row yorg
    int tiles_x, tiles_y, zchunk, ntiles;
    int bw, bh, nstage;  // staged box; nstage == 0 => direct gathers only
    unsigned stage_bytes, box_bytes;
    int rint, pad;       // 1: integer image, round half away from zero (see finish_f64)
    unsigned *sched;     // [0] next work item - gridDim.x, [1] CTAs done (the last one resets both)
    RadialDev rad;
    PerspDev per;
};

// ---------------------------------------------------------------------------
// tap fetchers
// ---------------------------------------------------------------------------
struct SmemFetch {
    const float *tile;  // staged box, row-major, bw floats per row
    int bw;
    int off;  // -(by0 * bw + bx0): image coordinates -> box index
    __device__ __forceinline__ float operator()(int y, int x) const {
        return tile[y * bw + x + off];
    }
};
struct GlobalFetch {
    const float *slice;
    long long pitch;
    __device__ __forceinline__ float operator()(int y, int x) const {
        return __ldg(slice + (long long)y * pitch + x);
    }
};

// ---------------------------------------------------------------------------
// output conversion.  float32 images: one IEEE rounding of the fp64 sum.
// Integer images (uint8/16, int8/16 travel as float32, exactly): SciPy adds
// +-0.5 to the fp64 sum and truncates ("round half away from zero"); doing it
// here, before the value ever becomes a float32, avoids a double rounding.
// `rint` is a kernel parameter (warp-uniform).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double round_half_away(double s) {
    return trunc(s + (s >= 0.0 ? 0.5 : -0.5));
}
__device__ __forceinline__ float finish_f64(double s, int rint) {
    return __double2float_rn(rint ? round_half_away(s) : s);
}
__device__ __forceinline__ float finish_f32(float v, int rint) {
    return rint ? truncf(v + (v >= 0.0f ? 0.5f : -0.5f)) : v;
}

// floor and fraction of a coordinate in [0, 2^23) without the XU pipe (no F2I / I2F): a
// round-down add of 2^23 (2^52 for float64 coordinates) leaves floor(c) in the low mantissa
// bits; both subtractions are exact
__device__ __forceinline__ void split_floor(float c, int &i, float &t) {
    const float m = __fadd_rd(c, 8388608.0f);
    i = __float_as_int(m) - 0x4B000000;
    t = c - (m - 8388608.0f);
}
__device__ __forceinline__ void split_floor(double c, int &i, double &t) {
    const double m = __dadd_rd(c, 4503599627370496.0);
    i = __double2loint(m);
    t = __dsub_rn(c, __dsub_rn(m, 4503599627370496.0));
}

// SciPy's order-1 value  sum_ij  rn(rn(m_ij * wy_i) * wx_j)  accumulated first
// tap to last (DCB_BLEND_EXACT), with fewer operations but the same roundings:
//   * wy_i and wx_j are fp32 fractions widened to fp64 (<= 24 significant
//     bits), so m * wy (24 + 24 bits) is exact and rn(rn(m wy) wx) equals
//     rn(m * W_ij) with W_ij = wy_i * wx_j, itself exact (<= 48 bits);
//   * W11 = ty*tx is one multiplication; W10 = ty - W11, W01 = tx - W11 and
//     W00 = (1 - ty) - W01 are exact because their results are the <= 48-bit
//     products ty(1-tx), (1-ty)tx, (1-ty)(1-tx).
// 5 + 4 + 3 = 12 fp64 operations instead of 13, every one of them rounding
// exactly where SciPy's does.
__device__ __forceinline__ double blend_exact(double a, double b, double c, double d, double tx,
                                              double ty) {
    const double w11 = __dmul_rn(ty, tx);
    const double w10 = __dsub_rn(ty, w11);
    const double w01 = __dsub_rn(tx, w11);
    const double w00 = __dsub_rn(__dsub_rn(1.0, ty), w01);
    double s = __dmul_rn(a, w00);
    s = __dadd_rn(s, __dmul_rn(b, w01));
    s = __dadd_rn(s, __dmul_rn(c, w10));
    s = __dadd_rn(s, __dmul_rn(d, w11));
    return s;
}

// ---------------------------------------------------------------------------
// certified blend (round 2): a cheaper evaluation of SciPy's order-1 sum plus one integer test
// that proves the float32 result is the same
// ---------------------------------------------------------------------------
// distance test of a double's low mantissa word from the float32 rounding boundary (bit 28 set,
// bits 27..0 clear): the shifted word is 0x80000000 there
__device__ __forceinline__ uint32_t cert_key(double v, uint32_t add) {
    return ((uint32_t)__double2loint(v) << 3) + add;
}
// blend certificate: 32 ulp64 around the boundary (the FMA form below is within 10 ulp64 of
// SciPy's sum for taps of one sign, see lerp_fma)
constexpr uint32_t kBlendCertAdd = 0x80000000u + 8u * 32u, kBlendCertLim = 16u * 32u;

// (1-ty)((1-tx) a + tx b) + ty ((1-tx) c + tx d) with six FMAs, every intermediate a positive
// combination of the taps: for taps of one sign nothing cancels, the result is within 2^-51
// relative of the exact value, SciPy's rn-sum of rn-products (blend_exact) within 2^-51 too.
__device__ __forceinline__ double lerp_fma(double a, double b, double c, double d, double tx,
                                           double ty) {
    const double top = fma(tx, b, fma(-tx, a, a));
    const double bot = fma(tx, d, fma(-tx, c, c));
    return fma(ty, bot, fma(-ty, top, top));
}

// ---------------------------------------------------------------------------
// one output pixel: the arithmetic of scipy.ndimage.map_coordinates(order 0|1)
// for a coordinate that already lies in [0, W-1] x [0, H-1]
// ---------------------------------------------------------------------------
template <int ORDER, int BLEND, class CT, class Fetch>
__device__ __forceinline__ float sample_px(const Fetch &fetch, CT x, CT y, int wmax, int ylo,
                                           int yhi, int rint = 0) {
    int x0, y0;
    CT tx, ty;
    split_floor(x, x0, tx);
    split_floor(y, y0, ty);
    if (ORDER == 0) {
        // SciPy: floor(c + 0.5) evaluated in double; the fractional part of a
        // float is exact, so comparing it with 0.5 is the same decision.
        if (tx >= (CT)0.5) ++x0;
        if (ty >= (CT)0.5) ++y0;
        return fetch(min(max(y0, ylo), yhi), x0);
    }
    const int x1 = min(x0 + 1, wmax);
    // rows are clamped into the window the caller holds (a no-op for whole
    // images); the +1 tap folds back onto the last row like SciPy's 'reflect'
    const int y1 = min(max(y0 + 1, ylo), yhi);
    y0 = min(max(y0, ylo), yhi);
    const float a = fetch(y0, x0);
    const float b = fetch(y0, x1);
    const float c = fetch(y1, x0);
    const float d = fetch(y1, x1);
    if (BLEND == DCB_BLEND_LERP32) {
        const float ftx = (float)tx, fty = (float)ty;
        const float top = fmaf(b - a, ftx, a);
        const float bot = fmaf(d - c, ftx, c);
        return finish_f32(fmaf(bot - top, fty, top), rint);
    } else if (BLEND == DCB_BLEND_LERP64) {
        const double dtx = (double)tx, dty = (double)ty;
        const double da = a, db = b, dc = c, dd = d;
        const double top = fma(db - da, dtx, da);
        const double bot = fma(dd - dc, dtx, dc);
        return finish_f64(fma(bot - top, dty, top), rint);
    } else {
        // SciPy's order: each tap times its y weight, then its x weight, the
        // four products summed first to last, every step rounded (no FMA).
        const double wx1 = (double)tx, wy1 = (double)ty;
        const double wx0 = __dsub_rn(1.0, wx1), wy0 = __dsub_rn(1.0, wy1);
        double t = __dmul_rn(__dmul_rn((double)a, wy0), wx0);
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)b, wy0), wx1));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)c, wy1), wx0));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)d, wy1), wx1));
        return finish_f64(t, rint);
    }
}

template <class CT>
__device__ __forceinline__ CT clamp_coord(double v, int vmax);
template <>
__device__ __forceinline__ float clamp_coord<float>(double v, int vmax) {
    // round to fp32 first, clip second: identical to the reference's
    // clip-then-round because rounding is monotone and 0 / vmax are exact.
    return fminf(fmaxf(__double2float_rn(v), 0.0f), (float)vmax);
}
template <>
__device__ __forceinline__ double clamp_coord<double>(double v, int vmax) {
    return fmin(fmax(v, 0.0), (double)vmax);
}

// ---------------------------------------------------------------------------
// caller-supplied coordinates (map_index= / _mapping): one thread per output
// ---------------------------------------------------------------------------
template <int ORDER, int BLEND, class CT>
__global__ void __launch_bounds__(256)
    map_coords_kernel(const float *__restrict__ src, float *__restrict__ dst, int H, int W,
                      long long pitch, const CT *__restrict__ yd, const CT *__restrict__ xd,
                      size_t n, unsigned *oob_count, int rint) {
    unsigned oob = 0;
    GlobalFetch fetch{src, pitch};
    const CT xmax = (CT)(W - 1), ymax = (CT)(H - 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        CT x = xd[i], y = yd[i];
        oob += !(x >= (CT)0 && x <= xmax && y >= (CT)0 && y <= ymax);
        x = x > (CT)0 ? x : (CT)0;  // NaN -> 0
        y = y > (CT)0 ? y : (CT)0;
        x = x < xmax ? x : xmax;
        y = y < ymax ? y : ymax;
        dst[i] = sample_px<ORDER, BLEND, CT>(fetch, x, y, W - 1, 0, H - 1, rint);
    }
    if (oob_count != nullptr && oob != 0) atomicAdd(oob_count, oob);
}

// ---------------------------------------------------------------------------
// synthetic input generator (bench only): splitmix64 counter hash -> [0,1)
// ---------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
    fill_synthetic_kernel(float *__restrict__ dst, size_t n, uint64_t seed, uint64_t offset) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t h = splitmix64(seed ^ (offset + i));
        dst[i] = (float)(h >> 40) * (1.0f / 16777216.0f);
    }
}

}  // namespace dcb
