// remap.cuh -- the backward-remap kernels of libdiscorpy_b200 (sm_100a).
//
// One kernel template serves every entry of the hot path:
//   a1 unwarp_image_backward          MAP_RADIAL, ROUND32, D = 1
//   a2 unwarp_slice_backward          MAP_RADIAL, !ROUND32 (fp64 coordinates), nrows = 1
//   a3 unwarp_chunk_slices_backward   MAP_RADIAL, ROUND32, rows start..stop, D slices
//   a4 correct_perspective_image      MAP_PERSP,  ROUND32, D = 1
// It restates, per output pixel, discorpy/post/postprocessing.py:138-147 /
// :214-228 / :302-312 / :448-457 plus the order-0/1 arithmetic of
// scipy.ndimage.map_coordinates -- see DESIGN.md "Numerics" for the exact
// operation order that is kept and why.
//
// Work decomposition (B200-first, nothing like it in the reference):
//   * a CTA of 256 threads owns an output tile of 128 x (8*RPT) pixels for a
//     chunk of Z slices; lane l of warp w owns columns x0+l+32k (k<4) of rows
//     w*RPT..w*RPT+RPT-1, so every global store is one full 128-byte line.
//   * the fp64 coordinate evaluation happens once per tile and stays in
//     registers for all slices of the chunk.
//   * the tile's source bounding box is reduced with redux.sync + one smem
//     exchange; if it fits the staged box, one elected thread issues a 3-D
//     TMA load (cp.async.bulk.tensor) per slice into an mbarrier-guarded
//     shared-memory ring and the four taps are read from shared memory;
//     otherwise (strong magnification, e.g. BASELINE config 1) that tile
//     gathers straight from global memory through the read-only path.
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
constexpr int kMaxStages = 4;

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
// one output pixel: the arithmetic of scipy.ndimage.map_coordinates(order 0|1)
// for a coordinate that already lies in [0, W-1] x [0, H-1]
// ---------------------------------------------------------------------------
template <int ORDER, int BLEND, class CT, class Fetch>
__device__ __forceinline__ float sample_px(const Fetch &fetch, CT x, CT y, int wmax, int ylo,
                                           int yhi) {
    int x0 = (int)x;  // truncation == floor, coordinates are >= 0
    int y0 = (int)y;
    const CT tx = x - (CT)x0;  // exact
    const CT ty = y - (CT)y0;
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
        return fmaf(bot - top, fty, top);
    } else if (BLEND == DCB_BLEND_LERP64) {
        const double dtx = (double)tx, dty = (double)ty;
        const double da = a, db = b, dc = c, dd = d;
        const double top = fma(db - da, dtx, da);
        const double bot = fma(dd - dc, dtx, dc);
        return (float)fma(bot - top, dty, top);
    } else {
        // SciPy's order: each tap times its y weight, then its x weight, the
        // four products summed first to last, every step rounded (no FMA).
        const double wx1 = (double)tx, wy1 = (double)ty;
        const double wx0 = __dsub_rn(1.0, wx1), wy0 = __dsub_rn(1.0, wy1);
        double t = __dmul_rn(__dmul_rn((double)a, wy0), wx0);
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)b, wy0), wx1));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)c, wy1), wx0));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)d, wy1), wx1));
        return __double2float_rn(t);
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
// the tile kernel
// ---------------------------------------------------------------------------
template <int MAP, int ORDER, int BLEND, bool ROUND32, int RPT>
__global__ void __launch_bounds__(kThreads, (RPT >= 4 ? 2 : 3))
    remap_tile_kernel(const __grid_constant__ RemapParams p,
                      const __grid_constant__ CUtensorMap tmap) {
    using CT = typename std::conditional<ROUND32, float, double>::type;
    constexpr int TH = kWarps * RPT;

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.nstage * p.stage_bytes);
    int *red = reinterpret_cast<int *>(full + kMaxStages);  // [2][4][kWarps]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool staged = p.nstage > 0;
    if (staged && threadIdx.x == 0) {
        for (int s = 0; s < p.nstage; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    uint32_t fills = 0;  // slices that went through the ring so far (CTA-uniform)
    int tile_par = 0;
    const int wmax = p.W - 1;
    const int y_end = p.row0 + p.nrows;

    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, tile_par ^= 1) {
        const int txi = t % p.tiles_x;
        const int rest = t / p.tiles_x;
        const int tyi = rest % p.tiles_y;
        const int zci = rest / p.tiles_y;
        const int x_base = txi * kTileW + lane;
        const int y_base = p.row0 + tyi * TH + warp * RPT;
        const int z0 = zci * p.zchunk;
        const int nz = min(p.zchunk, p.D - z0);

        // ---- coordinates: fp64, once per tile --------------------------------
        CT cx[RPT][kCols], cy[RPT][kCols];
        if (MAP == MAP_RADIAL) {
            double xu[kCols], xu2[kCols];
#pragma unroll
            for (int k = 0; k < kCols; ++k) {
                xu[k] = (double)(x_base + 32 * k) - p.rad.xc;  // :138
                xu2[k] = __dmul_rn(xu[k], xu[k]);
            }
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const double yu = (double)(y_base + j) - p.rad.yc;  // :139
                const double yu2 = __dmul_rn(yu, yu);
                double r[kCols], f[kCols];
#pragma unroll
                for (int k = 0; k < kCols; ++k) r[k] = dsqrt_pos(__dadd_rn(xu2[k], yu2));  // :141
                radial_factor<kCols>(p.rad.a, p.rad.n, r, f);  // :142-143
#pragma unroll
                for (int k = 0; k < kCols; ++k) {  // :144-145
                    cx[j][k] = clamp_coord<CT>(fma(f[k], xu[k], p.rad.xc), wmax);
                    cy[j][k] = clamp_coord<CT>(fma(f[k], yu, p.rad.yc), p.H - 1);
                }
            }
        } else {
            // projective map, postprocessing.py:450-457, same operation order
            double c1x[kCols], c4x[kCols], c7x[kCols];
#pragma unroll
            for (int k = 0; k < kCols; ++k) {
                const double x = (double)(x_base + 32 * k);
                c1x[k] = __dmul_rn(p.per.c[0], x);
                c4x[k] = __dmul_rn(p.per.c[3], x);
                c7x[k] = __dmul_rn(p.per.c[6], x);
            }
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const double y = (double)(y_base + j);
                const double c2y = __dmul_rn(p.per.c[1], y);
                const double c5y = __dmul_rn(p.per.c[4], y);
                const double c8y = __dmul_rn(p.per.c[7], y);
#pragma unroll
                for (int k = 0; k < kCols; ++k) {
                    const double den = __dadd_rn(__dadd_rn(c7x[k], c8y), 1.0);
                    const double nx = __dadd_rn(__dadd_rn(c1x[k], c2y), p.per.c[2]);
                    const double ny = __dadd_rn(__dadd_rn(c4x[k], c5y), p.per.c[5]);
                    cx[j][k] = clamp_coord<CT>(__ddiv_rn(nx, den), wmax);
                    cy[j][k] = clamp_coord<CT>(__ddiv_rn(ny, den), p.H - 1);
                }
            }
        }

        // ---- source bounding box of the tile ---------------------------------
        bool fits = false;
        int bx0 = 0, by0 = 0;
        if (staged) {
            CT fmnx = (CT)3.0e9, fmny = (CT)3.0e9, fmxx = (CT)-1, fmxy = (CT)-1;
#pragma unroll
            for (int j = 0; j < RPT; ++j)
#pragma unroll
                for (int k = 0; k < kCols; ++k) {
                    const bool valid = (x_base + 32 * k < p.W) && (y_base + j < y_end);
                    if (valid) {
                        fmnx = cx[j][k] < fmnx ? cx[j][k] : fmnx;
                        fmxx = cx[j][k] > fmxx ? cx[j][k] : fmxx;
                        fmny = cy[j][k] < fmny ? cy[j][k] : fmny;
                        fmxy = cy[j][k] > fmxy ? cy[j][k] : fmxy;
                    }
                }
            // cvt.rzi saturates: 3e9 -> INT_MAX, -1 -> -1
            int mnx = __reduce_min_sync(0xffffffffu, (int)fmnx);
            int mny = __reduce_min_sync(0xffffffffu, (int)fmny);
            int mxx = __reduce_max_sync(0xffffffffu, (int)fmxx);
            int mxy = __reduce_max_sync(0xffffffffu, (int)fmxy);
            int *rd = red + tile_par * 4 * kWarps;
            if (lane == 0) {
                rd[0 * kWarps + warp] = mnx;
                rd[1 * kWarps + warp] = mny;
                rd[2 * kWarps + warp] = mxx;
                rd[3 * kWarps + warp] = mxy;
            }
            __syncthreads();
            mnx = mny = INT_MAX;
            mxx = mxy = -1;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                mnx = min(mnx, rd[0 * kWarps + w]);
                mny = min(mny, rd[1 * kWarps + w]);
                mxx = max(mxx, rd[2 * kWarps + w]);
                mxy = max(mxy, rd[3 * kWarps + w]);
            }
            const int bx1 = min(mxx + 1, wmax);
            const int by1 = min(max(mxy + 1, p.yorg), p.ylast);
            // measured on B200: the box's innermost start coordinate must be a
            // multiple of 16 bytes, otherwise UTMALDG raises "illegal instruction"
            bx0 = mnx & ~3;
            by0 = min(max(mny, p.yorg), p.ylast);
            fits = (bx1 - bx0 + 1 <= p.bw) && (by1 - by0 + 1 <= p.bh);
        }

        if (fits && threadIdx.x == 0) {
            const int npre = min(nz, p.nstage);
            for (int s = 0; s < npre; ++s) {
                const uint32_t st = (fills + s) % p.nstage;
                mbar_expect_tx(&full[st], p.box_bytes);
                tma_load_3d(smem + (size_t)st * p.stage_bytes, &tmap, bx0, by0 - p.yorg, z0 + s,
                            &full[st]);
            }
        }

        // ---- slices of the chunk ----------------------------------------------
        for (int iz = 0; iz < nz; ++iz) {
            const int z = z0 + iz;
            float *out = p.dst + (long long)z * p.dst_slice;
            if (fits) {
                const uint32_t st = fills % p.nstage;
                mbar_wait(&full[st], (fills / p.nstage) & 1u);
                SmemFetch fetch{reinterpret_cast<const float *>(smem + (size_t)st * p.stage_bytes),
                                p.bw, -(by0 * p.bw + bx0)};
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const int y = y_base + j;
                    if (y < y_end) {
                        float *orow = out + (long long)(y - p.row0) * p.dst_pitch;
                        float v[kCols];
#pragma unroll
                        for (int k = 0; k < kCols; ++k)
                            v[k] = (x_base + 32 * k < p.W)
                                       ? sample_px<ORDER, BLEND, CT>(fetch, cx[j][k], cy[j][k], wmax,
                                                                     p.yorg, p.ylast)
                                       : 0.0f;
#pragma unroll
                        for (int k = 0; k < kCols; ++k)
                            if (x_base + 32 * k < p.W) __stcs(orow + x_base + 32 * k, v[k]);
                    }
                }
                ++fills;
                if (iz + p.nstage < nz) {  // CTA-uniform: refill the stage just drained
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        mbar_expect_tx(&full[st], p.box_bytes);
                        tma_load_3d(smem + (size_t)st * p.stage_bytes, &tmap, bx0, by0 - p.yorg,
                                    z + p.nstage, &full[st]);
                    }
                }
            } else {
                GlobalFetch fetch{p.src + (long long)z * p.src_slice - (long long)p.yorg * p.src_pitch,
                                  p.src_pitch};
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const int y = y_base + j;
                    if (y < y_end) {
                        float *orow = out + (long long)(y - p.row0) * p.dst_pitch;
                        float v[kCols];
#pragma unroll
                        for (int k = 0; k < kCols; ++k)
                            v[k] = (x_base + 32 * k < p.W)
                                       ? sample_px<ORDER, BLEND, CT>(fetch, cx[j][k], cy[j][k], wmax,
                                                                     p.yorg, p.ylast)
                                       : 0.0f;
#pragma unroll
                        for (int k = 0; k < kCols; ++k)
                            if (x_base + 32 * k < p.W) __stcs(orow + x_base + 32 * k, v[k]);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// caller-supplied coordinates (map_index= / _mapping): one thread per output
// ---------------------------------------------------------------------------
template <int ORDER, int BLEND, class CT>
__global__ void __launch_bounds__(256)
    map_coords_kernel(const float *__restrict__ src, float *__restrict__ dst, int H, int W,
                      long long pitch, const CT *__restrict__ yd, const CT *__restrict__ xd,
                      size_t n, unsigned *oob_count) {
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
        dst[i] = sample_px<ORDER, BLEND, CT>(fetch, x, y, W - 1, 0, H - 1);
    }
    if (oob_count != nullptr && oob != 0) atomicAdd(oob_count, oob);
}

// ---------------------------------------------------------------------------
// synthetic input generator (bench only): splitmix64 counter hash -> [0,1)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    fill_synthetic_kernel(float *__restrict__ dst, size_t n, uint64_t seed, uint64_t offset) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t h = splitmix64(seed ^ (offset + i));
        dst[i] = (float)(h >> 40) * (1.0f / 16777216.0f);
    }
}

}  // namespace dcb
