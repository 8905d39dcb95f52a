// remap_image.cuh -- the single-image backward-remap kernel (sm_100a), v2.
//
// Serves a1 `unwarp_image_backward` (postprocessing.py:111-148) and a4
// `correct_perspective_image` (:444-492) for one float32 image, or a range of
// its output rows.  Numerics are those of remap.cuh (same helpers); what is
// new is the schedule, shaped by what the first ncu capture showed
// (profiles/r1): the v1 kernel executed 131 thread-instructions per pixel, 14
// of them on the 16-lane/clk XU pipe (f32<->f64 conversions, F2I, MUFU), kept
// 16 coordinates live across the TMA wait (128 registers, spills) and exposed
// the whole TMA latency in every tile.
//
//   * persistent CTAs walk output tiles of 128 x 32 pixels; while tile i is
//     sampled, warp 0 estimates the source box of tile i+1 from 9 probe
//     points and issues its TMA load, so the copy is hidden behind a full
//     tile of fp64 work;
//   * the box lands as float32 and is widened ONCE per source pixel into a
//     float64 tile in shared memory (1.2 conversions per output pixel instead
//     of 4, the conversions being the scarcest issue slots);
//   * coordinates are evaluated four pixels at a time right before they are
//     used -- nothing stays live across a barrier except the per-thread column
//     terms;
//   * floor() of the clamped fp32 coordinate uses the 2^23 trick on the FP32
//     pipe instead of F2I/I2F on the XU pipe;
//   * every pixel checks that its 2x2 footprint lies inside the staged box and
//     strictly inside the image; the (rare) pixels that do not -- box estimate
//     too small, last row/column, clipped regions -- take the generic global
//     gather of remap.cuh, so correctness never depends on the estimate.
#pragma once
#include "remap.cuh"

namespace dcb {

constexpr int kImgTileH = 32;
constexpr int kImgRowsPerWarp = kImgTileH / kWarps;  // 4

struct ImageParams {
    const float *src;
    float *dst;
    long long src_pitch, dst_pitch;  // elements
    int H, W;
    int row0, nrows;   // output rows produced by this launch
    int yorg, ylast;   // image rows held by src (src points at row yorg)
    int tiles_x, ntiles;
    int bw, bh;        // staged box, bw % 4 == 0; bw == 0 => never stage
    unsigned box_bytes, stage_bytes;
    RadialDev rad;
    PerspDev per;
};

struct TileBox {
    int bx0, by0;  // image coordinates of box element (0,0)
    int use;       // 1: TMA issued for this tile, 0: sample straight from global
    int pad;
};

// floor of a float v in [0, 2^23) without the XU pipe: one round-toward-minus-
// infinity add puts floor(v) in the low mantissa bits of t = 2^23 + floor(v)
// (0x4B000000 is the bit pattern of 2^23).
__device__ __forceinline__ float floor_magic(float v) { return __fadd_rd(v, 8388608.0f); }

// clip(round_to_f32(d), 0, vmax) on the integer pipe: non-negative floats order
// like their bit patterns, negative ones (sign bit set) are negative integers.
__device__ __forceinline__ float clamp_coord_bits(double d, int vmax_bits) {
    return __int_as_float(__vimin_s32_relu(__float_as_int(__double2float_rn(d)), vmax_bits));
}

// sqrt for s > 0 (callers keep s away from 0, see MapEval<MAP_RADIAL>::row)
__device__ __forceinline__ double dsqrt_nz(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double g = s * y;
    double h = 0.5 * y;
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, s);
    return fma(d, h, g);
}

// Per-thread column terms of the map (constant down a tile) and the row
// evaluation producing clamped fp32 coordinates for 4 pixels.
template <int MAP, int NT>
struct MapEval;

template <int NT>
struct MapEval<MAP_RADIAL, NT> {
    double xu[kCols], xu2[kCols];
    __device__ __forceinline__ void set_columns(const ImageParams &p, const int (&x)[kCols]) {
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
            xu[k] = (double)x[k] - p.rad.xc;
            xu2[k] = __dmul_rn(xu[k], xu[k]);
        }
    }
    __device__ __forceinline__ void row(const ImageParams &p, int y, int wbits, int hbits,
                                        float (&xf)[kCols], float (&yf)[kCols]) const {
        const double yu = (double)y - p.rad.yc;
        double yu2 = __dmul_rn(yu, yu);
        // r = 0 only at a pixel sitting exactly on the centre.  Keeping s >= 1e-300
        // there gives r = 1e-150, hence the same F (= a0 after rounding) and the same
        // coordinates (F * 0), and it leaves every other s bit-identical -- so the
        // per-pixel zero test of dsqrt_pos is not needed in the hot loop.
        yu2 = (yu2 < 1e-300) ? 1e-300 : yu2;
        double r[kCols], f[kCols];
#pragma unroll
        for (int k = 0; k < kCols; ++k) r[k] = dsqrt_nz(__dadd_rn(xu2[k], yu2));
        if (NT > 0) {
#pragma unroll
            for (int k = 0; k < kCols; ++k) f[k] = horner<(NT > 0 ? NT : 1)>(p.rad.a, r[k]);
        } else {
            radial_factor<kCols>(p.rad.a, p.rad.n, r, f);
        }
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
            xf[k] = clamp_coord_bits(fma(f[k], xu[k], p.rad.xc), wbits);
            yf[k] = clamp_coord_bits(fma(f[k], yu, p.rad.yc), hbits);
        }
    }
};

template <int NT>
struct MapEval<MAP_PERSP, NT> {
    double c1x[kCols], c4x[kCols], c7x[kCols];
    __device__ __forceinline__ void set_columns(const ImageParams &p, const int (&x)[kCols]) {
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
            const double xd = (double)x[k];
            c1x[k] = __dmul_rn(p.per.c[0], xd);
            c4x[k] = __dmul_rn(p.per.c[3], xd);
            c7x[k] = __dmul_rn(p.per.c[6], xd);
        }
    }
    __device__ __forceinline__ void row(const ImageParams &p, int y, int wbits, int hbits,
                                        float (&xf)[kCols], float (&yf)[kCols]) const {
        const double yd = (double)y;
        const double c2y = __dmul_rn(p.per.c[1], yd);
        const double c5y = __dmul_rn(p.per.c[4], yd);
        const double c8y = __dmul_rn(p.per.c[7], yd);
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
            const double den = __dadd_rn(__dadd_rn(c7x[k], c8y), 1.0);
            const double nx = __dadd_rn(__dadd_rn(c1x[k], c2y), p.per.c[2]);
            const double ny = __dadd_rn(__dadd_rn(c4x[k], c5y), p.per.c[5]);
            xf[k] = clamp_coord_bits(__ddiv_rn(nx, den), wbits);
            yf[k] = clamp_coord_bits(__ddiv_rn(ny, den), hbits);
        }
    }
};

// single probe point (used by warp 0 to place the next tile's box)
template <int MAP>
__device__ __forceinline__ void map_point(const ImageParams &p, int x, int y, float &xf,
                                          float &yf) {
    if (MAP == MAP_RADIAL) {
        const double xu = (double)x - p.rad.xc, yu = (double)y - p.rad.yc;
        const double r = dsqrt_pos(__dadd_rn(__dmul_rn(xu, xu), __dmul_rn(yu, yu)));
        double f = 0.0;
        for (int i = p.rad.n - 1; i >= 0; --i) f = fma(f, r, p.rad.a[i]);
        xf = clamp_coord<float>(fma(f, xu, p.rad.xc), p.W - 1);
        yf = clamp_coord<float>(fma(f, yu, p.rad.yc), p.H - 1);
    } else {
        const double xd = (double)x, yd = (double)y;
        const double den = p.per.c[6] * xd + p.per.c[7] * yd + 1.0;
        xf = clamp_coord<float>((p.per.c[0] * xd + p.per.c[1] * yd + p.per.c[2]) / den, p.W - 1);
        yf = clamp_coord<float>((p.per.c[3] * xd + p.per.c[4] * yd + p.per.c[5]) / den, p.H - 1);
    }
}

template <int ORDER, int BLEND>
struct ImageKernelTraits {
    // bilinear in fp64: sample from the widened tile; otherwise from the raw
    // float32 box (two stages, no widening pass)
    static constexpr bool kWide = (ORDER == 1 && BLEND != DCB_BLEND_LERP32);
};

// NT > 0: number of polynomial terms known at compile time (coefficients become
// constant-bank operands of the DFMAs); NT == 0: any p.rad.n through a switch.
template <int MAP, int ORDER, int BLEND, int NT>
__global__ void __launch_bounds__(kThreads, 3)
    remap_image_kernel(const __grid_constant__ ImageParams p,
                       const __grid_constant__ CUtensorMap tmap) {
    constexpr bool WIDE = ImageKernelTraits<ORDER, BLEND>::kWide;
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [raw stage 0][raw stage 1 (only !WIDE)][wide tile (only WIDE)][mbar][TileBox x2]
    float *raw0 = reinterpret_cast<float *>(smem);
    unsigned char *after_raw = smem + (WIDE ? 1 : 2) * (size_t)p.stage_bytes;
    double *wide = reinterpret_cast<double *>(after_raw);
    unsigned char *tail = after_raw + (WIDE ? 2 * (size_t)p.stage_bytes : 0);
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);           // [2]
    TileBox *boxes = reinterpret_cast<TileBox *>(tail + 16);       // [2]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool staged = p.bw > 0;
    const int wmax = p.W - 1;
    const int y_end = p.row0 + p.nrows;

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
        if (staged) tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    // ---- warp 0: place the box of tile t and start its copy -------------------
    auto prefetch_tile = [&](int t, int slot, uint32_t fill_index) {
        // all 32 lanes of warp 0 execute this
        const int txi = t % p.tiles_x, tyi = t / p.tiles_x;
        const int x_lo = txi * kTileW, y_lo = p.row0 + tyi * kImgTileH;
        const int x_hi = min(x_lo + kTileW - 1, wmax), y_hi = min(y_lo + kImgTileH - 1, y_end - 1);
        const int q = lane % 9;
        const int px = x_lo + ((x_hi - x_lo) * (q % 3)) / 2;
        const int py = y_lo + ((y_hi - y_lo) * (q / 3)) / 2;
        float xf, yf;
        map_point<MAP>(p, px, py, xf, yf);
        const int mnx = __reduce_min_sync(0xffffffffu, (int)xf);
        const int mny = __reduce_min_sync(0xffffffffu, (int)yf);
        const int mxx = __reduce_max_sync(0xffffffffu, (int)xf);
        const int mxy = __reduce_max_sync(0xffffffffu, (int)yf);
        // one pixel of slack around the probes for curvature inside the tile
        const int bx0 = max(mnx - 1, 0) & ~3;
        const int by0 = min(max(mny - 1, p.yorg), p.ylast);
        const bool use = staged && (mxx + 2 - bx0 < p.bw) && (mxy + 2 - by0 < p.bh);
        if (lane == 0) {
            boxes[slot].bx0 = bx0;
            boxes[slot].by0 = by0;
            boxes[slot].use = use ? 1 : 0;
            if (use) {
                const int st = WIDE ? 0 : (int)(fill_index & 1u);
                mbar_expect_tx(&full[st], p.box_bytes);
                tma_load_3d(smem + (size_t)st * p.stage_bytes, &tmap, bx0, by0 - p.yorg, 0,
                            &full[st]);
            }
        }
    };

    uint32_t nfill = 0;  // TMA copies consumed so far (CTA-uniform)
    int t = blockIdx.x;
    if (t < p.ntiles && warp == 0) prefetch_tile(t, 0, 0);
    __syncthreads();

    for (int it = 0; t < p.ntiles; ++it, t += gridDim.x) {
        const TileBox box = boxes[it & 1];
        const int st = WIDE ? 0 : (int)(nfill & 1u);
        if (box.use) {
            // WIDE: one barrier, phase flips per fill.  !WIDE: two barriers used alternately.
            const uint32_t parity = WIDE ? (nfill & 1u) : ((nfill >> 1) & 1u);
            mbar_wait(&full[st], parity);
            if (WIDE) {
                // widen the landed box once: float32 -> float64 (exact)
                const float2 *src2 = reinterpret_cast<const float2 *>(raw0);
                double2 *dst2 = reinterpret_cast<double2 *>(wide);
                const int n2 = (p.bw * p.bh) >> 1;
                for (int i = threadIdx.x; i < n2; i += kThreads) {
                    const float2 v = src2[i];
                    dst2[i] = make_double2((double)v.x, (double)v.y);
                }
            }
        }
        __syncthreads();  // wide tile complete, raw stage reusable, boxes[(it+1)&1] free
        const int t_next = t + gridDim.x;
        const uint32_t fills_after = nfill + (box.use ? 1u : 0u);
        if (warp == 0 && t_next < p.ntiles) prefetch_tile(t_next, (it + 1) & 1, fills_after);

        // ---- sample tile t ---------------------------------------------------------
        {
            const int txi = t % p.tiles_x, tyi = t / p.tiles_x;
            const int x_base = txi * kTileW + lane;
            const int y_base = p.row0 + tyi * kImgTileH + warp * kImgRowsPerWarp;
            int xs[kCols];
#pragma unroll
            for (int k = 0; k < kCols; ++k) xs[k] = min(x_base + 32 * k, wmax);  // clamp: edge lanes recompute a valid pixel
            MapEval<MAP, NT> ev;
            ev.set_columns(p, xs);
            const int wbits = __float_as_int((float)wmax), hbits = __float_as_int((float)(p.H - 1));
            // bits(2^23 + n) - magic = n - box origin
            const int magic_x = 0x4B000000 + box.bx0, magic_y = 0x4B000000 + box.by0;
            // fast-path window: footprint inside the box and strictly inside the image
            const int lim_x = box.use ? min(p.bw - 1, wmax - box.bx0) : 0;
            const int lim_y = box.use ? min(p.bh - 1, p.ylast - box.by0) : 0;
            const float *rawt = reinterpret_cast<const float *>(smem + (size_t)st * p.stage_bytes);
            const GlobalFetch gfetch{p.src - (long long)p.yorg * p.src_pitch, p.src_pitch};
#pragma unroll 1
            for (int j = 0; j < kImgRowsPerWarp; ++j) {
                const int y = y_base + j;
                if (y >= y_end) break;  // warp-uniform
                float xf[kCols], yf[kCols];
                ev.row(p, y, wbits, hbits, xf, yf);
                float tfx[kCols], tfy[kCols];
                int ix[kCols], iy[kCols];
                bool ok = true;
#pragma unroll
                for (int k = 0; k < kCols; ++k) {
                    tfx[k] = floor_magic(xf[k]);
                    tfy[k] = floor_magic(yf[k]);
                    ix[k] = __float_as_int(tfx[k]) - magic_x;  // x0 - bx0
                    iy[k] = __float_as_int(tfy[k]) - magic_y;  // y0 - by0
                    ok = ok && ((unsigned)ix[k] < (unsigned)lim_x) &&
                         ((unsigned)iy[k] < (unsigned)lim_y);
                }
                float v[kCols];
                if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
                    for (int k = 0; k < kCols; ++k) {
                        const float tx = xf[k] - (tfx[k] - 8388608.0f);  // exact
                        const float ty = yf[k] - (tfy[k] - 8388608.0f);
                        const int idx = iy[k] * p.bw + ix[k];
                        if (ORDER == 0) {
                            const int sel = idx + (tx >= 0.5f ? 1 : 0) + (ty >= 0.5f ? p.bw : 0);
                            v[k] = WIDE ? (float)wide[sel] : rawt[sel];
                        } else if (!WIDE) {
                            const float a = rawt[idx], b = rawt[idx + 1];
                            const float c = rawt[idx + p.bw], d = rawt[idx + p.bw + 1];
                            const float top = fmaf(b - a, tx, a);
                            const float bot = fmaf(d - c, tx, c);
                            v[k] = fmaf(bot - top, ty, top);
                        } else {
                            const double a = wide[idx], b = wide[idx + 1];
                            const double c = wide[idx + p.bw], d = wide[idx + p.bw + 1];
                            const double wx1 = (double)tx, wy1 = (double)ty;
                            if (BLEND == DCB_BLEND_LERP64) {
                                const double top = fma(b - a, wx1, a);
                                const double bot = fma(d - c, wx1, c);
                                v[k] = (float)fma(bot - top, wy1, top);
                            } else {
                                const double wx0 = __dsub_rn(1.0, wx1), wy0 = __dsub_rn(1.0, wy1);
                                double s = __dmul_rn(__dmul_rn(a, wy0), wx0);
                                s = __dadd_rn(s, __dmul_rn(__dmul_rn(b, wy0), wx1));
                                s = __dadd_rn(s, __dmul_rn(__dmul_rn(c, wy1), wx0));
                                s = __dadd_rn(s, __dmul_rn(__dmul_rn(d, wy1), wx1));
                                v[k] = __double2float_rn(s);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kCols; ++k)
                        v[k] = sample_px<ORDER, BLEND, float>(gfetch, xf[k], yf[k], wmax, p.yorg,
                                                              p.ylast);
                }
                float *orow = p.dst + (long long)(y - p.row0) * p.dst_pitch;
#pragma unroll
                for (int k = 0; k < kCols; ++k)
                    if (x_base + 32 * k <= wmax) __stcs(orow + x_base + 32 * k, v[k]);
            }
        }
        nfill = fills_after;
        __syncthreads();  // everyone is done with the wide tile / raw stage and boxes[it&1]
    }
}

}  // namespace dcb
