// remap_stack.cuh -- the Z-stack / batch backward-remap kernel (sm_100a).
//
// Serves a2 `unwarp_slice_backward` (postprocessing.py:188-229, float64
// coordinates, one output row of every slice), a3 `unwarp_chunk_slices_backward`
// (:255-313, fp32-rounded coordinates, rows start..stop of every slice) and,
// with all rows, a stack or batch of independent images that share one radial
// model (BASELINE configs 4 and 5).
//
// The reference loops over slices in Python and re-derives the sampling
// weights inside SciPy for every slice.  Here the whole geometry of an output
// tile is evaluated ONCE and kept in registers -- the fp64 radial map, floor,
// the four fp64 bilinear weights (or the fp32 fractions) and the tap offset of
// each of the thread's 8 pixels -- and the slices of a Z-chunk then stream
// through a TMA ring: per slice and pixel what is left is 4 shared-memory
// loads, the blend and one coalesced store, which is below the cost of
// moving the 8 bytes through HBM.
//
//   * work item = (128 x 16 output tile, chunk of <= 64 slices); CTAs take
//     items round-robin with the tile index fastest, so the CTAs resident at
//     any time sweep the same slices of neighbouring tiles and the box halos
//     they share are L2 hits;
//   * the tile's exact source bounding box comes from redux.sync + one
//     shared-memory exchange; if it fits the staged box, thread 0 keeps
//     nstage-1 slices in flight with 3-D TMA loads (cp.async.bulk.tensor) into
//     an mbarrier ring: `full[s]` completes on the copy's bytes, `empty[s]`
//     collects one arrival per warp, and the refill of a stage waits only for
//     the slice consumed one iteration earlier -- no CTA-wide barrier in the
//     slice loop;
//   * because the bounding box is exact, the slice loop has no per-pixel
//     range test; tiles whose box does not fit (strong magnification) gather
//     straight from global memory with the arithmetic of remap.cuh.
#pragma once
#include "remap.cuh"

namespace dcb {

constexpr int kStkTileH = 16;
constexpr int kStkRows = kStkTileH / kWarps;  // 2 rows per warp
constexpr int kStkPx = kStkRows * kCols;      // 8 pixels per thread
constexpr int kStkMaxStages = 8;

// How the per-pixel sampling state is kept across the slices of a chunk.
template <int ORDER, int BLEND, bool ROUND32>
struct StackWeights {
    static constexpr bool kNearest = (ORDER == 0);
    static constexpr bool kF32 = (ORDER == 1 && BLEND == DCB_BLEND_LERP32);
    // fp32-rounded coordinates: the four weight products are exact in fp64
    // (see blend_exact in remap_image.cuh), so they are formed once per tile
    static constexpr bool kW4 = (ORDER == 1 && BLEND == DCB_BLEND_EXACT && ROUND32);
    // everything else keeps the two fractions as doubles
    static constexpr bool kT64 = (ORDER == 1 && !kF32 && !kW4);
    static constexpr int kDoubles = kW4 ? 4 : (kT64 ? 2 : 0);
    // Sampling the fp64 blends from a float64 copy of the staged box (one
    // conversion per source pixel instead of one per tap) was measured and
    // rejected: the CTA barrier it needs per slice costs more than the XU
    // conversions it saves (64 x 4096^2, exact: 48 % vs 56 % of the HBM peak).
    // The code path is kept behind this switch for the record.
    static constexpr bool kWiden = false;
};

template <int ORDER, int BLEND, bool ROUND32>
__global__ void __launch_bounds__(kThreads, 2)
    remap_stack_kernel(const __grid_constant__ RemapParams p,
                       const __grid_constant__ CUtensorMap tmap) {
    using CT = typename std::conditional<ROUND32, float, double>::type;
    using SW = StackWeights<ORDER, BLEND, ROUND32>;

    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [nstage raw boxes][two float64 tiles (fp64 blends only)][full][empty][red]
    unsigned char *wide_base = smem + (size_t)p.nstage * p.stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(wide_base + (SW::kWiden ? 4 : 0) * (size_t)p.stage_bytes);
    uint64_t *empty = full + kStkMaxStages;
    int *red = reinterpret_cast<int *>(empty + kStkMaxStages);  // [2][4][kWarps]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool staged = p.nstage > 0;
    const uint32_t S = (uint32_t)p.nstage;
    if (staged && threadIdx.x == 0) {
        for (int s = 0; s < p.nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarps);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    uint32_t fills = 0;  // slices that went through the ring so far (CTA-uniform)
    int par = 0;
    const int wmax = p.W - 1;
    const int hmax = p.H - 1;
    const int y_end = p.row0 + p.nrows;
    const int bw = p.bw;

    for (int item = blockIdx.x; item < p.ntiles; item += gridDim.x, par ^= 1) {
        const int txi = item % p.tiles_x;
        const int rest = item / p.tiles_x;
        const int tyi = rest % p.tiles_y;
        const int zci = rest / p.tiles_y;
        const int x_base = txi * kTileW + lane;
        const int y_base = p.row0 + tyi * kStkTileH + warp * kStkRows;
        const int z0 = zci * p.zchunk;
        const int nz = min(p.zchunk, p.D - z0);

        // ---- geometry of the thread's 8 pixels, once per item ----------------------
        int x0[kStkPx], r0[kStkPx];            // floor column, first tap row (window-clamped)
        unsigned dxm = 0, dym = 0;             // bit i: second tap is one column right / one row down
        CT tx[kStkPx], ty[kStkPx];
        {
            double xu[kCols], xu2[kCols];
#pragma unroll
            for (int k = 0; k < kCols; ++k) {
                xu[k] = (double)min(x_base + 32 * k, wmax) - p.rad.xc;  // :138 (edge lanes redo a valid pixel)
                xu2[k] = __dmul_rn(xu[k], xu[k]);
            }
#pragma unroll
            for (int j = 0; j < kStkRows; ++j) {
                const double yu = (double)min(y_base + j, y_end - 1) - p.rad.yc;  // :139
                const double yu2 = __dmul_rn(yu, yu);
                double r[kCols], f[kCols];
#pragma unroll
                for (int k = 0; k < kCols; ++k) r[k] = dsqrt_pos(__dadd_rn(xu2[k], yu2));  // :141
                radial_factor<kCols>(p.rad.a, p.rad.n, r, f);                              // :142-143
#pragma unroll
                for (int k = 0; k < kCols; ++k) {  // :144-145 (image, chunk) / :219-220 (slice)
                    const int i = j * kCols + k;
                    const CT cx = clamp_coord<CT>(fma(f[k], xu[k], p.rad.xc), wmax);
                    const CT cy = clamp_coord<CT>(fma(f[k], yu, p.rad.yc), hmax);
                    int xi = (int)cx, yi = (int)cy;  // truncation == floor, coordinates are >= 0
                    tx[i] = cx - (CT)xi;             // exact
                    ty[i] = cy - (CT)yi;
                    if (ORDER == 0) {
                        // SciPy: floor(c + 0.5) in double == compare the exact fraction with 0.5
                        if (tx[i] >= (CT)0.5) ++xi;
                        if (ty[i] >= (CT)0.5) ++yi;
                        x0[i] = xi;
                        r0[i] = min(max(yi, p.yorg), p.ylast);
                    } else {
                        // rows are clamped into the window the caller holds (a no-op for whole
                        // images); the +1 taps fold back onto the last row / column
                        const int x1 = min(xi + 1, wmax);
                        const int y1 = min(max(yi + 1, p.yorg), p.ylast);
                        yi = min(max(yi, p.yorg), p.ylast);
                        x0[i] = xi;
                        r0[i] = yi;
                        dxm |= (unsigned)(x1 - xi) << i;
                        dym |= (unsigned)(y1 - yi) << i;
                    }
                }
            }
        }

        // ---- exact source bounding box of the tile ----------------------------------
        bool fits = false, edge = false;
        int bx0 = 0, by0 = 0;
        if (staged) {
            int mnx = INT_MAX, mny = INT_MAX, mxx = -1, mxy = -1;
#pragma unroll
            for (int i = 0; i < kStkPx; ++i) {
                mnx = min(mnx, x0[i]);
                mny = min(mny, r0[i]);
                mxx = max(mxx, x0[i] + (int)((dxm >> i) & 1u));
                mxy = max(mxy, r0[i] + (int)((dym >> i) & 1u));
            }
            mnx = __reduce_min_sync(0xffffffffu, mnx);
            mny = __reduce_min_sync(0xffffffffu, mny);
            mxx = __reduce_max_sync(0xffffffffu, mxx);
            mxy = __reduce_max_sync(0xffffffffu, mxy);
            int *rd = red + par * 4 * kWarps;
            if (lane == 0) {
                rd[0 * kWarps + warp] = mnx;
                rd[1 * kWarps + warp] = mny;
                rd[2 * kWarps + warp] = mxx;
                rd[3 * kWarps + warp] = mxy;
            }
            // a pixel whose +1 tap was folded back needs the generic tap offsets
            const bool mine = (ORDER == 1) && (dxm != 0xffu || dym != 0xffu);
            edge = __syncthreads_or(mine ? 1 : 0) != 0;
            mnx = mny = INT_MAX;
            mxx = mxy = -1;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                mnx = min(mnx, rd[0 * kWarps + w]);
                mny = min(mny, rd[1 * kWarps + w]);
                mxx = max(mxx, rd[2 * kWarps + w]);
                mxy = max(mxy, rd[3 * kWarps + w]);
            }
            // measured on B200: the box's innermost start coordinate must be a
            // multiple of 16 bytes, otherwise UTMALDG raises "illegal instruction"
            bx0 = mnx & ~3;
            by0 = mny;
            fits = (mxx - bx0 + 1 <= bw) && (mxy - by0 + 1 <= p.bh);
        }

        if (fits) {
            // ---- per-pixel state kept across the chunk --------------------------------
            int off[kStkPx];
            double wd[kStkPx][SW::kDoubles > 0 ? SW::kDoubles : 1];
            float wf[kStkPx][SW::kF32 ? 2 : 1];
#pragma unroll
            for (int i = 0; i < kStkPx; ++i) {
                off[i] = (r0[i] - by0) * bw + (x0[i] - bx0);
                if (SW::kW4) {
                    const double dtx = (double)tx[i], dty = (double)ty[i];
                    const double w11 = __dmul_rn(dty, dtx);      // all four products are exact
                    wd[i][3] = w11;
                    wd[i][2] = __dsub_rn(dty, w11);              // w10 = ty (1 - tx)
                    wd[i][1] = __dsub_rn(dtx, w11);              // w01 = (1 - ty) tx
                    wd[i][0] = __dsub_rn(__dsub_rn(1.0, dty), wd[i][1]);
                } else if (SW::kT64) {
                    wd[i][0] = (double)tx[i];
                    wd[i][1] = (double)ty[i];
                } else if (SW::kF32) {
                    wf[i][0] = (float)tx[i];
                    wf[i][1] = (float)ty[i];
                }
            }
            const uint32_t base = fills;
            if (threadIdx.x == 0) {
                const int npre = min(nz, p.nstage - 1);
                for (int s = 0; s < npre; ++s) {
                    const uint32_t g = base + s, st = g % S;
                    if (g >= S) mbar_wait(&empty[st], ((g / S) - 1u) & 1u);
                    mbar_expect_tx(&full[st], p.box_bytes);
                    tma_load_3d(smem + (size_t)st * p.stage_bytes, &tmap, bx0, by0 - p.yorg, z0 + s,
                                &full[st]);
                }
            }
            float *orow = p.dst + (long long)z0 * p.dst_slice +
                          (long long)(y_base - p.row0) * p.dst_pitch + x_base;
            bool colok[kCols];
#pragma unroll
            for (int k = 0; k < kCols; ++k) colok[k] = (x_base + 32 * k <= wmax);

            for (int iz = 0; iz < nz; ++iz, orow += p.dst_slice) {
                const uint32_t g = base + iz, st = g % S;
                mbar_wait(&full[st], (g / S) & 1u);
                const float *tile = reinterpret_cast<const float *>(smem + (size_t)st * p.stage_bytes);
                const double *wtile = reinterpret_cast<const double *>(
                    wide_base + (size_t)(2 * (iz & 1)) * p.stage_bytes);
                if (SW::kWiden) {
                    // float32 box -> float64 tile (exact), then the raw stage can be refilled;
                    // the tile written here was last read two slices ago, behind the barrier
                    // of the previous slice
                    const float4 *src4 = reinterpret_cast<const float4 *>(tile);
                    double2 *dst2 = reinterpret_cast<double2 *>(
                        wide_base + (size_t)(2 * (iz & 1)) * p.stage_bytes);
                    const int n4 = (bw * p.bh) >> 2;  // bw % 4 == 0
                    int e = threadIdx.x;
                    for (; e + kThreads < n4; e += 2 * kThreads) {
                        const float4 u = src4[e], w = src4[e + kThreads];
                        dst2[2 * e] = make_double2((double)u.x, (double)u.y);
                        dst2[2 * e + 1] = make_double2((double)u.z, (double)u.w);
                        dst2[2 * (e + kThreads)] = make_double2((double)w.x, (double)w.y);
                        dst2[2 * (e + kThreads) + 1] = make_double2((double)w.z, (double)w.w);
                    }
                    if (e < n4) {
                        const float4 u = src4[e];
                        dst2[2 * e] = make_double2((double)u.x, (double)u.y);
                        dst2[2 * e + 1] = make_double2((double)u.z, (double)u.w);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[st]);
                    __syncthreads();
                }
                float v[kStkPx];
#pragma unroll
                for (int i = 0; i < kStkPx; ++i) {
                    if (ORDER == 0) {
                        v[i] = tile[off[i]];
                        continue;
                    }
                    const int ox = edge ? (int)((dxm >> i) & 1u) : 1;
                    const int oy = edge ? (((dym >> i) & 1u) ? bw : 0) : bw;
                    if (SW::kF32) {
                        const float *q = tile + off[i];
                        const float a = q[0], b = q[ox], c = q[oy], d = q[oy + ox];
                        const float top = fmaf(b - a, wf[i][0], a);
                        const float bot = fmaf(d - c, wf[i][0], c);
                        v[i] = finish_f32(fmaf(bot - top, wf[i][1], top), p.rint);
                        continue;
                    }
                    double a, b, c, d;
                    if (SW::kWiden) {
                        const double *q = wtile + off[i];
                        a = q[0], b = q[ox], c = q[oy], d = q[oy + ox];
                    } else {
                        const float *q = tile + off[i];
                        a = (double)q[0], b = (double)q[ox], c = (double)q[oy], d = (double)q[oy + ox];
                    }
                    if (SW::kW4) {
                        double s = __dmul_rn(a, wd[i][0]);
                        s = __dadd_rn(s, __dmul_rn(b, wd[i][1]));
                        s = __dadd_rn(s, __dmul_rn(c, wd[i][2]));
                        s = __dadd_rn(s, __dmul_rn(d, wd[i][3]));
                        v[i] = finish_f64(s, p.rint);
                    } else if (BLEND == DCB_BLEND_LERP64) {
                        const double top = fma(b - a, wd[i][0], a);
                        const double bot = fma(d - c, wd[i][0], c);
                        v[i] = finish_f64(fma(bot - top, wd[i][1], top), p.rint);
                    } else {
                        // float64 coordinates: SciPy's two-step products, every step rounded
                        const double wx1 = wd[i][0], wy1 = wd[i][1];
                        const double wx0 = __dsub_rn(1.0, wx1), wy0 = __dsub_rn(1.0, wy1);
                        double s = __dmul_rn(__dmul_rn(a, wy0), wx0);
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn(b, wy0), wx1));
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn(c, wy1), wx0));
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn(d, wy1), wx1));
                        v[i] = finish_f64(s, p.rint);
                    }
                }
                if (!SW::kWiden) {
                    // this warp is done with the stage: let the producer refill it
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[st]);
                }
#pragma unroll
                for (int j = 0; j < kStkRows; ++j) {
                    if (y_base + j < y_end) {
                        float *o = orow + (long long)j * p.dst_pitch;
#pragma unroll
                        for (int k = 0; k < kCols; ++k)
                            if (colok[k]) __stcs(o + 32 * k, v[j * kCols + k]);
                    }
                }
                if (threadIdx.x == 0 && iz + p.nstage - 1 < nz) {
                    // refill the stage consumed one iteration ago (all warps are past it or
                    // about to be): keeps nstage-1 slices in flight
                    const uint32_t gn = g + S - 1u, sn = gn % S;
                    if (gn >= S) mbar_wait(&empty[sn], ((gn / S) - 1u) & 1u);
                    mbar_expect_tx(&full[sn], p.box_bytes);
                    tma_load_3d(smem + (size_t)sn * p.stage_bytes, &tmap, bx0, by0 - p.yorg,
                                z0 + iz + p.nstage - 1, &full[sn]);
                }
            }
            fills += (uint32_t)nz;
        } else {
            // ---- direct gathers (strong magnification, or the layout is not TMA-able):
            //      same per-pixel state, taps come through the read-only path -----------
            int off[kStkPx];
            double wd[kStkPx][SW::kDoubles > 0 ? SW::kDoubles : 1];
            float wf[kStkPx][SW::kF32 ? 2 : 1];
            const int pitch = (int)p.src_pitch;
#pragma unroll
            for (int i = 0; i < kStkPx; ++i) {
                off[i] = (r0[i] - p.yorg) * pitch + x0[i];  // < 2^31, checked by the host
                if (SW::kW4) {
                    const double dtx = (double)tx[i], dty = (double)ty[i];
                    const double w11 = __dmul_rn(dty, dtx);
                    wd[i][3] = w11;
                    wd[i][2] = __dsub_rn(dty, w11);
                    wd[i][1] = __dsub_rn(dtx, w11);
                    wd[i][0] = __dsub_rn(__dsub_rn(1.0, dty), wd[i][1]);
                } else if (SW::kT64) {
                    wd[i][0] = (double)tx[i];
                    wd[i][1] = (double)ty[i];
                } else if (SW::kF32) {
                    wf[i][0] = (float)tx[i];
                    wf[i][1] = (float)ty[i];
                }
            }
            float *orow = p.dst + (long long)z0 * p.dst_slice +
                          (long long)(y_base - p.row0) * p.dst_pitch + x_base;
            const float *sl = p.src + (long long)z0 * p.src_slice;
            for (int iz = 0; iz < nz; ++iz, orow += p.dst_slice, sl += p.src_slice) {
                float v[kStkPx];
#pragma unroll
                for (int i = 0; i < kStkPx; ++i) {
                    const float *q = sl + off[i];
                    if (ORDER == 0) {
                        v[i] = __ldg(q);
                        continue;
                    }
                    const int ox = (int)((dxm >> i) & 1u);
                    const int oy = ((dym >> i) & 1u) ? pitch : 0;
                    const float a = __ldg(q), b = __ldg(q + ox);
                    const float c = __ldg(q + oy), d = __ldg(q + oy + ox);
                    if (SW::kF32) {
                        const float top = fmaf(b - a, wf[i][0], a);
                        const float bot = fmaf(d - c, wf[i][0], c);
                        v[i] = finish_f32(fmaf(bot - top, wf[i][1], top), p.rint);
                    } else if (SW::kW4) {
                        double s = __dmul_rn((double)a, wd[i][0]);
                        s = __dadd_rn(s, __dmul_rn((double)b, wd[i][1]));
                        s = __dadd_rn(s, __dmul_rn((double)c, wd[i][2]));
                        s = __dadd_rn(s, __dmul_rn((double)d, wd[i][3]));
                        v[i] = finish_f64(s, p.rint);
                    } else if (BLEND == DCB_BLEND_LERP64) {
                        const double da = a, db = b, dc = c, dd = d;
                        const double top = fma(db - da, wd[i][0], da);
                        const double bot = fma(dd - dc, wd[i][0], dc);
                        v[i] = finish_f64(fma(bot - top, wd[i][1], top), p.rint);
                    } else {
                        const double wx1 = wd[i][0], wy1 = wd[i][1];
                        const double wx0 = __dsub_rn(1.0, wx1), wy0 = __dsub_rn(1.0, wy1);
                        double s = __dmul_rn(__dmul_rn((double)a, wy0), wx0);
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn((double)b, wy0), wx1));
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn((double)c, wy1), wx0));
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn((double)d, wy1), wx1));
                        v[i] = finish_f64(s, p.rint);
                    }
                }
#pragma unroll
                for (int j = 0; j < kStkRows; ++j) {
                    if (y_base + j < y_end) {
                        float *o = orow + (long long)j * p.dst_pitch;
#pragma unroll
                        for (int k = 0; k < kCols; ++k)
                            if (x_base + 32 * k <= wmax) __stcs(o + 32 * k, v[j * kCols + k]);
                    }
                }
            }
        }
    }
}

}  // namespace dcb
