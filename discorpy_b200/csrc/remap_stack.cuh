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

#ifndef DCB_STK_MINB
#define DCB_STK_MINB 2
#endif
#ifndef DCB_STK_SMEM_KB
#define DCB_STK_SMEM_KB 100
#endif
// Output tile shapes (template parameter COLS = columns per thread): 128 x 16 (COLS 4, two rows per
// warp) for ordinary maps, 64 x 32 (COLS 2, four rows per warp) for strongly sheared ones, whose
// 128-pixel tile rows have source footprints too tall to stage (BASELINE config 5: median 24,
// 90th percentile 67 source rows for 16 output rows; half its tiles missed the staged box and
// gathered from global memory).  Every store instruction still writes one full 128-byte line.
constexpr int kStkPx = 8;                     // pixels per thread
constexpr int kStkTileH = 16;                 // the 128 x 16 shape (host side: plan_and_launch_stack)
constexpr int kStkMaxStages = 8;

// sqrt for the tile geometry: the 5-operation one-ulp form of remap_image.cuh (dsqrt_nz) with
// the guard for s == 0 (a pixel exactly on the centre) and subnormal s that dsqrt_pos has.
__device__ __forceinline__ double dsqrt_fast0(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double g = s * y;
    const double e = fma(-g, y, 1.0);
    const double q = fma(e, 0.375, 0.5) * e;
    const double r = fma(g, q, g);
    return (__double2hiint(s) < 0x00100000) ? 0.0 : r;
}

// L2 eviction priority of the source boxes: 2 = evict_last (default), 1 = evict_first, 0 = no hint.
// Box halos are read again by the neighbouring tiles; measured on 64 x 4096^2 (config-2 model):
// evict_last 0.679 / 0.967 of the HBM peak (exact / float32 blend) against 0.671 / 0.945 without a
// hint, evict_first 0.687 / 0.91; configs 4 and 5 do not move (profiles/r2/ab_stack_l2hint.txt).
#ifndef DCB_STK_L2HINT
#define DCB_STK_L2HINT 2
#endif
#ifndef DCB_STK_SCALED
#define DCB_STK_SCALED 1
#endif
// see scaled_f64 (remap_image.cuh): the double v * 2^-896 of a non-negative finite float v
__device__ __forceinline__ double scaled_tap(float f) {
    unsigned long long w;
    asm("mul.wide.u32 %0, %1, 536870912;" : "=l"(w) : "r"(__float_as_uint(f)));
    return __longlong_as_double((long long)w);
}

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
    // float64 coordinates with the exact blend (unwarp_slice_backward): the complements 1 - tx and
    // 1 - ty are kept too instead of being re-formed for every slice (2 of 13 fp64 operations per
    // pixel and slice)
#ifndef DCB_STK_T64W
#define DCB_STK_T64W 1
#endif
    static constexpr bool kT64W = kT64 && BLEND == DCB_BLEND_EXACT && DCB_STK_T64W;
    static constexpr int kDoubles = (kW4 || kT64W) ? 4 : (kT64 ? 2 : 0);
    // (Sampling the fp64 blends from a float64 copy of the staged box -- one
    // conversion per source pixel instead of one per tap -- was measured with a
    // CTA barrier per slice and rejected: 48 % vs 56 % of the HBM peak on
    // 64 x 4096^2, exact blend.)
};

// RINT: integer image, SciPy's round-half-away-from-zero on the fp64 sum (a
// template parameter: as a run-time flag it cost a DSETP, three FSEL and an
// XU-pipe FRND per pixel and slice, profiles/r1/ncu_stack_v7.txt).
template <int ORDER, int BLEND, bool ROUND32, bool RINT_, int COLS>
__global__ void __launch_bounds__(kThreads, DCB_STK_MINB)
    remap_stack_kernel(const __grid_constant__ RemapParams p,
                       const __grid_constant__ CUtensorMap tmap) {
    using CT = typename std::conditional<ROUND32, float, double>::type;
    using SW = StackWeights<ORDER, BLEND, ROUND32>;
    constexpr int RINT = RINT_ ? 1 : 0;
    constexpr int ROWS = kStkPx / COLS;      // rows per warp
    constexpr int TW = 32 * COLS;            // tile width
    constexpr int TH = kWarps * ROWS;        // tile height

    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [nstage raw boxes][full][empty][red]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.nstage * p.stage_bytes);
    uint64_t *empty = full + kStkMaxStages;
    int *red = reinterpret_cast<int *>(empty + kStkMaxStages);  // [2][4][kWarps]

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool staged = p.nstage > 0;
    const uint32_t S = (uint32_t)p.nstage;
    // Work items are CLAIMED from a counter, one ahead of the item being processed, instead of
    // being dealt round-robin: the CTAs in flight then always hold the most recent consecutive
    // items -- neighbouring tiles at the same slices, whose box halos are L2 hits -- however far
    // their speeds have drifted apart.  (Dealt statically, a 2048-slice stack ran at 0.46 of the
    // HBM peak where a 64-slice one reached 0.60 with the same tiles: after a few hundred items
    // per CTA the neighbours were no longer concurrent and every halo came from DRAM.)
    __shared__ int s_item[2];
    if (threadIdx.x == 0) {
        s_item[0] = (int)blockIdx.x;   // the first item needs no claim
        if (staged) {
            for (int s = 0; s < p.nstage; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], kWarps);
            }
            fence_mbar_init();
            tma_prefetch_desc(&tmap);
        }
    }
    __syncthreads();

    uint32_t fills = 0;  // slices that went through the ring so far (CTA-uniform)
    int par = 0;
    const int wmax = p.W - 1;
    const int hmax = p.H - 1;
    const int bw = p.bw;
    const int y_end = p.row0 + p.nrows;

    for (;; par ^= 1) {
        const int item = s_item[par];
        if (item >= p.ntiles) break;
        // (read by the other threads after the barrier every item has: the bounding-box exchange
        // of staged launches, the one at the end of the loop body otherwise)
        if (threadIdx.x == 0) s_item[par ^ 1] = (int)gridDim.x + (int)atomicAdd(&p.sched[0], 1u);
        // (Claiming the items in strips of 4 / 8 / 16 tile columns, row by row inside a strip -- so that
        // a tile and the one below it, whose boxes share most of their rows under a sheared map, read
        // the same slice within a slice or two of each other -- changes nothing on any BASELINE shape:
        // profiles/r2/ab_stack_strip.txt.)
        const int txi = item % p.tiles_x;
        const int rest = item / p.tiles_x;
        const int tyi = rest % p.tiles_y;
        const int zci = rest / p.tiles_y;
        const int x_base = txi * TW + lane;
        const int y_base = p.row0 + tyi * TH + warp * ROWS;
        const int z0 = zci * p.zchunk;
        const int nz = min(p.zchunk, p.D - z0);

        // ---- geometry of the thread's 8 pixels, once per item ----------------------
        int x0[kStkPx], r0[kStkPx];            // floor column, first tap row (window-clamped)
        unsigned dxm = 0, dym = 0;             // bit i: second tap is one column right / one row down
        CT tx[kStkPx], ty[kStkPx];
        {
            double xu[COLS], xu2[COLS];
#pragma unroll
            for (int k = 0; k < COLS; ++k) {
                xu[k] = (double)min(x_base + 32 * k, wmax) - p.rad.xc;  // :138 (edge lanes redo a valid pixel)
                xu2[k] = __dmul_rn(xu[k], xu[k]);
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const double yu = (double)min(y_base + j, y_end - 1) - p.rad.yc;  // :139
                const double yu2 = __dmul_rn(yu, yu);
                double r[COLS], f[COLS];
#pragma unroll
                for (int k = 0; k < COLS; ++k) r[k] = dsqrt_fast0(__dadd_rn(xu2[k], yu2));  // :141
                radial_factor<COLS>(p.rad.a, p.rad.n, r, f);                              // :142-143
#pragma unroll
                for (int k = 0; k < COLS; ++k) {  // :144-145 (image, chunk) / :219-220 (slice)
                    const int i = j * COLS + k;
                    const CT cx = clamp_coord<CT>(fma(f[k], xu[k], p.rad.xc), wmax);
                    const CT cy = clamp_coord<CT>(fma(f[k], yu, p.rad.yc), hmax);
                    // floor and fraction of a coordinate in [0, 2^23) without the XU pipe (F2I / I2F):
                    // a round-down add of 2^23 (2^52 for float64 coordinates) leaves floor(c) in the
                    // low mantissa bits; both subtractions are exact
                    int xi, yi;
                    split_floor(cx, xi, tx[i]);
                    split_floor(cy, yi, ty[i]);
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
            constexpr unsigned kAllPx = (1u << kStkPx) - 1u;
            const bool mine = (ORDER == 1) && (dxm != kAllPx || dym != kAllPx);
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
            // (A second, smaller box for the tiles whose footprint fits it -- two tensor maps,
            // box pitch chosen per item -- was measured: the source bytes moved per output tile
            // fall from 3x to 1.6x on configs 4 / 5 and nothing gets faster, the exact blends
            // lose 4-8 % to the extra per-item state; profiles/r2/bench_stack_two_boxes_s4.jsonl.)
            fits = (mxx - bx0 + 1 <= bw) && (mxy - by0 + 1 <= p.bh);
        }

        if (fits) {
            // the first slices start moving before the weights are formed (hides one TMA latency
            // per item, which matters when a chunk is only a few slices deep)
            const uint32_t base = fills;
            if (threadIdx.x == 0) {
                const int npre = min(nz, p.nstage - 1);
                for (int s = 0; s < npre; ++s) {
                    const uint32_t g = base + s, st = g % S;
                    if (g >= S) mbar_wait(&empty[st], ((g / S) - 1u) & 1u);
                    mbar_expect_tx(&full[st], p.box_bytes);
#if DCB_STK_L2HINT
                    tma_load_3d_hint(smem + (size_t)st * p.stage_bytes, &tmap, bx0, by0 - p.yorg, z0 + s,
                                     &full[st], l2_policy(DCB_STK_L2HINT - 1));
#else
                    tma_load_3d(smem + (size_t)st * p.stage_bytes, &tmap, bx0, by0 - p.yorg, z0 + s,
                                &full[st]);
#endif
                }
            }
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
                    if (SW::kT64W) {
                        wd[i][SW::kT64W ? 2 : 0] = __dsub_rn(1.0, (double)tx[i]);
                        wd[i][SW::kT64W ? 3 : 0] = __dsub_rn(1.0, (double)ty[i]);
                    }
                } else if (SW::kF32) {
                    wf[i][0] = (float)tx[i];
                    wf[i][1] = (float)ty[i];
                }
            }
            float *orow = p.dst + (long long)z0 * p.dst_slice +
                          (long long)(y_base - p.row0) * p.dst_pitch + x_base;
            bool colok[COLS];
#pragma unroll
            for (int k = 0; k < COLS; ++k) colok[k] = (x_base + 32 * k <= wmax);

            // ONE slice loop for every tile.  (Measured, profiles/r1/stack_loop_ab.txt: as soon
            // as the loop exists in several specialised copies under CTA-uniform branches the
            // compiler can no longer prove the warp converged at __syncwarp -- BRA.DIV instead
            // of WARPSYNC.ALL -- and the kernel runs 35-50 % slower although each copy executes
            // fewer instructions.)  Only the tap loads differ between interior tiles (taps at
            // +1 / +bw) and tiles where some +1 tap was folded back onto the last row / column
            // (per-pixel offsets): a uniform branch around phase 1.
            // ring position of slice iz: stage st, phase parity ph -- kept as counters (a
            // division by the run-time stage count per slice costs three XU-pipe operations)
            uint32_t st = base % S, ph = (base / S) & 1u;
            for (int iz = 0; iz < nz; ++iz, orow += p.dst_slice) {
                const uint32_t g = base + iz;
                // Measured oddity (profiles/r1/stack_loop_ab.txt): the fp32 / nearest paths,
                // which run at 85-90 % of the HBM peak, are 10-15 % FASTER when every warp
                // re-derives the ring position with a division here (23.1 vs 27.0 us per
                // 4096^2 slice) -- neither a plain delay before the wait nor keeping the
                // counters out of the uniform datapath reproduces it; the fp64 blends
                // (XU-bound) are 6 % faster with the counters.  Each takes what measured best.
                if (SW::kF32 || SW::kNearest) st = g % S, ph = (g / S) & 1u;
                mbar_wait(&full[st], ph);
                const float *tile = reinterpret_cast<const float *>(smem + (size_t)st * p.stage_bytes);
                const float *tile1 = tile + bw;  // the row below (interior tiles)
                // phase 1: every tap of the thread's 8 pixels into registers, then the
                // stage goes back to the producer BEFORE the arithmetic (the copy of a
                // later slice overlaps the blend of this one)
                float fa[kStkPx], fb[ORDER == 1 ? kStkPx : 1], fc[ORDER == 1 ? kStkPx : 1],
                    fd[ORDER == 1 ? kStkPx : 1];
                if (ORDER == 0) {
#pragma unroll
                    for (int i = 0; i < kStkPx; ++i) fa[i] = tile[off[i]];
                } else if (edge) {
#pragma unroll
                    for (int i = 0; i < kStkPx; ++i) {
                        const int ox = (int)((dxm >> i) & 1u);
                        const int oy = ((dym >> i) & 1u) ? bw : 0;
                        const float *q = tile + off[i];
                        fa[i] = q[0], fb[ORDER == 1 ? i : 0] = q[ox];
                        fc[ORDER == 1 ? i : 0] = q[oy], fd[ORDER == 1 ? i : 0] = q[oy + ox];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < kStkPx; ++i) {
                        fa[i] = tile[off[i]], fb[ORDER == 1 ? i : 0] = tile[off[i] + 1];
                        fc[ORDER == 1 ? i : 0] = tile1[off[i]], fd[ORDER == 1 ? i : 0] = tile1[off[i] + 1];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
                // phase 2: the blend
                float v[kStkPx];
                if (ORDER == 0) {
#pragma unroll
                    for (int i = 0; i < kStkPx; ++i) v[i] = fa[i];
                } else if (SW::kF32) {
#pragma unroll
                    for (int i = 0; i < kStkPx; ++i) {
                        const float top = fmaf(fb[i] - fa[i], wf[i][0], fa[i]);
                        const float bot = fmaf(fd[i] - fc[i], wf[i][0], fc[i]);
                        v[i] = finish_f32(fmaf(bot - top, wf[i][1], top), RINT);
                    }
                } else {
                    // The float64 blends with the taps as doubles a, b, c, d; `post` undoes the
                    // scale of the taps (1 for converted taps).  Every operation is SciPy's,
                    // in SciPy's order.
                    auto blend8 = [&](auto widen, const double post) {
#pragma unroll
                        for (int i = 0; i < kStkPx; ++i) {
                            const double a = widen(fa[i]), b = widen(fb[ORDER == 1 ? i : 0]),
                                         c = widen(fc[ORDER == 1 ? i : 0]), d = widen(fd[ORDER == 1 ? i : 0]);
                            double s;
                            if (SW::kW4) {
                                s = __dmul_rn(a, wd[i][0]);
                                s = __dadd_rn(s, __dmul_rn(b, wd[i][1]));
                                s = __dadd_rn(s, __dmul_rn(c, wd[i][2]));
                                s = __dadd_rn(s, __dmul_rn(d, wd[i][3]));
                            } else if (BLEND == DCB_BLEND_LERP64) {
                                const double top = fma(b - a, wd[i][0], a);
                                const double bot = fma(d - c, wd[i][0], c);
                                s = fma(bot - top, wd[i][1], top);
                            } else {
                                // float64 coordinates: SciPy's two-step products, every step rounded
                                const double wx1 = wd[i][0], wy1 = wd[i][1];
                                const double wx0 = SW::kT64W ? wd[i][SW::kT64W ? 2 : 0] : __dsub_rn(1.0, wx1);
                                const double wy0 = SW::kT64W ? wd[i][SW::kT64W ? 3 : 0] : __dsub_rn(1.0, wy1);
                                s = __dmul_rn(__dmul_rn(a, wy0), wx0);
                                s = __dadd_rn(s, __dmul_rn(__dmul_rn(b, wy0), wx1));
                                s = __dadd_rn(s, __dmul_rn(__dmul_rn(c, wy1), wx0));
                                s = __dadd_rn(s, __dmul_rn(__dmul_rn(d, wy1), wx1));
                            }
                            if (post != 1.0) s = __dmul_rn(s, post);   // exact: a power of two, no underflow
                            v[i] = finish_f64(s, RINT);
                        }
                    };
#if DCB_STK_SCALED
                    if (BLEND == DCB_BLEND_LERP64 || DCB_STK_SCALED == 2) {
                    // Taps widened on the integer pipe in the scaled domain (scaled_f64 in
                    // remap_image.cuh: u * 2^29 read as a double is v * 2^-896): the four
                    // conversions per pixel and slice kept the 16-lane XU pipe busier than HBM
                    // (5 XU operations per pixel; profiles/r1/ncu_stack_v14_exact_16x4096.txt).
                    // A power-of-two scale commutes with every rounding as long as nothing becomes
                    // a denormal double, which holds when every tap of the thread is a positive
                    // normal float of at least 2^-78 (fp32-rounded coordinates: weights >= 2^-48)
                    // or 2^-30 (float64 coordinates: each weight >= 2^-43): one min / max over the
                    // 32 taps; otherwise (zeros, negative values, Inf / NaN, tiny values) the warp
                    // converts this slice's taps the old way.
                    uint32_t hi = 0u, lo = 0xffffffffu;
#pragma unroll
                    for (int i = 0; i < kStkPx; ++i) {
                        const uint32_t ua = __float_as_uint(fa[i]), ub = __float_as_uint(fb[ORDER == 1 ? i : 0]);
                        const uint32_t uc = __float_as_uint(fc[ORDER == 1 ? i : 0]), ud = __float_as_uint(fd[ORDER == 1 ? i : 0]);
                        hi = __vimax3_u32(hi, ua, ub);
                        hi = __vimax3_u32(hi, uc, ud);
                        lo = __vimin3_u32(lo, ua, ub);
                        lo = __vimin3_u32(lo, uc, ud);
                    }
                    constexpr uint32_t kTapMin = SW::kW4 ? 0x18800000u : 0x30800000u;
                    if (__all_sync(0xffffffffu, lo >= kTapMin && hi < 0x7f800000u)) {
                        blend8([](float f) { return scaled_tap(f); }, 0x1p896);
                    } else {
                        blend8([](float f) { return (double)f; }, 1.0);
                    }
                    } else {
                        // (the exact blends: with their 4 + 4 weight registers per pixel the two
                        // copies of the blend spill -- 39.0 against 36.3 us per 4096^2 slice,
                        // profiles/r2/bench_stack_scaled_s1.jsonl)
                        blend8([](float f) { return (double)f; }, 1.0);
                    }
#else
                    blend8([](float f) { return (double)f; }, 1.0);
#endif
                }
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    if (y_base + j < y_end) {
                        float *o = orow + (long long)j * p.dst_pitch;
#pragma unroll
                        for (int k = 0; k < COLS; ++k)
                            if (colok[k]) __stcs(o + 32 * k, v[j * COLS + k]);
                    }
                }
                if (threadIdx.x == 0 && iz + p.nstage - 1 < nz) {
                    // refill the stage consumed one iteration ago (all warps are past it or
                    // about to be): keeps nstage-1 slices in flight
                    // global fill number g + S - 1 goes into the stage before st; it was last
                    // used by fill g - 1, whose release is completion (g + S - 1) / S - 1 of `empty`
                    const uint32_t sn = (st == 0u) ? S - 1u : st - 1u;
                    if (g > 0u) mbar_wait(&empty[sn], (st == 0u) ? (ph ^ 1u) : ph);
                    mbar_expect_tx(&full[sn], p.box_bytes);
#if DCB_STK_L2HINT
                    tma_load_3d_hint(smem + (size_t)sn * p.stage_bytes, &tmap, bx0, by0 - p.yorg,
                                     z0 + iz + p.nstage - 1, &full[sn], l2_policy(DCB_STK_L2HINT - 1));
#else
                    tma_load_3d(smem + (size_t)sn * p.stage_bytes, &tmap, bx0, by0 - p.yorg,
                                z0 + iz + p.nstage - 1, &full[sn]);
#endif
                }
                if (++st == S) {
                    st = 0u;
                    ph ^= 1u;
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
                        v[i] = finish_f32(fmaf(bot - top, wf[i][1], top), RINT);
                    } else if (SW::kW4) {
                        double s = __dmul_rn((double)a, wd[i][0]);
                        s = __dadd_rn(s, __dmul_rn((double)b, wd[i][1]));
                        s = __dadd_rn(s, __dmul_rn((double)c, wd[i][2]));
                        s = __dadd_rn(s, __dmul_rn((double)d, wd[i][3]));
                        v[i] = finish_f64(s, RINT);
                    } else if (BLEND == DCB_BLEND_LERP64) {
                        const double da = a, db = b, dc = c, dd = d;
                        const double top = fma(db - da, wd[i][0], da);
                        const double bot = fma(dd - dc, wd[i][0], dc);
                        v[i] = finish_f64(fma(bot - top, wd[i][1], top), RINT);
                    } else {
                        const double wx1 = wd[i][0], wy1 = wd[i][1];
                        const double wx0 = __dsub_rn(1.0, wx1), wy0 = __dsub_rn(1.0, wy1);
                        double s = __dmul_rn(__dmul_rn((double)a, wy0), wx0);
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn((double)b, wy0), wx1));
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn((double)c, wy1), wx0));
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn((double)d, wy1), wx1));
                        v[i] = finish_f64(s, RINT);
                    }
                }
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    if (y_base + j < y_end) {
                        float *o = orow + (long long)j * p.dst_pitch;
#pragma unroll
                        for (int k = 0; k < COLS; ++k)
                            if (x_base + 32 * k <= wmax) __stcs(o + 32 * k, v[j * COLS + k]);
                    }
                }
            }
        }
        if (!staged) __syncthreads();
    }
    // the last CTA through here leaves the counters as it found them
    if (threadIdx.x == 0) {
        const unsigned d = atomicAdd(&p.sched[1], 1u);
        if (d == gridDim.x - 1) {
            p.sched[0] = 0u;
            p.sched[1] = 0u;
        }
    }
}

}  // namespace dcb
