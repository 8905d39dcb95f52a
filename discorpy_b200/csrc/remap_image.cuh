// remap_image.cuh -- the single-image backward-remap kernel (sm_100a), v5 (round 2).
//
// Serves a1 `unwarp_image_backward` (postprocessing.py:111-148) and a4
// `correct_perspective_image` (:444-492) for one float32 image, or a range of
// its output rows.  Numerics are those of remap.cuh (same helpers); what is
// different is the schedule.  The kernel is bound by instruction issue, not by memory: at the
// 70 %-of-HBM target an SM must retire 2 output pixels per clock, and on this part an fp64
// instruction holds a sub-partition's issue port for its whole dispatch (2.1 cycles, 3.1 with three
// distinct source registers) -- measured: every DFMA removed from the per-pixel loop saves its
// pipe time in full, other instructions 1 cycle (profiles/r2/ablation_raw2_r2a1.txt).
//
//   * a PLAN per (model, geometry), built once and cached on the device (api.cu): per tile the
//     source box, per tile row a verified degree-5 interpolant of the radial factor (RowPatch),
//     per tile a cost estimate, per CTA a cost-balanced static range of tiles;
//   * persistent CTAs (2 per SM): eight sampling warps and one producer warp that streams the
//     tiles' boxes (TMA) and plan records (bulk copy) through four shared-memory stages;
//   * tiles come from the CTA's static range first, then from a pool claimed with an atomic
//     counter (whole tiles, halves at the very end) -- CTAs finish within about 1 us of each other;
//   * programmatic dependent launch: the prologue, and the plan half of the first four tiles, run
//     under the tail of the preceding launch in the stream;
//   * verified rows ("patch path"): 7 DFMA per pixel for the coordinates, rounding to the
//     float32 grid / floor / fraction by one magic addition each, taps from the raw float32
//     stage, the exact blend as six certified FMAs in the scaled domain (scaled_f64: one integer
//     multiply-wide per tap instead of a conversion), the float32 result by an integer shift;
//   * every other row (image borders, the distortion centre, failed certificates, negative or
//     non-finite pixels) takes the exact coordinate chain and SciPy's own sum, or the generic
//     clip + global gather of remap.cuh, so no result depends on the plan's estimates.
#pragma once
#include "remap.cuh"

namespace dcb {

// > 0: every staged box of the single-image kernel is this many floats wide, so the pitch of the
// staged tile is a compile-time constant in the sampling loop (tap addresses become immediates).
#ifndef DCB_IMG_BOXW
#define DCB_IMG_BOXW 144   // 0 in A/B builds: box as wide as the footprint bound, run-time pitch
#endif
constexpr int kImgBoxW = DCB_IMG_BOXW;

struct ImageParams {
    const float *src;
    float *dst;
    long long src_pitch, dst_pitch;  // elements
    int H, W;
    int row0, nrows;   // output rows produced by this launch
    int yorg, ylast;   // image rows held by src (src points at row yorg)
    int tiles_y, ntiles;  // tiles are numbered column-major: t = txi * tiles_y + tyi
    int bw, bh;        // staged box, bw % 4 == 0; bw == 0 => never stage
    unsigned box_bytes, stage_bytes;
    int dbg;           // A/B builds (-DDCB_AB): ablation switches, 0 otherwise
    int deal;          // 1: uneven tiles (clipped regions, fallback box): every tile comes from the pool
    int nstatic;       // tiles [0, nstatic) are split into one contiguous range per CTA
    int npool_full;    // pool: tiles [nstatic, nstatic + npool_full) are claimed whole,
    int npool_units;   //       the rest in halves; npool_units claims in total
    unsigned *sched;   // [0] next pool unit, [1] CTAs done (the last one resets both)
    int pool_depth;    // units a producer keeps in flight while it claims from the pool (2..NBUF)
    int plan_ready;    // 1: the plan was complete before this launch was enqueued (cache hit)
    int fast;          // 1: rows certified by the producer take the patch path (see RowPatch)
    int rint;          // 1: integer image, round half away from zero (finish_f64 in remap.cuh)
    unsigned long long *stats;   // diagnostics (dcb_image_stats), NULL normally
    const void *plan;            // TilePlan<TH>[ntiles] (image_plan_kernel), device memory; followed by
    const int2 *plan_boxes;      //   {bx0, use ? by0 : ~by0} per tile, compact (what the producer reads)
    const int *plan_starts;      //   first tile of every CTA's static range, gridDim.x + 1 entries
    RadialDev rad;
    PerspDev per;
};

// The tile prologue (run by every thread for every tile, ~10 % of all instructions):
// 1: the exact path's window constants are formed inside its branch and the output row address with
//    one 32 x 32 -> 64 bit multiply -- 12 of ~107 instructions (exact 42.9 -> 42.6 us, float32 blend
//    31.5 -> 30.8: profiles/r2/ab2_lazywin.txt);
// 2 (default): the row address as one UNSIGNED multiply-add and the row count with a shift instead
//    of a signed division -- 7 more (exact 42.7 -> 41.3 us: the register allocation of the exact
//    kernel's row loop changes with it; profiles/r2/ab2_lazy2.txt);
// 0: everything in the prologue, for A/B builds
#ifndef DCB_IMG_LAZYWIN
#define DCB_IMG_LAZYWIN 2
#endif
// A/B builds: L2 eviction priority of the source boxes (1 evict_first, 2 evict_last, 0 no hint);
// -1: the kernel's own choice (see issue_box)
#ifndef DCB_IMG_L2HINT
#define DCB_IMG_L2HINT (-1)
#endif
#ifndef DCB_IMG_EXACT_RAW
#define DCB_IMG_EXACT_RAW 2   // 0: float64 tiles widened by the producers (round 1), 1: raw tiles + F2F, 2: raw tiles, scaled domain
#endif
// Ablations of the patch path for attribution runs (A/B builds only, results are WRONG):
// 1 no stores, 2 every tap read from one address, 4 F = c0 (no Horner chain), 8 no blend (first tap)
#ifndef DCB_ABL
#define DCB_ABL 0
#endif
// output stores: streaming (evict-first) by default; -DDCB_IMG_STCS=0 for A/B builds
#ifndef DCB_IMG_STCS
#define DCB_IMG_STCS 1
#endif
#if DCB_IMG_STCS
#define DCB_IMG_STORE(ptr, val) __stcs((ptr), (val))
#else
#define DCB_IMG_STORE(ptr, val) (*(ptr) = (val))
#endif

struct TileBox {
    int bx0, by0;  // image coordinates of box element (0,0)
    int use;       // 1: TMA issued for this tile, 0: sample straight from global
    int shx;       // patch path: 23 - binade of the tile's x coordinates (0: no row of the tile is fast)
    uint32_t mhi_x;  // high word of 1.5 * 2^(52 - shx)
    int pad[3];
};

// floor of a float v in [0, 2^23) without the XU pipe: one round-toward-minus-
// infinity add puts floor(v) in the low mantissa bits of t = 2^23 + floor(v)
// (0x4B000000 is the bit pattern of 2^23).  For v < 0 the bits of t fall below
// 0x4B000000, for v >= 2^23 or NaN far above: both fail the box test.
__device__ __forceinline__ float floor_magic(float v) { return __fadd_rd(v, 8388608.0f); }

// clip(x, 0, vmax) of an fp32 coordinate on the integer pipe: non-negative
// floats order like their bit patterns, negative ones are negative integers.
__device__ __forceinline__ float clamp_bits(float v, int vmax_bits) {
    return __int_as_float(__vimin_s32_relu(__float_as_int(v), vmax_bits));
}

// sqrt for s > 0 (callers keep s away from 0, see MapEval<MAP_RADIAL>), within
// one ulp: with y0 = MUFU.RSQ64H(s) and e = 1 - s*y0^2 (|e| < 2^-19),
// sqrt(s) = s*y0 * (1 - e)^(-1/2) = g * (1 + e/2 + 3e^2/8 + O(e^3)).  5 fp64
// operations + 1 MUFU, every DFMA in a two-register form (the three-register
// form issues at 70 % rate, profiles/r1/microbench_fp64_operand_forms.txt).
// It need not be correctly rounded: the radius only enters F(r), whose
// sensitivity F'(r) r is <= a few percent of F, so one ulp of r moves the
// coordinate by far less than the rounding of the Horner steps themselves
// (dcb_selftest_sqrt_fast bounds the error; the 4096^2 / 8192^2 parity runs
// count the fp32 coordinates that differ from the reference: 0).
__device__ __forceinline__ double dsqrt_nz(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double g = s * y;
    const double e = fma(-g, y, 1.0);
    const double q = fma(e, 0.375, 0.5) * e;
    return fma(g, q, g);
}

// (nx / den, ny / den) with one reciprocal: MUFU.RCP64H seed (20 bits), two
// Newton steps (-> ~2^-80 before rounding), then one Markstein correction per
// quotient -- 10 fp64 operations + 1 MUFU for both quotients, correctly
// rounded except in rare last-bit cases (like dsqrt: far below what can flip an
// fp32-rounded coordinate).  A zero / non-finite / subnormal denominator takes
// the IEEE division, so the projective singularity behaves like the reference.
__device__ __forceinline__ bool ddiv_pair(double nx, double ny, double den, double &qx, double &qy) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    double e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    const double x0 = __dmul_rn(nx, r), y0 = __dmul_rn(ny, r);
    qx = fma(fma(-den, x0, nx), r, x0);
    qy = fma(fma(-den, y0, ny), r, y0);
    // den = 0, huge, tiny or NaN <=> the exponent of 1/den leaves [2^-995, 2^995]: the caller
    // redoes such pixels with the IEEE division (one integer test instead of two DSETP and a
    // divergent branch per pixel -- profiles/r1/ncu_persp_v10.txt)
    const unsigned ex = ((unsigned)__double2hiint(r) >> 20) & 0x7ffu;
    return (ex - 28u) > (2018u - 28u);
}

// ---------------------------------------------------------------------------------------------
// The patch path (round 2): verified cheap coordinates from a plan.
//
// The exact coordinate costs 12 fp64 operations, one MUFU and two conversions per pixel, then two
// more conversions and six FP32 operations for floor / fraction -- 31.7 us of the 57 us kernel
// (profiles/r1).  What the reference fixes is only the float32 ROUNDING of the float64 coordinate
// (postprocessing.py:144-145), and the map depends on the calibration and the image geometry
// alone, not on the pixels.  So the library keeps a PLAN per (model, geometry) -- built once by
// image_plan_kernel, cached on the device (api.cu), reused for every frame unwarped with that
// calibration, which is how the reference is used (one calibration, thousands of projections):
//   * per tile the staged box (what the producer warp used to derive from 9 probe points);
//   * per tile row a degree-5 interpolant of F(r(x)) along the 128 pixels (six Chebyshev nodes,
//     exact evaluations) and, per 32-pixel segment, a bit that says the interpolant was VERIFIED
//     pixel by pixel against the exact evaluation: same float32 coordinates (with a margin that
//     covers the last-ulp differences between the exact path's own variants), one float32
//     binade, 2x2 footprint inside the staged box.
// The sampling warps then spend 5 DFMA (Horner in tau) + 2 DFMA per pixel of a verified row;
// rounding to the float32 grid, floor and fraction come from ONE fp64 addition of
// 1.5 * 2^(52 - sh) (sh = 23 - binade: the sum's low word is the coordinate in units of its
// float32 ulp), a shift, a mask and one exact subtraction -- no conversion, no MUFU, no box
// test.  Rows that were not verified (image borders, the distortion centre, binade crossings,
// strong magnification) take the exact path below, so no result depends on the interpolant.
// ---------------------------------------------------------------------------------------------
#include "patch_const.inc"

struct __align__(16) RowPatch {   // one per tile row (64 bytes)
    double c[6];        // F(tau) = sum c_i tau^i, tau = (x - x_tile - 63.5) / 64
    uint32_t info;      // bits 0..3: segment k (pixels lane + 32 k) verified; bits 8..12: shy
    uint32_t mhi_y;     // high word of 1.5 * 2^(52 - shy)
    uint32_t mky;       // (1 << shy) - 1
    uint32_t e32y;      // (150 - shy) << 23: float exponent whose ulp is 2^-shy
};

// A non-negative finite float32 v with bit pattern u, read as the 64-bit integer u * 2^29 (high word
// u >> 3, low word u << 29), IS the double v * 2^-896: the float's exponent field lands in the low
// eight bits of the double's, its fraction in the top 23 bits of the double's fraction -- exactly,
// for zero, denormals (they become denormal doubles of the same relative value) and normals alike.
// One IMAD.WIDE on the FMA pipe instead of an F2F on the 16-lane XU pipe, no special cases; what
// it does not cover is a set sign bit, Inf and NaN (callers test the largest bit pattern).  Sums
// and products of such values by weights in [0, 1] are the true results times 2^-896 with the
// same roundings (a power-of-two scale commutes with every rounding; where a scaled intermediate
// is a denormal double its error is at most 2^-1075 absolute, i.e. 2^-179 of the unscaled value,
// far inside the blend certificate's 32 ulp).
__device__ __forceinline__ double scaled_f64(uint32_t u) {
    unsigned long long w;
    asm("mul.wide.u32 %0, %1, 536870912;" : "=l"(w) : "r"(u));
    return __longlong_as_double((long long)w);
}

// high word of 1.5 * 2^(52 - sh); the same constant with mantissa 1.0 is 2^(52 - sh) (0x80000 less)
__device__ __forceinline__ uint32_t magic_hi(int sh) { return ((uint32_t)(1075 - sh) << 20) | 0x80000u; }
__device__ __forceinline__ int binade_of(double v) {   // floor(log2 v) of a positive normal double
    return (int)(((uint32_t)__double2hiint(v) >> 20) & 0x7ffu) - 1023;
}

// Per-thread column terms of the map (constant down a tile) and the row
// evaluation producing UNCLIPPED fp32 coordinates for 4 pixels.
template <int MAP, int NT>
struct MapEval;

template <int NT>
struct MapEval<MAP_RADIAL, NT> {
    // (only xu is kept per tile: with tiles claimed from a pool the column changes with every tile,
    // and the verified rows of the patch path never need xu^2)
    double xu[kCols];
    __device__ __forceinline__ void set_columns(const ImageParams &p, const int (&x)[kCols]) {
#pragma unroll
        for (int k = 0; k < kCols; ++k) xu[k] = (double)x[k] - p.rad.xc;
    }
    // yd: the row as a double
    __device__ __forceinline__ void row(const ImageParams &p, double yd, float (&xf)[kCols],
                                        float (&yf)[kCols]) const {
        double xq[kCols], yq[kCols];
        row64(p, yd, xq, yq);
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
            xf[k] = __double2float_rn(xq[k]);
            yf[k] = __double2float_rn(yq[k]);
        }
    }
    // the unrounded float64 coordinates of the row (the patch path's exact redo uses them as is)
    __device__ __forceinline__ void row64(const ImageParams &p, double yd, double (&xq)[kCols],
                                          double (&yq)[kCols]) const {
        const double yu = __dsub_rn(yd, p.rad.yc);
        const double yu2 = __dmul_rn(yu, yu);
        double r[kCols], f[kCols], xu2[kCols];
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
            const double q = __dmul_rn(xu[k], xu[k]);
            // r = 0 only at a pixel sitting exactly on the centre.  Keeping
            // xu^2 >= 1e-300 gives r = 1e-150 there, hence the same F (= a0 after
            // rounding) and the same coordinates (F * 0 + centre), and it leaves
            // every other s = xu^2 + yu^2 bit-identical (1e-300 is far below half
            // an ulp of any non-zero yu^2) -- so the chain below needs no zero test.
            xu2[k] = (q < 1e-300) ? 1e-300 : q;
        }
        if (NT > 2) {
            // F(r) = E(s) + r O(s), s = r^2 (the rounded sum the sqrt is taken of): the two short
            // Horner chains in s do not wait for the sqrt -- same number of operations as Horner in
            // r, shorter dependent chain (57.5 instead of 58.6 us); the CPU emulation and the GPU
            // parity runs count 0 changed fp32 coordinates at 4096^2 and 8192^2
#pragma unroll
            for (int k = 0; k < kCols; ++k) {
                const double s = __dadd_rn(xu2[k], yu2);
                r[k] = dsqrt_nz(s);
                constexpr int NE = (NT + 1) / 2, NO = NT / 2;   // even / odd coefficient counts
                double e = p.rad.a[2 * (NE - 1)], o = p.rad.a[2 * (NO - 1) + 1];
#pragma unroll
                for (int i = NE - 2; i >= 0; --i) e = fma(e, s, p.rad.a[2 * i]);
#pragma unroll
                for (int i = NO - 2; i >= 0; --i) o = fma(o, s, p.rad.a[2 * i + 1]);
                f[k] = fma(o, r[k], e);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kCols; ++k) r[k] = dsqrt_nz(__dadd_rn(xu2[k], yu2));
            if (NT > 0) {
#pragma unroll
                for (int k = 0; k < kCols; ++k) f[k] = horner<(NT > 0 ? NT : 1)>(p.rad.a, r[k]);
            } else {
                radial_factor<kCols>(p.rad.a, p.rad.n, r, f);
            }
        }
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
            xq[k] = fma(f[k], xu[k], p.rad.xc);
            yq[k] = fma(f[k], yu, p.rad.yc);
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
    __device__ __forceinline__ void row(const ImageParams &p, double yd, float (&xf)[kCols],
                                        float (&yf)[kCols]) const {
        const double c2y = __dmul_rn(p.per.c[1], yd);
        const double c5y = __dmul_rn(p.per.c[4], yd);
        const double c8y = __dmul_rn(p.per.c[7], yd);
        bool odd = false;
#pragma unroll
        for (int k = 0; k < kCols; ++k) {
            const double den = __dadd_rn(__dadd_rn(c7x[k], c8y), 1.0);
            const double nx = __dadd_rn(__dadd_rn(c1x[k], c2y), p.per.c[2]);
            const double ny = __dadd_rn(__dadd_rn(c4x[k], c5y), p.per.c[5]);
            double qx, qy;
            odd |= ddiv_pair(nx, ny, den, qx, qy);
            xf[k] = __double2float_rn(qx);
            yf[k] = __double2float_rn(qy);
        }
        if (__any_sync(0xffffffffu, odd)) {   // projective singularity in this row: IEEE divisions
#pragma unroll
            for (int k = 0; k < kCols; ++k) {
                const double den = __dadd_rn(__dadd_rn(c7x[k], c8y), 1.0);
                const double nx = __dadd_rn(__dadd_rn(c1x[k], c2y), p.per.c[2]);
                const double ny = __dadd_rn(__dadd_rn(c4x[k], c5y), p.per.c[5]);
                xf[k] = __double2float_rn(__ddiv_rn(nx, den));
                yf[k] = __double2float_rn(__ddiv_rn(ny, den));
            }
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

// Exact F at (xu, yu) for the plan (any number of terms; correctly rounded sqrt).
__device__ __forceinline__ double radial_f(const RadialDev &m, double xu, double yu) {
    const double r = dsqrt_pos(fma(xu, xu, yu * yu));
    double f = 0.0;
    for (int t = m.n - 1; t >= 0; --t) f = fma(f, r, m.a[t]);
    return f;
}

// One record of the plan per tile: what the kernel needs to know about the tile before it
// touches a pixel.  The producer copies it next to the staged box with one bulk copy.
template <int TH>
struct __align__(16) TilePlan {
    TileBox box;
    RowPatch rows[TH];
};

// the source box of a tile from 9 probe points (all 32 lanes of one warp)
template <int MAP, int TH>
__device__ __forceinline__ TileBox place_tile_box(const ImageParams &p, int x_lo, int y_lo,
                                                  int lane) {
    const int wmax = p.W - 1, y_end = p.row0 + p.nrows;
    const int x_hi = min(x_lo + kTileW - 1, wmax), y_hi = min(y_lo + TH - 1, y_end - 1);
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
    const bool use = p.bw > 0 && (mxx + 2 - bx0 < p.bw) && (mxy + 2 - by0 < p.bh);
    return TileBox{bx0, by0, use ? 1 : 0, 0, 0u, {0, 0, 0}};
}

// distance of v from the nearest float32 rounding boundary (v > 0 normal in float32)
__device__ __forceinline__ double dist_to_f32_boundary(double v) {
    const float f = __double2float_rn(v);
    const uint32_t fb = (uint32_t)__float_as_int(f);
    double half_ulp = __hiloint2double((int)(((fb >> 23) & 0xffu) + 1023u - 127u - 24u) << 20, 0);
    // below a power of two the grid is twice as fine
    if ((fb & 0x7fffffu) == 0u && v < (double)f) half_ulp *= 0.5;
    return half_ulp - fabs(v - (double)f);
}

// Plan builder: one CTA of 8 warps per tile; warp w verifies tile rows w * TH/8 ... (one row at a
// time: lanes 0..5 evaluate the nodes, every lane then checks its four pixels).
template <int MAP, int TH>
__global__ void __launch_bounds__(kThreads)
    image_plan_kernel(const __grid_constant__ ImageParams p, TilePlan<TH> *__restrict__ plan,
                      int2 *__restrict__ boxes, unsigned *__restrict__ cost,
                      unsigned long long *__restrict__ stats) {
    __shared__ TileBox sbox;
    __shared__ unsigned scost;
    if (threadIdx.x == 0) scost = 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x;
    const int txi = t / p.tiles_y, tyi = t - txi * p.tiles_y;
    const int x_lo = txi * kTileW, y_lo = p.row0 + tyi * TH;
    const int wmax = p.W - 1, y_end = p.row0 + p.nrows;
    TilePlan<TH> &out = plan[t];
    if (warp == 0) {
        TileBox bx = place_tile_box<MAP, TH>(p, x_lo, y_lo, lane);
        if (p.fast && bx.use && x_lo + kTileW - 1 <= wmax) {
            // x binade of the tile: the one of its centre pixel
            double xd;
            if (MAP == MAP_RADIAL) {
                const double xu = ((double)x_lo + 63.5) - p.rad.xc;
                const double yu = ((double)y_lo + (TH - 1) * 0.5) - p.rad.yc;
                xd = fma(radial_f(p.rad, xu, yu), xu, p.rad.xc);
            } else {
                const double xm = (double)x_lo + 63.5, ym = (double)y_lo + (TH - 1) * 0.5;
                const double den = p.per.c[6] * xm + p.per.c[7] * ym + 1.0;
                xd = den > 0.0 ? (p.per.c[0] * xm + p.per.c[1] * ym + p.per.c[2]) / den : -1.0;
            }
            const int ex = binade_of(xd);
            if (xd > 0.0 && ex >= 0 && ex <= 22) {
                bx.shx = 23 - ex;
                bx.mhi_x = magic_hi(bx.shx);
            }
        }
        bx.pad[0] = txi, bx.pad[1] = tyi;   // the sampling warps need no division to find the tile
        if (lane == 0) {
            sbox = bx;
            out.box = bx;
            boxes[t] = make_int2(bx.bx0, bx.use ? bx.by0 : ~bx.by0);
        }
    }
    __syncthreads();
    const TileBox box = sbox;
    const int lim_x = box.use ? min(p.bw - 1, wmax - box.bx0) : 0;
    const int lim_y = box.use ? min(p.bh - 1, p.ylast - box.by0) : 0;
    constexpr int RPW = TH / kWarps;
    unsigned n_full = 0, n_part = 0, n_rows = 0, n_genx = 0;
    unsigned n_why[3] = {0, 0, 0};   // rows not verified: tile not eligible, row binade, pixels
    for (int rr = 0; rr < RPW; ++rr) {
        const int r = warp * RPW + rr, y = y_lo + r;
        RowPatch rp;
#pragma unroll
        for (int i = 0; i < 6; ++i) rp.c[i] = 0.0;
        rp.info = 0, rp.mhi_y = 0, rp.mky = 0, rp.e32y = 0;
        if (box.shx != 0 && y < y_end) {   // warp-uniform
            const double yrow = (double)y;
            const double yu = __dsub_rn(yrow, p.rad.yc);
            const double xm = ((double)x_lo + 63.5) - p.rad.xc;   // xu at tau = 0
            // projective map: what is interpolated along the row is w = 1 / (c6 x + c7 y + 1), the
            // numerators are linear in x (MapEval<MAP_PERSP> holds c0 x and c3 x per column)
            const double c8y = __dmul_rn(p.per.c[7], yrow);
            const double a_row = fma(p.per.c[1], yrow, p.per.c[2]);   // c1 y + c2
            const double e_row = fma(p.per.c[4], yrow, p.per.c[5]);   // c4 y + c5
            // node values on lanes 0..5, coefficient i on lane i, then broadcast
            double fn;
            if (MAP == MAP_RADIAL) {
                fn = radial_f(p.rad, fma(64.0, kPatchNodesX[lane % 6], xm), yu);
            } else {
                const double xn = fma(64.0, kPatchNodesX[lane % 6], (double)x_lo + 63.5);
                fn = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(p.per.c[6], xn), c8y), 1.0));
            }
            double ci = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j)
                ci = fma(kPatchVinvX[lane % 6][j], __shfl_sync(0xffffffffu, fn, j), ci);
            double c[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) c[i] = __shfl_sync(0xffffffffu, ci, i);
            // y binade of the row: the one of its centre pixel (from the interpolant)
            const double ydm = MAP == MAP_RADIAL
                                   ? fma(c[0], yu, p.rad.yc)
                                   : __dmul_rn(c[0], fma(p.per.c[3], (double)x_lo + 63.5, e_row));
            const int ey = binade_of(ydm);
            const bool row_ok = ydm > 0.0 && ey >= 0 && ey <= 22;   // warp-uniform
            const int shx = box.shx, shy = 23 - (row_ok ? ey : 0);
            const double My = __hiloint2double((int)magic_hi(shy), 0);
            const double gy = __hiloint2double((1023 - shy) << 20, 0);
            uint32_t mask = 0;
#pragma unroll
            for (int k = 0; k < kCols; ++k) {
                const int x = x_lo + lane + 32 * k;
                const double xu = (double)x - p.rad.xc;
                // the sampling warps' arithmetic, operation for operation
                const double tau = ((double)(lane + 32 * k) - 63.5) * (1.0 / 64.0);
                double f = fma(c[5], tau, c[4]);
                f = fma(f, tau, c[3]);
                f = fma(f, tau, c[2]);
                f = fma(f, tau, c[1]);
                f = fma(f, tau, c[0]);
                double xp, yp, xe, ye, mx, my;
                if (MAP == MAP_RADIAL) {
                    xp = fma(f, xu, p.rad.xc), yp = fma(f, yu, p.rad.yc);
                    // the exact coordinates
                    const double fe = radial_f(p.rad, xu, yu);
                    xe = fma(fe, xu, p.rad.xc), ye = fma(fe, yu, p.rad.yc);
                    // margin: the exact path's own variants (E/O split, one-ulp sqrt, Horner lengths
                    // known at compile time) differ from this evaluation by a few ulp of F
                    mx = 0x1p-48 * (fabs(xu) + fabs(xe)), my = 0x1p-48 * (fabs(yu) + fabs(ye));
                } else {
                    const double xdbl = (double)x;
                    const double c1x = __dmul_rn(p.per.c[0], xdbl), c4x = __dmul_rn(p.per.c[3], xdbl);
                    xp = __dmul_rn(f, __dadd_rn(c1x, a_row));
                    yp = __dmul_rn(f, __dadd_rn(c4x, e_row));
                    // the exact coordinates, MapEval<MAP_PERSP>'s operation order with IEEE divisions
                    const double den = __dadd_rn(__dadd_rn(__dmul_rn(p.per.c[6], xdbl), c8y), 1.0);
                    const double nxe = __dadd_rn(__dadd_rn(c1x, __dmul_rn(p.per.c[1], yrow)), p.per.c[2]);
                    const double nye = __dadd_rn(__dadd_rn(c4x, __dmul_rn(p.per.c[4], yrow)), p.per.c[5]);
                    const bool den_ok = den > 0x1p-900 && den < 0x1p900;
                    xe = den_ok ? __ddiv_rn(nxe, den) : -1.0;
                    ye = den_ok ? __ddiv_rn(nye, den) : -1.0;
                    // margin: the kernel's reciprocal + Markstein quotients are within an ulp
                    mx = 0x1p-48 * fabs(xe), my = 0x1p-48 * fabs(ye);
                }
                const uint32_t ny = (uint32_t)__double2loint(__dadd_rn(yp, My));
                bool oky = row_ok && xe > 0.0 && ye > 0.0 && xp > 0.0 &&
                           dist_to_f32_boundary(xe) > mx && dist_to_f32_boundary(ye) > my;
                // same float32 value (this also pins the binade: a float32 of another binade is
                // not a multiple of the grid or lies outside [2^23, 2^24] grid units)
                oky = oky && (double)ny * gy == (double)__double2float_rn(ye) && ny >= (1u << 23) &&
                      ny <= (1u << 24);
                const int iy = (int)(ny >> shy) - box.by0;
                oky = oky && (unsigned)iy < (unsigned)lim_y;
                // x with the tile's binade (variant 0) and with the pixel's own (variant 1)
#pragma unroll
                for (int var = 0; var < 2; ++var) {
                    const uint32_t hx = (uint32_t)__double2hiint(xp);
                    const int sx = var ? 1046 - (int)(hx >> 20) : shx;
                    const uint32_t mhx = var ? (hx & 0x7ff00000u) + 0x01d80000u : box.mhi_x;
                    bool ok = oky && sx >= 1 && sx <= 23;
                    const int sxc = min(max(sx, 1), 23);
                    const uint32_t nx =
                        (uint32_t)__double2loint(__dadd_rn(xp, __hiloint2double((int)mhx, 0)));
                    const double gxv = __hiloint2double((1023 - sxc) << 20, 0);   // 2^-sx
                    ok = ok && (double)nx * gxv == (double)__double2float_rn(xe) &&
                         nx >= (1u << 23) && nx <= (1u << 24);
                    // 2x2 footprint inside the staged box and the image
                    const int ix = (int)(nx >> sxc) - box.bx0;
                    ok = ok && (unsigned)ix < (unsigned)lim_x;
                    mask |= __all_sync(0xffffffffu, ok) ? (1u << (k + 4 * var)) : 0u;
                }
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) rp.c[i] = c[i];
            rp.info = mask | ((uint32_t)(shy & 31) << 8);
            rp.mhi_y = magic_hi(shy);
            rp.mky = (1u << shy) - 1u;
            rp.e32y = (uint32_t)(150 - shy) << 23;
            n_why[1] += !row_ok;
            n_why[2] += row_ok && mask == 0u;
            n_full += (mask & 0xfu) == 0xfu || (mask & 0xf0u) == 0xf0u;
            n_genx += (mask & 0xfu) != 0xfu && (mask & 0xf0u) == 0xf0u;
            n_part += !((mask & 0xfu) == 0xfu || (mask & 0xf0u) == 0xf0u) && mask != 0u;
        }
        n_why[0] += box.shx == 0 && y < y_end;
        n_rows += y < y_end;
        if (lane == 0) out.rows[r] = rp;
    }
    if (stats != nullptr && lane == 0) {
        for (int i = 0; i < 3; ++i) atomicAdd(stats + i, (unsigned long long)n_why[i]);
        atomicAdd(stats + 5, (unsigned long long)n_full);
        atomicAdd(stats + 6, (unsigned long long)n_part);
        atomicAdd(stats + 7, (unsigned long long)n_rows);
    }
    // what the tile will cost the sampling warps, in quarters of a verified row: rows that take
    // the per-pixel x binade a quarter more, rows on the exact path three times as much
    // (profiles/r2/timeline_r2v.txt: the CTAs owning the columns through the distortion centre and
    // the 2048 binade crossing finished 4 us after everyone else with equal tile counts)
    if (lane == 0) atomicAdd(&scost, 4u * (n_full - n_genx) + 5u * n_genx + 12u * (n_rows - n_full));
    __syncthreads();
    if (threadIdx.x == 0) cost[t] = scost;
}

// Cuts the first nstatic tiles into `nranges` contiguous ranges of equal cost (tile i goes to the
// range its cost midpoint falls into): starts[b] = first tile of range b, starts[nranges] = nstatic.
// One CTA of 1024 threads; every thread owns a chunk of consecutive tiles.
static __global__ void __launch_bounds__(1024)
    image_plan_ranges_kernel(const unsigned *__restrict__ cost, int nstatic, int nranges,
                             int *__restrict__ starts) {
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x;
    const int chunk = (nstatic + 1023) / 1024;
    const int lo = min(tid * chunk, nstatic), hi = min(lo + chunk, nstatic);
    unsigned long long sum = 0;
    for (int i = lo; i < hi; ++i) sum += cost[i];
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) {   // exclusive scan of 1024 partial sums: a few microseconds, once per plan
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            const unsigned long long v = part[i];
            part[i] = run;
            run += v;
        }
        starts[0] = 0;
        starts[nranges] = nstatic;
    }
    __syncthreads();
    unsigned long long total = part[1023];
    for (int i = min(1023 * chunk, nstatic); i < nstatic; ++i) total += cost[i];
    if (total == 0) total = 1;
    auto range_of = [&](unsigned long long before, unsigned c) -> int {
        const unsigned long long mid = 2ull * before + c;   // twice the cost midpoint
        return (int)min((unsigned long long)(nranges - 1), mid * (unsigned long long)nranges / (2ull * total));
    };
    unsigned long long before = part[tid];
    // range of the tile just before this chunk (range 0 "before" tile 0)
    int prev = 0;
    if (lo > 0 && lo < nstatic) {
        const unsigned cprev = cost[lo - 1];
        prev = range_of(before - cprev, cprev);
    }
    for (int i = lo; i < hi; ++i) {
        const unsigned c = cost[i];
        const int b = range_of(before, c);
        for (int r = prev + 1; r <= b; ++r) starts[r] = i;
        prev = b;
        before += c;
    }
    // ranges beyond the last tile's are empty
    if (hi == nstatic && lo < hi)
        for (int r = prev + 1; r < nranges; ++r) starts[r] = nstatic;
    if (nstatic == 0 && tid == 0)
        for (int r = 1; r < nranges; ++r) starts[r] = 0;
}

#ifndef DCB_IMG_STAGES
#define DCB_IMG_STAGES 4
#endif
constexpr int kRawStages = DCB_IMG_STAGES;   // raw float32 stages = tile buffers in flight (2 or 4)
// 8 sampling warps + one producer warp (copies only)
constexpr int kImgThreads = kThreads + 32;

// Shared-memory layout (host side must agree, see plan_and_launch_image in api.cu):
//   [raw 0] .. [raw 3][tail]
//   tail : uint64_t raw_full[4] (copy landed); 48 bytes unused (stages come back through named
//          barriers 1..4); uint32_t unit_g[4] (row groups of the unit in each stage);
//          TilePlan<TH> rec[4]
__host__ __device__ constexpr size_t image_rec_bytes(int th) {
    return sizeof(TileBox) + (size_t)th * sizeof(RowPatch);
}
__host__ __device__ constexpr size_t image_tail_bytes(int th) {
    return 96 + kRawStages * image_rec_bytes(th);
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr int kLogCtas = 4, kLogEvents = 64;   // per-warp event log of the first CTAs (timeline builds)
#ifdef DCB_IMG_TIMELINE
#define DCB_LOG(idx) \
    do { if (p.stats != nullptr && blockIdx.x < kLogCtas && lane == 0 && (idx) < kLogEvents) \
        p.stats[8 + 8 * 1024 + ((blockIdx.x * 10 + warp) * kLogEvents) + (idx)] = global_ns(); } while (0)
#else
#define DCB_LOG(idx) do { } while (0)
#endif
constexpr int kTimelineCtas = 1024;   // diagnostics: p.stats[8 + 8 b ..] = start, first tile, end, SM, waits of CTA b

// 1-D bulk copy global -> shared, completion on an mbarrier (16-byte aligned, size % 16 == 0)
// The plan records are loaded with the L2 evict_last hint (2; 1: evict_first, 0: none): the 8.6 MB plan of a
// calibration is read again by every launch while 128 MB of image data pass through the L2 in between
// (exact 43.00 -> 42.94 us, nearest 26.5 -> 26.3: profiles/r2/ab2_planhint.txt)
#ifndef DCB_IMG_PLANHINT
#define DCB_IMG_PLANHINT 2
#endif
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes,
                                          uint64_t *bar) {
#if DCB_IMG_PLANHINT
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_policy(DCB_IMG_PLANHINT - 1))
        : "memory");
#else
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
#endif
}

// NT > 0: number of polynomial terms known at compile time (coefficients become
// constant-bank operands of the DFMAs); NT == 0: any p.rad.n through a switch.
// TH: tile height (16 or 32 rows); MINB: resident CTAs per SM the register
// allocation is held to.
//
// Warp-specialised: warps 0..7 only evaluate coordinates and sample from the raw float32 stages;
// warp 8 (the producer) walks this CTA's tiles: per tile one TMA load of the staged box and one
// bulk copy of the tile's plan record, both completing on the stage's mbarrier (raw_full); the
// sampling warps hand a stage back with a `bar.arrive` on the stage's named barrier, on which the
// producer sleeps (`bar.sync`) instead of polling.  There is no CTA-wide
// barrier in the tile loop.  (Round 1 widened every landed box into float64 tiles with two
// producer warps; the per-warp event log showed that widening -- 1.7 us per tile -- was as slow as
// the sampling it fed, profiles/r2/timeline_r2t3_event_log.txt.  The exact blend now widens its
// taps itself, in the scaled domain, see scaled_f64.)
template <int MAP, int ORDER, int BLEND, int NT, int TH, int MINB>
__global__ void __launch_bounds__(kImgThreads, MINB)
    remap_image_kernel(const __grid_constant__ ImageParams p,
                       const __grid_constant__ CUtensorMap tmap) {
    constexpr bool PATCH = kImgBoxW > 0;   // the patch path exists
    constexpr int RPW = TH / kWarps;  // rows per sampling warp and tile
    constexpr int NBUF = kRawStages;   // stages = tile buffers the samplers read from
    constexpr int LOGB = kRawStages == 2 ? 1 : 2;
    static_assert((1 << LOGB) == NBUF, "buffer count");
    constexpr uint32_t kRecBytes = (uint32_t)sizeof(TilePlan<TH>);
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *tail = smem + kRawStages * (size_t)p.stage_bytes;
    uint64_t *raw_full = reinterpret_cast<uint64_t *>(tail);         // [4] TMA bytes landed
    TilePlan<TH> *rec = reinterpret_cast<TilePlan<TH> *>(tail + 96); // [NBUF]
    const TilePlan<TH> *plan = reinterpret_cast<const TilePlan<TH> *>(p.plan);

    const int lane = threadIdx.x & 31;
    // (through a shuffle: the compiler then knows the warp index is warp-uniform and keeps what
    // derives from it -- row counts, tile addresses, the stores' memory descriptor -- on the
    // uniform datapath)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const bool staged = p.bw > 0;
    const int wmax = p.W - 1;
    const int y_end = p.row0 + p.nrows;

    // Tile scheduling (round 2).  Tiles are numbered column-major.  The first p.nstatic of them
    // are split statically, one contiguous range per CTA, so that consecutive tiles sit below each
    // other and share their column terms; the rest form a POOL that the producer warps claim from
    // with an atomic counter once their range is exhausted -- whole tiles first, the last ones in
    // halves (RPW / 2 row groups).  The per-CTA timeline of the purely static split
    // (profiles/r2/timeline_r2t*.txt) showed CTAs finishing up to 11 us apart in a 53 us kernel:
    // 13 against 14 tiles, the second CTA of an SM running 4-9 % behind the first (oldest-first
    // warp scheduling), tiles with binade crossings.  While it claims from the pool a producer
    // keeps only one unit in flight beyond the one being sampled: a unit claimed is a unit this
    // CTA must do, and four claimed ahead would decide the balance 10 us before the end.
    // p.deal (uneven tiles: clipped regions, fallback boxes) => nstatic == 0, everything is pooled.
    // The static ranges are cut by estimated cost, not by tile count (image_plan_ranges_kernel).
    uint32_t *unit_g = reinterpret_cast<uint32_t *>(tail + 80);   // [4] row groups of the unit in stage b
    constexpr uint32_t kNoUnit = 0xffffffffu;

    if (threadIdx.x == 0) {
#ifdef DCB_IMG_TIMELINE
        if (p.stats != nullptr && blockIdx.x < kTimelineCtas) p.stats[8 + 8 * blockIdx.x] = global_ns();
#endif
        for (int b = 0; b < 4; ++b) {
            mbar_init(&raw_full[b], 1);
        }
        fence_mbar_init();
        if (staged) tma_prefetch_desc(&tmap);
    }
    __syncthreads();
    // Programmatic dependent launch (api.cu launches with programmatic stream serialization): the
    // CTAs of the next launch in the stream may take the SM slots this grid's CTAs free, and run
    // everything above, before this grid has finished; no read of the source and no store happens
    // before the preceding grid has completed and flushed (griddepcontrol.wait below); the plan is
    // read earlier only when it was complete before this launch was enqueued (p.plan_ready).
    // Without the launch attribute both instructions do nothing.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == kWarps) {
        // =========================== producer warp ===================================
        // lane 0 starts the copies of one tile in two steps: its plan record (and the read of its
        // box origin, from a compact array: sixteen tiles share a cache line, so along a static
        // range the read the TMA issue depends on is an L1 hit), then, when the tile is staged, its
        // box into stage st -- both completing on raw_full[st]
        auto issue_plan = [&](int tile, int st) -> int2 {
            const TilePlan<TH> *tp = plan + tile;
            const int2 bx = __ldg(p.plan_boxes + tile);   // bx0, use ? by0 : ~by0
            mbar_expect_tx(&raw_full[st], kRecBytes + (bx.y >= 0 ? p.box_bytes : 0u));
            bulk_load(&rec[st], tp, kRecBytes, &raw_full[st]);
            return bx;
        };
        auto issue_box = [&](const int2 bx, int st) {
            if (bx.y >= 0) {
                // L2 eviction priority of the source boxes: the bilinear kernels read them
                // evict_first (a box is needed once, its lines make room for the output's write
                // combining: exact 43.4 -> 43.0 us, float32 blend 32.1 -> 31.5), the nearest-neighbour
                // kernel -- the one closest to the memory roofline, where halo hits count -- without
                // a hint (26.5 us; 27.3 with evict_first).  profiles/r2/ab2_il2hint.txt
                constexpr int kHint = DCB_IMG_L2HINT >= 0 ? DCB_IMG_L2HINT : (ORDER == 1 ? 1 : 0);
                if constexpr (kHint != 0)
                    tma_load_3d_hint(smem + (size_t)st * p.stage_bytes, &tmap, bx.x, bx.y - p.yorg, 0,
                                     &raw_full[st], l2_policy(kHint - 1));
                else
                    tma_load_3d(smem + (size_t)st * p.stage_bytes, &tmap, bx.x, bx.y - p.yorg, 0,
                                &raw_full[st]);
            }
        };
        // A plan that was complete before this launch was enqueued may be read while the preceding
        // grid is still running: the static range and the plan half of its first NBUF tiles are
        // under way before the grid dependency is waited for, only the source boxes come after.
        int t0 = 0, n_static = 0, k0 = 0;
        int2 bxs[NBUF];
        auto read_range = [&]() {
            if (lane == 0 && p.nstatic > 0) {
                t0 = __ldg(p.plan_starts + blockIdx.x);
                n_static = __ldg(p.plan_starts + blockIdx.x + 1) - t0;
            }
        };
        if (p.plan_ready) {
            read_range();
            if (lane == 0) {
                k0 = min(n_static, NBUF);
#pragma unroll
                for (int k = 0; k < NBUF; ++k) {
                    if (k < k0) {
                        unit_g[k] = (uint32_t)RPW << 8;
                        bxs[k] = issue_plan(t0 + k, k);
                    }
                }
                // ... and the first box is asked into L2 already.  A prefetch delivers nothing to
                // this kernel -- the box is read after the grid dependency, through the same
                // coherent L2 -- so it is harmless even if the preceding grid is still writing the
                // source; when it is not, the DRAM latency of the box every sampling warp of this
                // CTA waits for passes under that grid's tail.  (44.8 -> 43.4 us; all NBUF boxes:
                // 44.2, they compete with the preceding grid's stores -- profiles/r2/ab2_l2pf*.txt)
                if (k0 > 0 && bxs[0].y >= 0) tma_prefetch_l2_3d(&tmap, bxs[0].x, bxs[0].y - p.yorg, 0);
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (!p.plan_ready) read_range();
        if (lane == 0 && k0 > 0) {
            // the first box alone, the prefetches once it has landed: all 296 CTAs start at the same
            // moment, and four boxes each would queue the one everybody waits for behind 19 MB
            issue_box(bxs[0], 0);
            if (k0 > 1) mbar_wait(&raw_full[0], 0u);
#pragma unroll
            for (int k = 1; k < NBUF; ++k)
                if (k < k0) issue_box(bxs[k], k);
        }
        t0 = __shfl_sync(0xffffffffu, t0, 0);
        n_static = __shfl_sync(0xffffffffu, n_static, 0);
        k0 = __shfl_sync(0xffffffffu, k0, 0);
        // raw stage k % NBUF is the data buffer of this CTA's k-th unit: the samplers wait on
        // raw_full[k % NBUF]; the producer runs at most NBUF units ahead of the slowest sampling
        // warp (with two stages the samplers waited 11 % of their time for copies issued only one
        // tile earlier, profiles/r2/ncu_image_r2h_lerp32_two_stages.txt), p.pool_depth while it
        // claims
        // Stages come back through NAMED BARRIERS 1..NBUF, not an mbarrier: the eight sampling
        // warps `bar.arrive` on barrier 1 + stage when they are done with a unit, the producer
        // `bar.sync`s on it and is descheduled until then.  (Polling data_empty with
        // mbarrier.try_wait -- neither its suspend hint nor a nanosleep kept the warp asleep --
        // was 4.7 of the kernel's 62 instructions per pixel, issued on the sub-partition it shares
        // with two sampling warps: profiles/r2/ncu_image_r2ac_exact.txt, SASS lines 320-383.)
        // Completions are consumed strictly in order, one per unit.
        int released = 0;   // units 0 .. released-1 are known to have been handed back
        for (int k = k0;; ++k) {
            const int b = k & (NBUF - 1);
            const int depth = k < n_static ? NBUF : p.pool_depth;
            // unit k - depth must have been released by all samplers
            while (released <= k - depth) {
                asm volatile("bar.sync %0, %1;" ::"r"(1 + (released & (NBUF - 1))), "n"(kImgThreads) : "memory");
                ++released;
            }
            DCB_LOG(k);
            int more = 1;
            if (lane == 0) {
                if (k < n_static) {
                    unit_g[b] = (uint32_t)RPW << 8;
                    issue_box(issue_plan(t0 + k, b), b);
                } else {
                    const int v = (int)atomicAdd(&p.sched[0], 1u);
                    if (v >= p.npool_units) {
                        unit_g[b] = kNoUnit;   // tells the samplers to leave
                        mbar_arrive(&raw_full[b]);
                        more = 0;
                    } else if (v < p.npool_full) {
                        unit_g[b] = (uint32_t)RPW << 8;
                        issue_box(issue_plan(p.nstatic + v, b), b);
                    } else {
                        constexpr int HG = RPW >= 2 ? RPW / 2 : 1;   // row groups of half a tile
                        const int w = v - p.npool_full;
                        const int ga = RPW >= 2 ? (w & 1) * HG : 0;
                        unit_g[b] = (uint32_t)ga | ((uint32_t)(ga + HG) << 8);
                        issue_box(issue_plan(p.nstatic + p.npool_full + (RPW >= 2 ? (w >> 1) : w), b), b);
                    }
                }
            }
            more = __shfl_sync(0xffffffffu, more, 0);
            if (!more) break;
        }
        // the last CTA through here leaves the counters as it found them
        if (lane == 0) {
            const unsigned d = atomicAdd(&p.sched[1], 1u);
            if (d == gridDim.x - 1) {
                p.sched[0] = 0u;
                p.sched[1] = 0u;
            }
        }
        return;
    }

    // ============================== sampling warps ====================================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int txi = -1;   // tile column the per-thread column terms are set for
    MapEval<MAP, NT> ev;
    // patch path: the thread's four columns in the tile's variable tau = (x - x_tile - 63.5) / 64
    // (opaque to the compiler: under the register cap it would otherwise re-derive them from the
    // lane index -- four I2F and eight fp64 operations -- in every row)
    double tau[kCols];
#pragma unroll
    for (int k = 0; k < kCols; ++k) {
        tau[k] = ((double)(lane + 32 * k) - 63.5) * (1.0 / 64.0);
        asm volatile("" : "+d"(tau[k]));
    }

#ifdef DCB_IMG_TIMELINE   // -DDCB_IMG_TIMELINE builds only (tools/timeline_probe.py)
    unsigned long long dbg_t = 0, dbg_wait = 0, dbg_wmax = 0;
#endif
    int n = 0;   // units sampled (diagnostics)
    for (int i = 0;; ++i) {
        // unit i is ready: its record and box have landed in stage i % NBUF
        mbar_wait(&raw_full[i & (NBUF - 1)], (uint32_t)(i >> LOGB) & 1u);
        const uint32_t ug = unit_g[i & (NBUF - 1)];
        if (ug == kNoUnit) break;
        n = i + 1;
        DCB_LOG(2 * i);
#ifdef DCB_IMG_TIMELINE
        if (p.stats != nullptr && threadIdx.x == 0 && blockIdx.x < kTimelineCtas) {
            const unsigned long long now = global_ns();
            if (i == 0) {
                p.stats[8 + 8 * blockIdx.x + 1] = now;
            } else {   // time warp 0 waited for tile i after finishing tile i - 1
                dbg_wait += now - dbg_t;
                dbg_wmax = max(dbg_wmax, now - dbg_t);
            }
        }
#endif
        const TilePlan<TH> &trec = rec[i & (NBUF - 1)];
        const TileBox box = trec.box;
        const int tyi = box.pad[1];
        if (box.pad[0] != txi) {   // CTA-uniform: a new tile column
            txi = box.pad[0];
            int xs[kCols];
#pragma unroll
            for (int k = 0; k < kCols; ++k) xs[k] = min(txi * kTileW + lane + 32 * k, wmax);
            ev.set_columns(p, xs);
        }
        // ---- sample tile i ---------------------------------------------------------
        {
            const int x_base = txi * kTileW + lane;
            // this warp's rows of the tile: warp + kWarps g for the row groups g of this CTA's share
            const int ga = (int)(ug & 0xffu), gb = (int)(ug >> 8);
            const int y_base = p.row0 + tyi * TH + ga * kWarps + warp;
            const int rows_left = y_end - y_base;   // rows y_base + kWarps j exist for kWarps j < rows_left
#if DCB_IMG_LAZYWIN >= 2
            const int nrow = min(gb - ga, (rows_left + kWarps - 1) >> 3);  // warp-uniform, may be <= 0
            static_assert(kWarps == 8, "row groups of eight warps");
#else
            const int nrow = min(gb - ga, (rows_left + kWarps - 1) / kWarps);  // warp-uniform, may be <= 0
#endif
#if !DCB_IMG_LAZYWIN
            // bits(2^23 + n) - magic = n - box origin
            const int magic_x = 0x4B000000 + box.bx0, magic_y = 0x4B000000 + box.by0;
            // fast-path window: footprint inside the box and strictly inside the image
            const int lim_x = box.use ? min(p.bw - 1, wmax - box.bx0) : 0;
            const int lim_y = box.use ? min(p.bh - 1, p.ylast - box.by0) : 0;
#endif
            const int bw = kImgBoxW > 0 ? kImgBoxW : p.bw;
            const int sb = i & (NBUF - 1);
            const float *rawt =
                reinterpret_cast<const float *>(smem + (size_t)sb * p.stage_bytes);
            const bool full_w = __all_sync(0xffffffffu, txi * kTileW + kTileW - 1 <= wmax);  // CTA-uniform
#if DCB_IMG_LAZYWIN
            // (row pitches are below 2^31 elements: one 32 x 32 -> 64 bit multiply)
#if DCB_IMG_LAZYWIN >= 2
            // (all three terms are non-negative and below 2^31: one unsigned multiply-add)
            float *orow = p.dst + ((unsigned long long)(unsigned)(y_base - p.row0) * (unsigned)p.dst_pitch +
                                   (unsigned)x_base);
#else
            float *orow = p.dst + (long long)(y_base - p.row0) * (long long)(int)p.dst_pitch + x_base;
#endif
#else
            float *orow = p.dst + (long long)(y_base - p.row0) * p.dst_pitch + x_base;
#endif
            double yd = (double)y_base;
            // patch path, per tile: the x rounding constants and the tap address of image pixel (0, 0)
            const int shx_t = box.shx;
            const uint32_t mkx_t = (1u << shx_t) - 1u, e32x_t = (uint32_t)(150 - shx_t) << 23;
            const uint32_t org = (uint32_t)(box.by0 * bw + box.bx0);
            const uint32_t base_s = smem_u32(rawt) - 4u * org;
            const RowPatch *prow = trec.rows + ga * kWarps + warp;
            unsigned n_bfail = 0;   // diagnostics, see p.stats
#ifndef DCB_IMG_UNROLL
#define DCB_IMG_UNROLL 1
#endif
            constexpr int kRowUnroll = DCB_IMG_UNROLL;   // A/B builds; 1 measured best
            // The row loop exists twice: for tiles that are 128 pixels wide (all but the last
            // column of an image whose width is not a multiple of 128) the four stores of a row
            // are unconditional -- predicated, ptxas moved their memory descriptor from a vector
            // register pair to uniform registers eight times per row (2 of 70 instructions per
            // pixel).
            auto tile_rows = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll kRowUnroll
            // (Rows claimed one at a time from a shared counter by whichever sampling warp is free,
            // instead of this static split, were measured: 61.5 / 59.4 / 49.4 us against 57.4 / 53.8 /
            // 44.0 -- the atomic and the lost incremental row state cost more than the 8 % the
            // warps wait for each other at tile boundaries.)
            for (int j = 0; j < nrow; ++j, yd += (double)kWarps, orow += kWarps * p.dst_pitch, prow += kWarps) {
                float v[kCols];
                bool done = false;
                if constexpr (PATCH) if (shx_t != 0) {
                    // ------------------- the patch path (see RowPatch) -------------------
                    // (all four loads before the branch on the first: one shared-memory round trip
                    // per row instead of two)
                    uint4 inf;
                    double2 c01, c23, c45;
                    {
                        const uint32_t pa = smem_u32(prow);
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+48];"
                                     : "=r"(inf.x), "=r"(inf.y), "=r"(inf.z), "=r"(inf.w) : "r"(pa));
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c01.x), "=d"(c01.y) : "r"(pa));
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(c23.x), "=d"(c23.y) : "r"(pa));
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+32];" : "=d"(c45.x), "=d"(c45.y) : "r"(pa));
                    }
                    // bits 0..3: verified with the tile's x binade; bits 4..7: verified with the
                    // x binade taken per pixel (rows crossing a power of two in x)
                    if ((inf.x & 0xfu) == 0xfu || (inf.x & 0xf0u) == 0xf0u) {
                        double xq[kCols], yq[kCols];
                        {
                            // per-row terms: radial yu; projective c1 y + c2 and c4 y + c5
                            double ra, rb;
                            if constexpr (MAP == MAP_RADIAL) {
                                ra = __dsub_rn(yd, p.rad.yc), rb = 0.0;
                            } else {
                                ra = fma(p.per.c[1], yd, p.per.c[2]), rb = fma(p.per.c[4], yd, p.per.c[5]);
                            }
#pragma unroll
                            for (int k = 0; k < kCols; ++k) {
                                double f = fma(c45.y, tau[k], c45.x);
                                if (DCB_ABL & 4) {
                                    f = c01.x + c45.y;
                                } else {
                                    f = fma(f, tau[k], c23.y);
                                    f = fma(f, tau[k], c23.x);
                                    f = fma(f, tau[k], c01.y);
                                    f = fma(f, tau[k], c01.x);
                                }
                                if constexpr (MAP == MAP_RADIAL) {
                                    xq[k] = fma(f, ev.xu[k], p.rad.xc);
                                    yq[k] = fma(f, ra, p.rad.yc);
                                } else {   // f interpolates 1 / (c6 x + c7 y + 1)
                                    xq[k] = __dmul_rn(f, __dadd_rn(ev.c1x[k], ra));
                                    yq[k] = __dmul_rn(f, __dadd_rn(ev.c4x[k], rb));
                                }
                            }
                        }
                        const int shy = (int)(inf.x >> 8) & 31;
                        const uint32_t mky = inf.z, e32y = inf.w;
                        const double My = __hiloint2double((int)inf.y, 0);
                        uint32_t accb = 0xffffffffu, tapmax = 0u;
                        // GENX: the x binade (shift, rounding constant, masks) from each pixel's own
                        // exponent instead of the tile's: 8 more integer operations per pixel
                        auto sample4 = [&](auto genx_tag) {
                            constexpr bool GENX = decltype(genx_tag)::value;
#pragma unroll
                            for (int k = 0; k < kCols; ++k) {
                                const uint32_t hx = (uint32_t)__double2hiint(xq[k]);
                                const int shx = GENX ? 1046 - (int)(hx >> 20) : shx_t;
                                const uint32_t mhx = GENX ? (hx & 0x7ff00000u) + 0x01d80000u : box.mhi_x;
                                const uint32_t mkx = GENX ? (1u << shx) - 1u : mkx_t;
                                const uint32_t e32x = GENX ? (uint32_t)(150 - shx) << 23 : e32x_t;
                                const double Mx = __hiloint2double((int)mhx, 0);
                                // round to the float32 grid: the low word is the coordinate in ulp32
                                const double ux = __dadd_rn(xq[k], Mx), uy = __dadd_rn(yq[k], My);
                                const uint32_t nx = (uint32_t)__double2loint(ux);
                                const uint32_t ny = (uint32_t)__double2loint(uy);
                                if (ORDER == 0) {
                                    const uint32_t xi = (nx + (1u << (shx - 1))) >> shx;
                                    const uint32_t yi = (ny + (1u << (shy - 1))) >> shy;
                                    float t;
                                    asm("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(base_s + 4u * (yi * bw + xi)));
                                    v[k] = t;
                                    continue;
                                }
                                const uint32_t xi = nx >> shx, yi = ny >> shy;
                                const uint32_t fx = nx & mkx, fy = ny & mky;
                                if (BLEND == DCB_BLEND_EXACT && DCB_IMG_EXACT_RAW == 2) {
                                    // The exact blend in the SCALED domain (see scaled_f64): taps are
                                    // widened by one integer multiply-wide each (no conversion, no
                                    // float64 tile), the certified FMA blend runs on v * 2^-896, and the
                                    // float32 result is the sum's bit pattern shifted back with a
                                    // half-ulp added -- the certificate excludes ties, and the float32
                                    // rounding boundary sits at bit 28 of the low word whether the
                                    // result is a float32 normal or denormal.
                                    // (fractions through I2F on the XU pipe plus an exponent step were
                                    // measured: 46.0 us against 44.6, profiles/r2/ab_frac_i2f.txt)
                                    const double tx = __dsub_rn(__hiloint2double(__double2hiint(ux), (int)fx), Mx);
                                    const double ty = __dsub_rn(__hiloint2double(__double2hiint(uy), (int)fy), My);
                                    const uint32_t qa = (DCB_ABL & 2) ? base_s + 4u * (org + (yi & 1u))
                                                                      : base_s + 4u * (yi * bw + xi);
                                    uint32_t a, b, c, d;
                                    asm("ld.shared.b32 %0, [%1];" : "=r"(a) : "r"(qa));
                                    if (DCB_ABL & 8) {
                                        tapmax = max(tapmax, a ^ (uint32_t)__double2loint(tx) ^ (uint32_t)__double2loint(ty));
                                        v[k] = __uint_as_float(a);
                                        continue;
                                    }
                                    asm("ld.shared.b32 %0, [%1+4];" : "=r"(b) : "r"(qa));
                                    asm("ld.shared.b32 %0, [%1+%2];" : "=r"(c) : "r"(qa), "n"(4 * kImgBoxW));
                                    asm("ld.shared.b32 %0, [%1+%2];" : "=r"(d) : "r"(qa), "n"(4 * kImgBoxW + 4));
                                    // a set sign bit, Inf or NaN among the taps: the exact row below
                                    tapmax = __vimax3_u32(tapmax, a, b);
                                    tapmax = __vimax3_u32(tapmax, c, d);
                                    const double sd = lerp_fma(scaled_f64(a), scaled_f64(b), scaled_f64(c),
                                                               scaled_f64(d), tx, ty);
                                    // + half a float32 ulp; the low 29 bits of the sum are then the
                                    // distance from the tie (mod 2^29): near 0 or 2^29 = not certified
                                    const unsigned long long r =
                                        (unsigned long long)__double_as_longlong(sd) + 0x10000000ull;
                                    accb = min(accb, ((uint32_t)r << 3) + 8u * 32u);
                                    v[k] = __uint_as_float((uint32_t)(r >> 29));
                                    continue;
                                }
                                if (BLEND != DCB_BLEND_LERP32) {
                                    // fp64 blend from the raw float32 box: four conversions per pixel
                                    const double tx = __dsub_rn(__hiloint2double(__double2hiint(ux), (int)fx), Mx);
                                    const double ty = __dsub_rn(__hiloint2double(__double2hiint(uy), (int)fy), My);
                                    const uint32_t qa = base_s + 4u * (yi * bw + xi);
                                    float a, b, c, d;
                                    asm("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(qa));
                                    asm("ld.shared.f32 %0, [%1+4];" : "=f"(b) : "r"(qa));
                                    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(c) : "r"(qa), "n"(4 * kImgBoxW));
                                    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(d) : "r"(qa), "n"(4 * kImgBoxW + 4));
                                    // a set sign bit, Inf or NaN among the taps: not what lerp_fma is
                                    // certified for (largest bit pattern >= 0x7f800000)
                                    if (BLEND == DCB_BLEND_LERP64) {   // three float64 lerps, no certificate needed
                                        const double da = a, db = b, dc = c, dd = d;
                                        const double top = fma(db - da, tx, da);
                                        const double bot = fma(dd - dc, tx, dc);
                                        v[k] = __double2float_rn(fma(bot - top, ty, top));
                                        continue;
                                    }
                                    tapmax = __vimax3_u32(tapmax, __float_as_uint(a), __float_as_uint(b));
                                    tapmax = __vimax3_u32(tapmax, __float_as_uint(c), __float_as_uint(d));
                                    const double sd = lerp_fma((double)a, (double)b, (double)c, (double)d, tx, ty);
                                    accb = min(accb, cert_key(sd, kBlendCertAdd));
                                    v[k] = __double2float_rn(sd);
                                    continue;
                                }
                                {
                                    // fraction as a float: exponent 150 - sh puts its ulp at 2^-sh
                                    const float tx = __uint_as_float(e32x | fx) - __uint_as_float(e32x);
                                    const float ty = __uint_as_float(e32y | fy) - __uint_as_float(e32y);
                                    const uint32_t qa = base_s + 4u * (yi * bw + xi);
                                    float a, b, c, d;
                                    asm("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(qa));
                                    asm("ld.shared.f32 %0, [%1+4];" : "=f"(b) : "r"(qa));
                                    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(c) : "r"(qa), "n"(4 * kImgBoxW));
                                    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(d) : "r"(qa), "n"(4 * kImgBoxW + 4));
                                    const float top = fmaf(b - a, tx, a);
                                    const float bot = fmaf(d - c, tx, c);
                                    v[k] = fmaf(bot - top, ty, top);
                                }
                            }
                        };
                        if ((inf.x & 0xfu) == 0xfu)
                            sample4(std::false_type{});
                        else
                            sample4(std::true_type{});
                        // (a blend within 32 ulp64 of a rounding boundary: the exact row below)
                        done = !__any_sync(0xffffffffu, accb < kBlendCertLim || tapmax >= 0x7f800000u);
                        n_bfail += done ? 0u : 1u;
                    }
                }
                if (!done) {
#if DCB_IMG_LAZYWIN
                // the exact path's window arithmetic, formed here and not in the tile prologue (the
                // empty asm keeps the compiler from hoisting it back: every tile would pay for what
                // 1 % of the rows use)
                int wbx0 = box.bx0, wby0 = box.by0;
                asm volatile("" : "+r"(wbx0), "+r"(wby0));
                const int magic_x = 0x4B000000 + wbx0, magic_y = 0x4B000000 + wby0;
                const int lim_x = box.use ? min(p.bw - 1, wmax - wbx0) : 0;
                const int lim_y = box.use ? min(p.bh - 1, p.ylast - wby0) : 0;
#endif
                float xf[kCols], yf[kCols];
                ev.row(p, yd, xf, yf);
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
                if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
                    for (int k = 0; k < kCols; ++k) {
                        const float tx = xf[k] - (tfx[k] - 8388608.0f);  // exact
                        const float ty = yf[k] - (tfy[k] - 8388608.0f);
                        const int idx = iy[k] * bw + ix[k];
                        if (ORDER == 0) {
                            const int sel = idx + (tx >= 0.5f ? 1 : 0) + (ty >= 0.5f ? bw : 0);
                            v[k] = rawt[sel];
                        } else if (BLEND != DCB_BLEND_LERP32) {
                            const float *q = rawt + idx;
                            const double a = q[0], b = q[1];
                            const double c = q[bw], d = q[bw + 1];
                            if (BLEND == DCB_BLEND_LERP64) {
                                const double top = fma(b - a, (double)tx, a);
                                const double bot = fma(d - c, (double)tx, c);
                                v[k] = finish_f64(fma(bot - top, (double)ty, top), p.rint);
                            } else {
                                v[k] = finish_f64(blend_exact(a, b, c, d, (double)tx, (double)ty), p.rint);
                            }
                        } else {
                            const float *q = rawt + idx;
                            const float a = q[0], b = q[1];
                            const float c = q[bw], d = q[bw + 1];
                            const float top = fmaf(b - a, tx, a);
                            const float bot = fmaf(d - c, tx, c);
                            v[k] = fmaf(bot - top, ty, top);
                        }
                    }
                    if (ORDER == 1 && BLEND == DCB_BLEND_LERP32 && p.rint) {
#pragma unroll
                        for (int k = 0; k < kCols; ++k) v[k] = finish_f32(v[k], 1);
                    }
                } else {
                    const GlobalFetch gfetch{p.src - (long long)p.yorg * p.src_pitch, p.src_pitch};
                    const int wbits = __float_as_int((float)wmax);
                    const int hbits = __float_as_int((float)(p.H - 1));
#pragma unroll
                    for (int k = 0; k < kCols; ++k)
                        v[k] = sample_px<ORDER, BLEND, float>(gfetch, clamp_bits(xf[k], wbits),
                                                              clamp_bits(yf[k], hbits), wmax,
                                                              p.yorg, p.ylast, p.rint);
                }
                }
                if (DCB_ABL & 1) {
#pragma unroll
                    for (int k = 0; k < kCols; ++k)
                        if (__float_as_uint(v[k]) == 0x7fc12345u) DCB_IMG_STORE(orow + 32 * k, v[k]);
                } else if (FULL) {
#pragma unroll
                    for (int k = 0; k < kCols; ++k) DCB_IMG_STORE(orow + 32 * k, v[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < kCols; ++k)
                        if (x_base + 32 * k <= wmax) DCB_IMG_STORE(orow + 32 * k, v[k]);
                }
            }
            };
            if (full_w)
                tile_rows(std::true_type{});
            else
                tile_rows(std::false_type{});
            if (n_bfail != 0 && p.stats != nullptr && lane == 0)
                atomicAdd(p.stats + 3, (unsigned long long)n_bfail);
        }
        // this warp is done with buffer i&1 (and with its plan record)
        __syncwarp();
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + (i & (NBUF - 1))), "n"(kImgThreads) : "memory");
        DCB_LOG(2 * i + 1);
#ifdef DCB_IMG_TIMELINE
        if (p.stats != nullptr && threadIdx.x == 0) dbg_t = global_ns();
#endif
    }
#ifdef DCB_IMG_TIMELINE
    if (p.stats != nullptr && threadIdx.x == 0 && blockIdx.x < kTimelineCtas) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long *q = p.stats + 8 + 8 * blockIdx.x;
        q[2] = global_ns();
        q[3] = (unsigned long long)smid | ((unsigned long long)n << 32);
        q[4] = dbg_wait;
        q[5] = dbg_wmax;
    }
#endif
}

}  // namespace dcb
