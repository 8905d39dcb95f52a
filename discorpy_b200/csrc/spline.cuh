// spline.cuh -- spline orders 2..5 (and float64 images at any order) of the
// backward-remap path: the float64 B-spline prefilter and the (order+1)^2-tap
// sampler of scipy.ndimage.map_coordinates, for coordinates that lie inside the
// image.
//
// The reference reaches this arithmetic through `order=` / `mode=` of
// discorpy/post/postprocessing.py:147 and :491 and discorpy/util/utility.py:333,
// :338 (examples/readthedocs_demo/demo_07.py:60 uses order 3); the arithmetic
// itself is SciPy's (scipy/ndimage/_interpolation.py:375-476 is the readable
// wrapper; pre-padding :212-227, prefilter :467-469).  oracle/oracle_spline.py
// restates it and is bit-identical to the installed SciPy for every order, mode
// and dtype; the kernels below follow the oracle operation by operation --
// separate multiplications and additions (no FMA contraction: SciPy's binary
// has none), true IEEE divisions, the same summation orders -- so prefiltered
// coefficients and samples are bit-identical too, given the same coordinates.
//
// Not tuned like the order-0/1 kernels: one thread per image line in the
// recursive filter (the recursion is sequential by nature; lines are
// independent), one thread per output pixel in the sampler, taps gathered
// through L1/L2.  DESIGN.md section 5.6 has the measured times.
#pragma once
#include "remap.cuh"
#include "remap_image.cuh"

namespace dcb {

enum { SPL_MIRROR = 0, SPL_REFLECT = 1, SPL_WRAP = 2 };   // prefilter boundary / tap folding
enum { SPL_MAP_RADIAL = 0, SPL_MAP_PERSP = 1, SPL_MAP_COORDS = 2 };

// ---------------------------------------------------------------------------
// prefilter
// ---------------------------------------------------------------------------

// dst (Hc x Wc, Hc = H + 2 npad) = gain * pad(src): 'edge' replication
// (mode nearest) or zeros (grid-constant); npad == 0 is a plain widening copy.
template <class T>
__global__ void __launch_bounds__(256)
    spline_pad_kernel(const T *__restrict__ src, long long spitch, int H, int W,
                      double *__restrict__ dst, long long dpitch, int npad, int pad_const,
                      double gain) {
    const int Wc = W + 2 * npad, Hc = H + 2 * npad;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y0 = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= Wc) return;
    for (int y = y0; y < Hc; y += gridDim.y * 8) {
        int sx = x - npad, sy = y - npad;
        double v;
        if (pad_const && (sx < 0 || sx >= W || sy < 0 || sy >= H)) {
            v = 0.0;
        } else {
            sx = min(max(sx, 0), W - 1);
            sy = min(max(sy, 0), H - 1);
            v = (double)src[(long long)sy * spitch + sx];
        }
        dst[(long long)y * dpitch + x] = __dmul_rn(v, gain);
    }
}

// out[x][y] = gain * in[y][x]   (rows x cols -> cols x rows), 32 x 32 tiles
__global__ void __launch_bounds__(256)
    spline_transpose_kernel(const double *__restrict__ in, long long ipitch, int rows, int cols,
                            double *__restrict__ out, long long opitch, double gain) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k, c = c0 + tx;
        if (r < rows && c < cols) tile[k][tx] = in[(long long)r * ipitch + c];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, r = r0 + tx;
        if (r < rows && c < cols) out[(long long)c * opitch + r] = __dmul_rn(tile[tx][k], gain);
    }
}

constexpr int kSplU = 16;   // rows loaded ahead of the dependent chain

struct SplinePoles {
    int npoles;
    double z[2];
    double zp[2];   // pow(z, n) (reflect), pow(z, n - 1) (mirror); unused for wrap
};

// In-place recursive filter of every column of c (n rows): per pole the causal
// initialisation, the forward recursion, the anticausal initialisation and the
// backward recursion, in SciPy's operation order (oracle_spline.filter_lines).
// One thread per column; neighbouring threads touch neighbouring addresses.
__global__ void __launch_bounds__(32)
    spline_filter_cols_kernel(double *__restrict__ c, long long pitch, int n, int ncols,
                              SplinePoles pl, int kind) {
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (j >= ncols || n < 2) return;
    double *col = c + j;
#define C_(i) col[(long long)(i) * pitch]
    for (int p = 0; p < pl.npoles; ++p) {
        const double z = pl.z[p], zp = pl.zp[p];
        // ---- causal initialisation -------------------------------------------------
        if (kind == SPL_MIRROR) {
            double acc = __dadd_rn(__dmul_rn(zp, C_(n - 1)), C_(0));
            double z_i = z;
            int i = 1;
            for (; i + (kSplU - 1) < n - 1; i += kSplU) {   // loads first: they do not depend on the chain
                double a[kSplU], b[kSplU];
#pragma unroll
                for (int k = 0; k < kSplU; ++k) a[k] = C_(n - 1 - i - k), b[k] = C_(i + k);
#pragma unroll
                for (int k = 0; k < kSplU; ++k) {
                    const double t = __dadd_rn(__dmul_rn(a[k], zp), b[k]);
                    acc = __dadd_rn(acc, __dmul_rn(t, z_i));
                    z_i = __dmul_rn(z_i, z);
                }
            }
            for (; i < n - 1; ++i) {
                const double t = __dadd_rn(__dmul_rn(C_(n - 1 - i), zp), C_(i));
                acc = __dadd_rn(acc, __dmul_rn(t, z_i));
                z_i = __dmul_rn(z_i, z);
            }
            C_(0) = __ddiv_rn(acc, __dsub_rn(1.0, __dmul_rn(zp, zp)));
        } else if (kind == SPL_REFLECT) {
            const double c0 = C_(0);
            double acc = __dadd_rn(__dmul_rn(C_(n - 1), zp), c0);
            double z_i = z;
            int i = 1;
            for (; i + (kSplU - 1) < n; i += kSplU) {   // C_(0) is still the original value throughout
                double a[kSplU], b[kSplU];
#pragma unroll
                for (int k = 0; k < kSplU; ++k) a[k] = C_(n - 1 - i - k), b[k] = C_(i + k);
#pragma unroll
                for (int k = 0; k < kSplU; ++k) {
                    const double t = __dadd_rn(__dmul_rn(a[k], zp), b[k]);
                    acc = __dadd_rn(acc, __dmul_rn(t, z_i));
                    z_i = __dmul_rn(z_i, z);
                }
            }
            for (; i < n; ++i) {
                const double t = __dadd_rn(__dmul_rn(C_(n - 1 - i), zp), C_(i));
                acc = __dadd_rn(acc, __dmul_rn(t, z_i));
                z_i = __dmul_rn(z_i, z);
            }
            const double q = __ddiv_rn(__dmul_rn(z, acc), __dsub_rn(1.0, __dmul_rn(zp, zp)));
            C_(0) = __dadd_rn(q, c0);
        } else {
            double acc = C_(0);
            double z_i = z;
            for (int i = 1; i < n; ++i) {
                acc = __dadd_rn(acc, __dmul_rn(C_(n - i), z_i));
                z_i = __dmul_rn(z_i, z);
            }
            C_(0) = __ddiv_rn(acc, __dsub_rn(1.0, z_i));
        }
        // ---- forward recursion: c[i] = z c[i-1] + c[i] (loads run ahead of the chain) --
        {
            double prev = C_(0);
            int i = 1;
            for (; i + (kSplU - 1) < n; i += kSplU) {
                double v[kSplU];
#pragma unroll
                for (int k = 0; k < kSplU; ++k) v[k] = C_(i + k);
#pragma unroll
                for (int k = 0; k < kSplU; ++k) {
                    prev = __dadd_rn(__dmul_rn(prev, z), v[k]);
                    C_(i + k) = prev;
                }
            }
            for (; i < n; ++i) {
                prev = __dadd_rn(__dmul_rn(prev, z), C_(i));
                C_(i) = prev;
            }
        }
        // ---- anticausal initialisation -------------------------------------------------
        if (kind == SPL_MIRROR) {
            const double t = __dadd_rn(__dmul_rn(C_(n - 2), z), C_(n - 1));
            C_(n - 1) = __ddiv_rn(__dmul_rn(t, z), __dsub_rn(__dmul_rn(z, z), 1.0));
        } else if (kind == SPL_REFLECT) {
            C_(n - 1) = __dmul_rn(__ddiv_rn(z, __dsub_rn(z, 1.0)), C_(n - 1));
        } else {
            double acc = C_(n - 1);
            double z_i = z;
            for (int i = 0; i < n - 1; ++i) {
                acc = __dadd_rn(acc, __dmul_rn(C_(i), z_i));
                z_i = __dmul_rn(z_i, z);
            }
            C_(n - 1) = __dmul_rn(__ddiv_rn(z, __dsub_rn(z_i, 1.0)), acc);
        }
        // ---- backward recursion: c[i] = z (c[i+1] - c[i]) ----------------------------------
        {
            double next = C_(n - 1);
            int i = n - 2;
            for (; i - (kSplU - 1) >= 0; i -= kSplU) {
                double v[kSplU];
#pragma unroll
                for (int k = 0; k < kSplU; ++k) v[k] = C_(i - k);
#pragma unroll
                for (int k = 0; k < kSplU; ++k) {
                    next = __dmul_rn(__dsub_rn(next, v[k]), z);
                    C_(i - k) = next;
                }
            }
            for (; i >= 0; --i) {
                next = __dmul_rn(__dsub_rn(next, C_(i)), z);
                C_(i) = next;
            }
        }
    }
#undef C_
}

// The same filter with the rows staged through shared memory: the one-thread-per-line kernel
// above keeps only 16 rows per line in flight (128 warps on the whole GPU for a 4096-wide
// image, ~0.5 TB/s); here a CTA owns 32 columns, seven warps move windows of kSplR rows
// between global and shared memory (double-buffered) while warp 0 runs the sequential
// recursions on the window that has landed.  Same operations in the same order per line.
constexpr int kSplR = 64;   // rows per window; shared memory = 2 stages x 2 streams x kSplR x 32 doubles
constexpr size_t kSplSmemBytes = (size_t)2 * 2 * kSplR * 32 * sizeof(double);

// One sweep over `count` steps: step j reads row firstA + dirA * j (stream A) and, if TWO, row
// firstB + dirB * j (stream B).  body(a, b, m) is run by warp 0 on every window (a[j * 32],
// b[j * 32] for j < m, this lane's column); if WRITE the window of stream A is written back.
template <bool TWO, bool WRITE, class Body>
__device__ __forceinline__ void spline_sweep(double *__restrict__ gcol, bool colok, long long pitch,
                                             double *sbuf, int firstA, int dirA, int firstB, int dirB,
                                             int count, Body &&body) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (count + kSplR - 1) / kSplR;
    auto win = [&](int st, int strm) { return sbuf + (size_t)((st * 2 + strm) * kSplR) * 32 + lane; };
    auto move = [&](int w, bool to_smem) {   // loader warps 1..7: window w <-> global
        const int j0 = w * kSplR, m = min(kSplR, count - j0);
        double *a = win(w & 1, 0), *b = win(w & 1, 1);
        if (!colok) return;
        // (issuing all loads of a window before the first shared-memory store was measured:
        //  no faster, 254 registers)
        for (int r = warp - 1; r < m; r += 7) {
            const long long ra = (long long)(firstA + dirA * (j0 + r)) * pitch;
            if (to_smem) {
                a[r * 32] = gcol[ra];
                if (TWO) b[r * 32] = gcol[(long long)(firstB + dirB * (j0 + r)) * pitch];
            } else {
                gcol[ra] = a[r * 32];
            }
        }
    };
    if (nw <= 0) return;
    if (warp != 0) move(0, true);
    __syncthreads();
    for (int w = 0; w < nw; ++w) {
        if (warp == 0) {
            body(win(w & 1, 0), win(w & 1, 1), min(kSplR, count - w * kSplR));
        } else {
            if (WRITE && w >= 1) move(w - 1, false);   // results of the previous window
            if (w + 1 < nw) move(w + 1, true);         // then refill that buffer
        }
        __syncthreads();
    }
    if (WRITE) {
        if (warp != 0) move(nw - 1, false);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
    spline_filter_cols_staged_kernel(double *__restrict__ c, long long pitch, int n, int ncols,
                                     SplinePoles pl, int kind) {
    extern __shared__ __align__(16) double spl_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const bool colok = col < ncols;
    double *gcol = c + (colok ? col : 0);
    if (n < 2) return;
#define C_(i) gcol[(long long)(i) * pitch]
    for (int p = 0; p < pl.npoles; ++p) {
        const double z = pl.z[p], zp = pl.zp[p];
        double acc = 0.0, z_i = z, prev = 0.0, prev2 = 0.0;   // live in warp 0 only
        // ---- causal initialisation -------------------------------------------------
        if (kind == SPL_MIRROR) {
            if (warp == 0) acc = __dadd_rn(__dmul_rn(zp, C_(n - 1)), C_(0));
            spline_sweep<true, false>(gcol, colok, pitch, spl_smem, 1, 1, n - 2, -1, n - 2,
                                      [&](double *a, double *b, int m) {
                                          for (int j = 0; j < m; ++j) {
                                              const double t = __dadd_rn(__dmul_rn(b[j * 32], zp), a[j * 32]);
                                              acc = __dadd_rn(acc, __dmul_rn(t, z_i));
                                              z_i = __dmul_rn(z_i, z);
                                          }
                                      });
            prev = __ddiv_rn(acc, __dsub_rn(1.0, __dmul_rn(zp, zp)));
        } else if (kind == SPL_REFLECT) {
            double c0 = 0.0;
            if (warp == 0) {
                c0 = C_(0);
                acc = __dadd_rn(__dmul_rn(C_(n - 1), zp), c0);
            }
            spline_sweep<true, false>(gcol, colok, pitch, spl_smem, 1, 1, n - 2, -1, n - 1,
                                      [&](double *a, double *b, int m) {
                                          for (int j = 0; j < m; ++j) {
                                              const double t = __dadd_rn(__dmul_rn(b[j * 32], zp), a[j * 32]);
                                              acc = __dadd_rn(acc, __dmul_rn(t, z_i));
                                              z_i = __dmul_rn(z_i, z);
                                          }
                                      });
            prev = __dadd_rn(__ddiv_rn(__dmul_rn(z, acc), __dsub_rn(1.0, __dmul_rn(zp, zp))), c0);
        } else {
            if (warp == 0) acc = C_(0);
            spline_sweep<false, false>(gcol, colok, pitch, spl_smem, n - 1, -1, 0, 0, n - 1,
                                       [&](double *a, double *, int m) {
                                           for (int j = 0; j < m; ++j) {
                                               acc = __dadd_rn(acc, __dmul_rn(a[j * 32], z_i));
                                               z_i = __dmul_rn(z_i, z);
                                           }
                                       });
            prev = __ddiv_rn(acc, __dsub_rn(1.0, z_i));
        }
        if (warp == 0 && colok) C_(0) = prev;
        __syncthreads();
        // ---- forward recursion over rows 1 .. n-1 -----------------------------------------
        prev2 = prev;
        spline_sweep<false, true>(gcol, colok, pitch, spl_smem, 1, 1, 0, 0, n - 1,
                                  [&](double *a, double *, int m) {
                                      int j = 0;
                                      for (; j + 7 < m; j += 8) {   // loads first: the stores alias them
                                          double v[8];
#pragma unroll
                                          for (int k = 0; k < 8; ++k) v[k] = a[(j + k) * 32];
#pragma unroll
                                          for (int k = 0; k < 8; ++k) {
                                              prev2 = prev;
                                              prev = __dadd_rn(__dmul_rn(prev, z), v[k]);
                                              a[(j + k) * 32] = prev;
                                          }
                                      }
                                      for (; j < m; ++j) {
                                          prev2 = prev;
                                          prev = __dadd_rn(__dmul_rn(prev, z), a[j * 32]);
                                          a[j * 32] = prev;
                                      }
                                  });
        // ---- anticausal initialisation (prev = c[n-1], prev2 = c[n-2]) ---------------------
        double next;
        if (kind == SPL_MIRROR) {
            const double t = __dadd_rn(__dmul_rn(prev2, z), prev);
            next = __ddiv_rn(__dmul_rn(t, z), __dsub_rn(__dmul_rn(z, z), 1.0));
        } else if (kind == SPL_REFLECT) {
            next = __dmul_rn(__ddiv_rn(z, __dsub_rn(z, 1.0)), prev);
        } else {
            acc = prev;
            z_i = z;
            spline_sweep<false, false>(gcol, colok, pitch, spl_smem, 0, 1, 0, 0, n - 1,
                                       [&](double *a, double *, int m) {
                                           for (int j = 0; j < m; ++j) {
                                               acc = __dadd_rn(acc, __dmul_rn(a[j * 32], z_i));
                                               z_i = __dmul_rn(z_i, z);
                                           }
                                       });
            next = __dmul_rn(__ddiv_rn(z, __dsub_rn(z_i, 1.0)), acc);
        }
        if (warp == 0 && colok) C_(n - 1) = next;
        __syncthreads();
        // ---- backward recursion over rows n-2 .. 0 -----------------------------------------
        spline_sweep<false, true>(gcol, colok, pitch, spl_smem, n - 2, -1, 0, 0, n - 1,
                                  [&](double *a, double *, int m) {
                                      int j = 0;
                                      for (; j + 7 < m; j += 8) {
                                          double v[8];
#pragma unroll
                                          for (int k = 0; k < 8; ++k) v[k] = a[(j + k) * 32];
#pragma unroll
                                          for (int k = 0; k < 8; ++k) {
                                              next = __dmul_rn(__dsub_rn(next, v[k]), z);
                                              a[(j + k) * 32] = next;
                                          }
                                      }
                                      for (; j < m; ++j) {
                                          next = __dmul_rn(__dsub_rn(next, a[j * 32]), z);
                                          a[j * 32] = next;
                                      }
                                  });
    }
#undef C_
}

// ---------------------------------------------------------------------------
// sampler
// ---------------------------------------------------------------------------
struct SplineParams {
    const double *coef;       // (Hc x Wc) prefiltered (or widened, order <= 1) image
    long long cpitch;         // elements
    int Hc, Wc, npad;
    void *dst;                // float or double, (H x W) or n points
    long long dst_pitch;      // elements
    int H, W;                 // image the coordinates refer to (clipping range)
    int tap_kind;             // SPL_MIRROR / SPL_REFLECT / SPL_WRAP
    int rint;                 // integer image: round half away from zero, saturate to [lo, hi]
    double lo, hi;
    int map;                  // SPL_MAP_*
    int coord_f64;            // SPL_MAP_COORDS: coordinate arrays are double
    const void *yd, *xd;      // SPL_MAP_COORDS
    unsigned long long n;     // SPL_MAP_COORDS: number of points
    unsigned *oob_count;      // SPL_MAP_COORDS: coordinates outside the image (clamped)
    RadialDev rad;
    PerspDev per;
};

// SciPy's B-spline basis at offset x from the middle knot (ni_splines.c), every
// operation rounded separately; the last weight is 1 minus the others in order.
template <int ORDER>
__device__ __forceinline__ void spline_weights(double x, double (&w)[ORDER + 1]) {
    const double y = x, z = __dsub_rn(1.0, x);
    if (ORDER == 1) {
        w[0] = z;
    } else if (ORDER == 2) {
        w[1] = __dsub_rn(0.75, __dmul_rn(x, x));
        const double u = __dsub_rn(0.5, x);
        w[0] = __dmul_rn(__dmul_rn(0.5, u), u);
    } else if (ORDER == 3) {
        w[1] = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(y, y), __dsub_rn(y, 2.0)), 3.0), 4.0), 6.0);
        w[2] = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(z, z), __dsub_rn(z, 2.0)), 3.0), 4.0), 6.0);
        w[0] = __ddiv_rn(__dmul_rn(__dmul_rn(z, z), z), 6.0);
    } else if (ORDER == 4) {
        double t = __dmul_rn(x, x);
        w[2] = __dadd_rn(__dmul_rn(t, __dsub_rn(__dmul_rn(t, 0.25), 0.625)), 115.0 / 192.0);
        double u = __dadd_rn(1.0, x);
        w[1] = __dadd_rn(
            __dmul_rn(u, __dadd_rn(__dmul_rn(u, __dsub_rn(__ddiv_rn(__dmul_rn(u, __dsub_rn(5.0, u)), 6.0), 1.25)),
                                   5.0 / 24.0)),
            55.0 / 96.0);
        w[3] = __dadd_rn(
            __dmul_rn(z, __dadd_rn(__dmul_rn(z, __dsub_rn(__ddiv_rn(__dmul_rn(z, __dsub_rn(5.0, z)), 6.0), 1.25)),
                                   5.0 / 24.0)),
            55.0 / 96.0);
        u = __dsub_rn(0.5, x);
        t = __dmul_rn(u, u);
        w[0] = __ddiv_rn(__dmul_rn(t, t), 24.0);
    } else if (ORDER == 5) {
        double t = __dmul_rn(y, y);
        w[2] = __dadd_rn(__dmul_rn(t, __dsub_rn(__dmul_rn(t, __dsub_rn(0.25, __ddiv_rn(y, 12.0))), 0.5)), 0.55);
        t = __dmul_rn(z, z);
        w[3] = __dadd_rn(__dmul_rn(t, __dsub_rn(__dmul_rn(t, __dsub_rn(0.25, __ddiv_rn(z, 12.0))), 0.5)), 0.55);
        const double y1 = __dadd_rn(y, 1.0);
        w[1] = __dadd_rn(
            __dmul_rn(y1, __dadd_rn(__dmul_rn(y1, __dsub_rn(__dmul_rn(y1, __dadd_rn(__dmul_rn(y1, __dsub_rn(__ddiv_rn(y1, 24.0), 0.375)), 1.25)), 1.75)), 0.625)),
            0.425);
        const double z1 = __dadd_rn(z, 1.0);
        w[4] = __dadd_rn(
            __dmul_rn(z1, __dadd_rn(__dmul_rn(z1, __dsub_rn(__dmul_rn(z1, __dadd_rn(__dmul_rn(z1, __dsub_rn(__ddiv_rn(z1, 24.0), 0.375)), 1.25)), 1.75)), 0.625)),
            0.425);
        const double z0 = __dsub_rn(z1, 1.0);
        t = __dmul_rn(z0, z0);
        w[0] = __ddiv_rn(__dmul_rn(__dmul_rn(z0, t), t), 120.0);
    }
    double last = 1.0;
#pragma unroll
    for (int i = 0; i < ORDER; ++i) last = __dsub_rn(last, w[i]);
    w[ORDER] = last;
}

// tap index -> array index for a tap that left [0, n) (the coordinate is in range)
__device__ __forceinline__ int spline_fold(int idx, int n, int kind) {
    if ((unsigned)idx < (unsigned)n) return idx;
    if (n <= 1) return 0;
    if (kind == SPL_MIRROR) {
        const int s2 = 2 * n - 2;
        int m = idx % s2;
        if (m < 0) m += s2;
        return m >= n ? s2 - m : m;
    }
    if (kind == SPL_REFLECT) {
        const int s2 = 2 * n;
        int m = idx % s2;
        if (m < 0) m += s2;
        return m >= n ? s2 - 1 - m : m;
    }
    int m = idx % n;
    return m < 0 ? m + n : m;
}

template <int ORDER>
__device__ __forceinline__ double spline_sample(const SplineParams &p, double cy, double cx) {
    // cy, cx: clipped coordinates (already fp32-rounded where the reference rounds)
    const double y = __dadd_rn(cy, (double)p.npad), x = __dadd_rn(cx, (double)p.npad);
    if (ORDER == 0) {
        const int yi = min((int)floor(__dadd_rn(y, 0.5)), p.Hc - 1);
        const int xi = min((int)floor(__dadd_rn(x, 0.5)), p.Wc - 1);
        return p.coef[(long long)yi * p.cpitch + xi];
    }
    const double fy = floor((ORDER & 1) ? y : __dadd_rn(y, 0.5));
    const double fx = floor((ORDER & 1) ? x : __dadd_rn(x, 0.5));
    double wy[ORDER + 1], wx[ORDER + 1];
    spline_weights<ORDER>(__dsub_rn(y, fy), wy);
    spline_weights<ORDER>(__dsub_rn(x, fx), wx);
    const int sy = (int)fy - ORDER / 2, sx = (int)fx - ORDER / 2;
    int xs[ORDER + 1];
#pragma unroll
    for (int j = 0; j <= ORDER; ++j) xs[j] = spline_fold(sx + j, p.Wc, p.tap_kind);
    double t = 0.0;
#pragma unroll
    for (int i = 0; i <= ORDER; ++i) {
        const double *row = p.coef + (long long)spline_fold(sy + i, p.Hc, p.tap_kind) * p.cpitch;
#pragma unroll
        for (int j = 0; j <= ORDER; ++j)
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(__ldg(row + xs[j]), wy[i]), wx[j]));
    }
    return t;
}

template <class OUT>
__device__ __forceinline__ void spline_store(const SplineParams &p, long long idx, double t) {
    if (p.rint) {  // SciPy: +-0.5, truncate, saturate at the integer type's range
        t = round_half_away(t);
        t = fmin(fmax(t, p.lo), p.hi);
    }
    if (std::is_same<OUT, double>::value)
        reinterpret_cast<double *>(p.dst)[idx] = t;
    else
        reinterpret_cast<float *>(p.dst)[idx] = __double2float_rn(t);
}

// one thread per output pixel of a 32 x 8 tile (images) or per point (coordinates)
template <int ORDER, class OUT>
__global__ void __launch_bounds__(256) spline_remap_kernel(const __grid_constant__ SplineParams p) {
    if (p.map == SPL_MAP_COORDS) {
        unsigned oob = 0;
        const double xmax = (double)(p.W - 1), ymax = (double)(p.H - 1);
        for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < p.n;
             i += (unsigned long long)gridDim.x * 256) {
            double x, y;
            if (p.coord_f64) {
                x = reinterpret_cast<const double *>(p.xd)[i];
                y = reinterpret_cast<const double *>(p.yd)[i];
            } else {
                x = (double)reinterpret_cast<const float *>(p.xd)[i];
                y = (double)reinterpret_cast<const float *>(p.yd)[i];
            }
            oob += !(x >= 0.0 && x <= xmax && y >= 0.0 && y <= ymax);
            x = x > 0.0 ? x : 0.0;  // NaN -> 0
            y = y > 0.0 ? y : 0.0;
            x = x < xmax ? x : xmax;
            y = y < ymax ? y : ymax;
            spline_store<OUT>(p, (long long)i, spline_sample<ORDER>(p, y, x));
        }
        if (p.oob_count != nullptr && oob != 0) atomicAdd(p.oob_count, oob);
        return;
    }
    const int px = blockIdx.x * 32 + (threadIdx.x & 31);
    const int py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px >= p.W || py >= p.H) return;
    float xf, yf;
    if (p.map == SPL_MAP_RADIAL) {
        // postprocessing.py:138-145: float64 map, clip, one rounding to float32
        const double xu = (double)px - p.rad.xc, yu = (double)py - p.rad.yc;
        const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(xu, xu), __dmul_rn(yu, yu)));
        double f = 0.0;
        for (int i = p.rad.n - 1; i >= 0; --i) f = fma(f, r, p.rad.a[i]);
        xf = clamp_coord<float>(fma(f, xu, p.rad.xc), p.W - 1);
        yf = clamp_coord<float>(fma(f, yu, p.rad.yc), p.H - 1);
    } else {
        // postprocessing.py:448-457
        const double xd = (double)px, yd = (double)py;
        const double den = __dadd_rn(__dadd_rn(__dmul_rn(p.per.c[6], xd), __dmul_rn(p.per.c[7], yd)), 1.0);
        const double nx = __dadd_rn(__dadd_rn(__dmul_rn(p.per.c[0], xd), __dmul_rn(p.per.c[1], yd)), p.per.c[2]);
        const double ny = __dadd_rn(__dadd_rn(__dmul_rn(p.per.c[3], xd), __dmul_rn(p.per.c[4], yd)), p.per.c[5]);
        xf = clamp_coord<float>(__ddiv_rn(nx, den), p.W - 1);
        yf = clamp_coord<float>(__ddiv_rn(ny, den), p.H - 1);
    }
    spline_store<OUT>(p, (long long)py * p.dst_pitch + px, spline_sample<ORDER>(p, (double)yf, (double)xf));
}

}  // namespace dcb
