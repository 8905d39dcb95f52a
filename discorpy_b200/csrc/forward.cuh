// forward.cuh -- unwarp_image_forward on the device.
//
// Reference: discorpy/post/postprocessing.py:151-185.  Every SOURCE pixel is moved
// to the integer position  round(clip(centre + F(rd) * (p - centre)))  of the output
// (np.round: half to even), pixels nobody lands on stay 0, and where several sources
// land on one output pixel NumPy's fancy assignment `out[yu, xu] = mat` keeps the
// LAST one in C order, i.e. the source with the largest linear index.  A scatter
// with that rule is made deterministic in two passes: (1) every source pixel
// proposes itself with atomicMax(winner[target], linear_index + 1); (2) every output
// pixel copies its winner (or writes 0).  "Only for assessment" in the reference;
// here for device-resident pipelines that want the vacancy pattern without a
// round trip to the host.
#pragma once
#include "remap.cuh"

namespace dcb {

__global__ void __launch_bounds__(256)
    forward_propose_kernel(int H, int W, RadialDev rad, unsigned *__restrict__ winner) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const double xd = (double)x - rad.xc, yd = (double)y - rad.yc;                 // :176-179
    const double rd = __dsqrt_rn(__dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd)));
    // :180-183 in the reference's own operation order -- sum_i a_i * rd**i term by term, then
    // centre + F * d with the product rounded before the sum: on lattice-symmetric inputs
    // (integer centre, short decimal coefficients) the argument of np.round is EXACTLY k + 0.5
    // for some pixels, and only the same roundings break those ties the same way.  (rd**i is
    // formed by repeated multiplication: identical to NumPy for i <= 2, within one ulp above.)
    double f = 0.0, pw = 1.0;
    for (int i = 0; i < rad.n; ++i) {
        f = __dadd_rn(f, __dmul_rn(rad.a[i], pw));
        pw = __dmul_rn(pw, rd);
    }
    const double xu = rint(fmin(fmax(__dadd_rn(rad.xc, __dmul_rn(f, xd)), 0.0), (double)(W - 1)));
    const double yu = rint(fmin(fmax(__dadd_rn(rad.yc, __dmul_rn(f, yd)), 0.0), (double)(H - 1)));
    const unsigned target = (unsigned)((long long)yu * W + (long long)xu);
    atomicMax(&winner[target], (unsigned)((long long)y * W + x) + 1u);
}

__global__ void __launch_bounds__(256)
    forward_gather_kernel(const float *__restrict__ src, long long spitch, float *__restrict__ dst,
                          long long dpitch, int H, int W, const unsigned *__restrict__ winner) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const unsigned w = winner[(long long)y * W + x];
    float v = 0.0f;                                                               // :184 zeros_like
    if (w != 0u) {
        const unsigned s = w - 1u;
        v = src[(long long)(s / (unsigned)W) * spitch + (s % (unsigned)W)];
    }
    dst[(long long)y * dpitch + x] = v;
}

}  // namespace dcb
