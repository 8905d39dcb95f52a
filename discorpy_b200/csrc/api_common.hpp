// api_common.hpp -- error plumbing shared by the translation units of libdiscorpy_b200.so
#pragma once
#include <cuda_runtime.h>
#include "../../include/discorpy_b200.h"

namespace dcb {
// records a printf-style message for dcb_last_error() on this thread and returns `code`
int fail(int code, const char *fmt, ...);
// adds n to the launch counter behind dcb_launch_count()
void count_launches(int n);
// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda), or NULL
void *tma_encode_fn();
// output rows [row0, row0 + nrows) of the projective remap, dst pointing at row row0 (api.cu; the
// band launches of the host-buffer pipeline in hostpipe.cu)
int persp_rows_f32(const float *src, float *dst, int H, int W, size_t src_pitch, size_t dst_pitch,
                   int row0, int nrows, const dcb_persp *model, const dcb_options *opt, void *stream);
}  // namespace dcb

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return dcb::fail(DCB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                           \
    } while (0)

#define REQUIRE(cond, ...)                                   \
    do {                                                     \
        if (!(cond)) return dcb::fail(DCB_ERR_ARG, __VA_ARGS__); \
    } while (0)
