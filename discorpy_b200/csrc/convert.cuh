// convert.cuh -- layout / dtype conversion kernels around the Z-stack kernel
// (sm_100a).  Camera frames arrive as interleaved (H, W, C) uint8 / uint16
// arrays (discorpy/util/utility.py:278-342 unwarp_color_image_backward loops
// over mat[:, :, i]); the remap kernels work on float32 planes.  Doing the
// de-interleave + widening on the device keeps the PCIe traffic at the frame's
// own size (1 or 2 bytes per sample) and removes the host-side transposes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcb {

template <class T>
__device__ __forceinline__ T narrow_from_f32(float v);
// the values are already integers in range (DCB_FLAG_ROUND_INT); cvt saturates anyway
template <> __device__ __forceinline__ uint8_t narrow_from_f32<uint8_t>(float v) {
    return (uint8_t)min(max(__float2int_rn(v), 0), 255);
}
template <> __device__ __forceinline__ int8_t narrow_from_f32<int8_t>(float v) {
    return (int8_t)min(max(__float2int_rn(v), -128), 127);
}
template <> __device__ __forceinline__ uint16_t narrow_from_f32<uint16_t>(float v) {
    return (uint16_t)min(max(__float2int_rn(v), 0), 65535);
}
template <> __device__ __forceinline__ int16_t narrow_from_f32<int16_t>(float v) {
    return (int16_t)min(max(__float2int_rn(v), -32768), 32767);
}
template <> __device__ __forceinline__ float narrow_from_f32<float>(float v) { return v; }

// (H, W, C) dense interleaved -> C float32 planes; one thread per pixel
template <class T>
__global__ void __launch_bounds__(256)
    unpack_hwc_kernel(const T *__restrict__ src, float *__restrict__ dst, int H, int W, int C,
                      long long pitch, long long plane) {
    const long long npx = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx;
         i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        const T *s = src + i * C;
        float *d = dst + (long long)y * pitch + x;
        for (int c = 0; c < C; ++c) d[(long long)c * plane] = (float)s[c];
    }
}

// C float32 planes -> (H, W, C) dense interleaved
template <class T>
__global__ void __launch_bounds__(256)
    pack_hwc_kernel(const float *__restrict__ src, T *__restrict__ dst, int H, int W, int C,
                    long long pitch, long long plane) {
    const long long npx = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx;
         i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        const float *s = src + (long long)y * pitch + x;
        T *d = dst + i * C;
        for (int c = 0; c < C; ++c) d[c] = narrow_from_f32<T>(s[(long long)c * plane]);
    }
}

}  // namespace dcb
