// diag.cu -- diagnostics of the C ABI that are not part of the unwarp path: self-tests of the
// custom fp64 square roots and of the TMA addressing convention, and the pipe-rate / latency
// micro-benchmarks behind the numbers quoted in DESIGN.md (profiles/r1/microbench_*.txt).
#include <cstring>
#include <cuda.h>
#include "api_common.hpp"
#include "common.cuh"
#include "remap_image.cuh"
#include "microbench.cuh"

using namespace dcb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encoder() { return reinterpret_cast<EncodeTiledFn>(dcb::tma_encode_fn()); }

extern "C" {

int dcb_selftest_sqrt(size_t n, uint64_t seed, uint64_t *mismatch) {
    REQUIRE(mismatch != nullptr, "mismatch is NULL");
    unsigned long long *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(*d)));
    CUDA_TRY(cudaMemset(d, 0, sizeof(*d)));
    selftest_sqrt_kernel<<<148 * 8, 256>>>(n, seed, d);
    dcb::count_launches(1);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(DCB_ERR_CUDA, "selftest: %s", cudaGetErrorString(e));
    *mismatch = h;
    return DCB_OK;
}

int dcb_selftest_sqrt_fast(size_t n, uint64_t seed, uint64_t *differ, uint64_t *beyond_one_ulp) {
    REQUIRE(differ != nullptr && beyond_one_ulp != nullptr, "null pointer");
    unsigned long long *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 2 * sizeof(*d)));
    CUDA_TRY(cudaMemset(d, 0, 2 * sizeof(*d)));
    selftest_sqrt_fast_kernel<<<148 * 8, 256>>>(n, seed, d);
    dcb::count_launches(1);
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(DCB_ERR_CUDA, "selftest: %s", cudaGetErrorString(e));
    *differ = h[0];
    *beyond_one_ulp = h[1];
    return DCB_OK;
}

int dcb_selftest_tma(const float *src, int D, int H, int W, size_t pitch, size_t slice_stride,
                     int box_w, int box_h, int x0, int y0, int z0, float *out, int *status) {
    REQUIRE(src && out && status, "null pointer");
    REQUIRE(encoder() != nullptr, "driver does not export cuTensorMapEncodeTiled");
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
    const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)slice_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult cr = encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)src, gdim, gstr,
                                box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS)
        return fail(DCB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    int *dstatus = nullptr;
    CUDA_TRY(cudaMalloc(&dstatus, sizeof(int)));
    CUDA_TRY(cudaMemset(dstatus, 0, sizeof(int)));
    const size_t smem = (size_t)box_w * box_h * 4;
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute((const void *)selftest_tma_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_tma_kernel<<<1, 256, smem>>>(tmap, box_w, box_h, x0, y0, z0, out, dstatus);
    dcb::count_launches(1);
    cudaError_t e = cudaMemcpy(status, dstatus, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(dstatus);
    if (e != cudaSuccess) return fail(DCB_ERR_CUDA, "selftest_tma: %s", cudaGetErrorString(e));
    return DCB_OK;
}

int dcb_microbench(int which, double *gops) {
    REQUIRE(gops != nullptr, "gops is NULL");
    REQUIRE(which >= 0 && which <= 36, "which must be 0..36");
    double *sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 2 * sizeof(double)));
    if (which >= 6 && which <= 9) {  // latency probes: cycles per dependent operation
        switch (which) {
            case 6: microbench_latency_kernel<6><<<1, 32>>>(sink); break;
            case 7: microbench_latency_kernel<7><<<1, 32>>>(sink); break;
            case 8: microbench_latency_kernel<8><<<1, 32>>>(sink); break;
            default: microbench_latency_kernel<9><<<1, 32>>>(sink); break;
        }
        dcb::count_launches(1);
        cudaError_t e = cudaMemcpy(gops, sink, sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(sink);
        if (e != cudaSuccess) return fail(DCB_ERR_CUDA, "microbench: %s", cudaGetErrorString(e));
        return DCB_OK;
    }
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int grid = 148 * 8, block = 256;
    double ops_per_thread = (double)kMbIters * kMbChains;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        switch (which) {
            case 0: microbench_kernel<0><<<grid, block>>>(sink, 1.0); break;
            case 1: microbench_kernel<1><<<grid, block>>>(sink, 1.0); break;
            case 2: microbench_kernel<2><<<grid, block>>>(sink, 1.0); break;
            case 3: microbench_coords_kernel<<<grid, block>>>(sink, 2050.37, 2040.81); break;
            case 10: microbench_kernel<10><<<grid, block>>>(sink, 1.0); break;
            case 20: microbench_kernel<20><<<grid, block>>>(sink, 1.0); break;
            case 21: microbench_kernel<21><<<grid, block>>>(sink, 1.0); break;
            case 22: microbench_kernel<22><<<grid, block>>>(sink, 1.0); break;
            case 23: microbench_kernel<23><<<grid, block>>>(sink, 1.0); break;
            case 24: microbench_mix_kernel<24><<<grid, block>>>(sink, 1.0); break;
            case 25: microbench_mix_kernel<25><<<grid, block>>>(sink, 1.0); break;
            case 26: microbench_mix_kernel<26><<<grid, block>>>(sink, 1.0); break;
            case 27: microbench_mix_kernel<27><<<grid, block>>>(sink, 1.0); break;
            case 28: microbench_mix_kernel<28><<<grid, block>>>(sink, 1.0); break;
            case 29: microbench_mix_kernel<29><<<grid, block>>>(sink, 1.0); break;
            case 11: case 12: case 13: case 14: case 15: case 16: case 17: case 18: {
                // the coordinate evaluation at 1..8 resident CTAs (8..64 warps) per SM
                const int ctas = which - 10;
                const size_t sm = (size_t)(220 * 1024) / ctas - 2048;
                CUDA_TRY(cudaFuncSetAttribute((const void *)microbench_coords_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                microbench_coords_kernel<<<grid, block, sm>>>(sink, 2050.37, 2040.81);
                break;
            }
            case 30: microbench_kernel<30><<<grid, block>>>(sink, 1.0); break;
            case 31: microbench_kernel<31><<<grid, block>>>(sink, 1.0); break;
            case 32: microbench_kernel<32><<<grid, block>>>(sink, 1.0); break;
            case 33: microbench_kernel<33><<<grid, block>>>(sink, 1.0); break;
            case 34: microbench_kernel<34><<<grid, block>>>(sink, 1.0); break;
            case 35: microbench_kernel<35><<<grid, block>>>(sink, 1.0); break;
            case 36: microbench_kernel<36><<<grid, block>>>(sink, 1.0); break;
            case 4: microbench_kernel<4><<<grid, block>>>(sink, 1.0); break;
            case 5: microbench_kernel<5><<<grid, block>>>(sink, 1.0); break;
        }
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
        dcb::count_launches(1);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (which == 1) ops_per_thread *= 2;       // two conversions per step
    if (which == 3 || (which >= 11 && which <= 18)) ops_per_thread = (double)kMbIters * 4;  // pixels
    if (which == 5) ops_per_thread *= 3;       // DFMA + two conversions
    if (which == 30 || which == 31) ops_per_thread *= 2;   // conversions counted (2 per step)
    if (which == 32) ops_per_thread *= 1;                  // conversions counted (1 per step)
    if (which == 33) ops_per_thread *= 1;
    if (which >= 24 && which <= 29) {  // mixes: cycles per warp-level group on one SM sub-partition
        const double groups = (double)kMbIters * grid * block / 32.0;          // warp-groups issued
        const double smsp_cycles = best * 1e-3 * 1.965e9 * 148 * 4;             // at 1965 MHz
        *gops = smsp_cycles / groups;
        return DCB_OK;
    }
    *gops = ops_per_thread * grid * block / (best * 1e-3) / 1e9;
    return DCB_OK;
}

}  // extern "C"
