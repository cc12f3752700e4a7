// One-shot depth pre-processing on the device: the parallel steps of SRPS::execute's init
// (reference SRPS.cu:117-149), which the reference runs through OpenCV on the host.
//
//   depth_mean_kernel     mean over the depth frames + "any frame is 0" flag      (reference K: devicecalls.cu:95-125)
//   -- Telea inpainting of the flagged pixels stays on the host (fast marching is a sequential front; src/host/Preprocess.cpp) --
//   max / normalise       depth / max                                              SRPS.cu:137-138
//   bilateral_kernel      cv::bilateralFilter(d = -1, sigmaColor, sigmaSpace): radius round(1.5 sigmaSpace), circular
//                         support, BORDER_REFLECT_101, weights exp(-r^2/2 ss^2 - dv^2/2 sc^2)                       SRPS.cu:139
//   cubic_{rows,cols}     cv::resize(INTER_CUBIC): Keys kernel a = -0.75, pixel-centre mapping, replicated border    SRPS.cu:149
// The arithmetic (operation order, single-precision accumulators) is that of src/host/Preprocess.cpp, which
// tests/test_cpp_host.py holds to python cv2 within a few ulp; tests/test_gpu_init.py holds these kernels to cv2 directly.
// Like the reference, everything runs on the TRANSPOSED low-resolution image: the column-major (z0_h x z0_w) buffer is a
// row-major image of rows = z0_w, cols = z0_h (SRPS.cu:130-132).
#pragma once
#include <cuda_runtime.h>

namespace srps {

constexpr int INIT_NT = 256;

__global__ void __launch_bounds__(INIT_NT) depth_mean_kernel(const float* __restrict__ z0, int n, int frames, float* __restrict__ mean,
                                                             unsigned char* __restrict__ hole) {
    const int p = blockIdx.x * INIT_NT + threadIdx.x;
    if (p >= n) return;
    float avg = 0.f;
    unsigned char h = 0;
    for (int c = 0; c < frames; c++) {
        const float v = z0[(size_t)c * n + p];
        if (v != 0.f) avg += v; else h = 1;                  // devicecalls.cu:100-105
    }
    mean[p] = avg / (float)frames;                           // always divided by the frame count
    hole[p] = h;
}

// max over positive floats: integer order == float order
__global__ void __launch_bounds__(INIT_NT) max_kernel(const float* __restrict__ v, int n, int* __restrict__ out_bits) {
    float m = 0.f;
    for (int p = blockIdx.x * INIT_NT + threadIdx.x; p < n; p += gridDim.x * INIT_NT) m = fmaxf(m, v[p]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_int(m));
}

__global__ void __launch_bounds__(INIT_NT) scale_kernel(const float* __restrict__ src, int n, const int* __restrict__ mx_bits, bool divide,
                                                        float* __restrict__ dst) {
    const int p = blockIdx.x * INIT_NT + threadIdx.x;
    if (p >= n) return;
    const float mx = __int_as_float(*mx_bits);
    dst[p] = divide ? src[p] / mx : src[p] * mx;            // SRPS.cu:138 / :140
}

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

__global__ void __launch_bounds__(INIT_NT) bilateral_kernel(const float* __restrict__ src, int rows, int cols, int radius, float gc, float gs,
                                                            float* __restrict__ dst) {
    const int e = blockIdx.x * INIT_NT + threadIdx.x;
    if (e >= rows * cols) return;
    const int i = e / cols, j = e - i * cols;
    const float c = src[e];
    float acc = 0.f, wsum = 0.f;
    for (int di = -radius; di <= radius; di++)
        for (int dj = -radius; dj <= radius; dj++) {
            const float r2 = (float)(di * di + dj * dj);
            if (sqrtf(r2) > (float)radius) continue;
            const float v = src[(size_t)reflect101(i + di, rows) * cols + reflect101(j + dj, cols)];
            const float w = expf(r2 * gs + (v - c) * (v - c) * gc);
            acc += w * v; wsum += w;
        }
    dst[e] = acc / wsum;
}

__device__ __forceinline__ void cubic_weights(float t, float w[4]) {
    const float A = -0.75f;
    w[0] = ((A * (t + 1) - 5 * A) * (t + 1) + 8 * A) * (t + 1) - 4 * A;
    w[1] = ((A + 2) * t - (A + 3)) * t * t + 1;
    w[2] = ((A + 2) * (1 - t) - (A + 3)) * (1 - t) * (1 - t) + 1;
    w[3] = 1.f - w[0] - w[1] - w[2];
}

// horizontal pass: tmp[i][x], x in [0, ocols)
__global__ void __launch_bounds__(INIT_NT) cubic_cols_kernel(const float* __restrict__ src, int rows, int cols, int ocols, float* __restrict__ tmp) {
    const long long e = (long long)blockIdx.x * INIT_NT + threadIdx.x;
    if (e >= (long long)rows * ocols) return;
    const int i = (int)(e / ocols), x = (int)(e - (long long)i * ocols);
    const double sx = (double)cols / ocols;
    const float fx = (float)((x + 0.5) * sx - 0.5);
    const int ix = (int)floorf(fx);
    float w[4]; cubic_weights(fx - ix, w);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) acc += w[k] * src[(size_t)i * cols + min(max(ix - 1 + k, 0), cols - 1)];
    tmp[e] = acc;
}

// vertical pass: dst[y][x], y in [0, orows)
__global__ void __launch_bounds__(INIT_NT) cubic_rows_kernel(const float* __restrict__ tmp, int rows, int orows, int ocols, float* __restrict__ dst) {
    const long long e = (long long)blockIdx.x * INIT_NT + threadIdx.x;
    if (e >= (long long)orows * ocols) return;
    const int y = (int)(e / ocols), x = (int)(e - (long long)y * ocols);
    const double sy = (double)rows / orows;
    const float fy = (float)((y + 0.5) * sy - 0.5);
    const int iy = (int)floorf(fy);
    float w[4]; cubic_weights(fy - iy, w);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) acc += w[k] * tmp[(size_t)min(max(iy - 1 + k, 0), rows - 1) * ocols + x];
    dst[e] = acc;
}

}  // namespace srps
