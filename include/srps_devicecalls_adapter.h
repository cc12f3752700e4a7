/*
 * srps_devicecalls_adapter.h -- the four hot-path entry points of the reference's devicecalls.cuh
 * (SRmeetsPS-GPU/devicecalls.cuh:26-37) with their ORIGINAL names and signatures, on top of libsrps_b200.so.
 *
 * A maintainer who wants to keep SRPS::execute's loop (SRPS.cu:276-317) untouched replaces the four definitions in
 * devicecalls.cu by this header and binds a context once after the device state exists:
 *
 *     srps_ctx* ctx;  srps_ctx_create(&prob, &ctx);            // geometry = mask (include/srps_c_api.h)
 *     srps_adapter_bind(ctx);
 *     ... SRPS.cu:276-317 unchanged: cuda_based_lightning_estimation(cublas_handle, cusp_handle, d_s, d_rho, d_N, d_I, ...) ...
 *
 * The cuBLAS / cuSPARSE handles and the CSR operands Dx, Dy, KT are accepted and ignored (the operators are matrix-free;
 * d_zx, d_zy, d_xx, d_yy of cuda_based_normal_init are recomputed from d_z and the mask geometry).  Errors follow the
 * reference's convention (Utilities.cpp:8-19): message to stdout, exit(1).
 * C++ only (the reference's functions have C++ linkage); no cuBLAS / cuSPARSE library is needed to use it.
 */
#ifndef SRPS_DEVICECALLS_ADAPTER_H
#define SRPS_DEVICECALLS_ADAPTER_H

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "srps_c_api.h"

/* opaque handle types exactly as cublas_api.h / cusparse.h declare them (a repeated identical typedef is legal C++) */
typedef struct cublasContext* cublasHandle_t;
typedef struct cusparseContext* cusparseHandle_t;

inline srps_ctx*& srps_adapter_ctx() { static srps_ctx* ctx = nullptr; return ctx; }
inline void srps_adapter_bind(srps_ctx* ctx) { srps_adapter_ctx() = ctx; }
inline void srps_adapter_check(int rc, const char* what) {
    if (rc != 0) { printf("\n%s: %s (%d)\n", what, srps_last_error(srps_adapter_ctx()), rc); exit(1); }
}

/* devicecalls.cu:408-444 */
inline void cuda_based_lightning_estimation(cublasHandle_t, cusparseHandle_t, float* d_s, float* d_rho, float* d_N, float* d_I,
                                            int /*npix*/, int /*nimages*/, int /*nchannels*/) {
    srps_adapter_check(srps_dev_lighting(srps_adapter_ctx(), d_s, d_rho, d_N, d_I), "cuda_based_lightning_estimation");
}

/* devicecalls.cu:513-548 */
inline void cuda_based_albedo_estimation(cublasHandle_t, cusparseHandle_t, float* d_s, float* d_rho, float* d_N, float* d_I,
                                         int /*npix*/, int /*nimages*/, int /*nchannels*/) {
    srps_adapter_check(srps_dev_albedo(srps_adapter_ctx(), d_s, d_rho, d_N, d_I), "cuda_based_albedo_estimation");
}

/* devicecalls.cu:636-786 */
inline float cuda_based_depth_estimation(cublasHandle_t, cusparseHandle_t, float* d_s, float* d_rho, float* d_N, float* d_I, float* /*d_xx*/,
                                         float* /*d_yy*/, float* d_dz, int*, int*, float*, int, int, int, int*, int*, float*, int, int, int,
                                         int*, int*, float*, int, int, int, float* d_z0s, float* d_z, float /*K00*/, float /*K11*/,
                                         int /*npix*/, int /*nimages*/, int /*nchannels*/) {
    float energy = 0.f;
    srps_adapter_check(srps_dev_depth(srps_adapter_ctx(), d_s, d_rho, d_N, d_I, d_dz, d_z0s, d_z, &energy, nullptr), "cuda_based_depth_estimation");
    return energy;
}

/* devicecalls.cu:171-223: returns a new N[4][npix] and *d_dz, both cudaMalloc'ed and owned by the caller, as the reference's does */
inline float* cuda_based_normal_init(cublasHandle_t, float* d_z, float* /*d_zx*/, float* /*d_zy*/, float* /*d_xx*/, float* /*d_yy*/, int npix,
                                     float /*K00*/, float /*K11*/, float** d_dz) {
    float* d_N = nullptr;
    if (cudaMalloc(&d_N, sizeof(float) * 4 * (size_t)npix) != cudaSuccess || cudaMalloc(d_dz, sizeof(float) * (size_t)npix) != cudaSuccess) {
        printf("\ncuda_based_normal_init: out of device memory\n");
        exit(1);
    }
    srps_adapter_check(srps_dev_normals(srps_adapter_ctx(), d_z, d_N, *d_dz), "cuda_based_normal_init");
    return d_N;
}

#endif /* SRPS_DEVICECALLS_ADAPTER_H */
