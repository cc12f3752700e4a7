// Test program: the reference's loop (SRmeetsPS-GPU/SRPS.cu:262-317) written against the reference's OWN operator
// names and signatures, which include/srps_devicecalls_adapter.h supplies on top of libsrps_b200.so.  State lives on the
// device in the reference's masked layouts; nothing here knows about dense planes.  Built and run by
// tests/test_gpu_adapter.py:   adapter_replay <in.snap (post-init state)> <out.snap> <iterations> <albedo mode 0|1>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "srps_devicecalls_adapter.h"
#include "srps_snapshot.h"

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { printf("%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 5) { printf("usage: adapter_replay in.snap out.snap iterations albedo_mode\n"); return 2; }
    const srps::Snapshot in = srps::Snapshot::load(argv[1]);
    const int iters = atoi(argv[3]);
    const int* dims = in.at("dims").i32();
    const int h = dims[0], w = dims[1], sf = dims[2];
    const srps::SnapArray& aI = in.at("I");
    const int n = (int)aI.dims[0], c = (int)aI.dims[1];
    const int npix = (int)in.at("z").count(), npixs = (int)in.at("z0s").count();
    const float* K = in.at("K").f32();

    srps_problem prob = {};
    prob.h = h; prob.w = w; prob.n_images = n; prob.n_channels = c; prob.sf = sf;
    prob.fx = K[0]; prob.fy = K[4]; prob.cx = K[6]; prob.cy = K[7];
    prob.mask = in.at("mask").u8();
    prob.albedo_mode = atoi(argv[4]);
    srps_ctx* ctx = nullptr;
    if (srps_ctx_create(&prob, &ctx)) { printf("%s\n", srps_last_error(nullptr)); return 1; }
    if (srps_npix(ctx) != npix || srps_npixs(ctx) != npixs) { printf("snapshot does not match the mask\n"); return 1; }
    srps_adapter_bind(ctx);

    // device state exactly as SRPS.cu:206-260 leaves it: s = (0,0,-1,0), rho = 0.5, masked I, z, z0s
    float *d_s, *d_rho, *d_I, *d_z, *d_z0s;
    std::vector<float> s0((size_t)n * c * 4, 0.f), rho0((size_t)c * npix, 0.5f);
    for (int e = 0; e < n * c; e++) s0[(size_t)e * 4 + 2] = -1.f;
    CU(cudaMalloc(&d_s, s0.size() * 4)); CU(cudaMemcpy(d_s, s0.data(), s0.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_rho, rho0.size() * 4)); CU(cudaMemcpy(d_rho, rho0.data(), rho0.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_I, aI.count() * 4)); CU(cudaMemcpy(d_I, aI.f32(), aI.count() * 4, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_z, (size_t)npix * 4)); CU(cudaMemcpy(d_z, in.at("z").f32(), (size_t)npix * 4, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_z0s, (size_t)std::max(npixs, 1) * 4)); CU(cudaMemcpy(d_z0s, in.at("z0s").f32(), (size_t)npixs * 4, cudaMemcpyHostToDevice));

    cublasHandle_t cublas_handle = nullptr;          // accepted and ignored by the adapter
    cusparseHandle_t cusp_handle = nullptr;
    float *d_zx = nullptr, *d_zy = nullptr, *d_xx = nullptr, *d_yy = nullptr;      // recomputed from d_z by the adapter
    float* d_dz = nullptr;
    float* d_N = cuda_based_normal_init(cublas_handle, d_z, d_zx, d_zy, d_xx, d_yy, npix, K[0], K[4], &d_dz);      // SRPS.cu:269

    std::vector<float> energies;
    for (int iteration = 1; iteration <= iters; iteration++) {                                                          // SRPS.cu:276-317
        cuda_based_lightning_estimation(cublas_handle, cusp_handle, d_s, d_rho, d_N, d_I, npix, n, c);
        cuda_based_albedo_estimation(cublas_handle, cusp_handle, d_s, d_rho, d_N, d_I, npix, n, c);
        const float error = cuda_based_depth_estimation(cublas_handle, cusp_handle, d_s, d_rho, d_N, d_I, d_xx, d_yy, d_dz, nullptr, nullptr, nullptr,
                                                        npix, npix, 0, nullptr, nullptr, nullptr, npix, npix, 0, nullptr, nullptr, nullptr, npixs,
                                                        npix, 0, d_z0s, d_z, K[0], K[4], npix, n, c);
        energies.push_back(error);
        cudaFree(d_dz); cudaFree(d_N);                                                                                 // SRPS.cu:312-313
        d_dz = nullptr;
        d_N = cuda_based_normal_init(cublas_handle, d_z, d_zx, d_zy, d_xx, d_yy, npix, K[0], K[4], &d_dz);         // SRPS.cu:315
    }

    std::vector<float> z(npix), rho((size_t)c * npix), N((size_t)4 * npix), s((size_t)n * c * 4);
    CU(cudaMemcpy(z.data(), d_z, z.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(rho.data(), d_rho, rho.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(N.data(), d_N, N.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(s.data(), d_s, s.size() * 4, cudaMemcpyDeviceToHost));
    srps::Snapshot out;
    out.put("z", 0, {(int64_t)npix}, z.data());
    out.put("rho", 0, {c, (int64_t)npix}, rho.data());
    out.put("N", 0, {4, (int64_t)npix}, N.data());
    out.put("s", 0, {n, c, 4}, s.data());
    out.put("energy", 0, {(int64_t)energies.size()}, energies.data());
    out.save(argv[2]);
    srps_ctx_destroy(ctx);
    printf("adapter_replay ok: %d iterations, last energy %.6f\n", iters, energies.empty() ? 0.f : energies.back());
    return 0;
}
