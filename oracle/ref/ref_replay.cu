/*
 * oracle/ref/ref_replay.cu -- TEST INFRASTRUCTURE ONLY (reference oracle + "B-ref" baseline).
 *
 * Drives the reference's UNMODIFIED device code (compiled from
 * /root/reference/SRmeetsPS-GPU/devicecalls.cu, with oracle/ref/cusparse_legacy_shim.h
 * force-included) through the same call sequence as the outer loop of
 * SRmeetsPS-GPU/SRPS.cu:206-317, starting from a post-init snapshot (SRPSNAP1) instead of the
 * OpenCV pre-processing (no OpenCV library exists here).  It dumps s, rho, z, N, dz
 * and the energy after every outer iteration and prints per-phase cudaEvent times.
 *
 * Deviations from SRPS.cu, all outside the hot path: no imshow/waitKey (SRPS.cu:319-327,338),
 * no dead MAT dumps (:330-333), xx/yy come from the snapshot (the reference's meshgrid launch
 * swaps w/h), Dx/Dy/KT come from the snapshot as COO triplets built by
 * oracle/srps_oracle.py exactly per SRPS.cu:23-71,172-193 and go through the reference's own
 * cuda_based_host_COO_to_device_CSR.
 *
 * usage: ref_replay <in.snap> <out_prefix> [--iters K | --ref-stop] [--no-dump] [--device D]
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "devicecalls.cuh"      /* the reference's own header (SRmeetsPS-GPU/devicecalls.cuh) */
#include "srps_snapshot.h"

/* --- symbols the reference defines in Main.cpp / Utilities.cpp (not compiled here: OpenCV) --- */
int Preferences::blockX = 256;      /* Main.cpp:5 */
int Preferences::blockY = 4;        /* Main.cpp:6 */
int Preferences::deviceId = 0;      /* Main.cpp:7 */

void cuda_check(std::string file, int line) {       /* behaviour of Utilities.cpp:8-19 */
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::cout << std::endl << file << ", line " << line << ": " << cudaGetErrorString(e) << " (" << e << ")" << std::endl;
        exit(1);
    }
}
void cusparse_check(cusparseStatus_t st) {          /* Utilities.cpp:21-25 */
    if (st != CUSPARSE_STATUS_SUCCESS) throw std::runtime_error("CUSPARSE ERROR " + std::to_string(st));
}
void cublas_check(cublasStatus_t st) {              /* Utilities.cpp:27-31 */
    if (st != CUBLAS_STATUS_SUCCESS) throw std::runtime_error("CUBLAS ERROR " + std::to_string(st));
}

struct DevCSR { int *row_ptr = NULL, *col_ind = NULL; float* val = NULL; int n_row = 0, n_col = 0, nnz = 0; };

static DevCSR upload_coo(cusparseHandle_t h, const srps::Snapshot& sn, const std::string& name) {
    const srps::SnapArray& shape = sn.at(name + "_shape");
    const srps::SnapArray& r = sn.at(name + "_row");
    const srps::SnapArray& c = sn.at(name + "_col");
    const srps::SnapArray& v = sn.at(name + "_val");
    DevCSR out;
    out.n_row = shape.i32()[0];
    out.n_col = shape.i32()[1];
    out.nnz = (int)r.count();
    SparseCOO<float> coo(out.n_row, out.n_col, out.nnz);
    memcpy(coo.row, r.i32(), sizeof(int) * out.nnz);
    memcpy(coo.col, c.i32(), sizeof(int) * out.nnz);
    memcpy(coo.val, v.f32(), sizeof(float) * out.nnz);
    cuda_based_host_COO_to_device_CSR(h, &coo, &out.row_ptr, &out.col_ind, &out.val);   /* SRPS.cu:192,200-201 */
    coo.freeMemory();
    return out;
}

static float* upload(const srps::SnapArray& a) {
    float* d = NULL;
    cudaMalloc(&d, a.count() * sizeof(float)); CUDA_CHECK;
    cudaMemcpy(d, a.f32(), a.count() * sizeof(float), cudaMemcpyHostToDevice); CUDA_CHECK;
    return d;
}

static void download(srps::Snapshot& out, const char* name, const float* d, std::initializer_list<int64_t> dims) {
    size_t n = 1;
    for (auto v : dims) n *= (size_t)v;
    std::vector<float> h(n);
    cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost); CUDA_CHECK;
    out.put(name, 0, dims, h.data());
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <in.snap> <out_prefix> [--iters K | --ref-stop] [--no-dump] [--device D]\n", argv[0]);
        return 2;
    }
    std::string in = argv[1], prefix = argv[2];
    int fixed_iters = 3;
    bool ref_stop = false, dump = true;
    for (int i = 3; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--iters" && i + 1 < argc) fixed_iters = atoi(argv[++i]);
        else if (a == "--ref-stop") ref_stop = true;
        else if (a == "--no-dump") dump = false;
        else if (a == "--device" && i + 1 < argc) Preferences::deviceId = atoi(argv[++i]);
    }
    try {
        srps::Snapshot sn = srps::Snapshot::load(in);
        const srps::SnapArray& I = sn.at("I");
        const int n = (int)I.dims[0], c = (int)I.dims[1], npix = (int)I.dims[2];
        const int npixs = (int)sn.at("z0s").count();
        const float K00 = sn.at("K").f32()[0], K11 = sn.at("K").f32()[4];
        const float TOLERANCE = 5e-3f;          /* SRPS.cu:85 */
        const int MAX_ITERATIONS = 10;          /* SRPS.cu:86 */

        cudaSetDevice(Preferences::deviceId);
        cusparseHandle_t cusp = 0;
        cublasHandle_t cublas = 0;
        if (cusparseCreate(&cusp) != CUSPARSE_STATUS_SUCCESS) throw std::runtime_error("cusparseCreate");
        if (cublasCreate(&cublas) != CUBLAS_STATUS_SUCCESS) throw std::runtime_error("cublasCreate");

        DevCSR KT = upload_coo(cusp, sn, "KT");
        DevCSR Dx = upload_coo(cusp, sn, "Dx");
        DevCSR Dy = upload_coo(cusp, sn, "Dy");
        if (KT.n_row != npixs || KT.n_col != npix || Dx.n_row != npix) throw std::runtime_error("operator shapes");

        /* loop-state initialisation, SRPS.cu:209-260 */
        std::vector<float> s0((size_t)n * c * 4, 0.f);
        for (int i = 0; i < n * c; i++) s0[(size_t)i * 4 + 2] = -1.f;
        float* d_s = NULL;
        cudaMalloc(&d_s, s0.size() * sizeof(float)); CUDA_CHECK;
        cudaMemcpy(d_s, s0.data(), s0.size() * sizeof(float), cudaMemcpyHostToDevice); CUDA_CHECK;
        thrust::host_vector<int> imask(npix);
        float* d_rho = cuda_based_rho_init(imask, c);
        float* d_I = upload(I);
        float* d_z0s = upload(sn.at("z0s"));
        float* d_z = upload(sn.at("z"));
        float* d_xx = upload(sn.at("xx"));
        float* d_yy = upload(sn.at("yy"));
        float* d_zx = cuda_based_sparsemat_densevec_mul(cusp, Dx.row_ptr, Dx.col_ind, Dx.val, Dx.n_row, Dx.n_col, Dx.nnz, d_z);
        float* d_zy = cuda_based_sparsemat_densevec_mul(cusp, Dy.row_ptr, Dy.col_ind, Dy.val, Dy.n_row, Dy.n_col, Dy.nnz, d_z);
        float* d_dz = NULL;
        float* d_N = cuda_based_normal_init(cublas, d_z, d_zx, d_zy, d_xx, d_yy, npix, K00, K11, &d_dz);
        cudaDeviceSynchronize(); CUDA_CHECK;

        cudaEvent_t ev[5];
        for (auto& e : ev) cudaEventCreate(&e);
        float last_error = NAN;
        int iteration = 1;
        bool stop = false;
        printf("{\"ref_replay\": {\"npix\": %d, \"npixs\": %d, \"n\": %d, \"c\": %d}}\n", npix, npixs, n, c);
        do {
            cudaEventRecord(ev[0]);
            cuda_based_lightning_estimation(cublas, cusp, d_s, d_rho, d_N, d_I, npix, n, c);                 /* SRPS.cu:281 */
            cudaEventRecord(ev[1]);
            cuda_based_albedo_estimation(cublas, cusp, d_s, d_rho, d_N, d_I, npix, n, c);                    /* SRPS.cu:287 */
            cudaEventRecord(ev[2]);
            float error = cuda_based_depth_estimation(cublas, cusp, d_s, d_rho, d_N, d_I, d_xx, d_yy, d_dz,  /* SRPS.cu:293 */
                                                      Dx.row_ptr, Dx.col_ind, Dx.val, Dx.n_row, Dx.n_col, Dx.nnz,
                                                      Dy.row_ptr, Dy.col_ind, Dy.val, Dy.n_row, Dy.n_col, Dy.nnz,
                                                      KT.row_ptr, KT.col_ind, KT.val, KT.n_row, KT.n_col, KT.nnz,
                                                      d_z0s, d_z, K00, K11, npix, n, c);
            cudaEventRecord(ev[3]);
            float rel_err = fabs(last_error - error) / fabs(error);                                        /* SRPS.cu:298-301 */
            if (ref_stop) { if (error > last_error || rel_err < TOLERANCE || iteration > MAX_ITERATIONS) stop = true; }
            else if (iteration >= fixed_iters) stop = true;
            last_error = error;
            cudaFree(d_zx); CUDA_CHECK;
            cudaFree(d_zy); CUDA_CHECK;
            d_zx = cuda_based_sparsemat_densevec_mul(cusp, Dx.row_ptr, Dx.col_ind, Dx.val, Dx.n_row, Dx.n_col, Dx.nnz, d_z);   /* :310 */
            d_zy = cuda_based_sparsemat_densevec_mul(cusp, Dy.row_ptr, Dy.col_ind, Dy.val, Dy.n_row, Dy.n_col, Dy.nnz, d_z);   /* :311 */
            cudaFree(d_dz); CUDA_CHECK;
            cudaFree(d_N); CUDA_CHECK;
            d_dz = NULL;
            d_N = cuda_based_normal_init(cublas, d_z, d_zx, d_zy, d_xx, d_yy, npix, K00, K11, &d_dz);       /* :315 */
            cudaEventRecord(ev[4]);
            cudaEventSynchronize(ev[4]); CUDA_CHECK;
            float ms[4];
            for (int k = 0; k < 4; k++) cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
            printf("{\"iteration\": %d, \"energy\": %.6f, \"rel_err\": %.6g, \"ms_lighting\": %.3f, \"ms_albedo\": %.3f, "
                   "\"ms_depth\": %.3f, \"ms_normals\": %.3f, \"ms_total\": %.3f}\n",
                   iteration, error, rel_err, ms[0], ms[1], ms[2], ms[3], ms[0] + ms[1] + ms[2] + ms[3]);
            fflush(stdout);
            if (dump) {
                srps::Snapshot out;
                download(out, "s", d_s, {n, c, 4});
                download(out, "rho", d_rho, {c, npix});
                download(out, "z", d_z, {npix});
                download(out, "N", d_N, {4, npix});
                download(out, "dz", d_dz, {npix});
                float e2[2] = {error, rel_err};
                out.put("energy", 0, {2}, e2);
                char nm[32];
                snprintf(nm, sizeof nm, "_it%02d.snap", iteration);
                out.save(prefix + nm);
            }
            iteration++;
        } while (!stop);
        printf("{\"done\": true, \"iterations\": %d}\n", iteration - 1);
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_replay: %s\n", e.what());
        return 1;
    }
    return 0;
}
