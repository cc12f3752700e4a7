// SRPS::execute (reference SRmeetsPS-GPU/SRPS.cu:84-370) on top of the C ABI.
//   SRPS.cu:105-149   LR mask + depth pre-processing      -> host (Preprocess.cpp), same progress strings
//   SRPS.cu:151-270   index loops, operator assembly, device state init -> srps_ctx_create + srps_upload_state
//   SRPS.cu:272-335   the outer loop                      -> srps_outer_iteration per pass, same stop rule and prints
//   SRPS.cu:319-338   imshow / waitKey                    -> dropped (headless); optional result dump instead
#include "SRPS.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

#include "../../include/srps_c_api.h"
#include "../../include/srps_snapshot.h"

SRPS::SRPS(DataHandler& dh) { this->dh = &dh; }
SRPS::SRPS(SnapshotState& st) { this->dh = nullptr; this->snap = &st; }
SRPS::~SRPS() {}

static void die(srps_ctx* ctx, const char* what) {      // error convention of Utilities.cpp:8-19: print and exit(1)
    std::cout << std::endl << what << ": " << srps_last_error(ctx) << std::endl;
    exit(1);
}

void SRPS::execute() {
    const float TOLERANCE = 5e-3f;                       // SRPS.cu:85
    const int MAX_ITERATIONS = Preferences::maxOuter > 0 ? Preferences::maxOuter : 10;   // SRPS.cu:86
    SnapshotState st;
    if (snap) {
        st = *snap;
    } else {
        const int h = dh->I_h, w = dh->I_w, sf = (int)dh->sf;
        if (dh->I_c != 3) throw std::runtime_error("nchannels must be 3 (reference: devicecalls.cu:615)");
        st.h = h; st.w = w; st.sf = sf; st.n = dh->I_n; st.c = dh->I_c;
        st.K.assign(dh->K, dh->K + 9);
        st.mask.resize((size_t)h * w);
        for (size_t i = 0; i < st.mask.size(); i++) st.mask[i] = dh->mask[i] != 0;            // SRPS.cu:158
        std::cout << "Small mask calculation" << std::endl;                                       // SRPS.cu:106
        const int hs = h / sf, ws = w / sf;
        std::vector<unsigned char> masks((size_t)hs * ws, 0);                                     // D*mask == 1  SRPS.cu:110-111
        for (int q = 0; q < ws; q++)
            for (int r = 0; r < hs; r++) {
                bool all = true;
                for (int l = 0; l < sf && all; l++)
                    for (int k = 0; k < sf; k++)
                        if (!st.mask[(size_t)(r * sf + k) + (size_t)(q * sf + l) * h]) { all = false; break; }
                masks[(size_t)r + (size_t)q * hs] = all;
            }
        std::cout << "Mean of depth values" << std::endl;                                         // SRPS.cu:119
        std::cout << "Inpainting depth values" << std::endl;                                      // SRPS.cu:129
        std::cout << "Smoothing depth" << std::endl;                                              // SRPS.cu:135
        std::vector<float> zs, z_full;
        preprocess_depth(dh->z0, dh->z0_h, dh->z0_w, dh->z0_n, h, w, zs, z_full);
        std::cout << "Resample depths" << std::endl;                                              // SRPS.cu:146
        std::cout << "Mask index calculation" << std::endl;                                       // SRPS.cu:152
        std::vector<size_t> imask;
        for (size_t i = 0; i < st.mask.size(); i++) if (st.mask[i]) imask.push_back(i);           // SRPS.cu:157-162
        std::cout << "Masked resample matrix" << std::endl;                                       // SRPS.cu:171
        std::cout << "Masked gradient matrix" << std::endl;                                       // SRPS.cu:196
        std::cout << "Initialization" << std::endl;                                               // SRPS.cu:206
        const size_t npix = imask.size();
        st.I.resize((size_t)st.n * st.c * npix);
        for (int n = 0; n < st.n; n++)                                                            // SRPS.cu:227-232
            for (int c = 0; c < st.c; c++) {
                const float* src = dh->I + ((size_t)n * st.c + c) * h * w;
                float* dst = st.I.data() + ((size_t)n * st.c + c) * npix;
                for (size_t p = 0; p < npix; p++) dst[p] = src[imask[p]];
            }
        st.z.resize(npix);
        for (size_t p = 0; p < npix; p++) st.z[p] = z_full[imask[p]];                             // SRPS.cu:246
        for (size_t i = 0; i < masks.size(); i++) if (masks[i]) st.z0s.push_back(zs[i]);          // SRPS.cu:239
    }
    if (!dump_init.empty()) st.save(dump_init);
    if (init_only) { std::cout << "Done!" << std::endl; return; }

    srps_problem prob = {};
    prob.h = st.h; prob.w = st.w; prob.n_images = st.n; prob.n_channels = st.c; prob.sf = st.sf;
    prob.fx = st.K[0]; prob.fy = st.K[4]; prob.cx = st.K[6]; prob.cy = st.K[7];
    prob.mask = st.mask.data();
    prob.device = Preferences::deviceId;                                                          // SRPS.cu:88
    prob.albedo_mode = Preferences::albedoMode;
    srps_ctx* ctx = nullptr;
    if (srps_ctx_create(&prob, &ctx)) die(nullptr, "srps_ctx_create");
    if ((size_t)srps_npix(ctx) != st.z.size() || (size_t)srps_npixs(ctx) != st.z0s.size()) {
        std::cout << "state does not match the mask (npix " << srps_npix(ctx) << " vs " << st.z.size() << ")" << std::endl;
        exit(1);
    }
    if (srps_upload_state(ctx, st.I.data(), st.z.data(), st.z0s.data())) die(ctx, "srps_upload_state");

    float last_error = NAN;                                                                       // SRPS.cu:273-275
    bool stop_loop = false;
    int iteration = 1;
    energies.clear();
    do {
        float error = 0.f;
        int cg = 0;
        if (srps_outer_iteration(ctx, &error, &cg)) die(ctx, "srps_outer_iteration");           // SRPS.cu:281-315
        srps_timings t;
        srps_get_timings(ctx, &t);
        printf("\n%-25s: %-6.6fs\n", "Lightning Estimation", t.ms_lighting * 1e-3f);             // SRPS.cu:283
        printf("%-25s: %-6.6fs\n", "Albedo Estimation", t.ms_albedo * 1e-3f);                     // SRPS.cu:289
        printf("%-25s: %-6.6fs\n", "Depth Estimation", t.ms_depth * 1e-3f);                       // SRPS.cu:295
        const float rel_err = fabsf(last_error - error) / fabsf(error);                           // SRPS.cu:298
        if (error > last_error || rel_err < TOLERANCE || iteration > MAX_ITERATIONS) stop_loop = true;   // SRPS.cu:299-301
        if (fixed_iters > 0) stop_loop = iteration >= fixed_iters;
        last_error = error;
        printf("\nIteration %02d summary\n", iteration);                                          // SRPS.cu:303-305
        printf("%-25s: %-6.3f\n", "Error", error);
        printf("%-25s: %-6.3f\n", "Relative Error", rel_err);
        energies.push_back(error);
        iteration++;
    } while (!stop_loop);
    std::cout << "Done!" << std::endl;                                                            // SRPS.cu:337

    if (!dump_result.empty() || !dump_dir.empty()) {
        const size_t npix = st.z.size();
        std::vector<float> z(npix), rho(3 * npix), N(4 * npix), s((size_t)st.n * 12);
        if (srps_download(ctx, SRPS_BUF_Z, z.data()) || srps_download(ctx, SRPS_BUF_RHO, rho.data()) ||
            srps_download(ctx, SRPS_BUF_N, N.data()) || srps_download(ctx, SRPS_BUF_S, s.data()))
            die(ctx, "srps_download");
        if (!dump_result.empty()) {
            srps::Snapshot out;
            const int32_t hw[2] = {st.h, st.w};
            out.put("hw", 1, {2}, hw);
            out.put("mask", 2, {(int64_t)st.w, (int64_t)st.h}, st.mask.data());       // column-major h x w
            out.put("z", 0, {(int64_t)npix}, z.data());
            out.put("rho", 0, {3, (int64_t)npix}, rho.data());
            out.put("N", 0, {4, (int64_t)npix}, N.data());
            out.put("s", 0, {st.n, 3, 4}, s.data());
            out.put("energy", 0, {(int64_t)energies.size()}, energies.data());
            out.save(dump_result);
        }
        if (!dump_dir.empty())                                                          // SRPS.cu:319-333, as files
            save_results(dump_dir, st.h, st.w, st.mask.data(), npix, st.n, z.data(), rho.data(), N.data(), s.data());
    }
    srps_ctx_destroy(ctx);
}
