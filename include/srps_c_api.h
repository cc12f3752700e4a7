/*
 * srps_c_api.h -- C ABI of the B200-native SRmeetsPS outer loop (libsrps_b200.so).
 *
 * This is the drop-in boundary for the hot path of nihalsid/SRmeetsPS-CUDA: the do-while body
 * of SRPS::execute (SRmeetsPS-GPU/SRPS.cu:276-317) and the device operators it calls
 * (SRmeetsPS-GPU/devicecalls.cuh:26-37).  Plain pointers and sizes only; no torch / thrust /
 * cuSPARSE / cuBLAS types.  Host arrays use the reference's layouts:
 *
 *   column-major images: pixel (row i, col j) <-> i + j*h              Utilities.cpp:330,343
 *   masked vectors: the npix mask pixels in ascending linear order      SRPS.cu:157-162
 *   I   [n][c][npix]   SRPS.cu:223-232         s  [n][c][4]   SRPS.cu:209-217
 *   rho [c][npix]      devicecalls.cu:133-149  N  [4][npix]   devicecalls.cu:194-223
 *   z   [npix]         SRPS.cu:242-248         z0s [npixs]    SRPS.cu:237-239
 *
 * The one deliberate signature change against devicecalls.cuh: the CSR operands (Dx, Dy, KT) are
 * gone -- geometry is (mask, h, w, sf) and the operators are matrix-free stencils.
 *
 * Every function returns 0 on success, a cudaError_t value (>0) for CUDA failures or a negative
 * SRPS_E_* code; srps_last_error() gives the message.  A context is single-threaded, owns one
 * device, one non-blocking stream and all device memory (no allocation inside the loop).
 */
#ifndef SRPS_C_API_H
#define SRPS_C_API_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRPS_E_INVALID   (-1)   /* bad argument / unsupported configuration */
#define SRPS_E_STATE     (-2)   /* call order (e.g. iterate before upload) */
#define SRPS_E_NOMEM     (-3)

/* albedo_mode */
#define SRPS_ALBEDO_CLOSED_FORM 0   /* rho = At b / At A per pixel, fused into the stack pass */
#define SRPS_ALBEDO_REFERENCE_CG 1  /* the reference's <=101-pass CG on the diagonal system (devicecalls.cu:531,540) */

/* srps_download / srps_set_state selectors */
#define SRPS_BUF_S    0   /* [n][c][4]    */
#define SRPS_BUF_RHO  1   /* [c][npix]    */
#define SRPS_BUF_Z    2   /* [npix]       */
#define SRPS_BUF_N    3   /* [4][npix]    */
#define SRPS_BUF_DZ   4   /* [npix]       */
#define SRPS_BUF_Z0S  5   /* [npixs]      */
/* read-only internals (srps_download only; tests): per-pixel depth coefficients and CG residual */
#define SRPS_BUF_W    16  /* [3][npix]  (rho_c/dz)^2                          */
#define SRPS_BUF_G    17  /* [3][npix]  g = sum t B      (devicecalls.cu:744) */
#define SRPS_BUF_E0   18  /* [npix]     sum B^2                                */
#define SRPS_BUF_R    19  /* [npix]     CG residual plane                      */

typedef struct srps_ctx srps_ctx;

/* Replaces the geometry half of DataHandler (Utilities.h:166-181) + Preferences::deviceId
 * (Utilities.h:224-230). */
typedef struct srps_problem {
    int h, w;                    /* HR image rows / cols                       DataHandler::I_h, I_w */
    int n_images;                /* n                                          DataHandler::I_n      */
    int n_channels;              /* must be 3 (devicecalls.cu:615)             DataHandler::I_c      */
    int sf;                      /* super-resolution factor, divides h and w   DataHandler::sf       */
    float fx, fy, cx, cy;        /* K[0], K[4], K[6], K[7] of the column-major 3x3 K                 */
    const unsigned char* mask;   /* host, h*w column-major, non-zero = inside  DataHandler::mask     */
    int device;                  /* CUDA ordinal                               Preferences::deviceId */
    int albedo_mode;             /* SRPS_ALBEDO_*                                                     */
    int cg_max_iter;             /* 0 -> 100, the reference's max_iter (devicecalls.cu:231)           */
    float cg_tol;                /* 0 -> 1e-9 (devicecalls.cu:230)                                    */
    /* Strip partition of ONE scene across the GPUs of a box (no reference equivalent: SRPS.cu:88 is
     * single-GPU).  world <= 1: the context owns the whole image.  world > 1: `mask` is still the GLOBAL
     * mask; this context owns image columns [strip_j0, strip_j1) (multiples of 4), rank in [0, world). */
    int strip_j0, strip_j1;
    int rank, world;
} srps_problem;

/* Per-phase device times (cudaEvent, ms) of the last srps_outer_iteration + launch counters. */
typedef struct srps_timings {
    float ms_lighting, ms_albedo, ms_depth, ms_normals, ms_total;
    float ms_depth_cg;           /* the CG iterations alone (inside ms_depth) */
    int cg_iters;                /* depth CG passes executed */
    int albedo_cg_iters[3];
    long long launches;          /* kernels launched by this context since creation */
    int cg_deferred;             /* fused CG: passes of the last depth solve that measured r.r instead of expanding it */
    int cg_zskip;                /* persistent fused CG: passes of the last depth solve that left the depth untouched (its step is
                                    applied together with the next one: 36 instead of 44 bytes per pixel in those passes) */
} srps_timings;

/* Context: replaces cudaSetDevice + handle creation + all one-shot device setup of
 * SRPS::execute (SRPS.cu:88-115, 151-203): mask indexing, LR mask, stencil-type map. */
int  srps_ctx_create(const srps_problem* prob, srps_ctx** out);
void srps_ctx_destroy(srps_ctx* ctx);
const char* srps_last_error(const srps_ctx* ctx);   /* ctx may be NULL: last create error */
int  srps_npix(const srps_ctx* ctx);                 /* imask.size()   SRPS.cu:167 */
int  srps_npixs(const srps_ctx* ctx);                /* imasks.size()  SRPS.cu:168 */

/* Loop-state upload from HOST buffers in the reference's masked layouts; initialises
 * s=(0,0,-1,0), rho=0.5 and the first normals exactly as SRPS.cu:209-270.  I may be NULL if
 * srps_upload_images_u8 is used instead. */
int  srps_upload_state(srps_ctx* ctx, const float* I, const float* z, const float* z0s);
/* Same stack as 8-bit samples (I = v/255, the image loader's arithmetic, Utilities.cpp:343).  The samples STAY 8-bit
 * in device memory (a quarter of the stack bytes per pass and per upload); the two stack passes divide by 255 in
 * registers, bit-identical to uploading the floats v/255.f.  The _strided form reads a strip's run out of the planes
 * of a global stack (see srps_upload_state_strided). */
int  srps_upload_images_u8(srps_ctx* ctx, const unsigned char* I8);
int  srps_upload_images_u8_strided(srps_ctx* ctx, const unsigned char* I8, long long plane_stride);
/* Overwrite one state buffer from host (tests / resume).  After SRPS_BUF_Z call srps_normals. */
int  srps_set_state(srps_ctx* ctx, int which, const float* host);
int  srps_download(srps_ctx* ctx, int which, float* host);

/* The four operators of the loop body, in the reference's order. */
int  srps_lighting(srps_ctx* ctx);                                  /* cuda_based_lightning_estimation  devicecalls.cu:408-444 */
int  srps_albedo(srps_ctx* ctx);                                    /* cuda_based_albedo_estimation     devicecalls.cu:513-548 */
int  srps_depth(srps_ctx* ctx, float* energy, int* cg_iters);       /* cuda_based_depth_estimation      devicecalls.cu:636-786 */
int  srps_normals(srps_ctx* ctx);                                   /* zx,zy + cuda_based_normal_init   SRPS.cu:310-315, devicecalls.cu:194-223 */

/* One pass of the do-while body (SRPS.cu:276-317): lighting, albedo, depth (+energy), normals. */
int  srps_outer_iteration(srps_ctx* ctx, float* energy, int* cg_iters);
/* The whole loop with the reference's stop rule (SRPS.cu:298-301): stop when the energy rises,
 * the relative change is < tol, or iteration > max_outer.  fixed_iters > 0 overrides the rule.
 * energies (may be NULL) receives up to cap values; returns the number of passes in *n_done. */
int  srps_run(srps_ctx* ctx, int max_outer, float tol, int fixed_iters, float* energies, int cap, int* n_done);

int  srps_get_timings(const srps_ctx* ctx, srps_timings* out);
int  srps_synchronize(srps_ctx* ctx);

/* ---- strip partition (one process per GPU; peers are mapped with CUDA IPC over NVLink) -------------
 * 1. every rank: srps_ctx_create with its strip, then srps_dist_export -> an opaque blob;
 * 2. the caller all-gathers the blobs (torch.distributed / MPI / files) in rank order;
 * 3. every rank: srps_dist_connect(ctx, all_blobs, world).
 * Afterwards the operators above are COLLECTIVE: every rank must call them in the same order.
 * Host vectors passed to a strip context cover only its pixels: srps_pixel_range gives the half-open
 * ranges [p0,p1) of the global masked vector and [q0,q1) of the global LR masked vector it owns
 * (strips cut the column-major masked order into contiguous runs). */
int  srps_dist_blob_size(void);
int  srps_dist_export(srps_ctx* ctx, void* blob);
int  srps_dist_connect(srps_ctx* ctx, const void* blobs, int world);
int  srps_pixel_range(const srps_ctx* ctx, long long* p0, long long* p1, long long* q0, long long* q1);
/* srps_upload_state with an explicit plane stride (floats) between the n*c planes of I: a strip context
 * reads its run [p0,p1) out of each plane of the global stack (pass I + p0, stride = global npix). */
int  srps_upload_state_strided(srps_ctx* ctx, const float* I, long long plane_stride, const float* z, const float* z0s);

/* Device-side stopwatch on the context's stream (cudaEvent): everything the context enqueues
 * between start and stop -- uploads, kernels, downloads -- is inside the measured interval. */
int  srps_timer_start(srps_ctx* ctx);
int  srps_timer_stop(srps_ctx* ctx, float* ms);

/* Measurement hook (bench.py roofline): average device time (ms, cudaEvent on the launching
 * stream) of `reps` back-to-back launches of one kernel alone.  out_ms[6]: [0] = CG operator kernel
 * (p <- r + beta p; y <- A p; p.y), [1] = CG update kernel, [2] = lighting stack pass,
 * [3] = stack-projection pass (+ fused albedo / depth coefficients), [4] = fused CG pass (operator +
 * the previous pass's update in one kernel; 0 if sf > 4), [5] = CG driver this context uses
 * (0 operator + update graph, 1 persistent cooperative kernel, 2 fused pass graph, 3 fused passes in one
 * persistent cooperative kernel).  Needs one completed
 * srps_outer_iteration; leaves the loop state UNDEFINED (re-upload before further use). */
int  srps_profile_kernels(srps_ctx* ctx, int reps, float* out_ms);

/* Test hook: y = (Kt K + G^T M G) p with the M of the current rho/dz/s (masked host vectors). */
int  srps_apply_depth_operator(srps_ctx* ctx, const float* p_host, float* y_host);

/* ---- device-pointer forms of the four operators: the seam of devicecalls.cuh:26-37 itself -------------------------
 * The caller keeps the loop state on the device in the reference's masked layouts and passes raw device pointers, as
 * SRPS.cu:276-317 does.  Operands are imported into the context (device-to-device), results written back in place:
 * srps_dev_lighting updates d_s, srps_dev_albedo d_rho, srps_dev_depth d_z (and returns the energy), srps_dev_normals
 * fills d_N[4][npix] and d_dz for the given d_z.  d_I is imported when first seen.  Call them in the reference's order
 * (lighting, albedo, depth, normals); single GPU.  include/srps_devicecalls_adapter.h gives them the reference's names. */
int  srps_dev_lighting(srps_ctx* ctx, float* d_s, const float* d_rho, const float* d_N, const float* d_I);
int  srps_dev_albedo(srps_ctx* ctx, const float* d_s, float* d_rho, const float* d_N, const float* d_I);
int  srps_dev_depth(srps_ctx* ctx, const float* d_s, const float* d_rho, const float* d_N, const float* d_I, const float* d_dz,
                    const float* d_z0s, float* d_z, float* energy, int* cg_iters);
int  srps_dev_normals(srps_ctx* ctx, const float* d_z, float* d_N, float* d_dz);

/* ---- one-shot depth pre-processing on the device (the parallel steps of SRPS.cu:117-149; no context needed) ----------
 * srps_init_depth_mean: mean over `frames` low-resolution depth frames (z0: [frames][n_lr], any pixel order) divided by
 * the frame count, hole = 1 where any frame is 0 (devicecalls.cu:95-125).  The flagged pixels are then inpainted by the
 * caller (Telea fast marching, SRPS.cu:133: a sequential front, host code in src/host/Preprocess.cpp).
 * srps_init_depth_smooth_upsample: depth / max -> cv::bilateralFilter(-1, sigma_color, sigma_space) -> * max (= zs, the
 * smoothed LR depth, rows x cols) -> cv::resize(INTER_CUBIC) to orows x ocols (= z_full).  Row-major images; the
 * reference passes its column-major buffers transposed: rows = z0_w, cols = z0_h, orows = I_w, ocols = I_h (SRPS.cu:130-149). */
int  srps_init_depth_mean(int device, const float* z0, int n_lr, int frames, float* mean_out, unsigned char* hole_out);
int  srps_init_depth_smooth_upsample(int device, const float* depth, int rows, int cols, int orows, int ocols,
                                     float sigma_color, float sigma_space, float* zs_out, float* z_full_out);

/* Build identification: "sm_100a;<git or date>" */
const char* srps_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* SRPS_C_API_H */
