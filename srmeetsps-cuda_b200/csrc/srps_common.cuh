// Shared device-side definitions for the B200 SRmeetsPS kernels (sm_100a).
//
// Data layout in HBM ("dense grid"): the mask's bounding box, rounded out to multiples of sf,
// viewed as `ny` lines (image columns j, the slow axis of the reference's column-major layout)
// of `nx` pixels (image rows i, contiguous).  Every per-pixel array is a plane of
// (ny + 2*GUARD_LINES) x pitch floats, pitch = round_up(nx + 1, 32): there is always >= 1 zero
// pad element after each line and 2 zero guard lines before/after the grid, so every radius-1
// (and float4-aligned radius-4) neighbour read is in bounds and reads 0 outside the image.
// Out-of-mask cells hold 0 in every plane and every kernel keeps them 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srps {

constexpr int GUARD_LINES = 2;

// stencil-type bits per pixel (make_gradient semantics, reference SRPS.cu:23-71)
constexpr unsigned char T_MASK = 1;   // pixel is in the mask
constexpr unsigned char T_XF = 2;     // x (image column j = line) forward difference   SRPS.cu:39
constexpr unsigned char T_XB = 4;     // x backward difference                           SRPS.cu:43
constexpr unsigned char T_YF = 8;     // y (image row i = fast axis) forward difference  SRPS.cu:31
constexpr unsigned char T_YB = 16;    // y backward difference                           SRPS.cu:35
constexpr unsigned char T_LR = 32;    // the pixel's sf x sf block is fully masked (has a depth prior, SRPS.cu:111)

struct Grid {
    int nx, ny;          // pixels per line / lines
    int pitch;           // floats per line
    int ib0, jb0;        // image row / column of dense (line 0, col 0)
    int sf;
    int lnx, lny, lpitch;  // LR grid (nx/sf, ny/sf) and its pitch
    float fx, fy, cx, cy;
    long long plane;     // floats per plane incl. guards
    __host__ __device__ long long origin() const { return (long long)GUARD_LINES * pitch; }
};

// Per-outer-iteration constants derived from the lighting vectors s[n][c][4].
struct LightConsts {
    float S3[3][6];    // per channel: sum_j s[0:3] s[0:3]^T   (00 01 02 11 12 22)
    float S4[3][10];   // per channel: sum_j s s^T (4x4)       (00 01 02 03 11 12 13 22 23 33)
};

// Device-resident CG scalars: the host never reads them inside the solve
// (the reference's loop control, devicecalls.cu:251-275, runs on the device).
struct CgScalars {
    double r1, r0, dot;
    float alpha, beta;
    int k;            // completed passes
    int active;       // r1 > tol^2 && k <= max_iter
    int max_iter;
    float tol2;
    int plane;        // fused CG: the ping-pong plane that holds the search direction of the pending z step
    int defer;        // fused CG: the expanded |r_{k+1}|^2 cancelled (< 1e-6 r.r): the next pass only applies the pending
                      // step and MEASURES r.r, beta comes from the measurement (see cg_fused_kernel)
    int profile;      // srps_profile_kernels: keep the scalars as set by the host (timing of one pass in isolation)
    int n_defer;      // deferred passes of this solve (diagnostics: srps_timings.cg_deferred)
    int n_zskip;      // persistent fused CG: passes of this solve that left z untouched (srps_timings.cg_zskip)
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming (read-once) 128-bit load: keep it out of L1
__device__ __forceinline__ float4 ld4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// system-scope relaxed 128-bit load: for lines that live in a PEER GPU's memory and change between passes
// (never served from this SM's L1)
__device__ __forceinline__ float4 ld4_sys(const float* p) {
    float4 r;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
// loads of planes that are REWRITTEN inside the running kernel (single-launch persistent CG): 0 = never (one launch per
// pass: read-only path), 1 = by this GPU only (plain load, ordered by the grid barrier's acquire), 2 = possibly by a peer
// GPU (ghost lines read in place over NVLink: system scope, never served from this SM's L1)
template <int COH>
__device__ __forceinline__ float4 ld4_coh(const float* p) {
    if (COH == 0) return ldg4(p);
    if (COH == 1) return ld4(p);
    return ld4_sys(p);
}
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4get(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
__device__ __forceinline__ void f4set(float4& v, int k, float s) { if (k == 0) v.x = s; else if (k == 1) v.y = s; else if (k == 2) v.z = s; else v.w = s; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic single-pass grid reduction of one double per thread.
// Every block publishes its partial, the last block to arrive (atomic ticket) sums the partials in
// a fixed order and returns true with the total in `total` (valid in thread 0 of that block).
// `red_smem` needs NT/32 doubles; `partials` needs gridDim.x doubles; `*ticket` must be 0 on entry
// and is reset to 0 by the last block.
template <int NT>
__device__ __forceinline__ bool grid_reduce_last(double v, double* partials, unsigned* ticket, double* red_smem,
                                                 double& total, bool peer_stores = false) {
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red_smem[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int i = 0; i < NT / 32; i++) b += red_smem[i];
        partials[blockIdx.x] = b;
        // a block that stored halo lines into PEER memory (srps_comm.cuh) needs the system-scope fence so that
        // "mailbox flag seen" implies "halo arrived"; everybody else only needs device scope
        if (peer_stores) __threadfence_system(); else __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    double acc = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += NT) acc += __ldcg(partials + i);
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) red_smem[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int i = 0; i < NT / 32; i++) b += red_smem[i];
        total = b;
        *ticket = 0u;
    }
    return true;
}

// Sum of partials[b * stride + off], b = lane, lane + 32, ... < n, in that order (the fixed order every reduction of
// this library uses), with sixteen independent L2 loads in flight per lane -- one round trip for the <= 512 blocks of a
// resident grid: the final sum of a single-pass grid reduction sits on the critical path of every CG pass (ncu, round 2,
// 512-line slab: 14 % of the warp samples of the persistent kernel waited on these adds with eight in flight).
__device__ __forceinline__ double lane_strided_sum(const double* partials, int n, int stride, int off, int lane) {
    constexpr int U = 16;
    double acc = 0.0;
    for (int b0 = 0; b0 < n; b0 += 32 * U) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int b = b0 + u * 32 + lane;
            v[u] = b < n ? __ldcg(partials + (long long)b * stride + off) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u];      // adding 0.0 for b >= n leaves the sum unchanged
    }
    return acc;
}

// q = M (dxp, dyp, pc) with M = T Q T^T, Q = sum_c w_c S3_c, T = [[fx,0,-xx],[0,fy,-yy],[0,0,-1]]
// (the reference's rows (rho_c/dz)[fx s0 - xx s2, fy s1 - yy s2, -s2],
//  devicecalls.cu:583-620, folded over images and channels).
struct Qm { float q00, q01, q02, q11, q12, q22; };
__device__ __forceinline__ Qm make_qm(const LightConsts& lc, float w0, float w1, float w2) {
    Qm m;
    m.q00 = w0 * lc.S3[0][0] + w1 * lc.S3[1][0] + w2 * lc.S3[2][0];
    m.q01 = w0 * lc.S3[0][1] + w1 * lc.S3[1][1] + w2 * lc.S3[2][1];
    m.q02 = w0 * lc.S3[0][2] + w1 * lc.S3[1][2] + w2 * lc.S3[2][2];
    m.q11 = w0 * lc.S3[0][3] + w1 * lc.S3[1][3] + w2 * lc.S3[2][3];
    m.q12 = w0 * lc.S3[0][4] + w1 * lc.S3[1][4] + w2 * lc.S3[2][4];
    m.q22 = w0 * lc.S3[0][5] + w1 * lc.S3[1][5] + w2 * lc.S3[2][5];
    return m;
}
__device__ __forceinline__ void apply_m(const Qm& m, float fx, float fy, float xx, float yy, float dxp, float dyp,
                                        float pc, float& q0, float& q1, float& q2) {
    const float u0 = fx * dxp, u1 = fy * dyp, u2 = -(xx * dxp + yy * dyp + pc);
    const float v0 = m.q00 * u0 + m.q01 * u1 + m.q02 * u2;
    const float v1 = m.q01 * u0 + m.q11 * u1 + m.q12 * u2;
    const float v2 = m.q02 * u0 + m.q12 * u1 + m.q22 * u2;
    q0 = fx * v0 - xx * v2;
    q1 = fy * v1 - yy * v2;
    q2 = -v2;
}

}  // namespace srps
