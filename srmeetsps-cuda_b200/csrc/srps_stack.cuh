// Passes over the image stack I[n][c] (the only O(n*c*npix) data of the loop) and the small
// per-pixel kernels around them.
//
//   lighting_gram_kernel / lighting_reduce_kernel : replace cuda_based_lightning_estimation
//       (SRmeetsPS-GPU/devicecalls.cu:408-444): K7 `A_for_lightning_estimation`, n*c cublasSgemm
//       Gram products (identical for every image), n*c cublasSgemv, ~80 4-byte memcpys per (i,c) and
//       a host-driven CG on a 4x4 "sparse" matrix -> ONE pass over the stack with register-blocked
//       accumulators, a single-pass grid reduction and the 4x4 reference CG run by the last block.
//   stack_project_kernel : U_p[c][k] = sum_j s_jc[k] I_jcp (and sum_j I^2) in ONE pass; its fused
//       epilogue replaces cuda_based_albedo_estimation (devicecalls.cu:497-548: sgemm shading,
//       (npix*n)-row CSR expansions, SpGEMM producing a diagonal, CG) and the dense B / a1,a2,a3
//       planes of cuda_based_depth_estimation (devicecalls.cu:550-620) by per-pixel w, g, e0.
#pragma once
#include "srps_comm.cuh"

namespace srps {

constexpr int ST_NT = 128;     // threads per CTA in the stack passes
#ifndef SRPS_LIGHT_IB
#define SRPS_LIGHT_IB 4          // measured at 4096^2 x 32 (round 1): IB=4 / 3 CTAs per SM (168 registers, no spills) 1.21 ms;
#endif                           // IB=8 / 2 CTAs (255 registers, spills) 1.50 ms; IB=4 / 4 CTAs (128 registers) 1.55 ms
constexpr int LIGHT_IB = SRPS_LIGHT_IB;    // images per CTA row in the lighting reduction
constexpr int MAX_IMAGES = 64; // n_images limit (shared-memory copy of s)

// ---------------------------------------------------------------------------------------------
// Storage type of the image stack.  float: intensities in [0,1] as the reference holds them (SRPS.cu:223-232).
// unsigned char: the 8-bit samples the image loader read (Utilities.cpp:343 divides them by 255 on the host); the
// division happens in registers instead, 4x fewer stack bytes per pass and per upload.  v/255.f is reproduced exactly
// (same bits as the float path) by one reciprocal multiply and one FMA correction step -- q = v*r; q += (v - 255 q) r,
// r = fl(1/255) -- which is the correctly rounded quotient for every v in 0..255 (checked exhaustively in exact
// arithmetic, and by test_u8_upload_equals_float_upload on the device): 3 FP instructions instead of an IEEE division.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float u8_unit_of(float f) {      // f = the sample as an exact float
    const float r = 1.f / 255.f;                    // folded at compile time
    const float q = f * r;
    return fmaf(fmaf(-255.f, q, f), r, q);
}
__device__ __forceinline__ float u8_to_unit(unsigned v) { return u8_unit_of((float)v); }
// the four samples of a 32-bit word: byte k is placed under the exponent of 2^23 by one PRMT and the bias subtracted
// (exact: no integer-to-float conversion instruction, which issues at a quarter of the FP rate)
__device__ __forceinline__ float4 u8x4_to_unit(unsigned u) {
    const float f0 = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7540)) - 8388608.f;
    const float f1 = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7541)) - 8388608.f;
    const float f2 = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7542)) - 8388608.f;
    const float f3 = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7543)) - 8388608.f;
    return make_float4(u8_unit_of(f0), u8_unit_of(f1), u8_unit_of(f2), u8_unit_of(f3));
}
template <typename T> __device__ __forceinline__ float4 ld_stack4(const T* plane, long long i4);
template <> __device__ __forceinline__ float4 ld_stack4<float>(const float* plane, long long i4) { return ld4_stream(plane + 4 * i4); }
template <> __device__ __forceinline__ float4 ld_stack4<unsigned char>(const unsigned char* plane, long long i4) {
    unsigned u;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(u) : "l"(plane + 4 * i4));
    return u8x4_to_unit(u);
}

// ---------------------------------------------------------------------------------------------
// generic multi-value single-pass grid reduction (values as float per thread, partials as double)
// Returns true in the last block; totals[0..NV) then valid in shared memory `tot`.
// ---------------------------------------------------------------------------------------------
// partials[0], partials[stride], ... (n terms) summed in that order with eight independent L2 loads in flight: the
// last block of a reduction is a serial tail of the whole phase
__device__ __forceinline__ double ordered_sum(const double* partials, int n, int stride) {
    double b = 0.0;
    for (int k0 = 0; k0 < n; k0 += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (k0 + u < n) ? __ldcg(partials + (long long)(k0 + u) * stride) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++) b += v[u];
    }
    return b;
}

// SUM = false: only detect the last block (the caller sums the partials its own way).
template <int NT, int NV, bool SUM = true>
__device__ __forceinline__ bool grid_reduce_multi(const float (&v)[NV], double* partials, unsigned* ticket,
                                                  double* tot /* smem [NV] */, float* wsm /* smem [NT/32][NV] */,
                                                  int nblocks, int block_linear) {
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const float s = warp_sum(v[i]);
        if (lane == 0) wsm[wid * NV + i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NV; i += NT) {
        double b = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) b += (double)wsm[w * NV + i];
        partials[(long long)block_linear * NV + i] = b;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == (unsigned)nblocks - 1u);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (SUM) {
        // every value by PARTS threads, each over a contiguous share of the blocks, the shares added in order: a fixed
        // order (deterministic), PARTS times shorter than one thread per value
        constexpr int PARTS = (NV * 4 <= NT) ? 4 : ((NV * 2 <= NT) ? 2 : 1);
        __shared__ double part_sm[NT];
        const int per = (nblocks + PARTS - 1) / PARTS;
        if (threadIdx.x < NV * PARTS) {
            const int i = threadIdx.x / PARTS, q = threadIdx.x % PARTS;
            const int b0 = q * per, cnt = max(0, min(per, nblocks - b0));
            part_sm[threadIdx.x] = ordered_sum(partials + (long long)b0 * NV + i, cnt, NV);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < NV; i += NT) {
            double b = 0.0;
#pragma unroll
            for (int q = 0; q < PARTS; q++) b += part_sm[i * PARTS + q];
            tot[i] = b;
        }
    }
    if (threadIdx.x == 0) *ticket = 0u;
    __syncthreads();
    return true;
}

// ---------------------------------------------------------------------------------------------
// Gram matrices  AtA_c = sum_p rho_c^2 [N;1][N;1]^T       (devicecalls.cu:381,422)
// ---------------------------------------------------------------------------------------------
struct GramArgs {
    const float* rho[3];
    const float* N[3];
    long long n4;
    double* partials;      // [grid][30]
    unsigned* ticket;
    double* gram;          // out: [30] this RANK's sums, upper triangles of the three 4x4 (the cross-rank sum happens together
                           //      with the right-hand sides in lighting_reduce_kernel: one exchange per lighting update, not two)
    PeerComm comm;
};

__global__ void __launch_bounds__(ST_NT, 4) lighting_gram_kernel(const GramArgs a) {
    __shared__ double tot[30];
    __shared__ float wsm[(ST_NT / 32) * 30];
    float acc[30];
#pragma unroll
    for (int i = 0; i < 30; i++) acc[i] = 0.f;
    const long long stride = (long long)gridDim.x * ST_NT;
    for (long long i = (long long)blockIdx.x * ST_NT + threadIdx.x; i < a.n4; i += stride) {
        const float4 n0 = ld4(a.N[0] + 4 * i), n1 = ld4(a.N[1] + 4 * i), n2 = ld4(a.N[2] + 4 * i);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float4 r = ld4(a.rho[c] + 4 * i);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float rr = f4get(r, k);
                const float v[4] = {rr * f4get(n0, k), rr * f4get(n1, k), rr * f4get(n2, k), rr};
                int q = 0;
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = x; y < 4; y++) acc[c * 10 + q++] += v[x] * v[y];
            }
        }
    }
    if (grid_reduce_multi<ST_NT, 30>(acc, a.partials, a.ticket, tot, wsm, gridDim.x, blockIdx.x)) {
        if (threadIdx.x < 30) a.gram[threadIdx.x] = tot[threadIdx.x];
    }
}

// ---------------------------------------------------------------------------------------------
// Stack pass 1: rhs_{i,c} = sum_p rho_c [N;1] I_icp  (devicecalls.cu:423), then -- in the last
// block -- the warm-started 4x4 reference CG for every (i,c) (devicecalls.cu:424-437, :229-279)
// and the per-iteration lighting constants S3/S4.
// ---------------------------------------------------------------------------------------------
struct LightArgs {
    const void* I;         // stack base (origin-offset; float or unsigned char samples), plane stride `plane` in samples
    long long plane;
    const float* rho[3];
    const float* N[3];
    long long n4;
    int n_images;
    double* partials;      // [gridDim.x*gridDim.y][96]
    unsigned* ticket;
    const double* gram;    // [30] this rank's Gram sums (lighting_gram_kernel)
    float* s;              // in/out [n][3][4]
    LightConsts* lc;       // out
    int max_iter; float tol2;
    PeerComm comm;
};

__device__ inline void cg4_reference(const float* A, float* x, float* b, int max_iter, float tol2) {
    float p[4] = {0.f, 0.f, 0.f, 0.f}, om[4];
    float r0 = 0.f, r1 = 0.f;
    int k = 0;
    for (int i = 0; i < 4; i++) r1 += b[i] * b[i];
    while (r1 > tol2 && k <= max_iter) {
        k++;
        if (k == 1) { for (int i = 0; i < 4; i++) p[i] = b[i]; }
        else { const float beta = r1 / r0; for (int i = 0; i < 4; i++) p[i] = beta * p[i] + b[i]; }
        float dot = 0.f;
        for (int i = 0; i < 4; i++) {
            float o = 0.f;
            for (int j = 0; j < 4; j++) o += A[i * 4 + j] * p[j];
            om[i] = o;
            dot += p[i] * o;
        }
        const float alpha = r1 / dot;
        for (int i = 0; i < 4; i++) { x[i] += alpha * p[i]; b[i] -= alpha * om[i]; }
        r0 = r1;
        r1 = 0.f;
        for (int i = 0; i < 4; i++) r1 += b[i] * b[i];
    }
}

__device__ inline void light_consts_from_s(const float* s, int n, LightConsts* lc, int tid, int nt) {
    // S4_c = sum_j s_jc s_jc^T (upper triangle, 10) ; S3_c = its leading 3x3 (6)
    for (int e = tid; e < 30; e += nt) {
        const int c = e / 10, q = e % 10;
        int x = 0, y = 0, cnt = 0;
        for (int xx = 0; xx < 4; xx++)
            for (int yy = xx; yy < 4; yy++) { if (cnt == q) { x = xx; y = yy; } cnt++; }
        float acc = 0.f;
        for (int j = 0; j < n; j++) acc += s[(j * 3 + c) * 4 + x] * s[(j * 3 + c) * 4 + y];
        lc->S4[c][q] = acc;
        if (x < 3 && y < 3) {
            const int q3 = x * 3 - (x * (x - 1)) / 2 + (y - x);
            lc->S3[c][q3] = acc;
        }
    }
}

#ifndef SRPS_LIGHT_MINB
#define SRPS_LIGHT_MINB 3
#endif
template <typename T>
__global__ void __launch_bounds__(ST_NT, SRPS_LIGHT_MINB) lighting_reduce_kernel(const LightArgs a) {
    const T* const stack = static_cast<const T*>(a.I);
    __shared__ double tot[LIGHT_IB * 12];
    __shared__ float wsm[(ST_NT / 32) * LIGHT_IB * 12];
    const int group = blockIdx.y;
    const int i0 = group * LIGHT_IB;
    const int nimg = min(LIGHT_IB, a.n_images - i0);
    float acc[LIGHT_IB * 12];
#pragma unroll
    for (int i = 0; i < LIGHT_IB * 12; i++) acc[i] = 0.f;
    const long long stride = (long long)gridDim.x * ST_NT;
    // one float4 of pixels: acc += over c, images -- the accumulation order every variant below keeps
    auto accumulate = [&](const float4& n0, const float4& n1, const float4& n2, const float4 (&r)[3], const float4 (&v)[3][LIGHT_IB]) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float4 a0 = make_float4(r[c].x * n0.x, r[c].y * n0.y, r[c].z * n0.z, r[c].w * n0.w);
            const float4 a1 = make_float4(r[c].x * n1.x, r[c].y * n1.y, r[c].z * n1.z, r[c].w * n1.w);
            const float4 a2 = make_float4(r[c].x * n2.x, r[c].y * n2.y, r[c].z * n2.z, r[c].w * n2.w);
#pragma unroll
            for (int ii = 0; ii < LIGHT_IB; ii++) {
                float* o = acc + (ii * 3 + c) * 4;
                o[0] += v[c][ii].x * a0.x + v[c][ii].y * a0.y + v[c][ii].z * a0.z + v[c][ii].w * a0.w;
                o[1] += v[c][ii].x * a1.x + v[c][ii].y * a1.y + v[c][ii].z * a1.z + v[c][ii].w * a1.w;
                o[2] += v[c][ii].x * a2.x + v[c][ii].y * a2.y + v[c][ii].z * a2.z + v[c][ii].w * a2.w;
                o[3] += v[c][ii].x * r[c].x + v[c][ii].y * r[c].y + v[c][ii].z * r[c].z + v[c][ii].w * r[c].w;
            }
        }
    };
    if (sizeof(T) == 4) {
        for (long long i = (long long)blockIdx.x * ST_NT + threadIdx.x; i < a.n4; i += stride) {
            const float4 n0 = ld4(a.N[0] + 4 * i), n1 = ld4(a.N[1] + 4 * i), n2 = ld4(a.N[2] + 4 * i);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float4 r = ld4(a.rho[c] + 4 * i);
                float4 v[LIGHT_IB];
#pragma unroll
                for (int ii = 0; ii < LIGHT_IB; ii++)
                    v[ii] = (ii < nimg) ? ld_stack4<T>(stack + ((long long)(i0 + ii) * 3 + c) * a.plane, i) : f4zero();
                const float4 a0 = make_float4(r.x * n0.x, r.y * n0.y, r.z * n0.z, r.w * n0.w);
                const float4 a1 = make_float4(r.x * n1.x, r.y * n1.y, r.z * n1.z, r.w * n1.w);
                const float4 a2 = make_float4(r.x * n2.x, r.y * n2.y, r.z * n2.z, r.w * n2.w);
#pragma unroll
                for (int ii = 0; ii < LIGHT_IB; ii++) {
                    float* o = acc + (ii * 3 + c) * 4;
                    o[0] += v[ii].x * a0.x + v[ii].y * a0.y + v[ii].z * a0.z + v[ii].w * a0.w;
                    o[1] += v[ii].x * a1.x + v[ii].y * a1.y + v[ii].z * a1.z + v[ii].w * a1.w;
                    o[2] += v[ii].x * a2.x + v[ii].y * a2.y + v[ii].z * a2.z + v[ii].w * a2.w;
                    o[3] += v[ii].x * r.x + v[ii].y * r.y + v[ii].z * r.z + v[ii].w * r.w;
                }
            }
        }
    } else {
        // 8-bit samples: a float4 of pixels is ONE 4-byte load per plane, so a thread keeps a quarter of the bytes in flight;
        // two pixel groups (i, i + stride) are loaded together -- all channels, all images -- before either is accumulated, in
        // the same order as above (the sums stay bit-identical to the float stack's).
        for (long long i = (long long)blockIdx.x * ST_NT + threadIdx.x; i < a.n4; i += 2 * stride) {
            const long long ib = i + stride;
            const bool hb = ib < a.n4;
            const long long jb = hb ? ib : i;
            unsigned ua[3][LIGHT_IB], ub[3][LIGHT_IB];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int ii = 0; ii < LIGHT_IB; ii++) {
                    const unsigned char* pl = reinterpret_cast<const unsigned char*>(stack) + ((long long)(i0 + min(ii, nimg - 1)) * 3 + c) * a.plane;
                    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(ua[c][ii]) : "l"(pl + 4 * i));
                    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(ub[c][ii]) : "l"(pl + 4 * jb));
                }
            const float4 na0 = ld4(a.N[0] + 4 * i), na1 = ld4(a.N[1] + 4 * i), na2 = ld4(a.N[2] + 4 * i);
            const float4 ra[3] = {ld4(a.rho[0] + 4 * i), ld4(a.rho[1] + 4 * i), ld4(a.rho[2] + 4 * i)};
            const float4 nb0 = ld4(a.N[0] + 4 * jb), nb1 = ld4(a.N[1] + 4 * jb), nb2 = ld4(a.N[2] + 4 * jb);
            const float4 rb[3] = {ld4(a.rho[0] + 4 * jb), ld4(a.rho[1] + 4 * jb), ld4(a.rho[2] + 4 * jb)};
            float4 v[3][LIGHT_IB];
            auto unpack = [&](const unsigned (&u)[3][LIGHT_IB]) {
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int ii = 0; ii < LIGHT_IB; ii++) {
                        const unsigned w = u[c][ii];
                        v[c][ii] = (ii < nimg) ? u8x4_to_unit(w) : f4zero();
                    }
            };
            unpack(ua);
            accumulate(na0, na1, na2, ra, v);
            if (hb) {
                unpack(ub);
                accumulate(nb0, nb1, nb2, rb, v);
            }
        }
    }
    const int nblocks = gridDim.x * gridDim.y;
    const int lin = blockIdx.y * gridDim.x + blockIdx.x;
    if (!grid_reduce_multi<ST_NT, LIGHT_IB * 12, false>(acc, a.partials, a.ticket, tot, wsm, nblocks, lin)) return;
    // ---- last block: the blocks of one image group hold the partials of its images: per-group sums, then solve.
    __shared__ double s_new[MAX_IMAGES * 12 + 30];
    const int gx = gridDim.x;
    for (int e = threadIdx.x; e < a.n_images * 12; e += ST_NT) {
        const int img = e / 12, ck = e % 12;
        const int grp = img / LIGHT_IB, ii = img % LIGHT_IB;
        s_new[e] = ordered_sum(a.partials + (long long)grp * gx * (LIGHT_IB * 12) + ii * 12 + ck, gx, LIGHT_IB * 12);   // Atb for (img, c = ck/4, k = ck%4)
    }
    const int nrhs = a.n_images * 12;
    for (int e = threadIdx.x; e < 30; e += ST_NT) s_new[nrhs + e] = a.gram[e];          // Gram sums ride along: one exchange
    __syncthreads();
    peer_allreduce<ST_NT>(a.comm, s_new, nrhs + 30);         // strip partition: sum over the ranks
    __shared__ float gram_sm[48];                            // the three full symmetric 4x4
    if (threadIdx.x < 48) {
        const int c = threadIdx.x / 16, e = threadIdx.x % 16, x = e / 4, y = e % 4;
        const int lo = x < y ? x : y, hi = x < y ? y : x;
        const int q = lo * 4 - (lo * (lo - 1)) / 2 + (hi - lo);     // index into the upper triangle
        gram_sm[threadIdx.x] = (float)s_new[nrhs + c * 10 + q];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < a.n_images * 3; e += ST_NT) {
        const int c = e % 3;
        float A[16], x[4], b[4];
        for (int t = 0; t < 16; t++) A[t] = gram_sm[c * 16 + t];
        for (int t = 0; t < 4; t++) x[t] = a.s[e * 4 + t];
        for (int t = 0; t < 4; t++) {                         // residual Atb - AtA s   devicecalls.cu:424
            float ax = 0.f;
            for (int u = 0; u < 4; u++) ax += A[t * 4 + u] * x[u];
            b[t] = (float)s_new[e * 4 + t] - ax;
        }
        cg4_reference(A, x, b, a.max_iter, a.tol2);           // devicecalls.cu:437
        for (int t = 0; t < 4; t++) a.s[e * 4 + t] = x[t];
    }
    __threadfence_block();
    __syncthreads();
    light_consts_from_s(a.s, a.n_images, a.lc, threadIdx.x, ST_NT);
}

// ---------------------------------------------------------------------------------------------
// Per-pixel algebra shared by the fused and the reference-CG albedo paths
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float quad4(const float* S /*10*/, float n0, float n1, float n2) {
    // [n0 n1 n2 1] S4 [n0 n1 n2 1]^T with S = (00 01 02 03 11 12 13 22 23 33)
    return n0 * (S[0] * n0 + 2.f * (S[1] * n1 + S[2] * n2 + S[3])) + n1 * (S[4] * n1 + 2.f * (S[5] * n2 + S[6])) +
           n2 * (S[7] * n2 + 2.f * S[8]) + S[9];
}

// w_c = (rho_c/dz)^2, g = sum_{c,j} t_cj B_cj, e0 = sum_{c,j} B_cj^2 from U, II   
__device__ __forceinline__ void depth_coeffs_px(const LightConsts& lc, const float U[3][4], const float II[3],
                                                const float rho[3], float dz, float fx, float fy, float xx, float yy,
                                                float w[3], float g[3], float& e0) {
    float gv0 = 0.f, gv1 = 0.f, gv2 = 0.f;
    e0 = 0.f;
    const float idz = 1.f / dz;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float rd = rho[c] * idz;
        w[c] = rd * rd;
        gv0 += rd * (U[c][0] - rho[c] * lc.S4[c][3]);
        gv1 += rd * (U[c][1] - rho[c] * lc.S4[c][6]);
        gv2 += rd * (U[c][2] - rho[c] * lc.S4[c][8]);
        e0 += II[c] - 2.f * rho[c] * U[c][3] + rho[c] * rho[c] * lc.S4[c][9];
    }
    g[0] = fx * gv0 - xx * gv2;
    g[1] = fy * gv1 - yy * gv2;
    g[2] = -gv2;
}

// ---------------------------------------------------------------------------------------------
// Stack pass 2: U, II and (FUSED) closed-form albedo + depth coefficients
// ---------------------------------------------------------------------------------------------
struct ProjectArgs {
    Grid g;
    const void* I; long long plane; int n_images;     // stack base (float or unsigned char samples), plane stride in samples
    const float* s;               // [n][3][4] (device)
    const LightConsts* lc;
    const unsigned char* types;
    const float* N[3]; const float* dz;
    float* rho[3];                // FUSED: in (fallback) / out
    float* w[3]; float* gq[3]; float* e0;     // FUSED: out
    float* U;                     // !FUSED: out 15 planes (U[c][k] at plane c*4+k, II[c] at 12+c), stride plane_u
    long long plane_u;
    long long n4;
};

#ifndef SRPS_PROJ_MINB
#define SRPS_PROJ_MINB 3
#endif
#ifndef SRPS_PROJ_UNROLL
#define SRPS_PROJ_UNROLL 4       // images in flight per thread; measured at 4096^2 x 32 (round 1): 4 -> 1.18 ms, 2 -> 1.38 ms
#endif
constexpr int PROJ_UNROLL = SRPS_PROJ_UNROLL;
template <bool FUSED, typename T>
__global__ void __launch_bounds__(ST_NT, SRPS_PROJ_MINB) stack_project_kernel(const ProjectArgs a) {
    const T* const stack = static_cast<const T*>(a.I);
    __shared__ float4 s_sm[MAX_IMAGES * 3];
    for (int e = threadIdx.x; e < a.n_images * 3; e += ST_NT) s_sm[e] = *reinterpret_cast<const float4*>(a.s + 4 * e);
    __syncthreads();
    const LightConsts lc = *a.lc;
    const Grid& g = a.g;
    const long long stride = (long long)gridDim.x * ST_NT;
    const int q_per_line = g.pitch / 4;
    for (long long i = (long long)blockIdx.x * ST_NT + threadIdx.x; i < a.n4; i += stride) {
        float4 U[3][4], II[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            II[c] = f4zero();
#pragma unroll
            for (int k = 0; k < 4; k++) U[c][k] = f4zero();
        }
        // 8-bit samples: twice the images in flight (a plane's float4 of pixels is a 4-byte load)
        constexpr int UN = sizeof(T) == 1 ? 2 * PROJ_UNROLL : PROJ_UNROLL;
#pragma unroll UN
        for (int j = 0; j < a.n_images; j++) {
            float4 v[3];
#pragma unroll
            for (int c = 0; c < 3; c++) v[c] = ld_stack4<T>(stack + ((long long)j * 3 + c) * a.plane, i);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float4 sv = s_sm[j * 3 + c];
                const float sk[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    U[c][k].x += sk[k] * v[c].x; U[c][k].y += sk[k] * v[c].y;
                    U[c][k].z += sk[k] * v[c].z; U[c][k].w += sk[k] * v[c].w;
                }
                II[c].x += v[c].x * v[c].x; II[c].y += v[c].y * v[c].y;
                II[c].z += v[c].z * v[c].z; II[c].w += v[c].w * v[c].w;
            }
        }
        if (!FUSED) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
#pragma unroll
                for (int k = 0; k < 4; k++) st4(a.U + (long long)(c * 4 + k) * a.plane_u + 4 * i, U[c][k]);
                st4(a.U + (long long)(12 + c) * a.plane_u + 4 * i, II[c]);
            }
            continue;
        }
        // ---- fused epilogue: closed-form albedo (the fixed point of the reference's diagonal CG,
        //      devicecalls.cu:531,540), then w, g, e0 with the NEW albedo (devicecalls.cu:550-620)
        const long long line = i / q_per_line;
        const int x = (int)(i - line * q_per_line) * 4;
        const uchar4 t4 = *reinterpret_cast<const uchar4*>(a.types + 4 * i);
        const unsigned char tv[4] = {t4.x, t4.y, t4.z, t4.w};
        const float4 n0 = ld4(a.N[0] + 4 * i), n1 = ld4(a.N[1] + 4 * i), n2 = ld4(a.N[2] + 4 * i), dz4 = ld4(a.dz + 4 * i);
        const float4 ro0 = ld4(a.rho[0] + 4 * i), ro1 = ld4(a.rho[1] + 4 * i), ro2 = ld4(a.rho[2] + 4 * i);
        const float xx = (float)(g.jb0 + (int)line) - g.cx;
        float4 orho[3], ow[3], og[3], oe = f4zero();
#pragma unroll
        for (int c = 0; c < 3; c++) { orho[c] = f4zero(); ow[c] = f4zero(); og[c] = f4zero(); }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!(tv[k] & T_MASK)) continue;
            const float N0 = f4get(n0, k), N1 = f4get(n1, k), N2 = f4get(n2, k), dz = f4get(dz4, k);
            const float rold[3] = {f4get(ro0, k), f4get(ro1, k), f4get(ro2, k)};
            float Up[3][4], IIp[3], rho[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
#pragma unroll
                for (int kk = 0; kk < 4; kk++) Up[c][kk] = f4get(U[c][kk], k);
                IIp[c] = f4get(II[c], k);
                const float d = quad4(lc.S4[c], N0, N1, N2);
                const float b = N0 * Up[c][0] + N1 * Up[c][1] + N2 * Up[c][2] + Up[c][3];
                rho[c] = d > 0.f ? b / d : rold[c];
            }
            const float yy = (float)(g.ib0 + x + k) - g.cy;
            float w[3], gg[3], e0;
            depth_coeffs_px(lc, Up, IIp, rho, dz, g.fx, g.fy, xx, yy, w, gg, e0);
#pragma unroll
            for (int c = 0; c < 3; c++) { f4set(orho[c], k, rho[c]); f4set(ow[c], k, w[c]); f4set(og[c], k, gg[c]); }
            f4set(oe, k, e0);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            st4(a.rho[c] + 4 * i, orho[c]);
            st4(a.w[c] + 4 * i, ow[c]);
            st4(a.gq[c] + 4 * i, og[c]);
        }
        st4(a.e0 + 4 * i, oe);
    }
}

// ---------------------------------------------------------------------------------------------
// Reference-faithful albedo: CG on the diagonal systems of the three channels at once
// (devicecalls.cu:513-548 with :229-279).  blockIdx.y = channel; one CgScalars per channel.
// ---------------------------------------------------------------------------------------------
constexpr int AL_NT = 256;

struct AlbedoArgs {
    const LightConsts* lc;
    const float* U; long long plane_u;      // U[c][k] planes from stack_project_kernel<false>
    const float* N[3];
    float* rho[3];
    float* d[3]; float* r[3]; float* p[3];
    long long n4;
    CgScalars* sc;            // [3]
    double* partials;         // [3][gridDim.x]
    unsigned* ticket;         // [3]
};

// d = sum_i (N.s_ic)^2, b = sum_i (N.s_ic) I_icp, r = b - d rho, r1 = r.r   (devicecalls.cu:395-406,251)
__global__ void __launch_bounds__(AL_NT, 4) albedo_init_kernel(const AlbedoArgs a) {
    __shared__ double red[AL_NT / 32];
    const int c = blockIdx.y;
    const LightConsts lc = *a.lc;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * AL_NT;
    for (long long i = (long long)blockIdx.x * AL_NT + threadIdx.x; i < a.n4; i += stride) {
        const float4 n0 = ld4(a.N[0] + 4 * i), n1 = ld4(a.N[1] + 4 * i), n2 = ld4(a.N[2] + 4 * i);
        const float4 u0 = ld4(a.U + (long long)(c * 4 + 0) * a.plane_u + 4 * i);
        const float4 u1 = ld4(a.U + (long long)(c * 4 + 1) * a.plane_u + 4 * i);
        const float4 u2 = ld4(a.U + (long long)(c * 4 + 2) * a.plane_u + 4 * i);
        const float4 u3 = ld4(a.U + (long long)(c * 4 + 3) * a.plane_u + 4 * i);
        const float4 ro = ld4(a.rho[c] + 4 * i);
        float4 d4, r4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float N0 = f4get(n0, k), N1 = f4get(n1, k), N2 = f4get(n2, k);
            // out-of-mask cells have N = 0 and U = 0: d = S4[3][3] > 0 but b = 0 and rho = 0 -> r = 0
            const float d = quad4(lc.S4[c], N0, N1, N2);
            const float b = N0 * f4get(u0, k) + N1 * f4get(u1, k) + N2 * f4get(u2, k) + f4get(u3, k);
            const float r = b - d * f4get(ro, k);
            f4set(d4, k, d);
            f4set(r4, k, r);
            acc += (double)(r * r);
        }
        st4(a.d[c] + 4 * i, d4);
        st4(a.r[c] + 4 * i, r4);
        st4(a.p[c] + 4 * i, f4zero());
    }
    double total;
    if (grid_reduce_last<AL_NT>(acc, a.partials + (long long)c * gridDim.x, a.ticket + c, red, total)) {
        if (threadIdx.x == 0) {
            CgScalars* s = a.sc + c;
            s->r1 = total; s->r0 = 0.0; s->k = 0; s->beta = 0.f; s->alpha = 0.f;
            s->active = ((float)total > s->tol2) && (0 <= s->max_iter);
        }
    }
}

// p = r + beta p ; dot = p.(d p) ; alpha = r1/dot          (devicecalls.cu:256-269)
__global__ void __launch_bounds__(AL_NT, 4) albedo_dir_kernel(const AlbedoArgs a) {
    __shared__ double red[AL_NT / 32];
    const int c = blockIdx.y;
    CgScalars* s = a.sc + c;
    if (!s->active) return;
    const float beta = s->beta;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * AL_NT;
    for (long long i = (long long)blockIdx.x * AL_NT + threadIdx.x; i < a.n4; i += stride) {
        const float4 r4 = ld4(a.r[c] + 4 * i), d4 = ld4(a.d[c] + 4 * i);
        float4 p4 = ld4(a.p[c] + 4 * i);
        p4.x = r4.x + beta * p4.x; p4.y = r4.y + beta * p4.y; p4.z = r4.z + beta * p4.z; p4.w = r4.w + beta * p4.w;
        st4(a.p[c] + 4 * i, p4);
        acc += (double)(p4.x * (d4.x * p4.x) + p4.y * (d4.y * p4.y)) + (double)(p4.z * (d4.z * p4.z) + p4.w * (d4.w * p4.w));
    }
    double total;
    if (grid_reduce_last<AL_NT>(acc, a.partials + (long long)c * gridDim.x, a.ticket + c, red, total)) {
        if (threadIdx.x == 0) { s->dot = total; s->alpha = (float)s->r1 / (float)total; }
    }
}

// rho += alpha p ; r -= alpha d p ; r1 = r.r ; beta, k, active   (devicecalls.cu:270-274,252,262)
__global__ void __launch_bounds__(AL_NT, 4) albedo_update_kernel(const AlbedoArgs a) {
    __shared__ double red[AL_NT / 32];
    const int c = blockIdx.y;
    CgScalars* s = a.sc + c;
    if (!s->active) return;
    const float alpha = s->alpha;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * AL_NT;
    for (long long i = (long long)blockIdx.x * AL_NT + threadIdx.x; i < a.n4; i += stride) {
        const float4 p4 = ld4(a.p[c] + 4 * i), d4 = ld4(a.d[c] + 4 * i);
        float4 x4 = ld4(a.rho[c] + 4 * i), r4 = ld4(a.r[c] + 4 * i);
        x4.x += alpha * p4.x; x4.y += alpha * p4.y; x4.z += alpha * p4.z; x4.w += alpha * p4.w;
        r4.x -= alpha * (d4.x * p4.x); r4.y -= alpha * (d4.y * p4.y); r4.z -= alpha * (d4.z * p4.z); r4.w -= alpha * (d4.w * p4.w);
        st4(a.rho[c] + 4 * i, x4);
        st4(a.r[c] + 4 * i, r4);
        acc += (double)(r4.x * r4.x + r4.y * r4.y) + (double)(r4.z * r4.z + r4.w * r4.w);
    }
    double total;
    if (grid_reduce_last<AL_NT>(acc, a.partials + (long long)c * gridDim.x, a.ticket + c, red, total)) {
        if (threadIdx.x == 0) {
            s->r0 = s->r1;
            s->r1 = total;
            s->k += 1;
            s->beta = (float)total / (float)s->r0;
            s->active = ((float)total > s->tol2) && (s->k <= s->max_iter);
        }
    }
}

// w, g, e0 from the stored U / II and the (CG-updated) albedo       (devicecalls.cu:550-620)
struct CoeffArgs {
    Grid g;
    const LightConsts* lc;
    const unsigned char* types;
    const float* U; long long plane_u;
    const float* rho[3]; const float* dz;
    float* w[3]; float* gq[3]; float* e0;
    long long n4;
};

__global__ void __launch_bounds__(AL_NT, 2) depth_coeffs_kernel(const CoeffArgs a) {
    const LightConsts lc = *a.lc;
    const Grid& g = a.g;
    const int q_per_line = g.pitch / 4;
    const long long stride = (long long)gridDim.x * AL_NT;
    for (long long i = (long long)blockIdx.x * AL_NT + threadIdx.x; i < a.n4; i += stride) {
        const long long line = i / q_per_line;
        const int x = (int)(i - line * q_per_line) * 4;
        const uchar4 t4 = *reinterpret_cast<const uchar4*>(a.types + 4 * i);
        const unsigned char tv[4] = {t4.x, t4.y, t4.z, t4.w};
        float4 U[3][4], II[3], ro[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
#pragma unroll
            for (int k = 0; k < 4; k++) U[c][k] = ld4(a.U + (long long)(c * 4 + k) * a.plane_u + 4 * i);
            II[c] = ld4(a.U + (long long)(12 + c) * a.plane_u + 4 * i);
            ro[c] = ld4(a.rho[c] + 4 * i);
        }
        const float4 dz4 = ld4(a.dz + 4 * i);
        const float xx = (float)(g.jb0 + (int)line) - g.cx;
        float4 ow[3], og[3], oe = f4zero();
#pragma unroll
        for (int c = 0; c < 3; c++) { ow[c] = f4zero(); og[c] = f4zero(); }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!(tv[k] & T_MASK)) continue;
            float Up[3][4], IIp[3], rho[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
#pragma unroll
                for (int kk = 0; kk < 4; kk++) Up[c][kk] = f4get(U[c][kk], k);
                IIp[c] = f4get(II[c], k);
                rho[c] = f4get(ro[c], k);
            }
            const float yy = (float)(g.ib0 + x + k) - g.cy;
            float w[3], gg[3], e0;
            depth_coeffs_px(lc, Up, IIp, rho, f4get(dz4, k), g.fx, g.fy, xx, yy, w, gg, e0);
#pragma unroll
            for (int c = 0; c < 3; c++) { f4set(ow[c], k, w[c]); f4set(og[c], k, gg[c]); }
            f4set(oe, k, e0);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) { st4(a.w[c] + 4 * i, ow[c]); st4(a.gq[c] + 4 * i, og[c]); }
        st4(a.e0 + 4 * i, oe);
    }
}

}  // namespace srps
