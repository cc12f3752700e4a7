// One-shot geometry analysis of the mask ON THE DEVICE.
//
// Replaces the reference's host loops over the mask -- index list SRPS.cu:157-162, LR mask :107-111 / :163-168,
// make_gradient :23-71 (four binary searches per pixel) and the masked KT build :172-193 -- which cost 0.2 s at
// 4096^2 on one host thread in round 1 (1.3 s with 8 ranks rescanning the global mask).  Here the mask is
// uploaded once (1 byte/pixel) and five small kernels produce what the loop kernels need:
//   mask_stats_kernel     bounding box, pixel counts (whole mask / before this rank's strip / inside it)
//   lr_mask_kernel        LR cell = 1 iff all sf x sf HR pixels are in the mask        (D*mask == 1, SRPS.cu:110-111)
//   stencil_types_kernel  forward-else-backward difference type per pixel               (make_gradient, SRPS.cu:31-46)
//   line_count / scan / line_fill   ordered stream compaction: dense offset of the p-th mask pixel in the
//                         reference's ascending column-major order (imask / imasks)
#pragma once
#include "srps_common.cuh"

namespace srps {

struct MaskStats {
    int imin, imax, jmin, jmax;                 // bounding box of the WHOLE mask (imin = h, imax = -1 if empty)
    unsigned long long total;                   // mask pixels in the whole image
    unsigned long long before;                  // ... in image columns j < j_lo
    unsigned long long inside;                  // ... in image columns j_lo <= j < j_hi
    unsigned long long lr_before;               // fully masked LR cells in LR columns q < j_lo / sf
};

__global__ void mask_stats_init_kernel(MaskStats* st, int h, int w) {
    st->imin = h; st->imax = -1; st->jmin = w; st->jmax = -1;
    st->total = 0ull; st->before = 0ull; st->inside = 0ull; st->lr_before = 0ull;
}

constexpr int GEO_NT = 256;

// one block per image column j (h contiguous bytes)
__global__ void __launch_bounds__(GEO_NT) mask_stats_kernel(const unsigned char* __restrict__ m, int h, int w, int j_lo, int j_hi,
                                                            MaskStats* st) {
    __shared__ int s_min[GEO_NT / 32], s_max[GEO_NT / 32], s_cnt[GEO_NT / 32];
    for (int j = blockIdx.x; j < w; j += gridDim.x) {
        int lo = h, hi = -1, cnt = 0;
        for (int i = threadIdx.x; i < h; i += GEO_NT)
            if (m[(size_t)i + (size_t)j * h]) { lo = min(lo, i); hi = max(hi, i); cnt++; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; s_cnt[threadIdx.x >> 5] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < GEO_NT / 32; k++) { lo = min(lo, s_min[k]); hi = max(hi, s_max[k]); cnt += s_cnt[k]; }
            if (cnt > 0) {
                atomicMin(&st->imin, lo); atomicMax(&st->imax, hi); atomicMin(&st->jmin, j); atomicMax(&st->jmax, j);
                atomicAdd(&st->total, (unsigned long long)cnt);
                if (j < j_lo) atomicAdd(&st->before, (unsigned long long)cnt);
                else if (j < j_hi) atomicAdd(&st->inside, (unsigned long long)cnt);
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ bool lr_cell_full(const unsigned char* __restrict__ m, int h, int w, int sf, int r, int q) {
    // HR block rows r*sf.., cols q*sf.. (Utilities.cpp:201-220); cells reaching outside the image are not full
    if (r < 0 || q < 0 || (r + 1) * sf > h || (q + 1) * sf > w) return false;
    for (int l = 0; l < sf; l++)
        for (int k = 0; k < sf; k++)
            if (!m[(size_t)(r * sf + k) + (size_t)(q * sf + l) * h]) return false;
    return true;
}

// fully masked LR cells in the LR columns q < q_hi of the GLOBAL image (strip partition: the rank's offset into imasks)
__global__ void __launch_bounds__(GEO_NT) lr_count_before_kernel(const unsigned char* __restrict__ m, int h, int w, int sf, int q_hi,
                                                                 MaskStats* st) {
    const int hs = h / sf;
    const long long ncell = (long long)hs * q_hi;
    unsigned cnt = 0;
    for (long long e = (long long)blockIdx.x * GEO_NT + threadIdx.x; e < ncell; e += (long long)gridDim.x * GEO_NT) {
        const int q = (int)(e / hs), r = (int)(e - (long long)q * hs);
        cnt += lr_cell_full(m, h, w, sf, r, q) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&st->lr_before, (unsigned long long)cnt);
}

// LR mask of the dense grid: lrmask[bl * lpitch + bx]
__global__ void __launch_bounds__(GEO_NT) lr_mask_kernel(const unsigned char* __restrict__ m, int h, int w, Grid g,
                                                         unsigned char* __restrict__ lrmask) {
    const long long ncell = (long long)g.lny * g.lnx;
    for (long long e = (long long)blockIdx.x * GEO_NT + threadIdx.x; e < ncell; e += (long long)gridDim.x * GEO_NT) {
        const int bl = (int)(e / g.lnx), bx = (int)(e - (long long)bl * g.lnx);
        lrmask[(size_t)bl * g.lpitch + bx] = lr_cell_full(m, h, w, g.sf, g.ib0 / g.sf + bx, g.jb0 / g.sf + bl) ? 1 : 0;
    }
}

// Stencil types of the dense grid, lines [-ghost, ny + ghost): ghost lines (strip partition) carry the neighbour
// strip's difference types but no T_LR.  types is the origin-offset plane (zero-initialised).
__global__ void __launch_bounds__(GEO_NT) stencil_types_kernel(const unsigned char* __restrict__ m, int h, int w, Grid g, int ghost,
                                                               const unsigned char* __restrict__ lrmask, unsigned char* __restrict__ types) {
    auto M = [&](int i, int j) -> bool { return i >= 0 && i < h && j >= 0 && j < w && m[(size_t)i + (size_t)j * h] != 0; };
    const long long lines = (long long)g.ny + 2 * ghost;
    const long long ncell = lines * g.nx;
    for (long long e = (long long)blockIdx.x * GEO_NT + threadIdx.x; e < ncell; e += (long long)gridDim.x * GEO_NT) {
        const int line = (int)(e / g.nx) - ghost, col = (int)(e % g.nx);
        const int i = g.ib0 + col, j = g.jb0 + line;
        if (!M(i, j)) continue;
        unsigned char t = T_MASK;
        if (M(i, j + 1)) t |= T_XF; else if (M(i, j - 1)) t |= T_XB;        // SRPS.cu:39-46
        if (M(i + 1, j)) t |= T_YF; else if (M(i - 1, j)) t |= T_YB;        // SRPS.cu:31-38
        if (line >= 0 && line < g.ny && lrmask[(size_t)(line / g.sf) * g.lpitch + col / g.sf]) t |= T_LR;
        types[(long long)line * g.pitch + col] = t;
    }
}

// ---- ordered compaction of flagged cells, line by line ------------------------------------------------------
// counts[line] = number of cells of the line with (flags & bit) != 0
__global__ void __launch_bounds__(GEO_NT) line_count_kernel(const unsigned char* __restrict__ flags, int pitch, int nx, int ny,
                                                            unsigned char bit, unsigned* __restrict__ counts) {
    __shared__ unsigned s_cnt[GEO_NT / 32];
    for (int line = blockIdx.x; line < ny; line += gridDim.x) {
        unsigned cnt = 0;
        for (int x = threadIdx.x; x < nx; x += GEO_NT) cnt += (flags[(long long)line * pitch + x] & bit) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < GEO_NT / 32; k++) cnt += s_cnt[k];
            counts[line] = cnt;
        }
        __syncthreads();
    }
}

// exclusive scan of counts[0..n) in place (one block), total to *total
__global__ void __launch_bounds__(1024) scan_counts_kernel(unsigned* __restrict__ counts, int n, unsigned long long* total) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned long long s_run;
    if (threadIdx.x == 0) s_run = 0ull;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < n ? counts[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            s_warp[threadIdx.x] = w;      // inclusive over the warps
        }
        __syncthreads();
        const unsigned warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        const unsigned long long run = s_run;
        if (i < n) counts[i] = (unsigned)(run + warp_off + inc - v);
        __syncthreads();
        if (threadIdx.x == 1023) s_run = run + warp_off + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_run;
}

// idx[offsets[line] + rank of the cell inside the line] = line * pitch + x, cells in ascending x
__global__ void __launch_bounds__(GEO_NT) line_fill_kernel(const unsigned char* __restrict__ flags, int pitch, int nx, int ny,
                                                           unsigned char bit, const unsigned* __restrict__ offsets, int* __restrict__ idx) {
    __shared__ unsigned s_warp[GEO_NT / 32];
    __shared__ unsigned s_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int line = blockIdx.x; line < ny; line += gridDim.x) {
        if (threadIdx.x == 0) s_base = offsets[line];
        __syncthreads();
        for (int x0 = 0; x0 < nx; x0 += GEO_NT) {
            const int x = x0 + threadIdx.x;
            const bool f = x < nx && (flags[(long long)line * pitch + x] & bit);
            const unsigned ballot = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_warp[wid] = __popc(ballot);
            __syncthreads();
            unsigned off = s_base;
            for (int k = 0; k < wid; k++) off += s_warp[k];
            if (f) idx[off + __popc(ballot & ((1u << lane) - 1u))] = line * pitch + x;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned t = 0;
                for (int k = 0; k < GEO_NT / 32; k++) t += s_warp[k];
                s_base += t;
            }
            __syncthreads();
        }
    }
}

}  // namespace srps
