// Depth-update epilogue and the layout conversion kernels at the C-ABI boundary.
//
//   normals_energy_kernel : replaces, after the depth CG, the two cusparseScsrmv for zx,zy
//       (SRmeetsPS-GPU/SRPS.cu:310-311), cuda_based_normal_init (devicecalls.cu:171-223: 2 saxpy +
//       K4-K6) and the photometric half of the energy (devicecalls.cu:763,765,767: a csrmv with the
//       (c*n*npix)-row A, squared_difference, thrust::reduce) -> one pass over z.
//   energy_depth_kernel   : ||K z - z0s||^2 (devicecalls.cu:762,764,766).
//   scatter/gather        : reference masked vectors <-> dense grid.
#pragma once
#include "srps_comm.cuh"

namespace srps {

constexpr int EP_NT = 256;

struct NormalsArgs {
    Grid g;
    const unsigned char* types;
    const float* z;
    float* N[3]; float* dz;
    // energy (ENERGY=true): lagged coefficients of this outer iteration
    const LightConsts* lc;
    const float* w[3]; const float* gq[3]; const float* e0;
    double* partials; unsigned* ticket; double* energy_out;   // energy_out[0] = photometric term
    long long n4;
    PeerComm comm;
};

template <bool ENERGY>
__global__ void __launch_bounds__(EP_NT, 3) normals_energy_kernel(const NormalsArgs a) {
    __shared__ double red[EP_NT / 32];
    const Grid& g = a.g;
    const int pitch = g.pitch, q_per_line = pitch / 4;
    LightConsts lc;
    if (ENERGY) lc = *a.lc;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * EP_NT;
    for (long long i = (long long)blockIdx.x * EP_NT + threadIdx.x; i < a.n4; i += stride) {
        const long long line = i / q_per_line;
        const int x = (int)(i - line * q_per_line) * 4;
        const float* zc = a.z + 4 * i;
        const uchar4 t4 = *reinterpret_cast<const uchar4*>(a.types + 4 * i);
        const unsigned char tv[4] = {t4.x, t4.y, t4.z, t4.w};
        const float4 c4 = ld4(zc), u4 = ld4(zc - pitch), d4 = ld4(zc + pitch);
        const float zl = zc[-1], zr = zc[4];
        const float zv[6] = {zl, c4.x, c4.y, c4.z, c4.w, zr};
        float4 w0, w1, w2, g0, g1, g2, e04;
        if (ENERGY) {
            w0 = ld4(a.w[0] + 4 * i); w1 = ld4(a.w[1] + 4 * i); w2 = ld4(a.w[2] + 4 * i);
            g0 = ld4(a.gq[0] + 4 * i); g1 = ld4(a.gq[1] + 4 * i); g2 = ld4(a.gq[2] + 4 * i);
            e04 = ld4(a.e0 + 4 * i);
        }
        const float xx = (float)(g.jb0 + (int)line) - g.cx;
        float4 o0 = f4zero(), o1 = f4zero(), o2 = f4zero(), od = f4zero();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned char t = tv[k];
            if (!(t & T_MASK)) continue;
            const float zz = zv[k + 1];
            const float zx = (t & T_XF) ? f4get(d4, k) - zz : ((t & T_XB) ? zz - f4get(u4, k) : 0.f);   // Dx z   SRPS.cu:310
            const float zy = (t & T_YF) ? zv[k + 2] - zz : ((t & T_YB) ? zz - zv[k] : 0.f);            // Dy z   SRPS.cu:311
            const float yy = (float)(g.ib0 + x + k) - g.cy;
            const float n0 = g.fx * zx, n1 = g.fy * zy;                      // devicecalls.cu:204,211
            const float n2 = -zz - xx * zx - yy * zy;                         // devicecalls.cu:174
            const float nrm = fmaxf(1e-10f, sqrtf(n0 * n0 + n1 * n1 + n2 * n2));   // devicecalls.cu:182
            f4set(o0, k, n0 / nrm); f4set(o1, k, n1 / nrm); f4set(o2, k, n2 / nrm); f4set(od, k, nrm);
            if (ENERGY) {
                // sum_{c,j} (t_cj . Gz - B_cj)^2 = Gz^T M Gz - 2 g.Gz + e0, lagged M,g,e0, new z (devicecalls.cu:763-767)
                const Qm m = make_qm(lc, f4get(w0, k), f4get(w1, k), f4get(w2, k));
                float q0, q1, q2;
                apply_m(m, g.fx, g.fy, xx, yy, zx, zy, zz, q0, q1, q2);
                const double quad = (double)zx * q0 + (double)zy * q1 + (double)zz * q2;
                const double lin = (double)zx * f4get(g0, k) + (double)zy * f4get(g1, k) + (double)zz * f4get(g2, k);
                acc += quad - 2.0 * lin + (double)f4get(e04, k);
            }
        }
        st4(a.N[0] + 4 * i, o0); st4(a.N[1] + 4 * i, o1); st4(a.N[2] + 4 * i, o2); st4(a.dz + 4 * i, od);
    }
    if (ENERGY) {
        double total;
        if (grid_reduce_last_world<EP_NT>(acc, a.partials, a.ticket, red, total, a.comm)) {
            if (threadIdx.x == 0) a.energy_out[0] = total;
        }
    }
}

// ||K z - z0s||^2 over the masked LR pixels                     devicecalls.cu:762,764,766
struct EnergyDepthArgs {
    Grid g;
    const float* z; const float* z0lr; const unsigned char* lrmask;
    double* partials; unsigned* ticket; double* energy_out;     // energy_out[1]
    PeerComm comm;
};

__global__ void __launch_bounds__(EP_NT, 4) energy_depth_kernel(const EnergyDepthArgs a) {
    __shared__ double red[EP_NT / 32];
    const Grid& g = a.g;
    const int sf = g.sf;
    const float inv2 = 1.f / (float)(sf * sf);
    double acc = 0.0;
    const long long ncell = (long long)g.lny * g.lnx;
    const long long stride = (long long)gridDim.x * EP_NT;
    for (long long e = (long long)blockIdx.x * EP_NT + threadIdx.x; e < ncell; e += stride) {
        const int bl = (int)(e / g.lnx), bx = (int)(e - (long long)bl * g.lnx);
        if (!a.lrmask[(long long)bl * g.lpitch + bx]) continue;
        float s = 0.f;
        for (int l = 0; l < sf; l++)
            for (int k = 0; k < sf; k++) s += a.z[(long long)(bl * sf + l) * g.pitch + bx * sf + k];
        const float df = s * inv2 - a.z0lr[(long long)bl * g.lpitch + bx];
        acc += (double)(df * df);
    }
    double total;
    if (grid_reduce_last_world<EP_NT>(acc, a.partials, a.ticket, red, total, a.comm)) {
        if (threadIdx.x == 0) a.energy_out[1] = total;
    }
}

// Strip partition: copy the first / last owned line of up to 8 planes into the neighbours' ghost lines,
// then a world barrier (mailbox all-reduce of a dummy) so every rank knows its ghosts have arrived.
struct HaloPushArgs {
    const float* plane[8];         // origin-offset local planes
    HaloPeers dst[8];
    int nplanes, pitch, ny;
    double* partials; unsigned* ticket;
    PeerComm comm;
};

__global__ void __launch_bounds__(EP_NT) halo_push_kernel(const HaloPushArgs a) {
    __shared__ double red[EP_NT / 32];
    const int q = a.pitch / 4;
    for (int e = blockIdx.x * EP_NT + threadIdx.x; e < a.nplanes * 2 * q; e += gridDim.x * EP_NT) {
        const int pl = e / (2 * q), rem = e - pl * 2 * q, side = rem / q, i = rem - side * q;
        float* dst = side ? a.dst[pl].next_ghost : a.dst[pl].prev_ghost;
        if (!dst) continue;
        const float* src = a.plane[pl] + (side ? (long long)(a.ny - 1) * a.pitch : 0);
        st4(dst + 4 * i, ld4(src + 4 * i));
    }
    __threadfence_system();
    double total;
    grid_reduce_last_world<EP_NT>(0.0, a.partials, a.ticket, red, total, a.comm, true, true);
}

// masked vector -> dense plane (idx = dense offset of masked pixel p)
__global__ void scatter_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = src[i];
}
// 8-bit samples stay 8-bit in HBM (the /255 of Utilities.cpp:343 happens in registers, srps_stack.cuh: u8_to_unit)
__global__ void scatter_u8_kernel(const unsigned char* __restrict__ src, const int* __restrict__ idx,
                                  unsigned char* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = src[i];
}
__global__ void gather_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void fill_linear_kernel(float* __restrict__ dst, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}
__global__ void fill_masked_kernel(const int* __restrict__ idx, float* __restrict__ dst, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = v;                              // rho = 0.5   devicecalls.cu:133-139
}

}  // namespace srps
