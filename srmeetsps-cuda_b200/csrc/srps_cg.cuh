// Depth update: matrix-free operator  y = (Kt K + G^T M G) p  and the two fused CG kernels.
//
// Replaces, for the depth solve of the reference (SRmeetsPS-GPU/devicecalls.cu:636-786):
//   * the assembly of A (c*n*npix rows) by 6 SpGEMM + 6 SpGEAM + csr2csc, KtK and AtA by two
//     more SpGEMM (:668-736)                       -> nothing is assembled; M_p is rebuilt per
//                                                     pixel from w_c = (rho_c/dz)^2 (12 B/pixel)
//   * cusparseScsrmv inside the CG (:267)          -> stencil_kernel<MODE_ITER>
//   * cublasSscal/Saxpy/Sdot/Scopy (:251-274)      -> fused into stencil_kernel / cg_update_kernel
//   * the host-side loop control with 3 blocking dots per pass -> device-resident CgScalars
// CG recurrences are exactly those of devicecalls.cu:252-275.
#pragma once
#include "srps_common.cuh"

namespace srps {

constexpr int TX = 128;            // tile: pixels along the contiguous axis
constexpr int TY = 16;             // tile: lines
constexpr int CG_NT = 256;         // threads per CTA
constexpr int RQ = TX / 4 + 2;     // float4 per region line (tile + one float4 halo each side)
constexpr int RL = TY + 2;         // region lines (tile + one halo line each side)
constexpr int SP = TX + 16;        // smem pitch of P/T: region col rc (0..TX+7) stored at rc+4
constexpr int Q1P = TX + 8;        // smem pitch of Q1: interior col ic (-1..TX) stored at ic+4

enum { MODE_ITER = 0, MODE_INIT = 1, MODE_APPLY = 2 };

struct StencilArgs {
    Grid g;
    const unsigned char* types;   // dense type map (origin-offset pointer)
    const float* w0; const float* w1; const float* w2;   // (rho_c/dz)^2 planes
    const LightConsts* lc;
    // MODE_ITER: p <- r + beta p ; y <- A p ; dot <- p.y
    // MODE_INIT: y(=r) <- Kt(z0s - K z) + G^T (g - M G z) with vin = z ; dot <- r.r
    // MODE_APPLY: y <- A vin (test hook)
    const float* vin;             // z (INIT) / p (APPLY)
    const float* r;               // ITER: residual (read)
    const float* p_in;            // ITER: previous search direction (read, tile + halo)
    float* p_out;                 // ITER: new search direction (written to the OTHER plane: neighbouring
                                  //       tiles recompute p on their halo from p_in, so it must stay intact)
    float* y;                     // output
    const float* g0; const float* g1; const float* g2;   // INIT: G^T g right-hand side planes
    const float* z0lr;            // INIT: dense LR depth
    CgScalars* sc;
    double* partials;
    unsigned* ticket;
    int tiles_x, tiles_y;
};

struct StencilSmem {
    float P[RL][SP];
    float Q0[RL][TX];
    float Q1[TY][Q1P];
    float BS[(TX * TY) / 4];
    unsigned char T[RL][SP];
    double red[CG_NT / 32];
};

template <int MODE>
__global__ void __launch_bounds__(CG_NT, 3) stencil_kernel(const StencilArgs a) {
    __shared__ StencilSmem sm;
    const Grid& g = a.g;
    const int tid = threadIdx.x;
    float beta = 0.f;
    if (MODE == MODE_ITER) {
        if (!a.sc->active) return;
        beta = a.sc->beta;
    }
    const LightConsts lc = *a.lc;
    const int pitch = g.pitch, ny = g.ny, sf = g.sf;
    const float inv2 = 1.f / (float)(sf * sf);
    const float inv4 = inv2 * inv2;
    const int ntiles = a.tiles_x * a.tiles_y;
    double dot = 0.0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int ty_i = tile / a.tiles_x, tx_i = tile - ty_i * a.tiles_x;
        const int x0 = tx_i * TX, y0 = ty_i * TY;

        // ---- phase A: stage the tile + halo of the operand (and the type map) in shared memory
        for (int idx = tid; idx < RL * RQ; idx += CG_NT) {
            const int ly = idx / RQ, q = idx - ly * RQ;
            const int j = y0 - 1 + ly, x = x0 - 4 + 4 * q;
            const bool ok = (j <= ny) && (x < pitch);
            const long long off = (long long)j * pitch + x;
            float4 v = f4zero();
            uchar4 t = make_uchar4(0, 0, 0, 0);
            if (ok) {
                t = *reinterpret_cast<const uchar4*>(a.types + off);
                if (MODE == MODE_ITER) {
                    const float4 r4 = ld4(a.r + off), p4 = ld4(a.p_in + off);
                    v.x = r4.x + beta * p4.x; v.y = r4.y + beta * p4.y;
                    v.z = r4.z + beta * p4.z; v.w = r4.w + beta * p4.w;
                    if (ly >= 1 && ly <= TY && q >= 1 && q <= TX / 4 && j < ny) st4(a.p_out + off, v);
                } else {
                    v = ld4(a.vin + off);
                }
            }
            *reinterpret_cast<float4*>(&sm.P[ly][4 * q + 4]) = v;
            *reinterpret_cast<uchar4*>(&sm.T[ly][4 * q + 4]) = t;
        }
        __syncthreads();

        // ---- phase B1: sf x sf block sums of the tile (the K part of Kt K)
        if (sf > 1) {
            const int nbx = TX / sf, nb = nbx * (TY / sf);
            for (int b = tid; b < nb; b += CG_NT) {
                const int by = b / nbx, bx = b - by * nbx;
                float s = 0.f;
                for (int l = 0; l < sf; l++)
                    for (int k = 0; k < sf; k++) s += sm.P[1 + by * sf + l][8 + bx * sf + k];
                sm.BS[b] = s;
            }
        }

        // ---- phase B2: q = M (G p) on the tile (q0,q1,q2), on the halo lines (q0) and halo columns (q1)
        float own[2][4], pc_keep[2][4];
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            const int it = tid + rep * CG_NT;
            const int ly = 1 + it / (TX / 4), q4 = it % (TX / 4);
            const int j = y0 + ly - 1, x = x0 + 4 * q4;
            const bool ok = (j < ny) && (x < pitch);
            const long long off = (long long)j * pitch + x;
            float4 w0 = f4zero(), w1 = f4zero(), w2 = f4zero(), gg0 = f4zero(), gg1 = f4zero(), gg2 = f4zero();
            if (ok) {
                w0 = ld4(a.w0 + off); w1 = ld4(a.w1 + off); w2 = ld4(a.w2 + off);
                if (MODE == MODE_INIT) { gg0 = ld4(a.g0 + off); gg1 = ld4(a.g1 + off); gg2 = ld4(a.g2 + off); }
            }
            const int sc0 = 8 + 4 * q4;
            const float4 c4 = *reinterpret_cast<const float4*>(&sm.P[ly][sc0]);
            const float4 u4 = *reinterpret_cast<const float4*>(&sm.P[ly - 1][sc0]);
            const float4 d4 = *reinterpret_cast<const float4*>(&sm.P[ly + 1][sc0]);
            const uchar4 t4 = *reinterpret_cast<const uchar4*>(&sm.T[ly][sc0]);
            const float lf = sm.P[ly][sc0 - 1], rt = sm.P[ly][sc0 + 4];
            const float pcv[6] = {lf, c4.x, c4.y, c4.z, c4.w, rt};
            const unsigned char tv[4] = {t4.x, t4.y, t4.z, t4.w};
            const float xx = (float)(g.jb0 + j) - g.cx;
            float4 q0v, q1v;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned char t = tv[k];
                const float pc = pcv[k + 1];
                const float dxp = (t & T_XF) ? f4get(d4, k) - pc : ((t & T_XB) ? pc - f4get(u4, k) : 0.f);
                const float dyp = (t & T_YF) ? pcv[k + 2] - pc : ((t & T_YB) ? pc - pcv[k] : 0.f);
                const float yy = (float)(g.ib0 + x + k) - g.cy;
                const Qm m = make_qm(lc, f4get(w0, k), f4get(w1, k), f4get(w2, k));
                float q0, q1, q2;
                apply_m(m, g.fx, g.fy, xx, yy, dxp, dyp, pc, q0, q1, q2);
                if (MODE == MODE_INIT) { q0 = f4get(gg0, k) - q0; q1 = f4get(gg1, k) - q1; q2 = f4get(gg2, k) - q2; }
                const float sgx = (t & T_XF) ? 1.f : ((t & T_XB) ? -1.f : 0.f);
                const float sgy = (t & T_YF) ? 1.f : ((t & T_YB) ? -1.f : 0.f);
                own[rep][k] = q2 - sgx * q0 - sgy * q1;
                pc_keep[rep][k] = pc;
                f4set(q0v, k, q0);
                f4set(q1v, k, q1);
            }
            *reinterpret_cast<float4*>(&sm.Q0[ly][4 * q4]) = q0v;
            *reinterpret_cast<float4*>(&sm.Q1[ly - 1][4 + 4 * q4]) = q1v;
        }
        if (tid < 2 * (TX / 4)) {
            // halo lines: only a forward row above (ly = 0) / a backward row below (ly = TY+1) reaches the tile
            const int which = tid / (TX / 4), q4 = tid % (TX / 4);
            const int ly = which ? TY + 1 : 0;
            const int j = y0 + ly - 1, x = x0 + 4 * q4;
            const bool ok = (j >= 0) && (j < ny) && (x < pitch);
            const long long off = (long long)j * pitch + x;
            float4 w0 = f4zero(), w1 = f4zero(), w2 = f4zero(), gg0 = f4zero();
            if (ok) {
                w0 = ld4(a.w0 + off); w1 = ld4(a.w1 + off); w2 = ld4(a.w2 + off);
                if (MODE == MODE_INIT) gg0 = ld4(a.g0 + off);
            }
            const int sc0 = 8 + 4 * q4;
            const float4 c4 = *reinterpret_cast<const float4*>(&sm.P[ly][sc0]);
            const float4 n4 = *reinterpret_cast<const float4*>(&sm.P[which ? TY : 1][sc0]);   // the tile line it couples to
            const uchar4 t4 = *reinterpret_cast<const uchar4*>(&sm.T[ly][sc0]);
            const float lf = sm.P[ly][sc0 - 1], rt = sm.P[ly][sc0 + 4];
            const float pcv[6] = {lf, c4.x, c4.y, c4.z, c4.w, rt};
            const unsigned char tv[4] = {t4.x, t4.y, t4.z, t4.w};
            const float xx = (float)(g.jb0 + j) - g.cx;
            float4 q0v;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned char t = tv[k];
                const float pc = pcv[k + 1];
                float dxp = 0.f;
                if (which == 0) { if (t & T_XF) dxp = f4get(n4, k) - pc; }
                else            { if (t & T_XB) dxp = pc - f4get(n4, k); }
                const float dyp = (t & T_YF) ? pcv[k + 2] - pc : ((t & T_YB) ? pc - pcv[k] : 0.f);
                const float yy = (float)(g.ib0 + x + k) - g.cy;
                const Qm m = make_qm(lc, f4get(w0, k), f4get(w1, k), f4get(w2, k));
                float q0, q1, q2;
                apply_m(m, g.fx, g.fy, xx, yy, dxp, dyp, pc, q0, q1, q2);
                if (MODE == MODE_INIT) q0 = f4get(gg0, k) - q0;
                f4set(q0v, k, q0);
            }
            *reinterpret_cast<float4*>(&sm.Q0[ly][4 * q4]) = q0v;
        } else if (tid < 2 * (TX / 4) + 2 * TY) {
            // halo columns: a forward pixel at ic = -1 / a backward pixel at ic = TX reaches the tile
            const int l = tid - 2 * (TX / 4);
            const int side = l / TY, ly = 1 + (l % TY);
            const int ic = side ? TX : -1;
            const int j = y0 + ly - 1, x = x0 + ic;
            const bool ok = (j < ny) && (x < pitch);      // x >= -1: the pad element of the previous line (zero)
            const long long off = (long long)j * pitch + x;
            float w0 = 0.f, w1 = 0.f, w2 = 0.f, gg1 = 0.f;
            if (ok) {
                w0 = a.w0[off]; w1 = a.w1[off]; w2 = a.w2[off];
                if (MODE == MODE_INIT) gg1 = a.g1[off];
            }
            const int sc = 8 + ic;
            const unsigned char t = sm.T[ly][sc];
            const float pc = sm.P[ly][sc];
            const float dxp = (t & T_XF) ? sm.P[ly + 1][sc] - pc : ((t & T_XB) ? pc - sm.P[ly - 1][sc] : 0.f);
            float dyp = 0.f;
            if (side == 0) { if (t & T_YF) dyp = sm.P[ly][sc + 1] - pc; }
            else           { if (t & T_YB) dyp = pc - sm.P[ly][sc - 1]; }
            const float xx = (float)(g.jb0 + j) - g.cx;
            const float yy = (float)(g.ib0 + x) - g.cy;
            const Qm m = make_qm(lc, w0, w1, w2);
            float q0, q1, q2;
            apply_m(m, g.fx, g.fy, xx, yy, dxp, dyp, pc, q0, q1, q2);
            if (MODE == MODE_INIT) q1 = gg1 - q1;
            sm.Q1[ly - 1][4 + ic] = q1;
        }
        __syncthreads();

        // ---- phase C: y = own + G^T gathers + Kt K term; fused dot product
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            const int it = tid + rep * CG_NT;
            const int ly = 1 + it / (TX / 4), q4 = it % (TX / 4);
            const int j = y0 + ly - 1, x = x0 + 4 * q4;
            const bool ok = (j < ny) && (x < pitch);
            const int sc0 = 8 + 4 * q4;
            const uchar4 t4 = *reinterpret_cast<const uchar4*>(&sm.T[ly][sc0]);
            const uchar4 tu4 = *reinterpret_cast<const uchar4*>(&sm.T[ly - 1][sc0]);
            const uchar4 td4 = *reinterpret_cast<const uchar4*>(&sm.T[ly + 1][sc0]);
            const unsigned char tl = sm.T[ly][sc0 - 1], tr = sm.T[ly][sc0 + 4];
            const float4 qu4 = *reinterpret_cast<const float4*>(&sm.Q0[ly - 1][4 * q4]);
            const float4 qd4 = *reinterpret_cast<const float4*>(&sm.Q0[ly + 1][4 * q4]);
            const float4 qc4 = *reinterpret_cast<const float4*>(&sm.Q1[ly - 1][4 + 4 * q4]);
            const float ql = sm.Q1[ly - 1][4 + 4 * q4 - 1], qr = sm.Q1[ly - 1][4 + 4 * q4 + 4];
            const unsigned char tv[6] = {tl, t4.x, t4.y, t4.z, t4.w, tr};
            const unsigned char tuv[4] = {tu4.x, tu4.y, tu4.z, tu4.w};
            const unsigned char tdv[4] = {td4.x, td4.y, td4.z, td4.w};
            const float q1v[6] = {ql, qc4.x, qc4.y, qc4.z, qc4.w, qr};
            float4 out = f4zero();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int ic = 4 * q4 + k;
                const unsigned char t = tv[k + 1];
                float yv = own[rep][k];
                if (tuv[k] & T_XF) yv += f4get(qu4, k);
                if (tdv[k] & T_XB) yv -= f4get(qd4, k);
                if (tv[k] & T_YF) yv += q1v[k];
                if (tv[k + 2] & T_YB) yv -= q1v[k + 2];
                if (t & T_LR) {
                    const float bs = (sf > 1) ? sm.BS[((ly - 1) / sf) * (TX / sf) + ic / sf] : pc_keep[rep][k];
                    if (MODE == MODE_INIT) {
                        const float z0 = a.z0lr[(long long)((y0 + ly - 1) / sf) * g.lpitch + (x0 + ic) / sf];
                        yv += (z0 - bs * inv2) * inv2;
                    } else {
                        yv += bs * inv4;
                    }
                }
                if (!(t & T_MASK)) yv = 0.f;
                f4set(out, k, yv);
                if (MODE == MODE_INIT) dot += (double)(yv * yv);
                else dot += (double)(pc_keep[rep][k] * yv);
            }
            if (ok) st4(a.y + (long long)j * pitch + x, out);
        }
        __syncthreads();
    }

    if (MODE == MODE_APPLY) return;
    double total;
    if (grid_reduce_last<CG_NT>(dot, a.partials, a.ticket, sm.red, total)) {
        if (threadIdx.x == 0) {
            CgScalars* s = a.sc;
            if (MODE == MODE_INIT) {                 // r1 = b.b ; k = 0          devicecalls.cu:242-252
                s->r1 = total; s->r0 = 0.0; s->k = 0; s->beta = 0.f; s->alpha = 0.f;
                s->active = ((float)total > s->tol2) && (0 <= s->max_iter);
            } else {                                 // alpha = r1 / (p.Ap)       devicecalls.cu:268-269
                s->dot = total;
                s->alpha = (float)s->r1 / (float)total;
            }
        }
    }
}

// x += alpha p ; r -= alpha y ; r1 = r.r ; beta = r1/r0 ; k++        devicecalls.cu:270-274,262
struct UpdateArgs {
    float* x; float* r; const float* p; const float* y;
    long long n4;              // float4 count of the interior (ny * pitch / 4)
    CgScalars* sc;
    double* partials;
    unsigned* ticket;
};

__global__ void __launch_bounds__(CG_NT, 4) cg_update_kernel(const UpdateArgs a) {
    __shared__ double red[CG_NT / 32];
    if (!a.sc->active) return;
    const float alpha = a.sc->alpha;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * CG_NT;
    for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < a.n4; i += stride) {
        const float4 p4 = ld4(a.p + 4 * i), y4 = ld4(a.y + 4 * i);
        float4 x4 = ld4(a.x + 4 * i), r4 = ld4(a.r + 4 * i);
        x4.x += alpha * p4.x; x4.y += alpha * p4.y; x4.z += alpha * p4.z; x4.w += alpha * p4.w;
        r4.x -= alpha * y4.x; r4.y -= alpha * y4.y; r4.z -= alpha * y4.z; r4.w -= alpha * y4.w;
        st4(a.x + 4 * i, x4);
        st4(a.r + 4 * i, r4);
        acc += (double)(r4.x * r4.x + r4.y * r4.y) + (double)(r4.z * r4.z + r4.w * r4.w);
    }
    double total;
    if (grid_reduce_last<CG_NT>(acc, a.partials, a.ticket, red, total)) {
        if (threadIdx.x == 0) {
            CgScalars* s = a.sc;
            s->r0 = s->r1;
            s->r1 = total;
            s->k += 1;
            s->beta = (float)total / (float)s->r0;                                  // devicecalls.cu:262
            s->active = ((float)total > s->tol2) && (s->k <= s->max_iter);          // devicecalls.cu:252
        }
    }
}

}  // namespace srps
