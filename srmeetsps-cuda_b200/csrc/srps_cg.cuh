// Depth update: matrix-free operator  y = (Kt K + G^T M G) p  and the CG drivers built on it.
//
// Replaces, for the depth solve of the reference (SRmeetsPS-GPU/devicecalls.cu:636-786):
//   * the assembly of A (c*n*npix rows) by 6 SpGEMM + 6 SpGEAM + csr2csc, KtK and AtA by two
//     more SpGEMM (:668-736)                       -> nothing is assembled; M_p is rebuilt per
//                                                     pixel from w_c = (rho_c/dz)^2 (12 B/pixel)
//   * cusparseScsrmv inside the CG (:267)          -> the stencil (strip_pass / stencil_kernel)
//   * cublasSscal/Saxpy/Sdot/Scopy (:251-274)      -> fused into the same kernels
//   * the host-side loop control with 3 blocking dots per pass -> device-resident CgScalars
// Four drivers run the reference's CG (devicecalls.cu:229-279: same alpha, beta, stop rule and pass count):
//   cg_fused_kernel             one kernel + one reduction per pass (default; sf <= 4)
//   cg_persistent_fused_kernel  the same passes in one cooperative launch (default below 3 M pixels on one GPU)
//   stencil_*_kernel<MODE_ITER> + cg_update_kernel   the textbook two-kernel pass (SRPS_CG=graph; sf 8, 16)
//   cg_persistent_kernel        the two-kernel pass in one cooperative launch (SRPS_CG=persistent)
#pragma once
#include "srps_comm.cuh"

namespace srps {

constexpr int TX = 128;            // tile: pixels along the contiguous axis
constexpr int TY = 16;             // tile: lines
constexpr int CG_NT = 256;         // threads per CTA
constexpr int RQ = TX / 4 + 2;     // float4 per region line (tile + one float4 halo each side)
constexpr int RL = TY + 2;         // region lines (tile + one halo line each side)
constexpr int SP = TX + 16;        // smem pitch of P/T: region col rc (0..TX+7) stored at rc+4
constexpr int Q1P = TX + 8;        // smem pitch of Q1: interior col ic (-1..TX) stored at ic+4

enum { MODE_ITER = 0, MODE_INIT = 1, MODE_APPLY = 2, MODE_FUSED = 3, MODE_FUSED0 = 4 };

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
    int strip_n, strip_chunks, strip_groups;   // warp-strip kernel: strips per line, chunks per strip, 4-line groups per strip (chunk_lines)
    PeerComm comm;                         // world == 1: single GPU
    // strip partition: the residual on the two ghost lines is PULLED from the neighbours' boundary lines (mapped peer
    // pointers; nullptr at the ends of the image).  The r.r all-reduce that ends every update kernel guarantees that the
    // neighbour finished writing r before this pass starts, so no halo push and no system-scope fence is needed.
    const float* r_prev_line;              // previous rank's line ny_prev-1 of r
    const float* r_next_line;              // next rank's line 0 of r
    // MODE_FUSED / MODE_FUSED0 (one kernel per CG pass, see cg_fused_kernel): the pending update of the previous pass
    //   r <- r - alpha y_in (written to r_out), z += alpha p_in, then p_out <- r + beta p_in, y <- A p_out
    const float* y_in;                     // previous pass's A p (read, window + halo; ghost lines pulled like r)
    float* r_out;                          // the other residual plane
    float* x;                              // z
    const float* y_prev_line;              // previous rank's line ny_prev-1 of y_in
    const float* y_next_line;              // next rank's line 0 of y_in
    int plane;                             // which ping-pong plane p_out is (recorded for cg_tail_kernel)
    GhostLL ll;                            // strip partition, fused pass: pushed ghost lines of r / y in LL format (srps_comm.cuh)
    int lc_slot;                           // warp-strip kernels: this context's slot of c_lc (constant-bank copy of *lc)
};

// Constant-bank copies of the per-iteration lighting constants, one slot per live context: the warp-strip kernels take
// the 18 S3 coefficients of make_qm as constant operands instead of holding them in 18 registers per thread (they run
// at the 168-register limit of 3 CTAs/SM).  Refreshed by a device-to-device copy whenever *lc changes (srps_api.cu).
constexpr int LC_SLOTS = 64;
__constant__ LightConsts c_lc[LC_SLOTS];

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
            // j == ny inside a partial tile: the zero guard line, or the ghost line of a strip partition whose
            // backward x-rows reach line ny-1 -> its coefficients are loaded, its output is discarded in phase C
            const bool ok = (j <= ny) && (x < pitch);
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
            const bool ok = (j <= ny) && (x < pitch);      // j = -1 / ny: guard lines (zero) or, in a strip partition, ghost lines
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
                if (!(t & T_MASK) || !ok) yv = 0.f;      // !ok: lines >= ny of a partial tile (a ghost line in a strip partition)
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
    if (grid_reduce_last_world<CG_NT>(dot, a.partials, a.ticket, sm.red, total, a.comm)) {
        if (threadIdx.x == 0) {
            CgScalars* s = a.sc;
            if (MODE == MODE_INIT) {                 // r1 = b.b ; k = 0          devicecalls.cu:242-252
                s->r1 = total; s->r0 = 0.0; s->k = 0; s->beta = 0.f; s->alpha = 0.f; s->defer = 0; s->n_defer = 0; s->n_zskip = 0;
                s->active = ((float)total > s->tol2) && (0 <= s->max_iter);
            } else {                                 // alpha = r1 / (p.Ap)       devicecalls.cu:268-269
                s->dot = total;
                s->alpha = (float)s->r1 / (float)total;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-strip variant of the operator (sf = 1, 2, 4; MODE_ITER / MODE_APPLY).
//
// One warp owns a strip of 30 float4 columns (+1 halo float4 column on each side, lanes 0 and 31)
// and marches down a chunk of lines in groups of 4, holding a 6-line register window of the search
// direction.  No shared memory, no block barriers: horizontal neighbours travel by warp shuffle,
// the line above hands its forward x-row contribution down in registers, backward rows (mask
// borders only) are evaluated under a warp vote.  Every global access is a full 512-byte warp
// request; the only redundancy is the two halo lanes (6.7 % more L2->SM traffic, no extra DRAM
// traffic) and one prologue line per chunk.
// ---------------------------------------------------------------------------------------------
#ifndef SRPS_STRIP_MINB
#define SRPS_STRIP_MINB 3        // CTAs per SM the register allocation aims at: 3 -> <= 168 registers, no spills.
#endif                           // Measured (round 1, 4096^2): 4 -> 128 registers with spills, 0.107 ms vs 0.092 ms; a register
                                 // software pipeline (next group's r,p prefetched, w one line ahead) was 0.097 ms: both rejected.
constexpr int SW_NT = 128;       // 4 independent warps per CTA
constexpr int SW_COLS = 30;      // output float4 columns per warp
constexpr int SW_G = 4;          // lines per group (multiple of sf)

#ifndef SRPS_ITER_CLC
#define SRPS_ITER_CLC 0          // two-kernel form: lighting constants from the constant bank (1) or 18 registers (0)
#endif
// Rejected (round 1, measured at 4096^2): staging the next group of lines in shared memory with per-lane cp.async copies
// (global -> shared while the current group is computed, 58 KB per CTA) instead of holding the loads in registers:
// two-kernel CG 18.2 ms against 15.5 ms, fused CG 16.8-21.3 ms against 14.5 ms -- 36 LDGSTS + 36 LDS per lane and group
// cost more than the latency they hide.
struct LineQ { float4 q0f, q0b, q1f, q1b, own; };

// Chunk c of the C = strip_chunks chunks of a strip covers the 4-line groups [c G / C, (c + 1) G / C) of the G = strip_groups
// groups of the grid: lengths differ by at most one group, so every resident warp gets an item and none a much longer
// one (a uniform chunk length rounded up to a group left 15 % of the warps idle at 1024 lines per GPU).
__device__ __forceinline__ void chunk_lines(const StencilArgs& a, int chunk, int ny, int& jA, int& jB) {
    jA = SW_G * (int)(((long long)chunk * a.strip_groups) / a.strip_chunks);
    jB = min(ny, SW_G * (int)(((long long)(chunk + 1) * a.strip_groups) / a.strip_chunks));
}

__device__ __forceinline__ unsigned ld_types(const unsigned char* t, long long off) {
    return *reinterpret_cast<const unsigned*>(t + off);
}

// q = M (G p) for the 4 pixels of one float4 of a line; returns the pieces the gather needs:
// q0f/q0b: x-row value where the pixel's x-row is forward / backward, q1f/q1b likewise for y,
// own = q2 - (own x-row) - (own y-row).
__device__ __forceinline__ LineQ line_q(const LightConsts& lc, float fx, float fy, float xx, float yy0, unsigned t4,
                                        const float4& pc, const float4& up, const float4& dn, float left, float right,
                                        const float4& w0, const float4& w1, const float4& w2) {
    LineQ o;
    const float pcv[6] = {left, pc.x, pc.y, pc.z, pc.w, right};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned t = (t4 >> (8 * k)) & 0xffu;
        const float c = pcv[k + 1];
        const float dxp = (t & T_XF) ? f4get(dn, k) - c : ((t & T_XB) ? c - f4get(up, k) : 0.f);
        const float dyp = (t & T_YF) ? pcv[k + 2] - c : ((t & T_YB) ? c - pcv[k] : 0.f);
        const Qm m = make_qm(lc, f4get(w0, k), f4get(w1, k), f4get(w2, k));
        float q0, q1, q2;
        apply_m(m, fx, fy, xx, yy0 + (float)k, dxp, dyp, c, q0, q1, q2);
        const float a0 = (t & T_XF) ? q0 : 0.f, b0 = (t & T_XB) ? q0 : 0.f;
        const float a1 = (t & T_YF) ? q1 : 0.f, b1 = (t & T_YB) ? q1 : 0.f;
        f4set(o.q0f, k, a0); f4set(o.q0b, k, b0); f4set(o.q1f, k, a1); f4set(o.q1b, k, b1);
        f4set(o.own, k, q2 - a0 + b0 - a1 + b1);
    }
    return o;
}

// One application of the operator over this block's share of the (strip, chunk) items.  `a.p_in` / `a.p_out`
// are the ping-pong planes of this pass; returns this thread's partial of p.y (MODE_ITER).
// COH: how planes that are rewritten between passes (r, y, p) are loaded -- 0: one launch per pass, read-only path;
// 1 / 2: inside a single-launch persistent CG (ld4_coh, srps_common.cuh).  types and w never change during a solve.
// LLG (fused pass of a strip partition): the boundary lines of r_out and y are PUSHED to the neighbours as LL words
// tagged tag_in + 1, and -- MODE_FUSED -- the ghost lines of r and y_in are read from this rank's LL buffer (tag_in)
// instead of being pulled from the neighbours' planes; MODE_FUSED0 (first pass of a solve) still pulls r, which the
// residual kernel wrote and ordered with its system-scope reduction.
// ZL (lazy z, persistent fused CG only): the depth is not touched in every pass.  A pass either skips z (`zskip`: no load,
// no store, the step stays pending) or adds `zc1 p_in + zc2 r_in`, which with zc2 != 0 applies TWO pending steps at once:
// the older one belongs to the direction before p_in, recovered as (p_in - r_in) / beta_prev from the two operands this
// pass loads anyway (p_in = r_in + beta_prev p_before was formed from exactly these stored values).  See
// cg_persistent_fused_kernel for the schedule; 44 -> 36 / 44 B per pixel in alternate passes.
template <int MODE, int SF, int COH = 0, bool LLG = false, int NT = SW_NT, bool ZL = false>
__device__ __forceinline__ double strip_pass(const StencilArgs& a, const LightConsts& lc, float beta, float alpha = 0.f,
                                             double* extra = nullptr /* FUSED: r.r, y_in.p, y.y */, unsigned tag_in = 0u,
                                             float zc1 = 0.f, float zc2 = 0.f, bool zskip = false) {
    static_assert(MODE == MODE_ITER || MODE == MODE_APPLY || MODE == MODE_FUSED || MODE == MODE_FUSED0,
                  "the warp-strip kernel implements ITER, APPLY and the fused pass");
    constexpr bool FUSED = (MODE == MODE_FUSED || MODE == MODE_FUSED0);
    constexpr bool KEEPS_P = (MODE == MODE_ITER) || FUSED;      // writes the new search direction
    static_assert(SF == 1 || SF == 2 || SF == 4, "sf must divide the group height");
    const Grid& g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = g.pitch, ny = g.ny;
    const float inv4 = 1.f / (float)(SF * SF * SF * SF);
    const unsigned FULL = 0xffffffffu;
    const int nitems = a.strip_n * a.strip_chunks;
    double dot = 0.0, s_rr = 0.0, s_yp = 0.0, s_yy = 0.0;

    for (int item = blockIdx.x * (NT / 32) + warp; item < nitems; item += gridDim.x * (NT / 32)) {
        const int strip = item % a.strip_n, chunk = item / a.strip_n;
        const int x = 4 * (strip * SW_COLS - 1 + lane);            // x = -4 on lane 0 of strip 0: the zero pad of the previous line
        const bool colok = x < pitch;
        const bool writer = (lane >= 1) && (lane <= SW_COLS) && colok;
        int jA, jB;
        chunk_lines(a, chunk, ny, jA, jB);
        const float yy0 = (float)(g.ib0 + x) - g.cy;

        // Loads are UNCONDITIONAL with a clamped address (the plane origin is always valid): no branch regions, so the
        // ~24 loads of a group are issued back to back.  What a clamped load returns is never zeroed: it can only belong to
        // (a) a lane beyond the line pitch -- a non-writing halo lane whose only consumer is the pad pixel of the last
        // valid float4 (type 0, forced to 0), or (b) a line beyond the guard/ghost line ny -- never an output line, never
        // the lower neighbour of one, and outside every sf x sf block of an output line (ny is a multiple of sf).
        // With one launch per pass (COH = 0) the operands are read-only for the whole kernel and use the non-coherent
        // path (the new p, r, y go to the other ping-pong planes).  Inside a single-launch persistent CG the same planes
        // are rewritten every pass, so r, y, p use coherent loads there (COH = 1; PTX defines .nc only for data that is
        // read-only for the whole kernel).  Strip partition: r / y on a ghost line (j = -1 / ny) are read in place from
        // the neighbour's boundary line over NVLink (pointer select); peer lines are cached in this SM's L1 only, which
        // is invalidated at a kernel boundary but not by the in-kernel barrier of the persistent form -> COH = 2 reads
        // r and y at system scope.
        auto load_pn = [&](int j) -> float4 {
            const bool ok = colok && j <= ny;                      // line ny is the zero guard (or ghost) line
            const long long off = ok ? (long long)j * pitch + x : 0;
            if (MODE == MODE_ITER) {
                const float* rs = a.r + off;
                rs = (ok && j < 0 && a.r_prev_line) ? a.r_prev_line + x : rs;
                rs = (ok && j == ny && a.r_next_line) ? a.r_next_line + x : rs;
                const float4 r4 = ld4_coh<COH>(rs), p4 = ld4_coh<(COH ? 1 : 0)>(a.p_in + off);      // p is never read from a peer
                return make_float4(r4.x + beta * p4.x, r4.y + beta * p4.y, r4.z + beta * p4.z, r4.w + beta * p4.w);
            }
            return ldg4(a.vin + off);
        };
        // fused pass: the previous pass's residual update is applied on the fly (window AND halo: the halo is recomputed
        // redundantly, like p), rn = r - alpha y_in ; returns p = rn + beta p_in ; pin = p_in (for the z update)
        auto load_fused = [&](int j, float4& rn, float4& pin, float& ypn) -> float4 {
            const bool ok = colok && j <= ny;
            const long long off = ok ? (long long)j * pitch + x : 0;
            const float* rs = a.r + off;
            if (MODE == MODE_FUSED0) {      // only the first pass of a solve pulls ghost lines out of the neighbours' planes
                rs = (ok && j < 0 && a.r_prev_line) ? a.r_prev_line + x : rs;
                rs = (ok && j == ny && a.r_next_line) ? a.r_next_line + x : rs;
            }
            const float4 r4 = ld4_coh<(COH ? 1 : 0)>(rs);
            if (MODE == MODE_FUSED0) { rn = r4; pin = f4zero(); ypn = 0.f; return r4; }       // first pass: p = r
            const float4 y4 = ld4_coh<(COH ? 1 : 0)>(a.y_in + off);
            pin = ld4_coh<(COH ? 1 : 0)>(a.p_in + off);
            rn = make_float4(r4.x - alpha * y4.x, r4.y - alpha * y4.y, r4.z - alpha * y4.z, r4.w - alpha * y4.w);
            const float4 pn = make_float4(rn.x + beta * pin.x, rn.y + beta * pin.y, rn.z + beta * pin.z, rn.w + beta * pin.w);
            ypn = (y4.x * pn.x + y4.y * pn.y) + (y4.z * pn.z + y4.w * pn.w);      // (A p_in).p : the conjugacy defect, see cg_fused_kernel
            if (ZL)     // `pin` leaves as the z increment of this pass (its only use): both pending steps, see the header
                pin = make_float4(zc1 * pin.x + zc2 * r4.x, zc1 * pin.y + zc2 * r4.y, zc1 * pin.z + zc2 * r4.z, zc1 * pin.w + zc2 * r4.w);
            return pn;
        };
        // LLG: line -1 (side 0) / ny (side 1) of r and y_in out of this rank's LL ghost buffer; p_in is local
        auto load_ghost = [&](int side, int j, float4& rn, float4& pin, float& ypn) -> float4 {
            const bool ok = colok && x >= 0;
            float4 r4 = f4zero(), y4 = f4zero();
            pin = ld4_coh<(COH ? 1 : 0)>(a.p_in + (ok ? (long long)j * pitch + x : 0));      // independent of the LL words: issued first
            if (ok) ll_load4x2(a.ll.in + a.ll.at(tag_in, side, 0), a.ll.in + a.ll.at(tag_in, side, 1), x, tag_in, r4, y4);
            rn = make_float4(r4.x - alpha * y4.x, r4.y - alpha * y4.y, r4.z - alpha * y4.z, r4.w - alpha * y4.w);
            const float4 pn = make_float4(rn.x + beta * pin.x, rn.y + beta * pin.y, rn.z + beta * pin.z, rn.w + beta * pin.w);
            ypn = (y4.x * pn.x + y4.y * pn.y) + (y4.z * pn.z + y4.w * pn.w);
            return pn;
        };
        // LLG: this pass's boundary lines of array `arr` (0 = r_out, 1 = y) to the neighbours, for their next pass.
        // Called only from the two places a boundary line can be (first line of chunk 0, line SW_G - 1 of the last group)
        // so that no test or store sits inside the batches of loads / stores of the other groups.
        auto push_first = [&](int arr, const float4& v) {       // this rank's line 0 = the previous rank's ghost line below
            if (writer && a.ll.out_prev) ll_store4(a.ll.out_prev + a.ll.at(tag_in + 1u, 1, arr), x, v, tag_in + 1u);
        };
        auto push_last = [&](int arr, const float4& v) {        // this rank's line ny-1 = the next rank's ghost line above
            if (writer && a.ll.out_next) ll_store4(a.ll.out_next + a.ll.at(tag_in + 1u, 0, arr), x, v, tag_in + 1u);
        };
        // z is read and written by this kernel (each float4 by its owner only): coherent load, clamped like the others
        auto load_x = [&](int j) -> float4 {
            const bool ok = colok && j <= ny;
            float4 v = f4zero();
            if (!(ZL && zskip)) v = ld4(a.x + (ok ? (long long)j * pitch + x : 0));      // uniform predicate, no branch region
            return v;
        };
        // owner's stores of a freshly loaded line: the new residual, the pending z step, and r.r
        auto store_owned = [&](int j, const float4& rn, const float4& xo, const float4& pin, float ypn) {
            if (writer && j < jB) {
                const long long off = (long long)j * pitch + x;
                st4(a.r_out + off, rn);
                if (MODE == MODE_FUSED && !ZL)
                    st4(a.x + off, make_float4(xo.x + alpha * pin.x, xo.y + alpha * pin.y, xo.z + alpha * pin.z, xo.w + alpha * pin.w));
                if (MODE == MODE_FUSED && ZL && !zskip)
                    st4(a.x + off, make_float4(xo.x + pin.x, xo.y + pin.y, xo.z + pin.z, xo.w + pin.w));
                s_rr += (double)((rn.x * rn.x + rn.y * rn.y) + (rn.z * rn.z + rn.w * rn.w));
                s_yp += (double)ypn;
            }
        };
        auto load_t = [&](int j) -> unsigned {
            const bool ok = colok && j <= ny;
            return __ldg(reinterpret_cast<const unsigned*>(a.types + (ok ? (long long)j * pitch + x : 0)));
        };
        auto load_w = [&](int j, float4& w0, float4& w1, float4& w2) {
            const bool ok = colok && j <= ny;
            const long long off = ok ? (long long)j * pitch + x : 0;
            w0 = ldg4(a.w0 + off); w1 = ldg4(a.w1 + off); w2 = ldg4(a.w2 + off);
        };

        float4 pl[SW_G + 1], wl0[SW_G + 1], wl1[SW_G + 1], wl2[SW_G + 1];
        unsigned tl[SW_G + 1];
        float4 pprev, q0f_prev;
        {   // prologue: the forward x-rows of line jA-1 reach line jA
            float4 w0, w1, w2;
            float4 pin0 = f4zero(), x0 = f4zero(), r0 = f4zero();
            float yp0 = 0.f;
            if (FUSED) {
                float4 rdum, pdum; float ydum;
                if (LLG && MODE == MODE_FUSED && chunk == 0 && a.comm.rank > 0) pprev = load_ghost(0, -1, rdum, pdum, ydum);
                else pprev = load_fused(jA - 1, rdum, pdum, ydum);
            } else pprev = load_pn(jA - 1);
            const unsigned tp = load_t(jA - 1);
            load_w(jA - 1, w0, w1, w2);
            if (FUSED) { pl[0] = load_fused(jA, r0, pin0, yp0); if (MODE == MODE_FUSED) x0 = load_x(jA); } else pl[0] = load_pn(jA);
            tl[0] = load_t(jA); load_w(jA, wl0[0], wl1[0], wl2[0]);
            if (FUSED) store_owned(jA, r0, x0, pin0, yp0);
            if (LLG && jA == 0) push_first(0, r0);
            const float left = __shfl_up_sync(FULL, pprev.w, 1), right = __shfl_down_sync(FULL, pprev.x, 1);
            const float xx = (float)(g.jb0 + jA - 1) - g.cx;
            const LineQ q = line_q(lc, g.fx, g.fy, xx, yy0, tp & 0xfbfbfbfbu /* backward x-rows not needed */, pprev, f4zero(),
                                   pl[0], left, right, w0, w1, w2);
            q0f_prev = q.q0f;
            // strip partition: keep the search direction on the ghost line above current (recomputed redundantly)
            if (KEEPS_P && a.comm.world > 1 && chunk == 0 && writer) st4(a.p_out - pitch + x, pprev);
        }
        for (int j0 = jA; j0 < jB; j0 += SW_G) {
            float4 pin[SW_G + 1], xo[SW_G + 1], rn[SW_G + 1];
            float ypn[SW_G + 1];
#pragma unroll
            for (int l = 1; l <= SW_G; l++) {
                if (FUSED) {
                    pl[l] = load_fused(j0 + l, rn[l], pin[l], ypn[l]);
                    xo[l] = (MODE == MODE_FUSED) ? load_x(j0 + l) : f4zero();
                } else {
                    pl[l] = load_pn(j0 + l);
                }
                tl[l] = load_t(j0 + l); load_w(j0 + l, wl0[l], wl1[l], wl2[l]);
            }
            if (LLG && MODE == MODE_FUSED && j0 + SW_G == ny && a.comm.rank + 1 < a.comm.world)
                // the ghost line below the strip comes from the LL buffer (the unconditional load above read the local guard
                // line): one warp-uniform branch AFTER the batch of loads, in the last group of the strip only
                pl[SW_G] = load_ghost(1, ny, rn[SW_G], pin[SW_G], ypn[SW_G]);
            if (FUSED) {                        // after the whole batch of loads: stores inside it would serialise them
#pragma unroll
                for (int l = 1; l <= SW_G; l++) store_owned(j0 + l, rn[l], xo[l], pin[l], ypn[l]);
                if (LLG && j0 + SW_G == ny) push_last(0, rn[SW_G - 1]);
            }
            if (KEEPS_P && a.comm.world > 1 && j0 + SW_G >= ny && writer) {      // ghost line below: keep p there too
#pragma unroll
                for (int l = 1; l <= SW_G; l++)
                    if (j0 + l == ny) st4(a.p_out + (long long)ny * pitch + x, pl[l]);
            }
            // sf x sf block sums of the group (the K of Kt K)
            float bs4 = 0.f, bs2[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
            if (SF == 4) {
#pragma unroll
                for (int l = 0; l < SW_G; l++) bs4 += (pl[l].x + pl[l].y) + (pl[l].z + pl[l].w);
            } else if (SF == 2) {
#pragma unroll
                for (int l = 0; l < SW_G; l++) { bs2[l / 2][0] += pl[l].x + pl[l].y; bs2[l / 2][1] += pl[l].z + pl[l].w; }
            }
#pragma unroll
            for (int l = 0; l < SW_G; l++) {
                const int j = j0 + l;
                const float4 pc = pl[l];
                const float4 up = (l == 0) ? pprev : pl[l > 0 ? l - 1 : 0];
                const float left = __shfl_up_sync(FULL, pc.w, 1), right = __shfl_down_sync(FULL, pc.x, 1);
                const float xx = (float)(g.jb0 + j) - g.cx;
                const LineQ q = line_q(lc, g.fx, g.fy, xx, yy0, tl[l], pc, up, pl[l + 1], left, right, wl0[l], wl1[l], wl2[l]);
                // backward x-rows of the line below (mask borders only): q0 of line j+1 where it is T_XB
                float4 q0b_dn = f4zero();
                const unsigned tn = tl[l + 1];
                if (__any_sync(FULL, (tn & 0x04040404u) != 0u)) {
                    const float4 pn = pl[l + 1];
                    const float ln = __shfl_up_sync(FULL, pn.w, 1), rn = __shfl_down_sync(FULL, pn.x, 1);
                    const LineQ qn = line_q(lc, g.fx, g.fy, xx + 1.f, yy0, tn & 0xfdfdfdfdu /* forward x-rows not needed */, pn, pc,
                                            f4zero(), ln, rn, wl0[l + 1], wl1[l + 1], wl2[l + 1]);
                    q0b_dn = qn.q0b;
                }
                const float q1f_left = __shfl_up_sync(FULL, q.q1f.w, 1), q1b_right = __shfl_down_sync(FULL, q.q1b.x, 1);
                const float q1fv[5] = {q1f_left, q.q1f.x, q.q1f.y, q.q1f.z, q.q1f.w};
                const float q1bv[5] = {q.q1b.x, q.q1b.y, q.q1b.z, q.q1b.w, q1b_right};
                float4 out;
                float dl = 0.f, dyy = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned t = (tl[l] >> (8 * k)) & 0xffu;
                    float yv = f4get(q.own, k) + f4get(q0f_prev, k) - f4get(q0b_dn, k) + q1fv[k] - q1bv[k + 1];
                    if (t & T_LR) {
                        const float bs = (SF == 4) ? bs4 : ((SF == 2) ? bs2[l / 2][k / 2] : f4get(pc, k));
                        yv += bs * inv4;
                    }
                    if (!(t & T_MASK)) yv = 0.f;
                    f4set(out, k, yv);
                    dl += f4get(pc, k) * yv;
                    if (FUSED) dyy += yv * yv;
                }
                if (writer && j < jB) {
                    const long long off = (long long)j * pitch + x;
                    st4(a.y + off, out);
                    if (LLG && l == 0 && j == 0) push_first(1, out);
                    if (LLG && l == SW_G - 1 && j == ny - 1) push_last(1, out);
                    if (KEEPS_P) { st4(a.p_out + off, pc); dot += (double)dl; }
                    if (FUSED) s_yy += (double)dyy;
                }
                q0f_prev = q.q0f;
            }
            pprev = pl[SW_G - 1];
            pl[0] = pl[SW_G]; tl[0] = tl[SW_G]; wl0[0] = wl0[SW_G]; wl1[0] = wl1[SW_G]; wl2[0] = wl2[SW_G];
        }
    }
    if (FUSED) { extra[0] = s_rr; extra[1] = s_yp; extra[2] = s_yy; }
    return dot;
}

template <int MODE, int SF>
__global__ void __launch_bounds__(SW_NT, SRPS_STRIP_MINB) stencil_strip_kernel(const StencilArgs a) {
    __shared__ double red[SW_NT / 32];
    float beta = 0.f;
    if (MODE == MODE_ITER) {
        if (!a.sc->active) return;
        beta = a.sc->beta;
    }
#if SRPS_ITER_CLC
    const LightConsts& lc = c_lc[a.lc_slot];
#else
    const LightConsts lc = *a.lc;
#endif
    const double dot = strip_pass<MODE, SF>(a, lc, beta);
    if (MODE == MODE_APPLY) return;
    double total;
    if (grid_reduce_last_world<SW_NT>(dot, a.partials, a.ticket, red, total, a.comm)) {
        if (threadIdx.x == 0) {                      // alpha = r1 / (p.Ap)       devicecalls.cu:268-269
            a.sc->dot = total;
            a.sc->alpha = (float)a.sc->r1 / (float)total;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The residual of the depth system in the warp-strip form (sf <= 4), once per outer iteration:
//     r = Kt (z0s - K z) + G^T (g - M G z) ,  r1 = r.r          devicecalls.cu:743-745,758 ; :242-252
// Same walk as strip_pass<MODE_APPLY> with z as the operand, plus the right-hand-side planes g0..2 of every line and the
// LR depth of every block.  Replaces stencil_kernel<MODE_INIT> (the shared-memory tile kernel, 0.44 of the HBM peak) on
// this path; 2 CTAs per SM (the three g planes of a 5-line window do not fit 168 registers), own chunk geometry.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ LineQ line_q_init(const LightConsts& lc, float fx, float fy, float xx, float yy0, unsigned t4,
                                             const float4& pc, const float4& up, const float4& dn, float left, float right,
                                             const float4& w0, const float4& w1, const float4& w2, const float4& g0, const float4& g1,
                                             const float4& g2) {
    LineQ o;
    const float pcv[6] = {left, pc.x, pc.y, pc.z, pc.w, right};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned t = (t4 >> (8 * k)) & 0xffu;
        const float c = pcv[k + 1];
        const float dxp = (t & T_XF) ? f4get(dn, k) - c : ((t & T_XB) ? c - f4get(up, k) : 0.f);
        const float dyp = (t & T_YF) ? pcv[k + 2] - c : ((t & T_YB) ? c - pcv[k] : 0.f);
        const Qm m = make_qm(lc, f4get(w0, k), f4get(w1, k), f4get(w2, k));
        float q0, q1, q2;
        apply_m(m, fx, fy, xx, yy0 + (float)k, dxp, dyp, c, q0, q1, q2);
        q0 = f4get(g0, k) - q0; q1 = f4get(g1, k) - q1; q2 = f4get(g2, k) - q2;          // g - M G z
        const float a0 = (t & T_XF) ? q0 : 0.f, b0 = (t & T_XB) ? q0 : 0.f;
        const float a1 = (t & T_YF) ? q1 : 0.f, b1 = (t & T_YB) ? q1 : 0.f;
        f4set(o.q0f, k, a0); f4set(o.q0b, k, b0); f4set(o.q1f, k, a1); f4set(o.q1b, k, b1);
        f4set(o.own, k, q2 - a0 + b0 - a1 + b1);
    }
    return o;
}

template <int SF>
__global__ void __launch_bounds__(SW_NT, 2) stencil_strip_init_kernel(const StencilArgs a) {
    static_assert(SF == 1 || SF == 2 || SF == 4, "sf must divide the group height");
    __shared__ double red[SW_NT / 32];
    const LightConsts& lc = c_lc[a.lc_slot];
    const Grid& g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = g.pitch, ny = g.ny;
    const float inv2 = 1.f / (float)(SF * SF);
    const unsigned FULL = 0xffffffffu;
    const int nitems = a.strip_n * a.strip_chunks;
    double dot = 0.0;
    for (int item = blockIdx.x * (SW_NT / 32) + warp; item < nitems; item += gridDim.x * (SW_NT / 32)) {
        const int strip = item % a.strip_n, chunk = item / a.strip_n;
        const int x = 4 * (strip * SW_COLS - 1 + lane);
        const bool colok = x < pitch;
        const bool writer = (lane >= 1) && (lane <= SW_COLS) && colok;
        int jA, jB;
        chunk_lines(a, chunk, ny, jA, jB);
        const float yy0 = (float)(g.ib0 + x) - g.cy;
        auto off_of = [&](int j) -> long long { return (colok && j <= ny) ? (long long)j * pitch + x : 0; };   // clamped, see strip_pass

        float4 pl[SW_G + 1], wl0[SW_G + 1], wl1[SW_G + 1], wl2[SW_G + 1], gl0[SW_G + 1], gl1[SW_G + 1], gl2[SW_G + 1];
        unsigned tl[SW_G + 1];
        float4 pprev, q0f_prev;
        {   // prologue: the forward x-rows of line jA-1 reach line jA
            const long long op = off_of(jA - 1), oa = off_of(jA);
            pprev = ldg4(a.vin + op);
            const unsigned tp = __ldg(reinterpret_cast<const unsigned*>(a.types + op));
            const float4 w0 = ldg4(a.w0 + op), w1 = ldg4(a.w1 + op), w2 = ldg4(a.w2 + op), gp0 = ldg4(a.g0 + op);
            pl[0] = ldg4(a.vin + oa);
            tl[0] = __ldg(reinterpret_cast<const unsigned*>(a.types + oa));
            wl0[0] = ldg4(a.w0 + oa); wl1[0] = ldg4(a.w1 + oa); wl2[0] = ldg4(a.w2 + oa);
            gl0[0] = ldg4(a.g0 + oa); gl1[0] = ldg4(a.g1 + oa); gl2[0] = ldg4(a.g2 + oa);
            const float left = __shfl_up_sync(FULL, pprev.w, 1), right = __shfl_down_sync(FULL, pprev.x, 1);
            const float xx = (float)(g.jb0 + jA - 1) - g.cx;
            const LineQ q = line_q_init(lc, g.fx, g.fy, xx, yy0, tp & 0xfbfbfbfbu, pprev, f4zero(), pl[0], left, right, w0, w1, w2, gp0,
                                        f4zero(), f4zero());
            q0f_prev = q.q0f;
        }
        for (int j0 = jA; j0 < jB; j0 += SW_G) {
#pragma unroll
            for (int l = 1; l <= SW_G; l++) {
                const long long o = off_of(j0 + l);
                pl[l] = ldg4(a.vin + o);
                tl[l] = __ldg(reinterpret_cast<const unsigned*>(a.types + o));
                wl0[l] = ldg4(a.w0 + o); wl1[l] = ldg4(a.w1 + o); wl2[l] = ldg4(a.w2 + o);
                gl0[l] = ldg4(a.g0 + o); gl1[l] = ldg4(a.g1 + o); gl2[l] = ldg4(a.g2 + o);
            }
            // LR depth of the blocks this lane's 4 pixels x 4 lines belong to (clamped like the planes: never used outside T_LR pixels)
            float z0v[SW_G][4];
#pragma unroll
            for (int l = 0; l < SW_G; l++) {
                const int jl = min(j0 + l, ny - 1) / SF;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int xc = min(max(x + k, 0), g.nx - 1) / SF;
                    z0v[l][k] = ((SF == 4 && (l > 0 || k > 0)) || (SF == 2 && ((l & 1) || (k & 1)))) ? 0.f : __ldg(a.z0lr + (long long)jl * g.lpitch + xc);
                }
            }
            float bs4 = 0.f, bs2[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
            if (SF == 4) {
#pragma unroll
                for (int l = 0; l < SW_G; l++) bs4 += (pl[l].x + pl[l].y) + (pl[l].z + pl[l].w);
            } else if (SF == 2) {
#pragma unroll
                for (int l = 0; l < SW_G; l++) { bs2[l / 2][0] += pl[l].x + pl[l].y; bs2[l / 2][1] += pl[l].z + pl[l].w; }
            }
#pragma unroll
            for (int l = 0; l < SW_G; l++) {
                const int j = j0 + l;
                const float4 pc = pl[l];
                const float4 up = (l == 0) ? pprev : pl[l > 0 ? l - 1 : 0];
                const float left = __shfl_up_sync(FULL, pc.w, 1), right = __shfl_down_sync(FULL, pc.x, 1);
                const float xx = (float)(g.jb0 + j) - g.cx;
                const LineQ q = line_q_init(lc, g.fx, g.fy, xx, yy0, tl[l], pc, up, pl[l + 1], left, right, wl0[l], wl1[l], wl2[l], gl0[l], gl1[l], gl2[l]);
                float4 q0b_dn = f4zero();
                const unsigned tn = tl[l + 1];
                if (__any_sync(FULL, (tn & 0x04040404u) != 0u)) {
                    const float4 pn = pl[l + 1];
                    const float ln = __shfl_up_sync(FULL, pn.w, 1), rn = __shfl_down_sync(FULL, pn.x, 1);
                    const LineQ qn = line_q_init(lc, g.fx, g.fy, xx + 1.f, yy0, tn & 0xfdfdfdfdu, pn, pc, f4zero(), ln, rn, wl0[l + 1], wl1[l + 1],
                                                 wl2[l + 1], gl0[l + 1], f4zero(), f4zero());
                    q0b_dn = qn.q0b;
                }
                const float q1f_left = __shfl_up_sync(FULL, q.q1f.w, 1), q1b_right = __shfl_down_sync(FULL, q.q1b.x, 1);
                const float q1fv[5] = {q1f_left, q.q1f.x, q.q1f.y, q.q1f.z, q.q1f.w};
                const float q1bv[5] = {q.q1b.x, q.q1b.y, q.q1b.z, q.q1b.w, q1b_right};
                float4 out;
                float dl = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned t = (tl[l] >> (8 * k)) & 0xffu;
                    float yv = f4get(q.own, k) + f4get(q0f_prev, k) - f4get(q0b_dn, k) + q1fv[k] - q1bv[k + 1];
                    if (t & T_LR) {
                        const float bs = (SF == 4) ? bs4 : ((SF == 2) ? bs2[l / 2][k / 2] : f4get(pc, k));
                        const float z0 = (SF == 4) ? z0v[0][0] : ((SF == 2) ? z0v[l & ~1][k & ~1] : z0v[l][k]);
                        yv += (z0 - bs * inv2) * inv2;
                    }
                    if (!(t & T_MASK)) yv = 0.f;
                    f4set(out, k, yv);
                    dl += yv * yv;
                }
                if (writer && j < jB) {
                    st4(a.y + (long long)j * pitch + x, out);
                    dot += (double)dl;
                }
                q0f_prev = q.q0f;
            }
            pprev = pl[SW_G - 1];
            pl[0] = pl[SW_G]; tl[0] = tl[SW_G]; wl0[0] = wl0[SW_G]; wl1[0] = wl1[SW_G]; wl2[0] = wl2[SW_G];
            gl0[0] = gl0[SW_G]; gl1[0] = gl1[SW_G]; gl2[0] = gl2[SW_G];
        }
    }
    double total;
    if (grid_reduce_last_world<SW_NT>(dot, a.partials, a.ticket, red, total, a.comm)) {
        if (threadIdx.x == 0) {                  // r1 = b.b ; k = 0          devicecalls.cu:242-252
            CgScalars* s = a.sc;
            s->r1 = total; s->r0 = 0.0; s->k = 0; s->beta = 0.f; s->alpha = 0.f; s->defer = 0; s->n_defer = 0; s->n_zskip = 0;
            s->active = ((float)total > s->tol2) && (0 <= s->max_iter);
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
    PeerComm comm;
};

#ifndef SRPS_UPD_UNROLL
#define SRPS_UPD_UNROLL 1
#endif
__global__ void __launch_bounds__(CG_NT, 4) cg_update_kernel(const UpdateArgs a) {
    __shared__ double red[CG_NT / 32];
    if (!a.sc->active) return;
    const float alpha = a.sc->alpha;
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * CG_NT;
    for (long long i0 = (long long)blockIdx.x * CG_NT + threadIdx.x; i0 < a.n4; i0 += stride * SRPS_UPD_UNROLL) {
        float4 p4[SRPS_UPD_UNROLL], y4[SRPS_UPD_UNROLL], x4[SRPS_UPD_UNROLL], r4[SRPS_UPD_UNROLL];
#pragma unroll
        for (int u = 0; u < SRPS_UPD_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < a.n4) { p4[u] = ld4(a.p + 4 * i); y4[u] = ld4(a.y + 4 * i); x4[u] = ld4(a.x + 4 * i); r4[u] = ld4(a.r + 4 * i); }
        }
#pragma unroll
        for (int u = 0; u < SRPS_UPD_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < a.n4) {
                x4[u].x += alpha * p4[u].x; x4[u].y += alpha * p4[u].y; x4[u].z += alpha * p4[u].z; x4[u].w += alpha * p4[u].w;
                r4[u].x -= alpha * y4[u].x; r4[u].y -= alpha * y4[u].y; r4[u].z -= alpha * y4[u].z; r4[u].w -= alpha * y4[u].w;
                st4(a.x + 4 * i, x4[u]);
                st4(a.r + 4 * i, r4[u]);
                acc += (double)(r4[u].x * r4[u].x + r4[u].y * r4[u].y) + (double)(r4[u].z * r4[u].z + r4[u].w * r4[u].w);
            }
        }
    }
    double total;
    if (grid_reduce_last_world<CG_NT>(acc, a.partials, a.ticket, red, total, a.comm)) {
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

// ---------------------------------------------------------------------------------------------
// Fused CG pass: ONE kernel and ONE reduction per pass (sf <= 4).
//
// The two-kernel form needs r.r of the updated residual before it can form beta, hence the separate update kernel
// and a second grid-wide (and, with a strip partition, cross-GPU) reduction per pass.  Here pass k applies the update
// of pass k-1 on the fly while it loads its operands,
//     r_k = r_{k-1} - alpha_{k-1} y_{k-1}        (window + halo, written to the other r plane)
//     z  += alpha_{k-1} p_{k-1}
//     p_k = r_k + beta_k p_{k-1} ,  y_k = A p_k
// and reduces four dots at once: S0 = r_k.r_k, S1 = p_k.y_k, S3 = y_k.y_k and C = y_{k-1}.p_k, from which
//     S2 = r_k.y_k = (p_k - beta_k p_{k-1}).A p_k = S1 - beta_k C        (A symmetric: p_{k-1}.A p_k = y_{k-1}.p_k)
// -- C is the conjugacy defect of consecutive directions (zero in exact arithmetic), measured where y_{k-1} and p_k are
// both in registers, so no residual window has to be carried to the point where y_k leaves the pipeline.  Then
//     alpha_k    = S0 / S1                                   devicecalls.cu:269  (r.r measured, as in the reference)
//     |r_{k+1}|^2 = S0 - 2 alpha_k S2 + alpha_k^2 S3          (= |r_k - alpha_k y_k|^2 expanded, fp64)
//     beta_{k+1} = |r_{k+1}|^2 / S0                          devicecalls.cu:262
// so only beta rests on the expanded norm, for one pass; the next pass measures r.r again (no drift).  The stop test
// r.r > tol^2 (devicecalls.cu:252) uses the MEASURED value one pass late: a pass that finds S0 <= tol^2 is void -- it has
// already applied the previous step to z, cancels its own, and leaves k where the reference's loop would.
// Guard: the expansion cancels when one step removes almost the whole residual (|r_{k+1}|^2 < 1e-6 r.r: early
// convergence, e.g. sf = 1 scenes where Kt K = I) -- its absolute error is ~1e-7 r.r, so beta would be noise, or
// negative.  Such a pass sets `defer`; the next pass slot then only applies the pending step (r, z; p and y are copied
// to the other planes so the ping-pong stays in phase), MEASURES r.r -- exactly the reference's r1 -- and forms beta
// from it; k does not advance, the solve gets two spare slots for it (fused_update_only).
// Algorithmic traffic: 44 B/pixel/pass (r, y, p, z, 3 w read; r, p, y, z written) against 52 for the two-kernel form.
// cg_tail_kernel applies the step that is still pending after the last pass.
// ---------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ bool grid_reduce_last4(const double (&v)[4], double* partials, unsigned* ticket, double* wsm /* [NT/32][4] */,
                                                  double* tot /* smem [4] */) {
    static_assert(NT == 128, "one warp per value in the final sum");
    __shared__ int s_last4;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double s = warp_sum(v[i]);
        if (lane == 0) wsm[wid * 4 + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double b = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) b += wsm[w * 4 + threadIdx.x];
        partials[(long long)blockIdx.x * 4 + threadIdx.x] = b;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        s_last4 = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last4) return false;
    __threadfence();
    double acc = lane_strided_sum(partials, (int)gridDim.x, 4, wid, lane);      // warp `wid` sums value `wid` over the blocks: fixed order
    acc = warp_sum(acc);
    if (lane == 0) tot[wid] = acc;
    if (threadIdx.x == 0) *ticket = 0u;
    __syncthreads();
    return true;
}

constexpr double FUSED_DEFER_REL = 1e-6;     // expanded |r_{k+1}|^2 below this fraction of r.r: measure instead
constexpr int FUSED_SPARE_PASSES = 2;        // pass slots a solve has for deferred (update-only) passes

// Deferred pass (see the header comment): r_out = r - alpha y ; z += alpha p ; p_out = p ; y_out = y ; returns this
// thread's share of |r_out|^2.  Element-wise over the owned lines; p is copied on the two ghost / guard lines as well
// (strip partition: the neighbours' p there is kept redundantly; r and y ghosts are pulled, not stored).
template <int COH, bool LLG>
__device__ __forceinline__ double fused_update_only(const StencilArgs& a, float alpha, unsigned tag_in, float zc1, float zc2 = 0.f) {
    const long long q = a.g.pitch / 4;
    const long long lo = -q, hi = (long long)(a.g.ny + 1) * q, own = (long long)a.g.ny * q;
    const long long stride = (long long)gridDim.x * blockDim.x;
    double s = 0.0;
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        const float4 p4 = ld4_coh<(COH ? 1 : 0)>(a.p_in + 4 * i);
        st4(a.p_out + 4 * i, p4);
        if (i >= 0 && i < own) {
            const float4 r4 = ld4_coh<(COH ? 1 : 0)>(a.r + 4 * i), y4 = ld4_coh<(COH ? 1 : 0)>(a.y_in + 4 * i);
            float4 x4 = ld4(a.x + 4 * i);
            const float4 rn = make_float4(r4.x - alpha * y4.x, r4.y - alpha * y4.y, r4.z - alpha * y4.z, r4.w - alpha * y4.w);
            // z += zc1 p + zc2 r: zc1 = alpha, zc2 = 0 unless an older step is pending as well (lazy z, see strip_pass)
            x4.x += zc1 * p4.x + zc2 * r4.x; x4.y += zc1 * p4.y + zc2 * r4.y; x4.z += zc1 * p4.z + zc2 * r4.z; x4.w += zc1 * p4.w + zc2 * r4.w;
            st4(a.r_out + 4 * i, rn);
            st4(a.y + 4 * i, y4);
            st4(a.x + 4 * i, x4);
            if (LLG) {               // the neighbours' next pass reads this rank's boundary lines of r_out / y as LL words
                const int xq = (int)(i % q) * 4;
                if (i < q && a.ll.out_prev) {
                    ll_store4(a.ll.out_prev + a.ll.at(tag_in + 1u, 1, 0), xq, rn, tag_in + 1u);
                    ll_store4(a.ll.out_prev + a.ll.at(tag_in + 1u, 1, 1), xq, y4, tag_in + 1u);
                }
                if (i >= own - q && a.ll.out_next) {
                    ll_store4(a.ll.out_next + a.ll.at(tag_in + 1u, 0, 0), xq, rn, tag_in + 1u);
                    ll_store4(a.ll.out_next + a.ll.at(tag_in + 1u, 0, 1), xq, y4, tag_in + 1u);
                }
            }
            s += (double)((rn.x * rn.x + rn.y * rn.y) + (rn.z * rn.z + rn.w * rn.w));
        }
    }
    return s;
}

// alpha, beta, k, active, defer of the next pass from the four world totals of this one (one thread per rank; identical
// bits on every rank).  tot = {r.r, p.y, y_prev.p, y.y}; `beta`, `deferred`: what this pass ran with.
__device__ __forceinline__ void fused_pass_scalars(const StencilArgs& a, const double* tot, float beta, bool deferred) {
    CgScalars* s = a.sc;
    const double S0 = tot[0], S1 = tot[1], S3 = tot[3];
    const double S2 = S1 - (double)beta * tot[2];          // r.y, see the header comment
    if (s->profile) {                        // srps_profile_kernels: scalars stay as the host set them
        s->k += 1;
    } else if (!((float)S0 > s->tol2)) {     // the reference left its loop before this pass (devicecalls.cu:252)
        s->r1 = S0;
        s->alpha = 0.f;                      // nothing pending: the previous step went into z above
        s->active = 0;
        s->defer = 0;
    } else if (deferred) {                   // S0 is the measured r1 of the reference's loop
        s->beta = (float)S0 / (float)s->r0;                                    // devicecalls.cu:262
        s->r1 = S0;
        s->alpha = 0.f;                      // applied above
        s->defer = 0;
        s->n_defer += 1;
    } else {
        const float al = (float)S0 / (float)S1;                                // devicecalls.cu:269
        const double rr = S0 - 2.0 * (double)al * S2 + (double)al * (double)al * S3;
        const bool cancelled = !(rr > FUSED_DEFER_REL * S0);                   // also catches a non-finite rr
        s->dot = S1;
        s->alpha = al;
        s->r0 = S0;
        s->r1 = rr;
        s->beta = cancelled ? 0.f : (float)rr / (float)S0;                     // devicecalls.cu:262
        s->defer = cancelled ? 1 : 0;
        s->k += 1;
        s->plane = a.plane;
        s->active = (s->k <= s->max_iter);                                     // devicecalls.cu:252 (k part)
    }
}

#ifndef SRPS_FUSED_MINB
#define SRPS_FUSED_MINB 3
#endif
template <int SF, bool FIRST, bool LLG>
__global__ void __launch_bounds__(SW_NT, SRPS_FUSED_MINB) cg_fused_kernel(const StencilArgs a) {
    __shared__ double wsm[(SW_NT / 32) * 4];
    __shared__ double tot[4];
    if (!a.sc->active) return;
    // LL ghost tag: the reduction sequence number this pass starts with (rewritten only by this kernel's last block,
    // after every block has read it: a block takes its reduction ticket after its own work)
    const unsigned tag_in = LLG ? (unsigned)__ldcg(a.comm.seq) : 0u;
    const float beta = FIRST ? 0.f : a.sc->beta;
    const float alpha = FIRST ? 0.f : a.sc->alpha;          // the step of the previous pass, still pending
    const bool deferred = !FIRST && a.sc->defer != 0;       // uniform over the grid: written by the previous launch
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (deferred) {
        v[0] = fused_update_only<0, LLG>(a, alpha, tag_in, alpha);
    } else {
        const LightConsts& lc = c_lc[a.lc_slot];
        double ex[3];
        v[1] = strip_pass<FIRST ? MODE_FUSED0 : MODE_FUSED, SF, 0, LLG>(a, lc, beta, alpha, ex, tag_in);
        v[0] = ex[0]; v[2] = ex[1]; v[3] = ex[2];
    }
    if (!grid_reduce_last4<SW_NT>(v, a.partials, a.ticket, wsm, tot)) return;
    // strip partition: the ghost lines travel as self-validating LL words (pushed above), so the reduction is a pure
    // scalar exchange -- no system-scope fence on the critical path of a pass
    peer_allreduce_small<SW_NT, 4>(a.comm, tot, false);
    if (threadIdx.x == 0) fused_pass_scalars(a, tot, beta, deferred);
}

// z += alpha p of the last valid pass (alpha == 0: the last pass was void or no pass ran)
struct TailArgs { float* x; const float* p[2]; long long n4; const CgScalars* sc; };
__global__ void __launch_bounds__(CG_NT, 4) cg_tail_kernel(const TailArgs a) {
    const float alpha = a.sc->alpha;
    if (alpha == 0.f) return;
    const float* p = (a.sc->plane & 1) ? a.p[1] : a.p[0];
    const long long stride = (long long)gridDim.x * CG_NT;
    for (long long i = (long long)blockIdx.x * CG_NT + threadIdx.x; i < a.n4; i += stride) {
        const float4 p4 = ld4(p + 4 * i);
        float4 x4 = ld4(a.x + 4 * i);
        x4.x += alpha * p4.x; x4.y += alpha * p4.y; x4.z += alpha * p4.z; x4.w += alpha * p4.w;
        st4(a.x + 4 * i, x4);
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent CG: ALL passes of one depth solve in ONE cooperative launch (sf <= 4).
//
// The two-kernel form pays, per pass, two launches and two "last block" reductions (~10-14 us): that is the
// whole cost of a pass on small scenes (Mitten: 148 600 pixels) and the scaling limit of the strip partition.
// Here every resident CTA runs  operator -> grid barrier -> update -> grid barrier  101 times.  The barrier is a
// monotonic counter in global memory; each block publishes its fp64 partial before arriving and, after the
// barrier, EVERY block sums all partials in the same fixed order, so all blocks hold bit-identical alpha / beta /
// active without a broadcast.  With a strip partition, block 0 additionally all-reduces the rank total with the
// peers (peer_allreduce_scalar) and publishes the world total to the other blocks through a generation-tagged slot.
// The CG recurrences are those of devicecalls.cu:252-275, as in the two-kernel form.
// ---------------------------------------------------------------------------------------------
struct PersistentArgs {
    StencilArgs st;                 // operator operands (p_in / p_out are set per pass)
    float* pp[2];                   // the two ping-pong planes of the search direction: pass k reads pp[k&1], writes pp[(k+1)&1]
    float* rr[2];                   // fused form: residual planes, same ping-pong
    float* yy[2];                   // fused form: A p planes, same ping-pong
    float* x;                       // z
    float* r;                       // residual (read + written)
    long long n4;
    int passes;                     // max_iter + 1
    int zlazy;                      // fused form: skip the depth in alternate passes (strip_pass, ZL)
    unsigned long long* bar;        // grid barrier counter, zero on entry
    double* part[2];                // per-block partials: the two reductions of a pass (gridDim.x doubles each) / fused form:
                                    // the four dots of a pass, buffers alternating between passes (4 * gridDim.x doubles each)
    double* world_tot;              // two-barrier form: [4] world totals published by block 0, slot gen & 3;
                                    // fused form: eight broadcast words {sequence tag | half of a world total} (strip partition)
    unsigned long long* world_gen;  // two-barrier form: generation of the last published world total
    // strip partition, fused form: the neighbours' boundary lines of the two r and the two y planes (read in place)
    const float* r_prev[2]; const float* r_next[2];
    const float* y_prev[2]; const float* y_next[2];
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// publish `v`, wait for every block, return the (world) total -- identical bits in every thread of every block
__device__ __forceinline__ double grid_allreduce(const PersistentArgs& a, double v, int which, unsigned long long& gen,
                                                 double* red_smem, bool peer_stores) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ double s_total;
    v = warp_sum(v);
    if (lane == 0) red_smem[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int i = 0; i < SW_NT / 32; i++) b += red_smem[i];
        a.part[which][blockIdx.x] = b;
        if (peer_stores) __threadfence_system(); else __threadfence();
        atomicAdd(a.bar, 1ull);
        const unsigned long long target = (gen + 1ull) * gridDim.x;
        while (ld_acquire_gpu(a.bar) < target) { poll_backoff(); }
    }
    __syncthreads();
    gen += 1ull;
    if (wid == 0) {       // fixed summation order: lane-strided partials, then the xor tree
        double t = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) t += __ldcg(a.part[which] + i);
        t = warp_sum(t);
        if (lane == 0) s_total = t;
    }
    __syncthreads();
    double total = s_total;
    if (a.st.comm.world > 1) {
        // one block talks to the peers; the others pick the world total up from a generation-tagged slot
        if (blockIdx.x == 0) {
            total = peer_allreduce_scalar<SW_NT>(a.st.comm, total, false);
            if (threadIdx.x == 0) {
                a.world_tot[gen & 3ull] = total;
                __threadfence();
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(a.world_gen), "l"(gen) : "memory");
            }
        } else {
            if (threadIdx.x == 0) {
                while (ld_acquire_gpu(a.world_gen) < gen) { }
                s_total = __ldcg(a.world_tot + (gen & 3ull));
            }
            __syncthreads();
            total = s_total;
        }
    }
    __syncthreads();
    return total;
}

template <int SF>
__global__ void __launch_bounds__(SW_NT, SRPS_STRIP_MINB) cg_persistent_kernel(const PersistentArgs a) {
    __shared__ double red[SW_NT / 32];
    CgScalars* sc = a.st.sc;
    if (!sc->active) return;                         // r.r <= tol^2 already after the residual kernel (uniform)
    const LightConsts lc = *a.st.lc;
    double r1 = sc->r1, r0 = 0.0;
    const float tol2 = sc->tol2;
    const int max_iter = sc->max_iter;
    float beta = 0.f;
    int k = 0;
    unsigned long long gen = 0ull;
    StencilArgs st = a.st;
    const long long stride = (long long)gridDim.x * SW_NT;
    for (int pass = 0; pass < a.passes; pass++) {
        // ---- p <- r + beta p ; y <- A p ; p.y                         devicecalls.cu:256-268
        st.p_in = a.pp[pass & 1];
        st.p_out = a.pp[(pass + 1) & 1];
        const double dot = grid_allreduce(a, strip_pass<MODE_ITER, SF, 1>(st, lc, beta), 0, gen, red, false);
        const float alpha = (float)r1 / (float)dot;                      // devicecalls.cu:269
        // ---- x += alpha p ; r -= alpha y ; r.r                          devicecalls.cu:270-274
        double acc = 0.0;
        const float* pn = st.p_out;
        for (long long i = (long long)blockIdx.x * SW_NT + threadIdx.x; i < a.n4; i += stride) {
            const float4 p4 = ld4(pn + 4 * i), y4 = ld4(st.y + 4 * i);
            float4 x4 = ld4(a.x + 4 * i), r4 = ld4(a.r + 4 * i);
            x4.x += alpha * p4.x; x4.y += alpha * p4.y; x4.z += alpha * p4.z; x4.w += alpha * p4.w;
            r4.x -= alpha * y4.x; r4.y -= alpha * y4.y; r4.z -= alpha * y4.z; r4.w -= alpha * y4.w;
            st4(a.x + 4 * i, x4);
            st4(a.r + 4 * i, r4);
            acc += (double)(r4.x * r4.x + r4.y * r4.y) + (double)(r4.z * r4.z + r4.w * r4.w);
        }
        const double rr = grid_allreduce(a, acc, 1, gen, red, false);
        r0 = r1; r1 = rr; k++;
        beta = (float)r1 / (float)r0;                                    // devicecalls.cu:262
        if (!(((float)r1 > tol2) && (k <= max_iter))) break;             // devicecalls.cu:252
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->r1 = r1; sc->r0 = r0; sc->k = k; sc->beta = beta;
        sc->active = ((float)r1 > tol2) && (k <= max_iter);
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent CG, fused form: the passes of cg_fused_kernel inside ONE cooperative launch -- one grid barrier per
// pass (it carries the four dots) instead of the two of cg_persistent_kernel.  Every block derives alpha / beta /
// active from the same rank-ordered totals, so all blocks (and, with a strip partition, all ranks) take the same
// decisions; the step still pending when the loop ends is applied by all blocks after the last barrier.
// Measured (round 1, Mitten, 148 600 pixels): 1.11 ms per outer iteration against 1.35 ms for cg_persistent_kernel.
//
// Single GPU: every block waits for the arrival counter and sums all partials itself (fixed order, no broadcast hop).
// Strip partition (world > 1): the LAST block to arrive sums the partials, exchanges the four rank totals with the
// peers (peer_allreduce_small: one NVLink store per peer and word) and publishes the world totals under a generation
// word; everybody else spins on that word.  The ghost lines of r / y do not depend on this barrier: every pass pushes
// its boundary lines to the neighbours as self-validating LL words (GhostLL, srps_comm.cuh).
// ---------------------------------------------------------------------------------------------
template <bool WORLD, int NT>
__device__ __forceinline__ void grid_allreduce4(const PersistentArgs& a, double (&v)[4], int which, unsigned long long& gen,
                                                double* wsm /* [NT/32][4] */, double* s_tot /* [4] */, unsigned long long seq) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ int s_islast;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double s = warp_sum(v[i]);
        if (lane == 0) wsm[wid * 4 + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double b = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) b += wsm[w * 4 + threadIdx.x];
        a.part[which][(long long)blockIdx.x * 4 + threadIdx.x] = b;
    }
    __syncthreads();
    const unsigned long long target = (gen + 1ull) * gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long t = atomicAdd(a.bar, 1ull);
        if (WORLD) {
            s_islast = (t == target - 1ull);
        } else {
            while (ld_acquire_gpu(a.bar) < target) { poll_backoff(); }
        }
    }
    __syncthreads();
    gen += 1ull;
    if (!WORLD) {
        // every block sums all partials itself: warp `wid` (< 4) sums value `wid` over the blocks in the fixed order
        if (wid < 4) {
            const double t = warp_sum(lane_strided_sum(a.part[which], (int)gridDim.x, 4, wid, lane));
            if (lane == 0) s_tot[wid] = t;
        }
        __syncthreads();
    } else {
        // The last block to arrive sums the partials and stores this rank's four totals, as self-validating words
        // {sequence tag | half a double}, into the mailbox of EVERY rank, its own included (one NVLink store per peer
        // and word).  Every block of every rank then polls its own GPU's mailbox for the words of all ranks and forms the
        // rank-ordered sum itself: bit-identical totals everywhere, no second hop from an exchanging block to the others.
        const PeerComm& c = a.st.comm;
        const unsigned long long tagw = (seq & 0xffffffffull) << 32;
        const int slot = (int)(seq & (MB_SLOTS - 1));
        const int t = threadIdx.x >> 3, w = threadIdx.x & 7;          // (rank, word): 8 words = 4 doubles per rank
        const bool talker = t < c.world && threadIdx.x < MAX_RANKS * 8;
        __shared__ unsigned s_word[MAX_RANKS][8];
        if (s_islast) {
            __threadfence();
            if (wid < 4) {
                const double tt = warp_sum(lane_strided_sum(a.part[which], (int)gridDim.x, 4, wid, lane));
                if (lane == 0) s_tot[wid] = tt;
            }
            __syncthreads();
            if (talker) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(s_tot[w >> 1]);
                st_relaxed_sys_u64(&c.peer[t]->ll[slot][c.rank][w], tagw | ((w & 1) ? (bits >> 32) : (bits & 0xffffffffull)));
            }
            if (threadIdx.x == 0) *c.seq = seq;
        }
        if (talker) {
            unsigned long long x;
            unsigned spins = 0u;
            do {
                x = ld_relaxed_sys_u64(&c.local->ll[slot][t][w]);
                if ((x & 0xffffffff00000000ull) != tagw) poll_backoff();
                if (++spins > SPIN_LIMIT) __trap();
            } while ((x & 0xffffffff00000000ull) != tagw);
            s_word[t][w] = (unsigned)(x & 0xffffffffull);
            // acquire: the other blocks' r / y / p of this pass (fenced before their arrival, which the sending block
            // observed) must be what the next pass reads -- this also drops this SM's stale L1 lines
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double total = 0.0;
            for (int r = 0; r < c.world; r++)                                   // rank order: identical bits on every rank
                total += __longlong_as_double((long long)((unsigned long long)s_word[r][2 * threadIdx.x] |
                                                          ((unsigned long long)s_word[r][2 * threadIdx.x + 1] << 32)));
            s_tot[threadIdx.x] = total;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = s_tot[i];
    __syncthreads();
}

// MINB: resident CTAs per SM the register allocation aims at.  3 (168 registers) is the point the fused pass runs at;
// 4 (128 registers, 0.9 KB of spills per thread) puts a third more warps on an SM and shortens the chunk a warp walks,
// but measured (round 2, SRPS_PF_MINB=4) 3.22 against 2.18 ms per outer iteration at 1080p and 3.57 against 2.73 ms on
// a 2 M-pixel strip: rejected, kept as a switch for the record.
// NT: threads per CTA.  128 x 3 CTAs per SM is the geometry of the one-launch-per-pass kernels; 384 x 1 (SRPS_PF_NT=384) puts the
// same twelve warps into ONE CTA per SM: a third of the arrivals at the grid barrier and a third of the partials every block
// sums per pass.  Measured (round 2): slower -- 1080p 2.23 against 2.20 ms per outer iteration, a 512^2 scene 1.38 against
// 1.02 ms (twelve warps wait at every __syncthreads of the barrier, and a small scene occupies a third of the SMs): the
// barrier's cost is its chain of dependent L2 round trips (fence, arrival, poll, partial sums), not the number of arrivals.
// A pause between two polls of the barrier words (SRPS_POLL_NS) changes nothing either.  Both kept as switches for the record.
template <int SF, int COH, int MINB = SRPS_FUSED_MINB, int NT = SW_NT>
__global__ void __launch_bounds__(NT, MINB) cg_persistent_fused_kernel(const PersistentArgs a) {
    static_assert(NT / 32 >= 4, "one warp per dot in grid_allreduce4");
    static_assert(COH == 1 || COH == 2, "1: single GPU, 2: strip partition (planes are rewritten inside this launch: coherent loads)");
    constexpr bool WORLD = (COH == 2);
    __shared__ double wsm[(NT / 32) * 4];
    __shared__ double s_tot[4];
    CgScalars* sc = a.st.sc;
    if (!sc->active) return;                         // r.r <= tol^2 already after the residual kernel (uniform, also over the ranks)
    const LightConsts& lc = c_lc[a.st.lc_slot];
    const float tol2 = sc->tol2;
    const int max_iter = sc->max_iter;
    double r1 = sc->r1, r0 = 0.0;
    float alpha = 0.f, beta = 0.f;                   // alpha: the step of the previous pass, still pending
    // Lazy z: `alpha` belongs to the direction the next pass loads as p_in; `alpha_old` (if != 0) to the direction before it,
    // which that pass recovers as (p_in - r_in) / beta_link.  A pass with nothing older pending may skip z altogether (36
    // instead of 44 B per pixel); the next one -- a normal pass, a deferred slot, or the tail below -- then applies both
    // steps.  Not after a tiny beta (the recovery divides by it): such a pass applies its step at once, as does every pass
    // with a.zlazy == 0.
    constexpr float ZL_MIN_BETA = 1e-3f;
    float alpha_old = 0.f, beta_link = 1.f;
    float t_c1 = 0.f, t_c2 = 0.f;                    // tail: z += t_c1 pp[t_plane] + t_c2 rr[t_plane]
    int t_plane = 0, n_zskip = 0;
    bool tail_set = false;
    int k = 0, plane = 0;
    bool deferred = false;
    int n_defer = 0;
    unsigned long long gen = 0ull;
    StencilArgs st = a.st;
    st.x = a.x;
    // LL ghost tags: one all-reduce per executed pass, so pass number `pass` starts with sequence number base + pass
    const unsigned long long tag0 = WORLD ? __ldcg(a.st.comm.seq) : 0ull;
    for (int pass = 0; pass < a.passes + FUSED_SPARE_PASSES; pass++) {
        st.r = a.rr[pass & 1];   st.r_out = a.rr[(pass + 1) & 1];
        st.y_in = a.yy[pass & 1]; st.y = a.yy[(pass + 1) & 1];
        st.p_in = a.pp[pass & 1]; st.p_out = a.pp[(pass + 1) & 1];
        if (WORLD) {
            // pass 0 pulls the ghost lines of r out of the neighbours' planes (written and ordered by the residual kernel,
            // first touch of those addresses in this launch: nothing stale in L1); later passes read pushed LL words
            st.r_prev_line = pass == 0 ? a.r_prev[0] : nullptr; st.r_next_line = pass == 0 ? a.r_next[0] : nullptr;
            st.y_prev_line = nullptr; st.y_next_line = nullptr;
        }
        const unsigned tag_in = (unsigned)(tag0 + (unsigned long long)pass);
        // what this pass does to z
        float zc1 = alpha, zc2 = 0.f;
        bool zskip = false;
        if (alpha_old != 0.f) { zc2 = -alpha_old / beta_link; zc1 = alpha - zc2; }            // two steps at once
        else if (alpha == 0.f) zskip = true;                                                  // nothing pending (first pass, after a deferred slot)
        else if (a.zlazy && !deferred && beta >= ZL_MIN_BETA) zskip = true;                   // leave it to the next pass
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (deferred) {                              // see cg_fused_kernel: apply the step(s), measure r.r
            v[0] = fused_update_only<1, WORLD>(st, alpha, tag_in, zc1, zc2);
        } else {
            double ex[3];
            v[1] = (pass == 0) ? strip_pass<MODE_FUSED0, SF, 1, WORLD, NT>(st, lc, 0.f, 0.f, ex, tag_in)
                               : strip_pass<MODE_FUSED, SF, 1, WORLD, NT, true>(st, lc, beta, alpha, ex, tag_in, zc1, zc2, zskip);
            v[0] = ex[0]; v[2] = ex[1]; v[3] = ex[2];
        }
        grid_allreduce4<WORLD, NT>(a, v, pass & 1, gen, wsm, s_tot, (unsigned long long)tag0 + (unsigned long long)pass + 1ull);
        const double S0 = v[0], S1 = v[1], S3 = v[3];
        if (!((float)S0 > tol2)) {                   // void pass, see cg_fused_kernel
            r1 = S0;
            // a void pass cancels its own step; if it skipped z, the step it ran with is still owed (its p_in plane is intact)
            tail_set = true;
            t_plane = pass & 1;
            t_c1 = (zskip && !deferred) ? alpha : 0.f;
            t_c2 = 0.f;
            alpha = 0.f;
            break;
        }
        if (deferred) {
            beta = (float)S0 / (float)r0;                                         // devicecalls.cu:262, r.r measured
            r1 = S0;
            alpha = 0.f; alpha_old = 0.f;            // the slot applied everything that was pending
            deferred = false;
            n_defer++;
            continue;
        }
        const double S2 = S1 - (double)beta * v[2];
        const float al = (float)S0 / (float)S1;                                   // devicecalls.cu:269
        const double rr = S0 - 2.0 * (double)al * S2 + (double)al * (double)al * S3;
        deferred = !(rr > FUSED_DEFER_REL * S0);
        r0 = S0; r1 = rr;
        if (zskip && alpha != 0.f) { alpha_old = alpha; beta_link = beta; n_zskip++; }      // p_out = r_out + beta p_in links the two directions
        else alpha_old = 0.f;
        beta = deferred ? 0.f : (float)rr / (float)S0;                            // devicecalls.cu:262
        alpha = al;
        k++;
        plane = (pass + 1) & 1;
        if (!(k <= max_iter)) break;                                              // devicecalls.cu:252 (k part)
    }
    if (!tail_set) {                                 // the step(s) still pending after the last valid pass (the last barrier ordered p, r)
        t_plane = plane;
        t_c2 = (alpha_old != 0.f) ? -alpha_old / beta_link : 0.f;
        t_c1 = alpha - t_c2;
    }
    if (t_c1 != 0.f || t_c2 != 0.f) {
        const float* p = a.pp[t_plane];
        const float* r = a.rr[t_plane];
        const long long stride = (long long)gridDim.x * NT;
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < a.n4; i += stride) {
            const float4 p4 = __ldcg(reinterpret_cast<const float4*>(p + 4 * i));
            float4 x4 = ld4(a.x + 4 * i);
            if (t_c2 != 0.f) {
                const float4 r4 = __ldcg(reinterpret_cast<const float4*>(r + 4 * i));
                x4.x += t_c1 * p4.x + t_c2 * r4.x; x4.y += t_c1 * p4.y + t_c2 * r4.y; x4.z += t_c1 * p4.z + t_c2 * r4.z; x4.w += t_c1 * p4.w + t_c2 * r4.w;
            } else {
                x4.x += t_c1 * p4.x; x4.y += t_c1 * p4.y; x4.z += t_c1 * p4.z; x4.w += t_c1 * p4.w;
            }
            st4(a.x + 4 * i, x4);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->r1 = r1; sc->r0 = r0; sc->k = k; sc->beta = beta; sc->alpha = 0.f; sc->defer = 0; sc->n_defer = n_defer; sc->n_zskip = n_zskip;
        sc->active = 0;
    }
}

}  // namespace srps
