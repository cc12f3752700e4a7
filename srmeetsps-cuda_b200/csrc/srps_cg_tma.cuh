// Fused CG pass with a TMA-fed shared-memory ring (sm_100a: cp.async.bulk + mbarrier).
//
// cg_fused_kernel (srps_cg.cuh) holds the 28 x 128-bit loads of a 4-line group in registers: 168 registers, 12 warps
// per SM, and a warp's loads, arithmetic and stores run strictly one after the other -- ncu (round 1): issue slots 45 %
// busy, 3.2 long-scoreboard stalls per issue, 0.79 of the HBM copy peak.  Here every warp owns a two-stage ring in
// shared memory; the operands that the stencil reads with a halo (r, y, p, w0..2: 6 arrays x 4 lines x 512 B = 12 KB
// per stage) are fetched by BULK ASYNC COPIES, one 512-byte copy per (array, line) issued by lanes 0..23, completion
// counted by one mbarrier per stage.  The copies for group g+2 are issued as soon as group g has been consumed, so a
// warp always has 12-24 KB in flight while it computes, independent of the register file; z (read-modify-write by its
// owner only) and the stencil-type words are plain loads prefetched one group ahead.  Arithmetic, operation order and
// the reductions are those of strip_pass<MODE_FUSED> -- the two kernels are bit-identical (tests: SRPS_CG=fused_tma
// against fused); only where the operands wait differs.
// Shared memory: 4 warps x 2 stages x 12 KB = 96 KB per CTA, 2 CTAs per SM (8 warps, registers no longer bind).
#pragma once
#include "srps_cg.cuh"

namespace srps {

constexpr int TMA_ARR = 6;                                        // r, y_in, p_in, w0, w1, w2
constexpr int TMA_LINE_BYTES = 32 * 16;                           // one warp-wide line segment
constexpr int TMA_STAGE_BYTES = TMA_ARR * SW_G * TMA_LINE_BYTES;  // 12288
constexpr int TMA_STAGES = 2;
constexpr int TMA_WARP_BYTES = TMA_STAGES * TMA_STAGE_BYTES;      // 24576
constexpr int TMA_SMEM_BYTES = (SW_NT / 32) * TMA_WARP_BYTES + (SW_NT / 32) * TMA_STAGES * 8;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One fused pass (MODE_FUSED: not the first pass of a solve) over this block's share of the (strip, chunk) items.
// ring: this warp's TMA_WARP_BYTES of shared memory; bars: its TMA_STAGES mbarriers (initialised, phase tracked in
// `gcount`, the number of groups this warp has consumed so far).
template <int SF, bool LLG>
__device__ __forceinline__ double strip_pass_tma(const StencilArgs& a, const LightConsts& lc, float beta, float alpha, double* extra,
                                                 unsigned tag_in, float4* ring, unsigned long long* bars) {
    static_assert(SF == 1 || SF == 2 || SF == 4, "sf must divide the group height");
    const Grid& g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = g.pitch, ny = g.ny;
    const float inv4 = 1.f / (float)(SF * SF * SF * SF);
    const unsigned FULL = 0xffffffffu;
    const int nitems = a.strip_n * a.strip_chunks;
    double dot = 0.0, s_rr = 0.0, s_yp = 0.0, s_yy = 0.0;
    unsigned gcount = 0u;

    // shared-memory view: ring[((stage * TMA_ARR + arr) * SW_G + line) * 32 + lane]
    auto slot = [&](int stage, int arr, int line) -> const float4& { return ring[((stage * TMA_ARR + arr) * SW_G + line) * 32 + lane]; };

    for (int item = blockIdx.x * (SW_NT / 32) + warp; item < nitems; item += gridDim.x * (SW_NT / 32)) {
        const int strip = item % a.strip_n, chunk = item / a.strip_n;
        const int x0w = 4 * (strip * SW_COLS - 1);                   // float offset of lane 0 inside a line (-4 for strip 0: the pad of the line before)
        const int x = x0w + 4 * lane;
        const bool colok = x < pitch;
        const bool writer = (lane >= 1) && (lane <= SW_COLS) && colok;
        int jA, jB;
        chunk_lines(a, chunk, ny, jA, jB);
        const int ngroups = (jB - jA + SW_G - 1) / SW_G;      // ny is a multiple of sf only: the last group of the grid may be partial
        const float yy0 = (float)(g.ib0 + x) - g.cy;

        // lanes 0..23 copy one (array, line) segment of 512 bytes each; a segment may run past the line pitch into the next
        // line (always inside the plane: two guard lines follow the grid) -- those floats belong to lanes that never write
        const int c_arr = lane >> 2, c_line = lane & 3;
        const float* c_src = c_arr == 0 ? a.r : (c_arr == 1 ? a.y_in : (c_arr == 2 ? a.p_in : (c_arr == 3 ? a.w0 : (c_arr == 4 ? a.w1 : a.w2))));
        auto issue = [&](unsigned gc, int j_first) {                 // lines j_first .. j_first + 3 into stage gc & 1
            const int stage = (int)(gc & 1u);
            if (lane == 0) mbar_arrive_expect_tx(bars + stage, TMA_STAGE_BYTES);
            __syncwarp();
            if (lane < TMA_ARR * SW_G)
                // lines beyond the guard line ny (partial last group) are never an output line nor the neighbour of one:
                // copy the guard line in their place, which keeps every copy inside the plane
                bulk_g2s(ring + ((stage * TMA_ARR + c_arr) * SW_G + c_line) * 32, c_src + (long long)min(j_first + c_line, ny) * pitch + x0w,
                         TMA_LINE_BYTES, bars + stage);
        };
        issue(gcount, jA + 1);
        if (ngroups > 1) issue(gcount + 1u, jA + 1 + SW_G);

        // ---- direct loads of the item's head (lines jA-1 and jA), as in strip_pass
        auto load_head = [&](int j, float4& rn, float4& pin, float& ypn) -> float4 {
            const bool ok = colok && j <= ny;
            const long long off = ok ? (long long)j * pitch + x : 0;
            const float4 r4 = ldg4(a.r + off), y4 = ldg4(a.y_in + off);
            pin = ldg4(a.p_in + off);
            rn = make_float4(r4.x - alpha * y4.x, r4.y - alpha * y4.y, r4.z - alpha * y4.z, r4.w - alpha * y4.w);
            const float4 pn = make_float4(rn.x + beta * pin.x, rn.y + beta * pin.y, rn.z + beta * pin.z, rn.w + beta * pin.w);
            ypn = (y4.x * pn.x + y4.y * pn.y) + (y4.z * pn.z + y4.w * pn.w);
            return pn;
        };
        auto load_ghost = [&](int side, int j, float4& rn, float4& pin, float& ypn) -> float4 {
            const bool ok = colok && x >= 0;
            float4 r4 = f4zero(), y4 = f4zero();
            pin = ldg4(a.p_in + (ok ? (long long)j * pitch + x : 0));                          // independent of the LL words: issued first
            if (ok) ll_load4x2(a.ll.in + a.ll.at(tag_in, side, 0), a.ll.in + a.ll.at(tag_in, side, 1), x, tag_in, r4, y4);
            rn = make_float4(r4.x - alpha * y4.x, r4.y - alpha * y4.y, r4.z - alpha * y4.z, r4.w - alpha * y4.w);
            const float4 pn = make_float4(rn.x + beta * pin.x, rn.y + beta * pin.y, rn.z + beta * pin.z, rn.w + beta * pin.w);
            ypn = (y4.x * pn.x + y4.y * pn.y) + (y4.z * pn.z + y4.w * pn.w);
            return pn;
        };
        auto push_first = [&](int arr, const float4& v) {
            if (writer && a.ll.out_prev) ll_store4(a.ll.out_prev + a.ll.at(tag_in + 1u, 1, arr), x, v, tag_in + 1u);
        };
        auto push_last = [&](int arr, const float4& v) {
            if (writer && a.ll.out_next) ll_store4(a.ll.out_next + a.ll.at(tag_in + 1u, 0, arr), x, v, tag_in + 1u);
        };
        auto load_x = [&](int j) -> float4 { return ld4(a.x + ((colok && j <= ny) ? (long long)j * pitch + x : 0)); };
        auto load_t = [&](int j) -> unsigned {
            return __ldg(reinterpret_cast<const unsigned*>(a.types + ((colok && j <= ny) ? (long long)j * pitch + x : 0)));
        };
        auto store_owned = [&](int j, const float4& rn, const float4& xo, const float4& pin, float ypn) {
            if (writer && j < jB) {
                const long long off = (long long)j * pitch + x;
                st4(a.r_out + off, rn);
                st4(a.x + off, make_float4(xo.x + alpha * pin.x, xo.y + alpha * pin.y, xo.z + alpha * pin.z, xo.w + alpha * pin.w));
                s_rr += (double)((rn.x * rn.x + rn.y * rn.y) + (rn.z * rn.z + rn.w * rn.w));
                s_yp += (double)ypn;
            }
        };

        float4 pl[SW_G + 1], w0c, w1c, w2c;          // p window; w of window line 0 (carried between groups)
        unsigned tl[SW_G + 1];
        float4 pprev, q0f_prev;
        float4 xo_n[SW_G + 1];                        // z of the NEXT group's lines (index 1..SW_G), prefetched
        unsigned tl_n[SW_G + 1];
        {
            float4 rdum, pdum, r0, pin0;
            float ydum, yp0;
            if (LLG && chunk == 0 && a.comm.rank > 0) pprev = load_ghost(0, -1, rdum, pdum, ydum);
            else pprev = load_head(jA - 1, rdum, pdum, ydum);
            const unsigned tp = load_t(jA - 1);
            const bool okp = colok && jA - 1 <= ny;
            const long long offp = okp ? (long long)(jA - 1) * pitch + x : 0;
            const float4 wp0 = ldg4(a.w0 + offp), wp1 = ldg4(a.w1 + offp), wp2 = ldg4(a.w2 + offp);
            pl[0] = load_head(jA, r0, pin0, yp0);
            const float4 x0 = load_x(jA);
            tl[0] = load_t(jA);
            const long long offa = colok ? (long long)jA * pitch + x : 0;
            w0c = ldg4(a.w0 + offa); w1c = ldg4(a.w1 + offa); w2c = ldg4(a.w2 + offa);
#pragma unroll
            for (int l = 1; l <= SW_G; l++) { xo_n[l] = load_x(jA + l); tl_n[l] = load_t(jA + l); }
            store_owned(jA, r0, x0, pin0, yp0);
            if (LLG && jA == 0) push_first(0, r0);
            const float left = __shfl_up_sync(FULL, pprev.w, 1), right = __shfl_down_sync(FULL, pprev.x, 1);
            const float xx = (float)(g.jb0 + jA - 1) - g.cx;
            const LineQ q = line_q(lc, g.fx, g.fy, xx, yy0, tp & 0xfbfbfbfbu, pprev, f4zero(), pl[0], left, right, wp0, wp1, wp2);
            q0f_prev = q.q0f;
            if (a.comm.world > 1 && chunk == 0 && writer) st4(a.p_out - pitch + x, pprev);
        }

        for (int gi = 0; gi < ngroups; gi++, gcount++) {
            const int j0 = jA + gi * SW_G;
            const int stage = (int)(gcount & 1u);
            float4 xo[SW_G + 1];
#pragma unroll
            for (int l = 1; l <= SW_G; l++) { xo[l] = xo_n[l]; tl[l] = tl_n[l]; }
            if (gi + 1 < ngroups) {                       // z and types of the next group: in flight during this group's arithmetic
#pragma unroll
                for (int l = 1; l <= SW_G; l++) { xo_n[l] = load_x(j0 + SW_G + l); tl_n[l] = load_t(j0 + SW_G + l); }
            }
            mbar_wait(bars + stage, (gcount >> 1) & 1u);
            float4 rn[SW_G + 1], pin[SW_G + 1];
            float ypn[SW_G + 1];
#pragma unroll
            for (int l = 1; l <= SW_G; l++) {
                const float4 r4 = slot(stage, 0, l - 1), y4 = slot(stage, 1, l - 1);
                pin[l] = slot(stage, 2, l - 1);
                rn[l] = make_float4(r4.x - alpha * y4.x, r4.y - alpha * y4.y, r4.z - alpha * y4.z, r4.w - alpha * y4.w);
                pl[l] = make_float4(rn[l].x + beta * pin[l].x, rn[l].y + beta * pin[l].y, rn[l].z + beta * pin[l].z, rn[l].w + beta * pin[l].w);
                ypn[l] = (y4.x * pl[l].x + y4.y * pl[l].y) + (y4.z * pl[l].z + y4.w * pl[l].w);
            }
            if (LLG && j0 + SW_G == ny && a.comm.rank + 1 < a.comm.world)
                pl[SW_G] = load_ghost(1, ny, rn[SW_G], pin[SW_G], ypn[SW_G]);
#pragma unroll
            for (int l = 1; l <= SW_G; l++) store_owned(j0 + l, rn[l], xo[l], pin[l], ypn[l]);
            if (LLG && j0 + SW_G == ny) push_last(0, rn[SW_G - 1]);
            if (a.comm.world > 1 && j0 + SW_G >= ny && writer) {      // ghost line below: keep p there too
#pragma unroll
                for (int l = 1; l <= SW_G; l++)
                    if (j0 + l == ny) st4(a.p_out + (long long)ny * pitch + x, pl[l]);
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
                const float4 wa0 = (l == 0) ? w0c : slot(stage, 3, l > 0 ? l - 1 : 0);
                const float4 wa1 = (l == 0) ? w1c : slot(stage, 4, l > 0 ? l - 1 : 0);
                const float4 wa2 = (l == 0) ? w2c : slot(stage, 5, l > 0 ? l - 1 : 0);
                const float left = __shfl_up_sync(FULL, pc.w, 1), right = __shfl_down_sync(FULL, pc.x, 1);
                const float xx = (float)(g.jb0 + j) - g.cx;
                const LineQ q = line_q(lc, g.fx, g.fy, xx, yy0, tl[l], pc, up, pl[l + 1], left, right, wa0, wa1, wa2);
                float4 q0b_dn = f4zero();
                const unsigned tn = tl[l + 1];
                if (__any_sync(FULL, (tn & 0x04040404u) != 0u)) {
                    const float4 pn = pl[l + 1];
                    const float ln = __shfl_up_sync(FULL, pn.w, 1), rnn = __shfl_down_sync(FULL, pn.x, 1);
                    const LineQ qn = line_q(lc, g.fx, g.fy, xx + 1.f, yy0, tn & 0xfdfdfdfdu, pn, pc, f4zero(), ln, rnn,
                                            slot(stage, 3, l), slot(stage, 4, l), slot(stage, 5, l));
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
                    dyy += yv * yv;
                }
                if (writer && j < jB) {
                    const long long off = (long long)j * pitch + x;
                    st4(a.y + off, out);
                    if (LLG && l == 0 && j == 0) push_first(1, out);
                    if (LLG && l == SW_G - 1 && j == ny - 1) push_last(1, out);
                    st4(a.p_out + off, pc);
                    dot += (double)dl;
                    s_yy += (double)dyy;
                }
                q0f_prev = q.q0f;
            }
            pprev = pl[SW_G - 1];
            pl[0] = pl[SW_G]; tl[0] = tl[SW_G];
            w0c = slot(stage, 3, SW_G - 1); w1c = slot(stage, 4, SW_G - 1); w2c = slot(stage, 5, SW_G - 1);
            __syncwarp();                                  // every lane has consumed the stage (the values above are in registers)
            if (gi + 2 < ngroups) issue(gcount + 2u, j0 + 2 * SW_G + 1);
        }
    }
    extra[0] = s_rr; extra[1] = s_yp; extra[2] = s_yy;
    return dot;
}

#ifndef SRPS_TMA_MINB
#define SRPS_TMA_MINB 2
#endif
template <int SF, bool LLG>
__global__ void __launch_bounds__(SW_NT, SRPS_TMA_MINB) cg_fused_tma_kernel(const StencilArgs a) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ double wsm[(SW_NT / 32) * 4];
    __shared__ double tot[4];
    if (!a.sc->active) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* ring = reinterpret_cast<float4*>(tma_smem + (size_t)warp * TMA_WARP_BYTES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tma_smem + (size_t)(SW_NT / 32) * TMA_WARP_BYTES) + warp * TMA_STAGES;
    if (lane == 0) {
        mbar_init(bars + 0, 1u);
        mbar_init(bars + 1, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const float beta = a.sc->beta, alpha = a.sc->alpha;
    const bool deferred = a.sc->defer != 0;
    const unsigned tag_in = LLG ? (unsigned)__ldcg(a.comm.seq) : 0u;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (deferred) {
        v[0] = fused_update_only<0, LLG>(a, alpha, tag_in, alpha);
    } else {
        const LightConsts& lc = c_lc[a.lc_slot];
        double ex[3];
        v[1] = strip_pass_tma<SF, LLG>(a, lc, beta, alpha, ex, tag_in, ring, bars);
        v[0] = ex[0]; v[2] = ex[1]; v[3] = ex[2];
    }
    if (!grid_reduce_last4<SW_NT>(v, a.partials, a.ticket, wsm, tot)) return;
    peer_allreduce_small<SW_NT, 4>(a.comm, tot, false);
    if (threadIdx.x == 0) fused_pass_scalars(a, tot, beta, deferred);
}

}  // namespace srps
