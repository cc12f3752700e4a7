// Cross-GPU plumbing for the strip partition (one process per GPU, peers mapped with CUDA IPC).
//
// Nothing here exists in the reference (single GPU, SRPS.cu:88).  A scene is cut into strips of
// lines (image columns); the guard lines of the dense layout (srps_common.cuh) become ghost lines
// that hold the neighbour's boundary line.  Two exchange mechanisms, both inside our own kernels over
// NVLink peer memory (no NCCL call on the critical path, no host involvement):
//
//   * scalar / small-vector all-reduce by MAILBOX: the last block of a reduction writes the rank's
//     partial into slot [seq&3][rank] of EVERY peer's mailbox, publishes it with a system-scope
//     release store of the sequence number, then acquires the world flags of its own mailbox and
//     sums the partials in rank order -> bit-identical totals on all ranks, so every rank takes the
//     same CG step and the same `active` decision.
//   * halo push: boundary lines are stored straight into the neighbour's ghost line through the
//     mapped peer pointer by the kernel that produces them; the stores are ordered before the
//     mailbox release by a system-scope fence, so "flag seen" implies "halo arrived".
#pragma once
#include "srps_common.cuh"

namespace srps {

constexpr int MAX_RANKS = 8;
constexpr int MB_SLOTS = 4;
constexpr int MB_VALS = 800;      // doubles per (slot, source rank): lighting needs n*12 + 30 <= 798
constexpr int MB_SMALL = 4;       // doubles per (slot, source rank) on the latency-optimised path (fused CG pass: 4 dots)
constexpr unsigned SPIN_LIMIT = 1u << 26;   // polls of one word before a waiting thread traps (tens of seconds)
#ifndef SRPS_POLL_NS
#define SRPS_POLL_NS 0           // pause between two polls of a barrier word (0: spin): hundreds of blocks poll the same L2 lines
#endif
__device__ __forceinline__ void poll_backoff() {
#if SRPS_POLL_NS > 0
    __nanosleep(SRPS_POLL_NS);
#endif
}

struct Mailbox {
    unsigned long long flag[MB_SLOTS][MAX_RANKS];
    unsigned long long ll[MB_SLOTS][MAX_RANKS][2 * MB_SMALL];   // fast path: per fp64 value {seq32 | low word}, {seq32 | high word}
    double val[MB_SLOTS][MAX_RANKS][MB_VALS];
};

struct PeerComm {
    int rank, world;
    Mailbox* local;                      // this rank's mailbox (device memory, IPC-exported)
    Mailbox* peer[MAX_RANKS];            // every rank's mailbox as mapped here (peer[rank] == local)
    unsigned long long* seq;             // device-resident reduction counter (same sequence on every rank)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// All-reduce (sum, rank order) of n <= MB_VALS doubles held in `vals` (shared or global memory of the
// calling block).  Called by ALL threads of exactly one block per rank (the reduction's last block);
// on return vals[0..n) holds the world total.  Needs one __syncthreads-compatible block.
template <int NT>
__device__ __forceinline__ void peer_allreduce(const PeerComm& c, double* vals, int n) {
    if (c.world <= 1) { __syncthreads(); return; }
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) s_seq = __ldcg(c.seq) + 1ull;
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int slot = (int)(seq & (MB_SLOTS - 1));
    // 1. deposit this rank's partial in every rank's mailbox (its own included)
    for (int e = threadIdx.x; e < n * c.world; e += NT) {
        const int r = e / n, i = e - r * n;
        st_relaxed_sys(&c.peer[r]->val[slot][c.rank][i], vals[i]);
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish, 3. wait for the world
    if (threadIdx.x < c.world) {
        st_release_sys(&c.peer[threadIdx.x]->flag[slot][c.rank], seq);
        while (ld_acquire_sys(&c.local->flag[slot][threadIdx.x]) != seq) { }
    }
    __syncthreads();
    // 4. rank-ordered sum: identical bits on every rank
    for (int i = threadIdx.x; i < n; i += NT) {
        double t = 0.0;
        for (int r = 0; r < c.world; r++) t += ld_relaxed_sys(&c.local->val[slot][r][i]);
        vals[i] = t;
    }
    if (threadIdx.x == 0) *c.seq = seq;
    __syncthreads();
}

// Scalar all-reduce, latency-optimised (the two per CG pass): the fp64 partial travels as two 8-byte words that
// each carry the 32-bit sequence number next to 32 bits of payload, so data and flag arrive in ONE store (no
// fence + flag round trip); 8-byte stores are single transactions.  st.release / ld.acquire at system scope
// keep "word seen" => "the sender's earlier halo stores are visible".  Called by all threads of the last block;
// returns the rank-ordered world total to every thread.
// `release` = the ranks exchange PLANE data around this reduction (ghost lines pushed into, or pulled out of, a
// neighbour's planes): the words then form a system-scope release/acquire chain -- the writer blocks fenced their
// stores (device scope) before taking their reduction ticket, the last block observed every ticket, fences at SYSTEM
// scope (cumulative) and stores the word; the receiver polls the word and fences at system scope before anything
// that reads the planes.  Pure scalar exchanges (release = false) need neither fence.
template <int NT>
__device__ __forceinline__ double peer_allreduce_scalar(const PeerComm& c, double v, bool release) {
    if (c.world <= 1) return v;
    __shared__ double s_part[MAX_RANKS];
    __shared__ unsigned long long s_seq1;
    if (threadIdx.x == 0) { s_seq1 = __ldcg(c.seq) + 1ull; s_part[0] = v; }
    __syncthreads();
    const unsigned long long seq = s_seq1;
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const int slot = (int)(seq & (MB_SLOTS - 1));
    const double mine = s_part[0];
    __syncthreads();
    if (threadIdx.x < c.world) {
        const int t = threadIdx.x;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(mine);
        // `release`: halo lines were stored into peer memory by this kernel (by blocks that fenced them before taking
        // their reduction ticket): ONE system fence orders them before the words below.  The words themselves are
        // self-validating (sequence tag inside), so no second fence / flag round trip is needed.
        if (release) __threadfence_system();
        st_relaxed_sys_u64(&c.peer[t]->ll[slot][c.rank][0], tag | (bits & 0xffffffffull));
        st_relaxed_sys_u64(&c.peer[t]->ll[slot][c.rank][1], tag | (bits >> 32));
        unsigned long long w0, w1;
        do { w0 = ld_relaxed_sys_u64(&c.local->ll[slot][t][0]); } while ((w0 & 0xffffffff00000000ull) != tag);
        do { w1 = ld_relaxed_sys_u64(&c.local->ll[slot][t][1]); } while ((w1 & 0xffffffff00000000ull) != tag);
        s_part[t] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        // acquire side of the chain: "word of rank t seen" + this fence => rank t's plane stores that preceded its
        // release fence are visible to everything ordered after the barrier below (the next pass / the next kernel)
        if (release) __threadfence_system();
    }
    __syncthreads();
    double total = 0.0;
    for (int r = 0; r < c.world; r++) total += s_part[r];      // rank order: identical bits on every rank
    if (threadIdx.x == 0) *c.seq = seq;
    return total;
}

// The same for NV <= MB_SMALL values at once (the four dot products of a fused CG pass): thread (r, w) sends word w to
// rank r and waits for rank r's word w.  vals: shared memory of the calling block, world totals on return.
// seq_known != 0: the caller knows the sequence number of this reduction (persistent CG: start value + pass), which
// saves the dependent L2 read of the counter on the critical path of a pass.
template <int NT, int NV>
__device__ __forceinline__ void peer_allreduce_small(const PeerComm& c, double* vals, bool release, unsigned long long seq_known = 0ull) {
    static_assert(NV <= MB_SMALL && MAX_RANKS * 2 * NV <= NT, "one thread per (rank, word)");
    if (c.world <= 1) return;
    __shared__ double s_part[MAX_RANKS][NV];
    __shared__ unsigned s_lo[MAX_RANKS][NV];
    __shared__ unsigned long long s_seqn;
    if (seq_known == 0ull) {
        if (threadIdx.x == 0) s_seqn = __ldcg(c.seq) + 1ull;   // L2: the previous caller may have been another block
        __syncthreads();
    }
    const unsigned long long seq = seq_known ? seq_known : s_seqn;
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const int slot = (int)(seq & (MB_SLOTS - 1));
    const int t = threadIdx.x / (2 * NV), w = threadIdx.x % (2 * NV);
    const bool talker = t < c.world && threadIdx.x < MAX_RANKS * 2 * NV;
    unsigned got = 0u;
    if (talker) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[w >> 1]);
        if (release) __threadfence_system();
        st_relaxed_sys_u64(&c.peer[t]->ll[slot][c.rank][w], tag | ((w & 1) ? (bits >> 32) : (bits & 0xffffffffull)));
        unsigned long long v;
        unsigned spins = 0u;
        do {
            v = ld_relaxed_sys_u64(&c.local->ll[slot][t][w]);
            if (++spins > SPIN_LIMIT) __trap();
        } while ((v & 0xffffffff00000000ull) != tag);
        got = (unsigned)(v & 0xffffffffull);
        if (!(w & 1)) s_lo[t][w >> 1] = got;
        if (release) __threadfence_system();      // acquire side, see peer_allreduce_scalar
    }
    __syncthreads();
    if (talker && (w & 1))
        s_part[t][w >> 1] = __longlong_as_double((long long)((unsigned long long)s_lo[t][w >> 1] | ((unsigned long long)got << 32)));
    __syncthreads();
    if (threadIdx.x < NV) {
        double total = 0.0;
        for (int r = 0; r < c.world; r++) total += s_part[r][threadIdx.x];      // rank order: identical bits on every rank
        vals[threadIdx.x] = total;
    }
    if (threadIdx.x == 0) *c.seq = seq;
    __syncthreads();
}

// grid_reduce_last + the cross-rank sum: returns true in the last block of every rank with the WORLD total
// in `total` (all threads of that block).
template <int NT>
__device__ __forceinline__ bool grid_reduce_last_world(double v, double* partials, unsigned* ticket, double* red_smem,
                                                       double& total, const PeerComm& c, bool block_pushed = false,
                                                       bool kernel_pushes = false) {
    // block_pushed: THIS block stored halo lines into peer memory (system fence before its ticket);
    // kernel_pushes: some block of this kernel did (the last block releases at system scope before the mailbox words)
    if (!grid_reduce_last<NT>(v, partials, ticket, red_smem, total, block_pushed)) return false;
    // world > 1: every reduction is also the ordering point for the planes this kernel wrote (the neighbours PULL their
    // ghost lines of r / y out of them in the next kernel), so the words always form a release/acquire chain
    (void)kernel_pushes;
    total = peer_allreduce_scalar<NT>(c, total, c.world > 1);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Ghost lines of the fused CG pass in LL format (data and flag in ONE 8-byte word, as in the mailbox):
// every pass PUSHES its two boundary lines of r_out and y into the neighbours' ghost buffers as words
// {sequence tag | float bits}; the next pass reads them from LOCAL memory and validates each word by its tag.
// An 8-byte word is written and read by single-copy-atomic accesses, so a reader sees the old word or the new one,
// never a mix, and needs no ordering against any other location: no system-scope fence and no remote read sits on
// the critical path of a pass (round 1 pulled the lines over NVLink and ordered them by the all-reduce, which is a
// release/acquire chain only with two system fences per pass -- measured +4.7 us per pass at 2 GPUs).
// tag = low 32 bits of the reduction sequence number the CONSUMING pass starts with (one all-reduce per pass, the
// same count on every rank); slot = tag & 1 (the all-reduce between two passes keeps a writer at most one pass ahead
// of its reader, so two slots suffice).  Layout: [slot][side][array r|y][pitch] words; side 0 = from the previous
// rank (this rank's line -1), side 1 = from the next rank (line ny).
// ---------------------------------------------------------------------------------------------
struct GhostLL {
    unsigned long long* in;          // this rank's buffer (nullptr: single GPU)
    unsigned long long* out_prev;    // the previous rank's buffer as mapped here, or nullptr at the first strip
    unsigned long long* out_next;    // the next rank's buffer, or nullptr at the last strip
    int pitch;
    __device__ __forceinline__ long long at(unsigned tag, int side, int arr) const {
        return (long long)((((int)(tag & 1u) * 2 + side) * 2 + arr)) * pitch;
    }
};
constexpr size_t ghost_ll_bytes(int pitch) { return (size_t)2 * 2 * 2 * (size_t)pitch * sizeof(unsigned long long); }

__device__ __forceinline__ void ll_store4(unsigned long long* line, int x, const float4& v, unsigned tag) {
    const unsigned long long t = (unsigned long long)tag << 32;
    unsigned long long w0 = t | __float_as_uint(v.x), w1 = t | __float_as_uint(v.y);
    unsigned long long w2 = t | __float_as_uint(v.z), w3 = t | __float_as_uint(v.w);
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(line + x), "l"(w0), "l"(w1) : "memory");
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(line + x + 2), "l"(w2), "l"(w3) : "memory");
}
__device__ __forceinline__ float4 ll_load4(const unsigned long long* line, int x, unsigned tag) {
    unsigned long long w0, w1, w2, w3;
    unsigned spins = 0u;
    do {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(line + x) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(line + x + 2) : "memory");
        if (++spins > SPIN_LIMIT) __trap();      // a neighbour died or the ranks left lock-step: fail the launch, do not hang the GPU
    } while ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag || (unsigned)(w2 >> 32) != tag || (unsigned)(w3 >> 32) != tag);
    return make_float4(__uint_as_float((unsigned)w0), __uint_as_float((unsigned)w1), __uint_as_float((unsigned)w2),
                       __uint_as_float((unsigned)w3));
}

// two lines (r and y of the same ghost line) at once: all four 16-byte loads are in flight together, one round trip to L2
// instead of two on the critical path of the strip's first / last chunk
__device__ __forceinline__ void ll_load4x2(const unsigned long long* line_a, const unsigned long long* line_b, int x, unsigned tag,
                                           float4& a, float4& b) {
    unsigned long long w[8];
    unsigned spins = 0u;
    bool ok;
    do {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(line_a + x) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[2]), "=l"(w[3]) : "l"(line_a + x + 2) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[4]), "=l"(w[5]) : "l"(line_b + x) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[6]), "=l"(w[7]) : "l"(line_b + x + 2) : "memory");
        ok = true;
#pragma unroll
        for (int i = 0; i < 8; i++) ok = ok && ((unsigned)(w[i] >> 32) == tag);
        if (++spins > SPIN_LIMIT) __trap();
    } while (!ok);
    a = make_float4(__uint_as_float((unsigned)w[0]), __uint_as_float((unsigned)w[1]), __uint_as_float((unsigned)w[2]), __uint_as_float((unsigned)w[3]));
    b = make_float4(__uint_as_float((unsigned)w[4]), __uint_as_float((unsigned)w[5]), __uint_as_float((unsigned)w[6]), __uint_as_float((unsigned)w[7]));
}

// Ghost-line destinations in the neighbours' planes (mapped peer pointers), per plane kind.
struct HaloPeers {
    float* prev_ghost;     // address of the previous rank's ghost line `ny_prev` of this plane, or nullptr
    float* next_ghost;     // address of the next rank's ghost line -1 of this plane, or nullptr
};

}  // namespace srps
