// Building blocks shared by the TMA pipeline kernels (sgns_pipe.cu, sgns_pipe_multi.cu).
//
// Work unit = one (centre, context) PAIR: 1 context row + n negative rows of syn1neg and the centre
// row of syn0.  A pair is processed by a GROUP of G lanes (G = 8, 16 or 32, chosen so that
// G * 4 * CHUNKS >= K); a warp therefore works on P = 32 / G pairs at once ("step").  All pairs
// have the same shape (n + 2 rows), so the P groups of a warp never diverge.
#pragma once
#include "common.cuh"
#include "sgns_params.h"

namespace gw2v {

constexpr int PIPE_ENTRY = 12;     // ints per pair descriptor: wtok, ctok, centre index, slot, negs[<=8]
constexpr int PIPE_RING = 128;     // descriptors per warp
constexpr int PIPE_GEN = 8;        // centres expanded per generation round (one lane per centre)
constexpr int PIPE_MAXNEG = 8;

__device__ __forceinline__ uint32_t pp_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void pp_bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pp_smem(sdst)), "l"(gsrc), "r"(bytes), "r"(pp_smem(bar)) : "memory");
}
__device__ __forceinline__ void pp_bulk_reduce_add(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                 ::"l"(gdst), "r"(pp_smem(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pp_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pp_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void pp_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void pp_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void pp_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pp_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pp_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pp_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "PP_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni PP_DONE;\n\t"
        "bra.uni PP_WAIT;\n\t"
        "PP_DONE:\n\t"
        "}\n" ::"r"(pp_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void pp_lds4(const float* p, float (&o)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void pp_sts4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// ---------------------------------------------------------------------------------------------
// Transposed butterfly inside a group of G lanes: reduce 8 per-lane values with 7..9 shuffles
// instead of 8 * log2(G).  The pairing order (xor G/2, G/4, ..., 1) is the one a plain butterfly
// uses, so every total is bit-identical to it.  Afterwards the lane holds the total of row
// row_of_lane<G>(lane); row r's total lives in lane lane_of_row<G>(r) of each group.
template <int G> __device__ __forceinline__ int row_of_lane(int lane) {
    const int lg = lane & (G - 1);
    return ((lg / (G / 2)) & 1) * 4 + ((lg / (G / 4)) & 1) * 2 + ((lg / (G / 8)) & 1);
}
template <int G> __device__ __forceinline__ int lane_of_row(int r) {     // lane index inside the group
    return ((r >> 2) & 1) * (G / 2) + ((r >> 1) & 1) * (G / 4) + (r & 1) * (G / 8);
}
template <int G>
__device__ __forceinline__ float group_reduce8(const float (&f)[8], int lane) {
    float a[4];
    const bool h1 = lane & (G / 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float send = h1 ? f[j] : f[j + 4];
        const float keep = h1 ? f[j + 4] : f[j];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, G / 2);
    }
    float b[2];
    const bool h2 = lane & (G / 4);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float send = h2 ? a[j] : a[j + 2];
        const float keep = h2 ? a[j + 2] : a[j];
        b[j] = keep + __shfl_xor_sync(0xffffffffu, send, G / 4);
    }
    const bool h3 = lane & (G / 8);
    const float send = h3 ? b[0] : b[1];
    const float keep = h3 ? b[1] : b[0];
    float c = keep + __shfl_xor_sync(0xffffffffu, send, G / 8);
#pragma unroll
    for (int o = G / 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

// ---------------------------------------------------------------------------------------------
// Pair generation: PIPE_GEN centres per round, one lane per centre for the window / context scan,
// then all 32 lanes share the Philox + alias-table work of the negatives.  Appends descriptors to
// the warp's ring; returns the number of pairs appended.  Decisions are bit-identical to the CPU
// oracle (models/sgns.py) and to the v1 kernels.
template <int GEN = PIPE_GEN, int RING = PIPE_RING>
__device__ __forceinline__ int generate_pairs(const SgnsParams& p, int T, int& gen_i, int n_warps, int* ring,
                                              int head, int lane) {
    const int n = p.negatives;
    const int ncalls = (n + 1) >> 1;
    int i = -1, lo = 0, hi = -1, wtok = 0;
    unsigned cmask = 0;                                  // valid context offsets of this lane's centre
    if (lane < GEN) {
        const long long ci = (long long)gen_i + (long long)lane * n_warps;
        if (ci < T) {
            i = (int)ci;
            uint4 rw = rand4(p.seed_lo, p.seed_hi, stream_word(STREAM_WINDOW, p.iteration),
                             p.pos0 + (unsigned long long)i, 0u);
            const int b = (int)__umulhi(rw.x, (uint32_t)p.window);
            if (p.window_mode == 0) { lo = -b; hi = b - 1; } else { const int rad = p.window - b; lo = -rad; hi = rad; }
            lo = max(lo, -i);
            hi = min(hi, T - 1 - i);
            if (hi >= lo) {
                wtok = __ldg(p.tokens + i);
                const int sid = __ldg(p.sent_id + i);
                for (int off = lo; off <= hi; ++off)
                    if (off != 0 && __ldg(p.sent_id + i + off) == sid) cmask |= 1u << (off - lo);
            }
        }
    }
    {   // advance the warp's centre cursor past this round (clamped; uniform)
        const long long nx = (long long)gen_i + (long long)GEN * n_warps;
        gen_i = nx > (long long)T ? T : (int)nx;
    }
    const int cnt = __popc(cmask);
    // exclusive prefix over the GEN lanes
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < GEN; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int total = __shfl_sync(0xffffffffu, incl, GEN - 1);
    if (total == 0) return 0;
    int base = head + incl - cnt;
    // descriptors: centre lanes write {wtok, ctok, i, slot}
    {
        unsigned m = cmask;
        while (m) {
            const int q = __ffs(m) - 1;
            m &= m - 1;
            int* e = ring + (base % RING) * PIPE_ENTRY;
            e[0] = wtok;
            e[1] = __ldg(p.tokens + i + lo + q);
            e[2] = i;
            e[3] = lo + q + p.window;                    // relative-offset slot (Philox sub-counter)
            ++base;
        }
    }
    __syncwarp();
    // negatives: (pair, philox call) items spread over all lanes
    const uint32_t sw_neg = stream_word(STREAM_NEG, p.iteration);
    for (int item = lane; item < total * ncalls; item += 32) {
        const int pr = item / ncalls, c = item - pr * ncalls;
        int* e = ring + ((head + pr) % RING) * PIPE_ENTRY;
        uint4 r = rand4(p.seed_lo, p.seed_hi, sw_neg, p.pos0 + (unsigned long long)e[2], (uint32_t)(e[3] * ncalls + c));
        e[4 + 2 * c] = alias_sample(p.alias, (uint32_t)p.vocab, r.x, r.y);
        if (2 * c + 1 < n) e[4 + 2 * c + 1] = alias_sample(p.alias, (uint32_t)p.vocab, r.z, r.w);
    }
    __syncwarp();
    return total;
}

}  // namespace gw2v
