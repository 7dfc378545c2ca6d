// Pair generation kernels (SURVEY.md 2.5 K5 + K7): sentence -> (centre, context) windowing and the
// n private negatives of every pair, materialised as a flat descriptor array for the step.
//
// Reference: the Spark workers draw the window per token (MLLIB:381-390) and the Glint servers draw
// the negatives of a request from its seed [G].  Keeping this work inside the training kernel puts
// its dependent loads (sentence ids, tokens, alias table) on every warp's critical path and costs
// ~40 registers; as separate, embarrassingly parallel kernels it is a few microseconds per step and
// the training kernels become purely pair-parallel (any warp can take any pair).
//
//   pair_count_scan : one thread per centre: window radius (Philox), valid-context bitmask, pair
//                     count; single-pass chained block scan -> exclusive offsets + total
//   pair_fill       : one thread per (centre, offset slot): descriptor {centre word, context word,
//                     negatives[n]} at offset[centre] + rank
// All decisions are bit-identical to models/sgns.py (enumerate_pairs / draw_negatives).
#include "common.cuh"
#include "launchers.h"

namespace gw2v {

constexpr int PC_THREADS = 256;
constexpr int PC_ITEMS = 4;
constexpr int PC_TILE = PC_THREADS * PC_ITEMS;

// cinfo[i] = valid-context bitmask (bits 0..23, bit q <-> offset lo + q) | (-lo) << 24
__global__ void __launch_bounds__(PC_THREADS)
pair_count_scan_kernel(const int* __restrict__ tokens, const int* __restrict__ sent_id, const int* __restrict__ n_tokens,
                       uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration, unsigned long long pos0,
                       int window, int window_mode, uint32_t* __restrict__ cinfo, int* __restrict__ pair_off,
                       int* __restrict__ n_pairs, unsigned int* ticket, unsigned long long* chain, uint32_t epoch) {
    __shared__ unsigned int bid_s;
    __shared__ int warp_tot[PC_THREADS / 32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = *n_tokens;
    if (tid == 0) bid_s = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int bid = bid_s;
    const int start = (int)bid * PC_TILE + tid * PC_ITEMS;
    const uint32_t sw = stream_word(STREAM_WINDOW, iteration);

    int cnt[PC_ITEMS];
    uint32_t info[PC_ITEMS];
    int local = 0;
#pragma unroll
    for (int e = 0; e < PC_ITEMS; ++e) {
        const int i = start + e;
        cnt[e] = 0; info[e] = 0;
        if (i < T) {
            uint4 r = rand4(seed_lo, seed_hi, sw, pos0 + (unsigned long long)i, 0u);
            const int b = (int)__umulhi(r.x, (uint32_t)window);
            int lo, hi;
            if (window_mode == 0) { lo = -b; hi = b - 1; } else { const int rad = window - b; lo = -rad; hi = rad; }
            lo = max(lo, -i);
            hi = min(hi, T - 1 - i);
            uint32_t mask = 0;
            if (hi >= lo) {
                const int sid = __ldg(sent_id + i);
                for (int off = lo; off <= hi; ++off)
                    if (off != 0 && __ldg(sent_id + i + off) == sid) mask |= 1u << (off - lo);
            }
            cnt[e] = __popc(mask);
            info[e] = mask | ((uint32_t)(-lo) << 24);
            local += cnt[e];
        }
    }
    int x = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < PC_THREADS / 32 ? warp_tot[lane] : 0;
        int xs = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, xs, o); if (lane >= o) xs += y; }
        if (lane < PC_THREADS / 32) warp_tot[lane] = xs - w;
        if (lane == PC_THREADS / 32 - 1) {
            const int block_total = xs;
            unsigned long long prev = 0;
            if (bid > 0) {
                volatile unsigned long long* c = chain + (bid - 1);
                unsigned long long v;
                do { v = *c; } while ((uint32_t)(v >> 32) != epoch);
                prev = v & 0xFFFFFFFFull;
            }
            base_s = (int)prev;
            __threadfence();
            atomicExch(chain + bid, ((unsigned long long)epoch << 32) | (prev + (unsigned long long)block_total));
            if (bid == gridDim.x - 1) {
                *n_pairs = (int)(prev + block_total);
                *ticket = 0u;
            }
        }
    }
    __syncthreads();
    int o = base_s + warp_tot[warp] + (x - local);
#pragma unroll
    for (int e = 0; e < PC_ITEMS; ++e) {
        const int i = start + e;
        if (i < T) { cinfo[i] = info[e]; pair_off[i] = o; }
        o += cnt[e];
    }
}

// one thread per (centre, offset slot q); PD ints per descriptor: {wtok, ctok, negs[n], pad}
__global__ void pair_fill_kernel(const int* __restrict__ tokens, const int* __restrict__ n_tokens,
                                 const uint32_t* __restrict__ cinfo, const int* __restrict__ pair_off,
                                 const int2* __restrict__ alias, int vocab, uint32_t seed_lo, uint32_t seed_hi,
                                 uint32_t iteration, unsigned long long pos0, int window, int negatives, int slots,
                                 int pd, int* __restrict__ desc) {
    const int T = *n_tokens;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = gid / slots;
    const int q = (int)(gid - i * slots);
    if (i >= T) return;
    const uint32_t info = __ldg(cinfo + i);
    const uint32_t mask = info & 0xFFFFFFu;
    if (!((mask >> q) & 1u)) return;
    const int lo = -(int)(info >> 24);
    const int off = lo + q;
    const int rank = __popc(mask & ((1u << q) - 1u));
    int* e = desc + (size_t)(__ldg(pair_off + i) + rank) * pd;
    const int ctok = __ldg(tokens + i + off);
    e[0] = __ldg(tokens + i);
    e[1] = ctok;
    const int ncalls = (negatives + 1) >> 1;
    const uint32_t sw = stream_word(STREAM_NEG, iteration);
    const int slot = off + window;
    for (int c = 0; c < ncalls; ++c) {
        uint4 r = rand4(seed_lo, seed_hi, sw, pos0 + (unsigned long long)i, (uint32_t)(slot * ncalls + c));
        e[2 + 2 * c] = alias_sample(alias, (uint32_t)vocab, r.x, r.y);
        if (2 * c + 1 < negatives) e[2 + 2 * c + 1] = alias_sample(alias, (uint32_t)vocab, r.z, r.w);
    }
}

int pairgen_max_blocks(int max_tokens) { return (max_tokens + PC_TILE - 1) / PC_TILE + 1; }
int pairgen_desc_ints(int negatives) { return ((2 + negatives) + 3) / 4 * 4; }

// grid is sized from the host-side upper bound `max_tokens` (the device count may be smaller after
// sub-sampling); blocks beyond the device count contribute zero pairs.
void launch_pairgen(const int* tokens, const int* sent_id, const int* n_tokens, int max_tokens, const int2* alias,
                    int vocab, uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration, unsigned long long pos0,
                    int window, int window_mode, int negatives, uint32_t* cinfo, int* pair_off, int* n_pairs,
                    int* desc, unsigned int* ticket, unsigned long long* chain, uint32_t epoch, cudaStream_t stream) {
    if (max_tokens <= 0) { cudaMemsetAsync(n_pairs, 0, sizeof(int), stream); return; }
    const int grid = (max_tokens + PC_TILE - 1) / PC_TILE;
    pair_count_scan_kernel<<<grid, PC_THREADS, 0, stream>>>(tokens, sent_id, n_tokens, seed_lo, seed_hi, iteration,
                                                            pos0, window, window_mode, cinfo, pair_off, n_pairs,
                                                            ticket, chain, epoch);
    const int slots = 2 * window + 1;
    const long long total = (long long)max_tokens * slots;
    pair_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(tokens, n_tokens, cinfo, pair_off, alias,
                                                                         vocab, seed_lo, seed_hi, iteration, pos0,
                                                                         window, negatives, slots, pairgen_desc_ints(negatives),
                                                                         desc);
}

}  // namespace gw2v
