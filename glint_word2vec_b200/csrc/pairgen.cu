// Pair generation kernels (SURVEY.md 2.5 K5 + K7): sentence -> (centre, context) windowing and the
// n private negatives of every pair, materialised as a flat descriptor array for the step.
//
// Reference: the Spark workers draw the window per token (MLLIB:381-390) and the Glint servers draw
// the negatives of a request from its seed [G].  Keeping this work inside the training kernel puts
// its dependent loads (sentence ids, tokens, alias table) on every warp's critical path and costs
// ~40 registers; as separate, embarrassingly parallel kernels it is a few microseconds per step and
// the training kernels become purely pair-parallel (any warp can take any pair).
//
//   pair_count      : one thread per centre: window radius (Philox), valid-context bitmask, pair
//                     count, exclusive offset inside the 1024-centre tile, tile sums
//   pair_tile_scan  : one CTA scans the (<= 1024) tile sums and writes the total pair count
//   pair_fill       : one thread per (centre, offset slot): descriptor {centre word, context word,
//                     negatives[n]} at tile_prefix + offset[centre] + rank
// (a single-pass chained scan was measured first: its serial CTA-to-CTA hand-off cost ~100 us per step)
// All decisions are bit-identical to models/sgns.py (enumerate_pairs / draw_negatives).
#include "common.cuh"
#include "launchers.h"
#include "sgns_tile.h"

namespace gw2v {

constexpr int PC_THREADS = 256;
constexpr int PC_ITEMS = 2;          // 512-centre tiles: 256 CTAs for a 131 k-token step (4 items left 128 CTAs on 148 SMs)
constexpr int PC_TILE = PC_THREADS * PC_ITEMS;

// cinfo[i] = valid-context bitmask (bits 0..23, bit q <-> offset lo + q) | (-lo) << 24
// pair_off[i] = exclusive offset of centre i INSIDE its tile; tile_sum[b] = pairs of tile b
__global__ void __launch_bounds__(PC_THREADS)
pair_count_kernel(const int* __restrict__ tokens, const int* __restrict__ sent_id, const int* __restrict__ n_tokens,
                  uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration, unsigned long long pos0,
                  int window, int window_mode, uint32_t* __restrict__ cinfo, int* __restrict__ pair_off,
                  int* __restrict__ tile_sum) {
    __shared__ int warp_tot[PC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = *n_tokens;
    const int start = (int)blockIdx.x * PC_TILE + tid * PC_ITEMS;
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
        if (lane == PC_THREADS / 32 - 1) tile_sum[blockIdx.x] = xs;
    }
    __syncthreads();
    int o = warp_tot[warp] + (x - local);
#pragma unroll
    for (int e = 0; e < PC_ITEMS; ++e) {
        const int i = start + e;
        if (i < T) { cinfo[i] = info[e]; pair_off[i] = o; }
        o += cnt[e];
    }
}

// exclusive scan of the (<= 1024) tile sums by one CTA; writes the total pair count
__global__ void __launch_bounds__(1024)
pair_tile_scan_kernel(int* __restrict__ tile_sum, int ntiles, int* __restrict__ n_pairs, float* __restrict__ stats) {
    if (stats != nullptr && threadIdx.x < 4) stats[threadIdx.x] = 0.f;      // step statistics start at zero
    __shared__ int wt[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int v = tid < ntiles ? tile_sum[tid] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wt[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = wt[lane];
        int xs = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, xs, o); if (lane >= o) xs += y; }
        wt[lane] = xs - w;
        if (lane == 31) *n_pairs = xs;
    }
    __syncthreads();
    if (tid < ntiles) tile_sum[tid] = wt[warp] + x - v;          // exclusive prefix of tile tid
}

// one thread per (centre, offset slot q).  Descriptor = PD ints {wtok, ctok, negs[<= 7], pad}; a negative slot that
// is not used holds -1.  More than 7 negatives per pair (the reference accepts any n > 0, MLLIB:184-190; BASELINE.json
// config 4 uses n = 10) are spread over `splits` consecutive descriptors of the same pair: the first carries the context
// and negatives 0..6, the following ones carry the next 7 negatives each and mark their context word as inactive (bit 31:
// it is still needed for the "negative == positive is skipped" rule).  The training kernels therefore never see more than
// 1 + 7 rows per descriptor whatever n is.
constexpr int PG_MAXNEG = 21;
__global__ void pair_fill_kernel(const int* __restrict__ tokens, const int* __restrict__ n_tokens,
                                 const uint32_t* __restrict__ cinfo, const int* __restrict__ pair_off,
                                 const int* __restrict__ tile_prefix,
                                 const int2* __restrict__ alias, int vocab, uint32_t seed_lo, uint32_t seed_hi,
                                 uint32_t iteration, unsigned long long pos0, int window, int negatives, int slots,
                                 int pd, int splits, int share_centre, int* __restrict__ desc) {
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
    const size_t pair = (size_t)(__ldg(tile_prefix + (int)(i / PC_TILE)) + __ldg(pair_off + i) + rank);
    int* e = desc + pair * (size_t)splits * pd;
    const int ctok = __ldg(tokens + i + off);
    const int wtok = __ldg(tokens + i);
    const int ncalls = (negatives + 1) >> 1;
    const uint32_t sw = stream_word(STREAM_NEG, iteration);
    const int slot = share_centre ? 0 : off + window;     // neg_sharing="centre": one draw per centre
    // all Philox calls first, then all alias-table reads in flight together (they are independent random
    // 8-byte reads into an 80 MB table), then the selects; the descriptor leaves as 16-byte stores
    if (splits == 1) {
        uint32_t idx[8], sel[8];
        int2 ent[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < ncalls) {
                const uint4 r = rand4(seed_lo, seed_hi, sw, pos0 + (unsigned long long)i, (uint32_t)(slot * ncalls + c));
                idx[2 * c] = __umulhi(r.x, (uint32_t)vocab); sel[2 * c] = r.y;
                idx[2 * c + 1] = __umulhi(r.z, (uint32_t)vocab); sel[2 * c + 1] = r.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < negatives) ent[j] = __ldg(alias + idx[j]);
        int w[12];
        w[0] = wtok; w[1] = ctok;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            w[2 + j] = (j < negatives) ? ((sel[j] < (uint32_t)ent[j].x) ? (int)idx[j] : ent[j].y) : -1;
        w[10] = w[11] = -1;
        int4* e4 = reinterpret_cast<int4*>(e);                 // pd is a multiple of 4 ints: 16-byte aligned
        e4[0] = make_int4(w[0], w[1], w[2], w[3]);
        e4[1] = make_int4(w[4], w[5], w[6], w[7]);
        if (pd > 8) e4[2] = make_int4(w[8], w[9], w[10], w[11]);
        return;
    }
    int neg[PG_MAXNEG + 1];
    for (int c = 0; c < ncalls; ++c) {
        const uint4 r = rand4(seed_lo, seed_hi, sw, pos0 + (unsigned long long)i, (uint32_t)(slot * ncalls + c));
        neg[2 * c] = alias_sample(alias, (uint32_t)vocab, r.x, r.y);
        neg[2 * c + 1] = alias_sample(alias, (uint32_t)vocab, r.z, r.w);
    }
    for (int sp = 0; sp < splits; ++sp) {
        int w[12];
        w[0] = wtok;
        w[1] = sp == 0 ? ctok : (int)((uint32_t)ctok | 0x80000000u);
        for (int j = 0; j < 7; ++j) w[2 + j] = (7 * sp + j < negatives) ? neg[7 * sp + j] : -1;
        w[9] = w[10] = w[11] = -1;
        int4* e4 = reinterpret_cast<int4*>(e + (size_t)sp * pd);
        e4[0] = make_int4(w[0], w[1], w[2], w[3]);
        e4[1] = make_int4(w[4], w[5], w[6], w[7]);
        e4[2] = make_int4(w[8], w[9], w[10], w[11]);
    }
}

int pairgen_max_blocks(int max_tokens) { return (max_tokens + PC_TILE - 1) / PC_TILE + 1; }
int pairgen_splits(int negatives) { return negatives <= 7 ? 1 : (negatives + 6) / 7; }
int pairgen_desc_ints(int negatives) { return negatives <= 7 ? ((2 + negatives) + 3) / 4 * 4 : 12; }
int pairgen_max_negatives() { return PG_MAXNEG; }
int pairgen_max_tokens() { return 1024 * PC_TILE; }

// grid is sized from the host-side upper bound `max_tokens` (the device count may be smaller after
// sub-sampling); tiles beyond the device count contribute zero pairs.  Three launches, no serial chain:
// per-tile counts + local offsets -> scan of the tile sums -> descriptor fill.
void launch_pairgen(const int* tokens, const int* sent_id, const int* n_tokens, int max_tokens, const int2* alias,
                    int vocab, uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration, unsigned long long pos0,
                    int window, int window_mode, int negatives, int share_centre, uint32_t* cinfo, int* pair_off,
                    int* n_pairs, int* desc, int* tile_ws, float* stats, cudaStream_t stream) {
    if (max_tokens <= 0) {
        cudaMemsetAsync(n_pairs, 0, sizeof(int), stream);
        if (stats) cudaMemsetAsync(stats, 0, 4 * sizeof(float), stream);
        return;
    }
    const int grid = (max_tokens + PC_TILE - 1) / PC_TILE;          // <= 1024 (checked by the binding)
    int* tile_sum = tile_ws;                                         // >= grid ints
    pair_count_kernel<<<grid, PC_THREADS, 0, stream>>>(tokens, sent_id, n_tokens, seed_lo, seed_hi, iteration, pos0,
                                                       window, window_mode, cinfo, pair_off, tile_sum);
    pair_tile_scan_kernel<<<1, 1024, 0, stream>>>(tile_sum, grid, n_pairs, stats);
    // offset slots a centre can use: reference window (Q2) spans [-b, b-1] with b <= window-1 -> 2(window-1) slots
    int slots = (window_mode == 0) ? 2 * (window - 1) : 2 * window + 1;
    if (slots < 1) slots = 1;
    const long long total = (long long)max_tokens * slots;
    pair_fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(tokens, n_tokens, cinfo, pair_off, tile_sum,
                                                                         alias, vocab, seed_lo, seed_hi, iteration,
                                                                         pos0, window, negatives, slots,
                                                                         pairgen_desc_ints(negatives), pairgen_splits(negatives),
                                                                         share_centre, desc);
}

// window masks + total pair count only: what the tensor-core tile kernel needs (it derives the pairs from the masks)
void launch_paircount(const int* tokens, const int* sent_id, const int* n_tokens, int max_tokens, uint32_t seed_lo,
                      uint32_t seed_hi, uint32_t iteration, unsigned long long pos0, int window, int window_mode,
                      uint32_t* cinfo, int* pair_off, int* n_pairs, int* tile_ws, float* stats, cudaStream_t stream) {
    if (max_tokens <= 0) {
        cudaMemsetAsync(n_pairs, 0, sizeof(int), stream);
        if (stats) cudaMemsetAsync(stats, 0, 4 * sizeof(float), stream);
        return;
    }
    const int grid = (max_tokens + PC_TILE - 1) / PC_TILE;
    pair_count_kernel<<<grid, PC_THREADS, 0, stream>>>(tokens, sent_id, n_tokens, seed_lo, seed_hi, iteration, pos0,
                                                       window, window_mode, cinfo, pair_off, tile_ws);
    pair_tile_scan_kernel<<<1, 1024, 0, stream>>>(tile_ws, grid, n_pairs, stats);
}

}  // namespace gw2v
