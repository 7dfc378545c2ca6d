// Pre-processing kernels (SURVEY.md 2.5 K6, K14 + weight init):
//   subsample_compact : frequent-word sub-sampling (MLLIB:371-379) as an
//                       order-preserving stream compaction (count / tile scan /
//                       scatter), sentence ids ride along.
//   zipf_stream       : synthetic Zipf token stream from the alias table.
//   init_syn0         : syn0 ~ U(-0.5,0.5)/d as a pure function of (row, col).
// Windowing (K7) and negative sampling (K5): pairgen.cu.
#include "common.cuh"
#include "launchers.h"

namespace gw2v {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 2;          // 512-token tiles: 256 CTAs for a 131 k-token step (8 items left 64 CTAs on 148 SMs)
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ bool sc_keep(const int* __restrict__ tok_in, const uint32_t* __restrict__ keep_thresh,
                                        uint32_t seed_lo, uint32_t seed_hi, uint32_t sw, unsigned long long raw_pos0,
                                        int i, int& tok) {
    tok = __ldg(tok_in + i);
    uint4 r = rand4(seed_lo, seed_hi, sw, raw_pos0 + (unsigned long long)i, 0u);
    return r.x <= __ldg(keep_thresh + tok);
}

// pass 1: kept tokens per tile
__global__ void __launch_bounds__(SC_THREADS)
subsample_count_kernel(const int* __restrict__ tok_in, int T, const uint32_t* __restrict__ keep_thresh,
                       uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration, unsigned long long raw_pos0,
                       int* __restrict__ tile_sum) {
    __shared__ int wt[SC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start = (int)blockIdx.x * SC_TILE + tid * SC_ITEMS;
    const uint32_t sw = stream_word(STREAM_SUBSAMPLE, iteration);
    int local = 0;
#pragma unroll
    for (int e = 0; e < SC_ITEMS; ++e) {
        const int i = start + e;
        int tok;
        if (i < T && sc_keep(tok_in, keep_thresh, seed_lo, seed_hi, sw, raw_pos0, i, tok)) ++local;
    }
    local = (int)warp_sum((float)local);           // counts <= 256: exact in fp32
    if (lane == 0) wt[warp] = local;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int w = 0; w < SC_THREADS / 32; ++w) s += wt[w];
        tile_sum[blockIdx.x] = s;
    }
}

// pass 2: one CTA turns the (<= 1024) tile sums into exclusive prefixes and writes the total
__global__ void __launch_bounds__(1024)
subsample_tile_scan_kernel(int* __restrict__ tile_sum, int ntiles, int* __restrict__ count_out) {
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
        if (lane == 31) *count_out = xs;
    }
    __syncthreads();
    if (tid < ntiles) tile_sum[tid] = wt[warp] + x - v;
}

// pass 3: order-preserving scatter (decisions recomputed, identical by construction)
__global__ void __launch_bounds__(SC_THREADS)
subsample_scatter_kernel(const int* __restrict__ tok_in, const int* __restrict__ sid_in, int T,
                         const uint32_t* __restrict__ keep_thresh, uint32_t seed_lo, uint32_t seed_hi,
                         uint32_t iteration, unsigned long long raw_pos0, const int* __restrict__ tile_prefix,
                         int* __restrict__ tok_out, int* __restrict__ sid_out) {
    __shared__ int warp_tot[SC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int start = (int)blockIdx.x * SC_TILE + tid * SC_ITEMS;
    const uint32_t sw = stream_word(STREAM_SUBSAMPLE, iteration);
    int tok[SC_ITEMS];
    bool keep[SC_ITEMS];
    int local = 0;
#pragma unroll
    for (int e = 0; e < SC_ITEMS; ++e) {
        const int i = start + e;
        keep[e] = false; tok[e] = 0;
        if (i < T) {
            keep[e] = sc_keep(tok_in, keep_thresh, seed_lo, seed_hi, sw, raw_pos0, i, tok[e]);
            local += keep[e] ? 1 : 0;
        }
    }
    int x = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < SC_THREADS / 32 ? warp_tot[lane] : 0;
        int xs = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, xs, o); if (lane >= o) xs += y; }
        if (lane < SC_THREADS / 32) warp_tot[lane] = xs - w;
    }
    __syncthreads();
    int o = tile_prefix[blockIdx.x] + warp_tot[warp] + (x - local);
#pragma unroll
    for (int e = 0; e < SC_ITEMS; ++e) {
        if (keep[e]) {
            tok_out[o] = tok[e];
            sid_out[o] = __ldg(sid_in + start + e);
            ++o;
        }
    }
}

// three launches, no CTA-to-CTA chain (a single-pass chained scan cost ~100 us per step on B200)
void launch_subsample_compact(const int* tok_in, const int* sid_in, int T, const uint32_t* keep_thresh,
                              uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration,
                              unsigned long long raw_pos0, int* tok_out, int* sid_out, int* count_out,
                              int* tile_ws,
                              cudaStream_t stream) {
    if (T <= 0) { cudaMemsetAsync(count_out, 0, sizeof(int), stream); return; }
    const int grid = (T + SC_TILE - 1) / SC_TILE;                  // <= 1024 tiles (512 k tokens per step)
    int* tile_sum = tile_ws;                                       // >= grid ints
    subsample_count_kernel<<<grid, SC_THREADS, 0, stream>>>(tok_in, T, keep_thresh, seed_lo, seed_hi, iteration,
                                                            raw_pos0, tile_sum);
    subsample_tile_scan_kernel<<<1, 1024, 0, stream>>>(tile_sum, grid, count_out);
    subsample_scatter_kernel<<<grid, SC_THREADS, 0, stream>>>(tok_in, sid_in, T, keep_thresh, seed_lo, seed_hi,
                                                              iteration, raw_pos0, tile_sum, tok_out, sid_out);
}

int subsample_max_tokens() { return 1024 * SC_TILE; }
int subsample_max_blocks(int max_tokens) { return (max_tokens + SC_TILE - 1) / SC_TILE + 1; }

// ---------------------------------------------------------------------------

__global__ void zipf_stream_kernel(const int2* __restrict__ alias, int vocab, uint32_t seed_lo, uint32_t seed_hi,
                                   unsigned long long pos0, int n, int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = rand4(seed_lo, seed_hi, stream_word(STREAM_ZIPF, 0), pos0 + (unsigned long long)i, 0u);
    out[i] = alias_sample(alias, (uint32_t)vocab, r.x, r.y);
}

void launch_zipf_stream(const int2* alias, int vocab, uint32_t seed_lo, uint32_t seed_hi,
                        unsigned long long pos0, int n, int* out, cudaStream_t stream) {
    if (n <= 0) return;
    zipf_stream_kernel<<<(n + 255) / 256, 256, 0, stream>>>(alias, vocab, seed_lo, seed_hi, pos0, n, out);
}

// ---------------------------------------------------------------------------

// element (row, global col c) = (unit(philox(seed, INIT, row, c/4)[c%4]) - 0.5) / d ; zero beyond d
__global__ void init_syn0_kernel(float* __restrict__ syn0, long long vocab, int K, int col_start, int vector_size,
                                 uint32_t seed_lo, uint32_t seed_hi) {
    const int groups = K >> 2;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= vocab * groups) return;
    long long row = gid / groups;
    int g = (int)(gid - row * groups);
    int gcol = col_start + g * 4;
    uint4 r = rand4(seed_lo, seed_hi, stream_word(STREAM_INIT, 0), (unsigned long long)row, (uint32_t)(gcol >> 2));
    const float inv = 1.0f / (float)vector_size;
    float4 o;
    o.x = (gcol + 0 < vector_size) ? (u32_to_unit_float(r.x) - 0.5f) / (float)vector_size : 0.f;
    o.y = (gcol + 1 < vector_size) ? (u32_to_unit_float(r.y) - 0.5f) / (float)vector_size : 0.f;
    o.z = (gcol + 2 < vector_size) ? (u32_to_unit_float(r.z) - 0.5f) / (float)vector_size : 0.f;
    o.w = (gcol + 3 < vector_size) ? (u32_to_unit_float(r.w) - 0.5f) / (float)vector_size : 0.f;
    (void)inv;
    reinterpret_cast<float4*>(syn0)[gid] = o;
}

void launch_init_syn0(float* syn0, long long vocab, int K, int col_start, int vector_size, uint32_t seed_lo,
                      uint32_t seed_hi, cudaStream_t stream) {
    long long total = vocab * (K >> 2);
    if (total <= 0) return;
    long long blocks = (total + 255) / 256;
    init_syn0_kernel<<<(unsigned)blocks, 256, 0, stream>>>(syn0, vocab, K, col_start, vector_size, seed_lo, seed_hi);
}

}  // namespace gw2v
