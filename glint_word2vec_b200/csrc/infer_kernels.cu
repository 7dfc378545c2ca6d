// Serving-side kernels on a column shard:
//   gather_rows        pull(rows)                       MLLIB:514,539,639,652
//   segment_mean_rows  pullAverage(sentences)           ML:453
//   row_sqnorm         norms() partial (sum of squares) MLLIB:486
//   scores_rows        multiply(): syn0_shard . q_shard for a batch of queries (CUDA-core path;
//                      the tcgen05 path lives in nn_tc.cu)
//   topk_merge         final k of a candidate list (candidate selection: serve_fused.cu); together they
//                      replace the driver-side loop + BoundedPriorityQueue of MLLIB:600-617
#include "common.cuh"
#include "launchers.h"
#include <float.h>

namespace gw2v {

__global__ void gather_rows_kernel(const float* __restrict__ syn0, const long long* __restrict__ rows, int R,
                                   int K, float* __restrict__ out) {
    const int groups = K >> 2;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)R * groups) return;
    int r = (int)(gid / groups);
    int g = (int)(gid - (long long)r * groups);
    const float4* src = reinterpret_cast<const float4*>(syn0 + (size_t)rows[r] * K);
    reinterpret_cast<float4*>(out)[gid] = __ldg(src + g);
}

void launch_gather_rows(const float* syn0, const long long* rows, int R, int K, float* out, cudaStream_t s) {
    long long total = (long long)R * (K >> 2);
    if (total <= 0) return;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(syn0, rows, R, K, out);
}

// one warp per sentence; lanes stride over float4 column groups
__global__ void segment_mean_rows_kernel(const float* __restrict__ syn0, const long long* __restrict__ rows,
                                         const long long* __restrict__ offsets, int NS, int K,
                                         float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= NS) return;
    const long long b = offsets[warp], e = offsets[warp + 1];
    const float inv = (e > b) ? 1.0f / (float)(e - b) : 0.f;
    const int groups = K >> 2;
    for (int g = lane; g < groups; g += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long t = b; t < e; ++t) {
            float4 v = __ldg(reinterpret_cast<const float4*>(syn0 + (size_t)rows[t] * K) + g);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        reinterpret_cast<float4*>(out + (size_t)warp * K)[g] = acc;
    }
}

void launch_segment_mean_rows(const float* syn0, const long long* rows, const long long* offsets, int NS, int K,
                              float* out, cudaStream_t s) {
    if (NS <= 0) return;
    int warps_per_block = 8;
    segment_mean_rows_kernel<<<(NS + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, s>>>(
        syn0, rows, offsets, NS, K, out);
}

// G lanes (power of two <= 32) per row, 32/G rows per warp; streaming, HBM bound
__global__ void row_sqnorm_kernel(const float* __restrict__ syn0, long long V, int K, int G, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int rpw = 32 / G;
    const int sub = lane / G, lig = lane % G;
    const int groups = K >> 2;
    for (long long r0 = warp * rpw; r0 < V; r0 += nwarps * rpw) {
        long long r = r0 + sub;
        float s = 0.f;
        if (r < V) {
            const float4* row = reinterpret_cast<const float4*>(syn0 + (size_t)r * K);
            for (int g = lig; g < groups; g += G) {
                float4 v = __ldcs(row + g);
                s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        }
        for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (r < V && lig == 0) out[r] = s;
    }
}

void launch_row_sqnorm(const float* syn0, long long V, int K, float* out, int sms, cudaStream_t s) {
    if (V <= 0) return;
    int groups = K >> 2;
    int G = 1;
    while (G < groups && G < 32) G <<= 1;
    row_sqnorm_kernel<<<sms * 8, 256, 0, s>>>(syn0, V, K, G, out);
}

// scores[q, v] = sum_k syn0[v, k] * qs[q, k]   (CUDA-core path, queries in shared memory)
// one warp per row v, loops over queries; Q small (<= 64) is the intended regime
__global__ void scores_rows_kernel(const float* __restrict__ syn0, long long V, int K, const float* __restrict__ qs,
                                   int Q, float* __restrict__ out) {
    extern __shared__ float qsm[];        // [Q, K]
    for (int i = threadIdx.x; i < Q * K; i += blockDim.x) qsm[i] = qs[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int groups = K >> 2;
    for (long long v = warp; v < V; v += nwarps) {
        const float4* row = reinterpret_cast<const float4*>(syn0 + (size_t)v * K);
        for (int q0 = 0; q0 < Q; q0 += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            for (int g = lane; g < groups; g += 32) {
                float4 x = __ldg(row + g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (q0 + j < Q) {
                        float4 y = reinterpret_cast<const float4*>(qsm + (size_t)(q0 + j) * K)[g];
                        acc[j] += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s = warp_sum(acc[j]);
                if (lane == 0 && q0 + j < Q) out[(size_t)(q0 + j) * V + v] = s;
            }
        }
    }
}

void launch_scores_rows(const float* syn0, long long V, int K, const float* qs, int Q, float* out, int sms,
                        cudaStream_t s) {
    if (V <= 0 || Q <= 0) return;
    size_t smem = (size_t)Q * K * sizeof(float);
    cudaFuncSetAttribute(scores_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scores_rows_kernel<<<sms * 4, 256, smem, s>>>(syn0, V, K, qs, Q, out);
}

// ------------------------------------------------------------------ cosine + top-k
constexpr int TK_THREADS = 256;
constexpr int TK_CHUNK = 4096;

// block-wide argmax over values held in shared memory; returns (value, index) to all threads
__device__ __forceinline__ void block_argmax(const float* vals, int n, float& best, int& besti, float* red_v,
                                             int* red_i) {
    float bv = -FLT_MAX; int bi = -1;
    for (int i = threadIdx.x; i < n; i += TK_THREADS) {
        float v = vals[i];
        if (v > bv) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bv = lane < TK_THREADS / 32 ? red_v[lane] : -FLT_MAX;
        bi = lane < TK_THREADS / 32 ? red_i[lane] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[0] = bv; red_i[0] = bi; }
    }
    __syncthreads();
    best = red_v[0]; besti = red_i[0];
    __syncthreads();
}

// (stage 1 - candidate selection by threshold filter or per-chunk arg-max - lives in serve_fused.cu and serves
// both the single-shard and the column-shard path)

// stage 2: one block per query merges nchunks*k candidates
__global__ void __launch_bounds__(TK_THREADS)
topk_merge_kernel(float* __restrict__ cand_v, const long long* __restrict__ cand_i, int ncand, int k,
                  float* __restrict__ out_v, long long* __restrict__ out_i) {
    __shared__ float red_v[TK_THREADS / 32];
    __shared__ int red_i[TK_THREADS / 32];
    const int q = blockIdx.x;
    float* cv = cand_v + (size_t)q * ncand;
    const long long* ci = cand_i + (size_t)q * ncand;
    for (int j = 0; j < k; ++j) {
        float bv; int bi;
        block_argmax(cv, ncand, bv, bi, red_v, red_i);
        if (threadIdx.x == 0) {
            out_v[(size_t)q * k + j] = (bi >= 0) ? bv : 0.f;
            out_i[(size_t)q * k + j] = (bi >= 0) ? ci[bi] : -1;
            if (bi >= 0) cv[bi] = -FLT_MAX;
        }
        __syncthreads();
    }
}

void launch_topk_merge(float* cand_v, const long long* cand_i, int ncand, int Q, int k, float* out_v,
                       long long* out_i, cudaStream_t s) {
    if (Q <= 0) return;
    topk_merge_kernel<<<Q, TK_THREADS, 0, s>>>(cand_v, cand_i, ncand, k, out_v, out_i);
}

}  // namespace gw2v
