// Serving kernels fused with their collectives over NVLink peer memory (world > 1).
//
//   reference op                         collective that followed it      here (one kernel: compute + push + signal)
//   pull(rows)          MLLIB:514,639    gather of column slices          gather_rows_push        -> every rank's [R, S*K] buffer
//   pullAverage(sent.)  ML:453           gather of column slices          segment_mean_push       -> every rank's [NS, S*K] buffer
//   norms()             MLLIB:486        sum of S partial V-vectors       row_sqnorm_push         -> reduce-scatter: owner's slab
//                                                                         reduce_finish_push      -> sqrt, all-gather of the owner slices
//   multiply(q)+top-k   MLLIB:598-617    sum of S partial V x Q scores    scores_rows_push / scores_tc (nn_tc.cu) epilogue
//                                                                            -> reduce-scatter: partial score tiles are stored
//                                                                               straight into the owner rank's slab
//                                                                         topk_owned + merge_push -> owner sums the S partials,
//                                                                               / norm, top-k of its V/S rows, candidates pushed to all
//                                                                         topk_merge (infer_kernels.cu) -> final k of S*k candidates
// Row v is owned by rank v / vown (vown = ceil(V/S) rounded up to 128).  Synchronisation: serve_common.cuh.
#include "serve_common.cuh"
#include "launchers.h"
#include <float.h>

namespace gw2v {

// ------------------------------------------------------------------ wait / barrier
__global__ void serve_wait_kernel(const uint32_t* __restrict__ flags_local, int world, uint32_t seq,
                                  int* __restrict__ error_flag) {
    const int r = threadIdx.x;
    if (r < world) {
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(flags_local + r) - seq) < 0) {
            if (globaltimer_ns() - t0 > 20000000000ull) {      // a peer died: fail loudly instead of hanging
                *error_flag = 2;
                __threadfence_system();
                __trap();
            }
        }
    }
}

// pure barrier: "I have consumed everything pushed so far" -> peers may overwrite my buffers
__global__ void serve_barrier_kernel(ServeSync s) {
    const int r = threadIdx.x;
    if (r < s.world) {
        __threadfence_system();
        st_release_sys(s.flags[r] + s.rank, s.seq);
        const unsigned long long t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(s.flags[s.rank] + r) - s.seq) < 0) {
            if (globaltimer_ns() - t0 > 20000000000ull) {
                *s.error_flag = 2;
                __threadfence_system();
                __trap();
            }
        }
    }
}

void launch_serve_wait(const uint32_t* flags_local, int world, uint32_t seq, int* error_flag, cudaStream_t st) {
    serve_wait_kernel<<<1, 32, 0, st>>>(flags_local, world, seq, error_flag);
}
void launch_serve_barrier(const ServeSync& s, cudaStream_t st) { serve_barrier_kernel<<<1, 32, 0, st>>>(s); }

// ------------------------------------------------------------------ K8: pull
__global__ void gather_rows_push_kernel(const float* __restrict__ syn0, const long long* __restrict__ rows, int R,
                                        int K, PeerPtrs out, int ldo, ServeSync s) {
    const int groups = K >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < (long long)R * groups) {
        const int r = (int)(gid / groups);
        const int g = (int)(gid - (long long)r * groups);
        const float4 v = __ldg(reinterpret_cast<const float4*>(syn0 + (size_t)rows[r] * K) + g);
        const size_t o = (size_t)r * ldo + (size_t)s.rank * K + (size_t)g * 4;
        for (int p = 0; p < s.world; ++p) *reinterpret_cast<float4*>(out.p[p] + o) = v;
    }
    serve_cta_done(s);
}

void launch_gather_rows_push(const float* syn0, const long long* rows, int R, int K, const PeerPtrs& out, int ldo,
                             const ServeSync& s, cudaStream_t st) {
    long long total = (long long)R * (K >> 2);
    unsigned grid = (unsigned)((total + 255) / 256);
    if (grid == 0) grid = 1;                                   // the signal must still be published
    gather_rows_push_kernel<<<grid, 256, 0, st>>>(syn0, rows, R, K, out, ldo, s);
}

// ------------------------------------------------------------------ K9: pullAverage (one warp per sentence)
__global__ void segment_mean_push_kernel(const float* __restrict__ syn0, const long long* __restrict__ rows,
                                         const long long* __restrict__ offsets, int NS, int K, PeerPtrs out, int ldo,
                                         ServeSync s) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp < NS) {
        const long long b = offsets[warp], e = offsets[warp + 1];
        const float inv = (e > b) ? 1.0f / (float)(e - b) : 0.f;
        const int groups = K >> 2;
        for (int g = lane; g < groups; g += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (long long t = b; t < e; ++t) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(syn0 + (size_t)rows[t] * K) + g);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
            const size_t o = (size_t)warp * ldo + (size_t)s.rank * K + (size_t)g * 4;
            for (int p = 0; p < s.world; ++p) *reinterpret_cast<float4*>(out.p[p] + o) = acc;
        }
    }
    serve_cta_done(s);
}

void launch_segment_mean_push(const float* syn0, const long long* rows, const long long* offsets, int NS, int K,
                              const PeerPtrs& out, int ldo, const ServeSync& s, cudaStream_t st) {
    const int wpb = 8;
    int grid = (NS + wpb - 1) / wpb;
    if (grid == 0) grid = 1;
    segment_mean_push_kernel<<<grid, wpb * 32, 0, st>>>(syn0, rows, offsets, NS, K, out, ldo, s);
}

// ------------------------------------------------------------------ K10: norms, step 1 (partial -> owner slab)
// slab layout on the owner: [src rank][vown]
__global__ void row_sqnorm_push_kernel(const float* __restrict__ syn0, long long V, int K, int G, PeerPtrs slab,
                                       long long vown, ServeSync s) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int rpw = 32 / G;
    const int sub = lane / G, lig = lane % G;
    const int groups = K >> 2;
    for (long long r0 = warp * rpw; r0 < V; r0 += nwarps * rpw) {
        const long long r = r0 + sub;
        float acc = 0.f;
        if (r < V) {
            const float4* row = reinterpret_cast<const float4*>(syn0 + (size_t)r * K);
            for (int g = lig; g < groups; g += G) {
                const float4 v = __ldcs(row + g);
                acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        }
        for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (r < V && lig == 0) {
            const int owner = (int)(r / vown);
            slab.p[owner][(size_t)s.rank * vown + (size_t)(r - (long long)owner * vown)] = acc;
        }
    }
    serve_cta_done(s);
}

void launch_row_sqnorm_push(const float* syn0, long long V, int K, const PeerPtrs& slab, long long vown,
                            const ServeSync& s, int sms, cudaStream_t st) {
    int groups = K >> 2;
    int G = 1;
    while (G < groups && G < 32) G <<= 1;
    row_sqnorm_push_kernel<<<sms * 8, 256, 0, st>>>(syn0, V, K, G, slab, vown, s);
}

// ------------------------------------------------------------------ K10/K11: owner sums S partial slices and
// all-gathers its slice of the result (norms: with sqrt; multiply: plain) into every rank's full vector
__global__ void reduce_finish_push_kernel(const float* __restrict__ slab_local, int nsrc, long long vown,
                                          long long nvalid, int take_sqrt, PeerPtrs full, ServeSync s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nvalid) {
        float acc = 0.f;
        for (int r = 0; r < nsrc; ++r) acc += slab_local[(size_t)r * vown + i];      // fixed rank order
        if (take_sqrt) acc = sqrtf(acc);
        const size_t o = (size_t)s.rank * vown + (size_t)i;
        for (int p = 0; p < s.world; ++p) full.p[p][o] = acc;
    }
    serve_cta_done(s);
}

void launch_reduce_finish_push(const float* slab_local, int nsrc, long long vown, long long nvalid, int take_sqrt,
                               const PeerPtrs& full, const ServeSync& s, cudaStream_t st) {
    unsigned grid = (unsigned)((nvalid + 255) / 256);
    if (grid == 0) grid = 1;
    reduce_finish_push_kernel<<<grid, 256, 0, st>>>(slab_local, nsrc, vown, nvalid, take_sqrt, full, s);
}

// ------------------------------------------------------------------ K11: scores, CUDA-core path (small Q)
// partial[q, v] = syn0_shard[v, :] . qs[q, :]  stored into the OWNER's slab [src rank][Q][vown]
__global__ void scores_rows_push_kernel(const float* __restrict__ syn0, long long V, int K,
                                        const float* __restrict__ qs, int Q, PeerPtrs slab, long long vown,
                                        ServeSync s) {
    extern __shared__ float qsm[];        // [Q, K]
    for (int i = threadIdx.x; i < Q * K; i += blockDim.x) qsm[i] = qs[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int groups = K >> 2;
    for (long long v = warp; v < V; v += nwarps) {
        const float4* row = reinterpret_cast<const float4*>(syn0 + (size_t)v * K);
        const int owner = (int)(v / vown);
        float* dst = slab.p[owner] + (size_t)s.rank * Q * vown + (size_t)(v - (long long)owner * vown);
        for (int q0 = 0; q0 < Q; q0 += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            for (int g = lane; g < groups; g += 32) {
                const float4 x = __ldg(row + g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (q0 + j < Q) {
                        const float4 y = reinterpret_cast<const float4*>(qsm + (size_t)(q0 + j) * K)[g];
                        acc[j] += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = warp_sum(acc[j]);
                if (lane == 0 && q0 + j < Q) dst[(size_t)(q0 + j) * vown] = t;
            }
        }
    }
    serve_cta_done(s);
}

void launch_scores_rows_push(const float* syn0, long long V, int K, const float* qs, int Q, const PeerPtrs& slab,
                             long long vown, const ServeSync& s, int sms, cudaStream_t st) {
    size_t smem = (size_t)Q * K * sizeof(float);
    cudaFuncSetAttribute(scores_rows_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scores_rows_push_kernel<<<sms * 4, 256, smem, st>>>(syn0, V, K, qs, Q, slab, vown, s);
}

// ------------------------------------------------------------------ K11: top-k of the owned rows
constexpr int TKO_THREADS = 256;
constexpr int TKO_CHUNK = 4096;

__device__ __forceinline__ void tko_block_argmax(const float* vals, int n, float& best, int& besti, float* red_v,
                                                 int* red_i) {
    float bv = -FLT_MAX; int bi = -1;
    for (int i = threadIdx.x; i < n; i += TKO_THREADS) {
        const float v = vals[i];
        if (v > bv) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bv = lane < TKO_THREADS / 32 ? red_v[lane] : -FLT_MAX;
        bi = lane < TKO_THREADS / 32 ? red_i[lane] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[0] = bv; red_i[0] = bi; }
    }
    __syncthreads();
    best = red_v[0]; besti = red_i[0];
    __syncthreads();
}

// stage 1: per (query, chunk of owned rows): cos = (sum over src of partial score) / norm; emit the chunk's top-k
__global__ void __launch_bounds__(TKO_THREADS)
topk_owned_stage1_kernel(const float* __restrict__ slab_local, int nsrc, int Q, long long vown, long long nvalid,
                         const float* __restrict__ norms_owned, long long row_base, int k,
                         float* __restrict__ cand_v, long long* __restrict__ cand_i, int nchunks, int chunk_stride) {
    __shared__ float vals[TKO_CHUNK];
    __shared__ float red_v[TKO_THREADS / 32];
    __shared__ int red_i[TKO_THREADS / 32];
    const int q = blockIdx.y, chunk = blockIdx.x;
    const long long base = (long long)chunk * chunk_stride * TKO_CHUNK;      // stride > 1: evenly spread sample
    const int n = (int)max(0ll, min((long long)TKO_CHUNK, nvalid - base));
    for (int i = threadIdx.x; i < n; i += TKO_THREADS) {
        const float nr = __ldg(norms_owned + base + i);
        float sc = 0.f;
        for (int r = 0; r < nsrc; ++r) sc += slab_local[((size_t)r * Q + q) * vown + base + i];   // fixed rank order
        vals[i] = nr > 0.f ? sc / nr : 0.f;
    }
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        float bv; int bi;
        tko_block_argmax(vals, n, bv, bi, red_v, red_i);
        if (threadIdx.x == 0) {
            const size_t o = ((size_t)q * nchunks + chunk) * k + j;
            cand_v[o] = (bi >= 0) ? bv : -FLT_MAX;
            cand_i[o] = (bi >= 0) ? row_base + base + bi : -1;
            if (bi >= 0) vals[bi] = -FLT_MAX;
        }
        __syncthreads();
    }
}

// stage 2: one block per query merges the rank's chunk candidates and pushes its k winners to every rank:
// destination layout [q][src rank][k]  (so the final merge sees S*k contiguous candidates per query)
__global__ void __launch_bounds__(TKO_THREADS)
topk_merge_push_kernel(float* __restrict__ cand_v, const long long* __restrict__ cand_i, int ncand, int k,
                       PeerPtrs out_v, PeerIdx out_i, ServeSync s) {
    __shared__ float red_v[TKO_THREADS / 32];
    __shared__ int red_i[TKO_THREADS / 32];
    const int q = blockIdx.x;
    float* cv = cand_v + (size_t)q * ncand;
    const long long* ci = cand_i + (size_t)q * ncand;
    for (int j = 0; j < k; ++j) {
        float bv; int bi;
        tko_block_argmax(cv, ncand, bv, bi, red_v, red_i);
        if (threadIdx.x == 0) {
            const float v = (bi >= 0) ? bv : -FLT_MAX;
            const long long id = (bi >= 0) ? ci[bi] : -1;
            const size_t o = ((size_t)q * s.world + s.rank) * k + j;
            for (int p = 0; p < s.world; ++p) { out_v.p[p][o] = v; out_i.p[p][o] = id; }
            if (bi >= 0) cv[bi] = -FLT_MAX;
        }
        __syncthreads();
    }
    serve_cta_done(s);
}

int topk_owned_num_chunks(long long nvalid);

// ---- threshold selection (replaces k block-wide arg-max passes over EVERY chunk, which dominated the search:
// 8 of 11.6 ms at V = 10 M, Q = 64).  tau[q] = k-th best cosine among a prefix sample of the rows is a lower
// bound of the k-th best overall, so one streaming pass that keeps `cos >= tau[q]` keeps every true winner;
// with a sample of V/32 rows about 32 k elements per query survive.
__global__ void topk_fill_kernel(float* __restrict__ cand_v, long long* __restrict__ cand_i, long long n,
                                 int* __restrict__ counts, int Q) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { cand_v[i] = -FLT_MAX; cand_i[i] = -1; }
    if (i < Q) counts[i] = 0;
}

__global__ void __launch_bounds__(256)
topk_filter_kernel(const float* __restrict__ slab_local, int nsrc, int Q, long long vown, long long nvalid,
                   const float* __restrict__ norms_owned, long long row_base, const float* __restrict__ tau_v,
                   const long long* __restrict__ tau_i, int k, float* __restrict__ cand_v,
                   long long* __restrict__ cand_i, int* __restrict__ counts, int cap) {
    const int q = blockIdx.y;
    const float tau = (tau_i[(size_t)q * k + (k - 1)] >= 0) ? tau_v[(size_t)q * k + (k - 1)] : -FLT_MAX;
    const long long base = (long long)blockIdx.x * 1024;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long i = base + j * 256 + threadIdx.x;
        if (i >= nvalid) break;
        float sc = 0.f;
        for (int r = 0; r < nsrc; ++r) sc += __ldcs(slab_local + ((size_t)r * Q + q) * vown + i);   // fixed rank order
        const float nr = __ldg(norms_owned + i);
        const float cs = nr > 0.f ? sc / nr : 0.f;
        if (cs >= tau) {
            const int pos = atomicAdd(counts + q, 1);
            if (pos < cap) { cand_v[(size_t)q * cap + pos] = cs; cand_i[(size_t)q * cap + pos] = row_base + i; }
        }
    }
}

// number of 4096-row chunks sampled for the threshold: ~1/32 of the rows, at least 16 chunks, spread evenly
int topk_sample_chunks(long long nvalid) {
    const int total = topk_owned_num_chunks(nvalid);
    int ns = total / 32;
    if (ns < 16) ns = 16;
    if (ns > total) ns = total;
    return ns;
}

// sample top-k -> tau; fill; filter.  cand_s_* : [Q, topk_sample_chunks * k] scratch; tau_* : [Q, k]; cand_f_* : [Q, cap]
void launch_topk_select(const float* slab_local, int nsrc, int Q, long long vown, long long nvalid,
                        const float* norms_owned, long long row_base, int k, float* cand_s_v, long long* cand_s_i,
                        float* tau_v, long long* tau_i, float* cand_f_v, long long* cand_f_i, int* counts, int cap,
                        cudaStream_t st) {
    const int total = topk_owned_num_chunks(nvalid);
    const int nchunks = topk_sample_chunks(nvalid);
    if (nchunks > 0) {
        dim3 grid(nchunks, Q);
        topk_owned_stage1_kernel<<<grid, TKO_THREADS, 0, st>>>(slab_local, nsrc, Q, vown, nvalid, norms_owned, row_base,
                                                               k, cand_s_v, cand_s_i, nchunks, total / nchunks);
    }
    launch_topk_merge(cand_s_v, cand_s_i, nchunks * k, Q, k, tau_v, tau_i, st);
    const long long nfill = (long long)Q * cap;
    topk_fill_kernel<<<(unsigned)((nfill + 255) / 256), 256, 0, st>>>(cand_f_v, cand_f_i, nfill, counts, Q);
    if (nvalid > 0) {
        dim3 grid((unsigned)((nvalid + 1023) / 1024), Q);
        topk_filter_kernel<<<grid, 256, 0, st>>>(slab_local, nsrc, Q, vown, nvalid, norms_owned, row_base, tau_v, tau_i,
                                                 k, cand_f_v, cand_f_i, counts, cap);
    }
}

// push the rank's k winners out of an arbitrary candidate list [Q, ncand]
void launch_topk_merge_push(float* cand_v, const long long* cand_i, int ncand, int Q, int k, const PeerPtrs& out_v,
                            const PeerIdx& out_i, const ServeSync& s, cudaStream_t st) {
    topk_merge_push_kernel<<<Q, TKO_THREADS, 0, st>>>(cand_v, cand_i, ncand, k, out_v, out_i, s);
}

// every chunk ranked by k arg-max passes (exact for any data; the fallback when the filter overflows)
void launch_topk_owned_stage1(const float* slab_local, int nsrc, int Q, long long vown, long long nvalid,
                              const float* norms_owned, long long row_base, int k, float* cand_v, long long* cand_i,
                              cudaStream_t st) {
    const int nchunks = topk_owned_num_chunks(nvalid);
    if (nchunks > 0) {
        dim3 grid(nchunks, Q);
        topk_owned_stage1_kernel<<<grid, TKO_THREADS, 0, st>>>(slab_local, nsrc, Q, vown, nvalid, norms_owned,
                                                               row_base, k, cand_v, cand_i, nchunks, 1);
    }
}

int topk_owned_num_chunks(long long nvalid) { return (int)((nvalid + TKO_CHUNK - 1) / TKO_CHUNK); }

void launch_topk_owned_push(const float* slab_local, int nsrc, int Q, long long vown, long long nvalid,
                            const float* norms_owned, long long row_base, int k, float* cand_v, long long* cand_i,
                            const PeerPtrs& out_v, const PeerIdx& out_i, const ServeSync& s, cudaStream_t st) {
    const int nchunks = topk_owned_num_chunks(nvalid);
    if (nchunks > 0) {
        dim3 grid(nchunks, Q);
        topk_owned_stage1_kernel<<<grid, TKO_THREADS, 0, st>>>(slab_local, nsrc, Q, vown, nvalid, norms_owned,
                                                               row_base, k, cand_v, cand_i, nchunks, 1);
    }
    topk_merge_push_kernel<<<Q, TKO_THREADS, 0, st>>>(cand_v, cand_i, nchunks * k, k, out_v, out_i, s);
}

// ------------------------------------------------------------------ small all-gather: n floats of every rank -> [S, n]
__global__ void push_block_kernel(const float* __restrict__ src, long long n, PeerPtrs dst, ServeSync s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = src[i];
        const size_t o = (size_t)s.rank * n + (size_t)i;
        for (int p = 0; p < s.world; ++p) dst.p[p][o] = v;
    }
    serve_cta_done(s);
}

void launch_push_block(const float* src, long long n, const PeerPtrs& dst, const ServeSync& s, cudaStream_t st) {
    unsigned grid = (unsigned)((n + 255) / 256);
    if (grid == 0) grid = 1;
    push_block_kernel<<<grid, 256, 0, st>>>(src, n, dst, s);
}

}  // namespace gw2v
