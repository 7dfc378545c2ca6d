// Nearest-neighbour search with the candidate selection fused into the tcgen05 epilogue.
//
// findSynonyms in the reference = one server-side sgemv per query (`multiply`, MLLIB:598 [G]) that returns all V
// scores to the driver, which then divides by the norms and scans for the top `num` (MLLIB:600-617).  The round-1
// GPU path kept that shape: a [Q, V] score matrix written to HBM (and, over column shards, pushed across NVLink)
// and read back by a top-k kernel -- at Q = 64 the V x Q round trip cost as much as the sweep of the matrix itself
// and the sharded version scaled negatively.
//
// Here the score matrix never exists.  The kernel streams a row-major [rows, Kp] matrix (the single-GPU syn0, or
// this rank's row shard of the serving replica, ops/serving.py) through TMA once, computes the Q dot products per
// row on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM, double buffered) and its epilogue
// turns them into cosines (x 1/|row|) and keeps only those that reach the query's threshold:
//     hit -> slot = atomicAdd(count[q]);  cand[q][slot] = row
// The thresholds come from the same kernel in `dense` mode on every S-th row (a strided tensor map: 1/S of the
// bytes): the k-th best cosine of a subset is a lower bound of the k-th best overall, so every true top-k row
// passes; about k*S rows per query do, whatever V is.  The survivors are re-scored in exact fp32 by
// nn_rerank_kernel, so tf32 rounding never reaches a reported similarity (the caller subtracts a margin that
// covers it from the thresholds).
//
// Pipeline (persistent, one CTA per SM, 6 warps): warp 0 TMA producer (SWIZZLE_128B boxes of 128 rows x 32 floats
// and Q x 32 floats into a smem ring), warp 1 MMA issuer (one elected lane, 4 x K=8 MMAs per k-block), warps 2-5
// epilogue (tcgen05.ld 32x32b, one matrix row per thread).
#include "nn_tc.h"
#include "serve_common.cuh"
#include "tc_common.cuh"

namespace gw2v {

namespace {

using namespace tc;

constexpr int SEL_BM = 128;
constexpr int SEL_BK = 32;
constexpr int SEL_THREADS = 192;
constexpr int SEL_A_BYTES = SEL_BM * SEL_BK * 4;

struct SelLayout { int stages; int b_bytes; size_t total; };
__host__ __device__ inline SelLayout sel_layout(int BN) {
    SelLayout l;
    l.b_bytes = BN * SEL_BK * 4;
    const int per = SEL_A_BYTES + l.b_bytes;
    const int st = (200 * 1024) / per;
    l.stages = st > 8 ? 8 : (st < 2 ? 2 : st);
    l.total = (size_t)l.stages * per + 1024 + 256 + 1024 /* thresholds */;
    return l;
}

struct SelArgs {
    long long rows;               // matrix rows seen through the tensor map
    int Q, K;                     // queries, padded columns (multiple of 32)
    int num_tiles;
    const float* inv_norm;        // 1/|row| (0 for zero rows), indexed by row * inv_stride
    long long inv_stride;
    int dense;                    // 1: out[q * ld + row] = cosine;  0: threshold select
    float* out; long long ld;
    const float* thr;             // [Q] thresholds
    int* cand;                    // [Q, cap] local row ids
    int* count;                   // [Q]
    int cap;
};

template <int BN>
__global__ void __launch_bounds__(SEL_THREADS, 1)
nn_select_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SelArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const SelLayout L = sel_layout(BN);
    const int STAGES = L.stages;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smA = base;
    uint8_t* smB = base + (size_t)STAGES * SEL_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)STAGES * L.b_bytes);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + 8;
    uint64_t* tfull = bars + 16;           // [2]
    uint64_t* tempty = bars + 18;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* thr_s = reinterpret_cast<float*>(bars + 32);          // [BN]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int KB = a.K / SEL_BK;
    constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

    if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 1) {
        if (elect_one()) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
            for (int x = 0; x < 2; ++x) { mbar_init(tfull + x, 1); mbar_init(tempty + x, 4); }
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS>(tmem_slot);
    }
    if (!a.dense)
        for (int q = threadIdx.x; q < BN; q += SEL_THREADS) thr_s[q] = q < a.Q ? a.thr[q] : 3.0e38f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1, 1);
                    mbar_expect_tx(full + stage, (uint32_t)(SEL_A_BYTES + L.b_bytes));
                    tma_load_2d(smA + (size_t)stage * SEL_A_BYTES, &tmA, kb * SEL_BK, tile * SEL_BM, full + stage);
                    tma_load_2d(smB + (size_t)stage * L.b_bytes, &tmB, kb * SEL_BK, 0, full + stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_tf32(SEL_BM, BN);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            mbar_wait(tempty + acc, acc_phase ^ 1, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(full + stage, phase, 3);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smA + (size_t)stage * SEL_A_BYTES));
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smB + (size_t)stage * L.b_bytes));
#pragma unroll
                    for (int k = 0; k < SEL_BK / 8; ++k)
                        umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                  (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(empty + stage);
                    if (kb == KB - 1) umma_commit(tfull + acc);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        const int quad = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const long long row = (long long)tile * SEL_BM + quad * 32 + lane;
            const bool rvalid = row < a.rows;
            const float inv = rvalid ? __ldg(a.inv_norm + row * a.inv_stride) : 0.f;     // in flight during the wait
            mbar_wait(tfull + acc, acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + (uint32_t)c0, r);
                tmem_ld_wait();
                if (rvalid) {
                    if (a.dense) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < a.Q) a.out[(size_t)(c0 + j) * (size_t)a.ld + (size_t)row] = __uint_as_float(r[j]) * inv;
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float cosv = __uint_as_float(r[j]) * inv;
                            if (cosv >= thr_s[c0 + j]) {                 // rare: ~k*S hits per query in the whole sweep
                                const int slot = atomicAdd(a.count + c0 + j, 1);
                                if (slot < a.cap) a.cand[(size_t)(c0 + j) * a.cap + slot] = (int)row;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int BN>
int launch_sel(const float* mat, long long rows, int K, long long pitch, const float* qpad, const SelArgs& a, int sms,
               cudaStream_t s) {
    CUtensorMap tmA, tmB;
    if (!make_tensormap_f32(&tmA, mat, (uint64_t)rows, (uint64_t)K, (uint64_t)pitch, SEL_BK, SEL_BM)) return 2;
    if (!make_tensormap_f32(&tmB, qpad, (uint64_t)BN, (uint64_t)K, (uint64_t)K, SEL_BK, BN)) return 2;
    const SelLayout L = sel_layout(BN);
    cudaFuncSetAttribute(nn_select_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    const int grid = a.num_tiles < sms ? a.num_tiles : sms;
    nn_select_kernel<BN><<<grid, SEL_THREADS, L.total, s>>>(tmA, tmB, a);
    return 0;
}

// exact fp32 re-score of the selected rows: one warp per candidate slot
__global__ void __launch_bounds__(256)
nn_rerank_kernel(const float* __restrict__ mat, int K, const float* __restrict__ inv_norm, const float* __restrict__ qpad,
                 const int* __restrict__ cand, const int* __restrict__ count, int cap, long long row_base,
                 float* __restrict__ out_v, long long* __restrict__ out_i) {
    const int q = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int n = min(count[q], cap);
    const float* qv = qpad + (size_t)q * K;
    for (int slot = blockIdx.x * 8 + (threadIdx.x >> 5); slot < cap; slot += gridDim.x * 8) {
        float v = -3.0e38f;
        long long idx = -1;
        if (slot < n) {
            const int row = cand[(size_t)q * cap + slot];
            const float* m = mat + (size_t)row * K;
            float acc = 0.f;
            for (int c = lane * 4; c < K; c += 128) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(m + c));
                const float4 y = __ldg(reinterpret_cast<const float4*>(qv + c));
                acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            v = acc * __ldg(inv_norm + row);
            idx = row_base + row;
        }
        if (lane == 0) { out_v[(size_t)q * cap + slot] = v; out_i[(size_t)q * cap + slot] = idx; }
    }
}

// one rank's column block -> the owners' row shards of the serving replica (peer stores over NVLink)
__global__ void __launch_bounds__(256)
rowshard_push_kernel(const float* __restrict__ syn0, long long V, int K, PeerPtrs dst, long long vown, int ldr,
                     int col0, ServeSync sync) {
    const int k4 = K / 4;
    const long long total = V * k4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / k4;
        const int c = (int)(i - row * k4) * 4;
        const int owner = (int)(row / vown);
        const float4 v = __ldg(reinterpret_cast<const float4*>(syn0 + (size_t)row * K + c));
        *reinterpret_cast<float4*>(dst.p[owner] + (size_t)(row - (long long)owner * vown) * ldr + col0 + c) = v;
    }
    serve_cta_done(sync);
}

}  // namespace

bool nn_select_supported(int K, int Q) { return K >= SEL_BK && K % SEL_BK == 0 && Q >= 1 && Q <= 256; }

// dense = 1: out[q, row] = cosine of every row seen through (rows, pitch);  dense = 0: threshold select
int launch_nn_select(const float* mat, long long rows, int K, long long pitch, const float* inv_norm, long long inv_stride,
                     const float* qpad, int Q, int dense, float* out, long long ld, const float* thr, int* cand, int* count,
                     int cap, int sms, cudaStream_t stream) {
    if (!nn_select_supported(K, Q) || rows <= 0) return 1;
    SelArgs a{};
    a.rows = rows; a.Q = Q; a.K = K;
    a.num_tiles = (int)((rows + SEL_BM - 1) / SEL_BM);
    a.inv_norm = inv_norm; a.inv_stride = inv_stride;
    a.dense = dense; a.out = out; a.ld = ld; a.thr = thr; a.cand = cand; a.count = count; a.cap = cap;
    switch (scores_tc_padded_queries(Q)) {
        case 16: return launch_sel<16>(mat, rows, K, pitch, qpad, a, sms, stream);
        case 32: return launch_sel<32>(mat, rows, K, pitch, qpad, a, sms, stream);
        case 64: return launch_sel<64>(mat, rows, K, pitch, qpad, a, sms, stream);
        case 128: return launch_sel<128>(mat, rows, K, pitch, qpad, a, sms, stream);
        default: return launch_sel<256>(mat, rows, K, pitch, qpad, a, sms, stream);
    }
}

void launch_nn_rerank(const float* mat, int K, const float* inv_norm, const float* qpad, int Q, const int* cand,
                      const int* count, int cap, long long row_base, float* out_v, long long* out_i, cudaStream_t stream) {
    dim3 grid((unsigned)((cap + 7) / 8 < 64 ? (cap + 7) / 8 : 64), (unsigned)Q);
    nn_rerank_kernel<<<grid, 256, 0, stream>>>(mat, K, inv_norm, qpad, cand, count, cap, row_base, out_v, out_i);
}

void launch_rowshard_push(const float* syn0, long long V, int K, const PeerPtrs& dst, long long vown, int ldr, int col0,
                          const ServeSync& sync, int sms, cudaStream_t stream) {
    rowshard_push_kernel<<<sms * 8, 256, 0, stream>>>(syn0, V, K, dst, vown, ldr, col0, sync);
}

}  // namespace gw2v
