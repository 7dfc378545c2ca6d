// nn_scores on the 5th-generation tensor cores: out[q, v] = sum_k syn0[v, k] * Q[q, k]
//
// This is the "multiply" of findSynonyms (MLLIB:598, server-side sgemv per query [G]) for a BATCH
// of queries: the shard's [V x K] matrix is streamed from HBM exactly once per batch while
// tcgen05.mma (kind::tf32, fp32 accumulate in TMEM) computes all Q scores per row.  On CUDA cores
// the same sweep is compute bound beyond ~8 queries; on tcgen05 it stays HBM bound up to Q = 256.
//
// Structure (persistent, one CTA per SM, warp specialised):
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (SWIZZLE_128B) of a [128 rows x 32 floats] A
//                              block and the matching [BN x 32] query block -> smem ring, mbarrier tx
//   warp 1      MMA issuer     one elected lane: 4 x tcgen05.mma (K=8 each) per k-block into a
//                              [128 x BN] fp32 TMEM accumulator (double buffered); tcgen05.commit
//                              releases smem stages / publishes the accumulator
//   warps 2..5  epilogue       tcgen05.ld 32x32b -> registers -> coalesced global stores
// A = syn0 rows (M, K-major), B = queries (N, K-major): both operands are exactly their natural
// row-major layout, no transposes anywhere.
#include "nn_tc.h"
#include "serve_common.cuh"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <mutex>

namespace gw2v {

namespace {

constexpr int BM = 128;             // rows per tile (UMMA M)
constexpr int BK = 32;              // floats per k-block: 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 8;           // tf32
constexpr int NUM_THREADS = 192;    // 6 warps
constexpr int A_STAGE_BYTES = BM * BK * 4;          // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (SM100 UMMA):
//   start address >> 4 | LBO (ignored for swizzled K-major; 1) << 16 | SBO = 1024 B (8 rows x 128 B) >> 4 << 32
//   | version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

struct SmemLayout {
    int stages;
    int b_stage_bytes;
    size_t total;
};

__host__ __device__ inline SmemLayout smem_layout(int BN) {
    SmemLayout l;
    l.b_stage_bytes = BN * BK * 4;
    int per_stage = A_STAGE_BYTES + l.b_stage_bytes;
    int st = (200 * 1024) / per_stage;
    l.stages = st > 8 ? 8 : (st < 2 ? 2 : st);
    l.total = (size_t)l.stages * per_stage + 1024 /*align*/ + 256 /*barriers*/;
    return l;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
scores_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const PeerPtrs outp, const long long vown, const ServeSync sync, long long V, int Q, int K,
                 int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    const SmemLayout L = smem_layout(BN);
    const int STAGES = L.stages;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smA = base;                                             // STAGES x 16 KB (1024-aligned)
    uint8_t* smB = base + (size_t)STAGES * A_STAGE_BYTES;            // STAGES x BN*128 B (1024-aligned: BN*128 % 1024 == 0 for BN % 8 == 0)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)STAGES * L.b_stage_bytes);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + 8;            // [STAGES]
    uint64_t* tfull = bars + 16;           // [2]
    uint64_t* tempty = bars + 18;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int KB = K / BK;
    constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 4); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    mbar_expect_tx(full + stage, (uint32_t)(A_STAGE_BYTES + L.b_stage_bytes));
                    tma_load_2d(smA + (size_t)stage * A_STAGE_BYTES, &tmA, kb * BK, tile * BM, full + stage);
                    tma_load_2d(smB + (size_t)stage * L.b_stage_bytes, &tmB, kb * BK, 0, full + stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_tf32(BM, BN);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(tempty + acc, acc_phase ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(full + stage, phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smA + (size_t)stage * A_STAGE_BYTES));
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smB + (size_t)stage * L.b_stage_bytes));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 bytes (8 tf32) inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
                        umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                  (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty + stage);                       // frees the smem stage when the MMAs retire
                    if (kb == KB - 1) umma_commit(tfull + acc);       // accumulator complete
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quad = warp & 3;                                    // TMEM lane quadrant this warp may access
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(tfull + acc, acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long row = (long long)tile * BM + quad * 32 + lane;
            // reduce-scatter epilogue: row v belongs to rank v / vown; the partial score goes straight into that
            // rank's slab [src rank][Q][vown] over NVLink (single shard: owner 0, vown >= V, plain [Q, V] output)
            const int owner = (int)(row / vown);
            float* __restrict__ out = outp.p[row < V ? owner : 0] + (size_t)sync.rank * (size_t)Q * (size_t)vown +
                                      (size_t)(row - (long long)owner * vown);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + (uint32_t)c0, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < V) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < Q) out[(size_t)(c0 + j) * (size_t)vown] = __uint_as_float(r[j]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
    if (sync.world > 1) serve_cta_done(sync);       // last CTA publishes the sequence number to every rank
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

bool make_map(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    auto fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN>
int launch_bn(const float* syn0, long long V, int K, const float* qpad, int Q, const PeerPtrs& outp, long long vown,
              const ServeSync& sync, int sms, cudaStream_t s) {
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, syn0, (uint64_t)V, (uint64_t)K, BM)) return 2;
    if (!make_map(&tmB, qpad, (uint64_t)BN, (uint64_t)K, BN)) return 2;
    const int num_tiles = (int)((V + BM - 1) / BM);
    SmemLayout L = smem_layout(BN);
    cudaFuncSetAttribute(scores_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    int grid = num_tiles < sms ? num_tiles : sms;
    scores_tc_kernel<BN><<<grid, NUM_THREADS, L.total, s>>>(tmA, tmB, outp, vown, sync, V, Q, K, num_tiles);
    return 0;
}

}  // namespace

int scores_tc_padded_queries(int Q) {
    if (Q <= 16) return 16;
    if (Q <= 32) return 32;
    if (Q <= 64) return 64;
    if (Q <= 128) return 128;
    return 256;
}

bool scores_tc_supported(int K, int Q) { return K >= BK && K % BK == 0 && Q >= 1 && Q <= 256; }

// qpad: [scores_tc_padded_queries(Q), K] row-major, zero padded
int launch_scores_tc_push(const float* syn0, long long V, int K, const float* qpad, int Q, const PeerPtrs& slab,
                          long long vown, const ServeSync& sync, int sms, cudaStream_t stream) {
    if (!scores_tc_supported(K, Q) || V <= 0) return 1;
    switch (scores_tc_padded_queries(Q)) {
        case 16: return launch_bn<16>(syn0, V, K, qpad, Q, slab, vown, sync, sms, stream);
        case 32: return launch_bn<32>(syn0, V, K, qpad, Q, slab, vown, sync, sms, stream);
        case 64: return launch_bn<64>(syn0, V, K, qpad, Q, slab, vown, sync, sms, stream);
        case 128: return launch_bn<128>(syn0, V, K, qpad, Q, slab, vown, sync, sms, stream);
        default: return launch_bn<256>(syn0, V, K, qpad, Q, slab, vown, sync, sms, stream);
    }
}

// single shard: plain [Q, V] output (owner 0 for every row, no signal)
int launch_scores_tc(const float* syn0, long long V, int K, const float* qpad, int Q, float* out, int sms,
                     cudaStream_t stream) {
    PeerPtrs o{};
    o.p[0] = out;
    ServeSync sync{};
    sync.world = 1; sync.rank = 0;
    return launch_scores_tc_push(syn0, V, K, qpad, Q, o, V, sync, sms, stream);
}

}  // namespace gw2v
