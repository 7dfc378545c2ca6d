// sgns_fused_pipe: the single-shard SGNS step as a per-warp TMA pipeline (v2).
//
// v1 (sgns_kernels.cu) stages the 1+n rows of a pair in registers: 189 registers/thread at K=512,
// one CTA of 8 warps per SM, every load latency exposed (ncu: warps active 11 %, DRAM 44 %).
// Here rows never touch registers on their way in or out:
//
//   generate  the warp expands centres into a ring of pair descriptors {centre, context, negatives}
//             (window + negatives from Philox, bit-identical to v1 / the oracle)
//   issue     for pair p+S-1: one cp.async.bulk (TMA, UBLKCP) per row, global -> this warp's shared
//             memory stage, completion on an mbarrier (tx bytes)
//   compute   for pair p: wait the mbarrier, dots from shared memory, sigmoid/alpha, du accumulation,
//             overwrite each row IN PLACE with g*u, fence.proxy.async, one
//             cp.reduce.async.bulk.add.f32 (TMA reduce) per row back to global (the scatter-add of
//             Glint's `adjust`, MLLIB:425) - no per-lane RED instructions at all.
//
// Each warp owns S stages x (n+2) rows (context, n negatives, centre row for the first pair of a
// centre; the same slot carries du for the last pair), so loads of later pairs are in flight while
// the current one computes.  Warps are independent: no __syncthreads in the steady state.
#include "pipe_common.cuh"
#include "sgns_params.h"

namespace gw2v {

constexpr int V2_RING = 32;          // pair descriptors per warp
constexpr int V2_ENTRY = 20;         // ints per descriptor: wtok, ctok, flags, pad, negs[<=16]
constexpr int V2_MAXNEG = 16;
constexpr int V2_RB = 8;

__device__ __forceinline__ uint32_t smem_a(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void v2_bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_a(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_a(bar)) : "memory");
}
__device__ __forceinline__ void v2_bulk_reduce_add(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_a(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void v2_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void v2_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void v2_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void v2_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void v2_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_a(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void v2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_a(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void v2_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "V2_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni V2_DONE;\n\t"
        "bra.uni V2_WAIT;\n\t"
        "V2_DONE:\n\t"
        "}\n" ::"r"(smem_a(bar)), "r"(parity) : "memory");
}

template <int VEC>
__device__ __forceinline__ void lds_vec(const float* p, float (&out)[VEC]) {
    if constexpr (VEC == 4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else {
        float2 v = *reinterpret_cast<const float2*>(p);
        out[0] = v.x; out[1] = v.y;
    }
}
template <int VEC>
__device__ __forceinline__ void sts_vec(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}

__device__ __forceinline__ void v2_window_bounds(const SgnsParams& p, int i, int T, int& lo, int& hi) {
    uint4 r = rand4(p.seed_lo, p.seed_hi, stream_word(STREAM_WINDOW, p.iteration), p.pos0 + (unsigned long long)i, 0u);
    int b = (int)__umulhi(r.x, (uint32_t)p.window);
    if (p.window_mode == 0) { lo = -b; hi = b - 1; }
    else { int rad = p.window - b; lo = -rad; hi = rad; }
    lo = max(lo, -i);
    hi = min(hi, T - 1 - i);
}

template <int VEC, int CHUNKS>
__global__ void __launch_bounds__((CHUNKS >= 3) ? 256 : 512)
sgns_fused_pipe_kernel(const SgnsParams p, const int nstage, const int stage_floats, const int warp_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp_cta = blockDim.x >> 5;
    unsigned char* wbase = smem_raw + (size_t)warp * warp_bytes;
    float* stages = reinterpret_cast<float*>(wbase);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + (size_t)nstage * stage_floats * 4);
    int* ring = reinterpret_cast<int*>(wbase + (size_t)nstage * stage_floats * 4 + 64);

    const int K = p.K;
    const int n = p.negatives;
    const int ncalls = (n + 1) >> 1;
    const uint32_t row_bytes = (uint32_t)K * 4u;
    const int T = *p.n_tokens;
    const int maxctx = 2 * p.window;
    const uint32_t sw_neg = stream_word(STREAM_NEG, p.iteration);

    if (lane == 0) {
        for (int s = 0; s < nstage; ++s) v2_mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    bool act[CHUNKS];
    int coff[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) { coff[c] = (c * 32 + lane) * VEC; act[c] = coff[c] < K; }

    const int n_warps = gridDim.x * nwarp_cta;
    int gen_i = blockIdx.x * nwarp_cta + warp;
    int head = 0, issued = 0, done = 0;
    float u[CHUNKS][VEC], du[CHUNKS][VEC];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) { u[c][e] = 0.f; du[c][e] = 0.f; }
    float loss = 0.f, maxdot = 0.f;
    unsigned pairs = 0;

    while (true) {
        // ------------------------------------------------------------ (1) generate pair descriptors
        while (gen_i < T && (V2_RING - (head - done)) >= maxctx) {
            const int i = gen_i;
            gen_i += n_warps;
            int lo, hi;
            v2_window_bounds(p, i, T, lo, hi);
            if (hi < lo) continue;
            const int span = hi - lo + 1;
            const int wi = __ldg(p.tokens + i);
            const int sid = __ldg(p.sent_id + i);
            bool valid = false;
            int ctok = 0;
            if (lane < span) {
                const int off = lo + lane;
                if (off != 0) {
                    valid = __ldg(p.sent_id + i + off) == sid;
                    if (valid) ctok = __ldg(p.tokens + i + off);
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, valid);
            const int np = __popc(mask);
            if (np == 0) continue;
            if (valid) {
                const int rank = __popc(mask & ((1u << lane) - 1u));
                int* e = ring + ((head + rank) % V2_RING) * V2_ENTRY;
                e[0] = wi; e[1] = ctok;
                e[2] = (rank == 0 ? 1 : 0) | (rank == np - 1 ? 2 : 0);
            }
            const int total = span * ncalls;
            const unsigned long long pos = p.pos0 + (unsigned long long)i;
            for (int item = lane; item < total; item += 32) {
                const int q = item / ncalls, c = item - q * ncalls;
                if ((mask >> q) & 1u) {
                    const int slot = lo + q + p.window;
                    uint4 r = rand4(p.seed_lo, p.seed_hi, sw_neg, pos, (uint32_t)(slot * ncalls + c));
                    const int rq = __popc(mask & ((1u << q) - 1u));
                    int* e = ring + ((head + rq) % V2_RING) * V2_ENTRY;
                    e[4 + 2 * c] = alias_sample(p.alias, (uint32_t)p.vocab, r.x, r.y);
                    if (2 * c + 1 < n) e[4 + 2 * c + 1] = alias_sample(p.alias, (uint32_t)p.vocab, r.z, r.w);
                }
            }
            head += np;
        }
        __syncwarp();

        // ------------------------------------------------------------ (2) issue loads (TMA bulk copies)
        while (issued < head && issued - done < nstage) {
            const int s = issued % nstage;
            const int* e = ring + (issued % V2_RING) * V2_ENTRY;
            float* stage = stages + (size_t)s * stage_floats;
            if (issued >= nstage) v2_wait_read0();         // the reduces that read this stage have drained it
            const int wtok = e[0], ctok = e[1], flags = e[2];
            if (!(p.debug & 4)) {
                if (lane == 0) {
                    int nact = 1 + (flags & 1);
                    for (int k = 0; k < n; ++k) nact += (e[4 + k] != ctok) ? 1 : 0;
                    v2_mbar_expect_tx(bars + s, (uint32_t)nact * row_bytes);
                }
                __syncwarp();
                if (lane <= n) {
                    const int row = (lane == 0) ? ctok : e[4 + lane - 1];
                    if (lane == 0 || row != ctok)
                        v2_bulk_load(stage + (size_t)lane * K, p.syn1 + (size_t)row * K, row_bytes, bars + s);
                } else if (lane == n + 1 && (flags & 1)) {
                    v2_bulk_load(stage + (size_t)(n + 1) * K, p.syn0 + (size_t)wtok * K, row_bytes, bars + s);
                }
            }
            ++issued;
        }
        if (done == head) {
            if (gen_i >= T) break;
            continue;
        }

        // ------------------------------------------------------------ (3) compute pair `done`
        {
            const int s = done % nstage;
            const int* e = ring + (done % V2_RING) * V2_ENTRY;
            float* stage = stages + (size_t)s * stage_floats;
            const int wtok = e[0], ctok = e[1], flags = e[2];
            if (!(p.debug & 4)) v2_mbar_wait(bars + s, (uint32_t)((done / nstage) & 1));
            if (flags & 1) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                    for (int el = 0; el < VEC; ++el) { u[c][el] = 0.f; du[c][el] = 0.f; }
                    if (act[c] && !(p.debug & 4)) lds_vec<VEC>(stage + (size_t)(n + 1) * K + coff[c], u[c]);
                }
            }
            ++pairs;
            for (int rb = 0; rb <= n; rb += V2_RB) {
                bool ract[V2_RB];
                float v[V2_RB][CHUNKS][VEC];
#pragma unroll
                for (int r = 0; r < V2_RB; ++r) {
                    const int k = rb + r;
                    ract[r] = (k <= n) && (k == 0 || e[4 + k - 1] != ctok);
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                        for (int el = 0; el < VEC; ++el) v[r][c][el] = 0.f;
                        if (ract[r] && act[c] && !(p.debug & 4)) lds_vec<VEC>(stage + (size_t)k * K + coff[c], v[r][c]);
                    }
                }
                float f[V2_RB];
#pragma unroll
                for (int r = 0; r < V2_RB; ++r) {
                    float sacc = 0.f;
                    if (ract[r]) {
#pragma unroll
                        for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
                            for (int el = 0; el < VEC; ++el) sacc = fmaf(u[c][el], v[r][c][el], sacc);
                    }
                    f[r] = sacc;
                }
                // 8 dots reduced together; the lane owning row r computes its coefficient (and loss) once
                const float ftot = reduce8_transposed(f, lane);
                const int myrow = rb + row_of_lane(lane);
                const float mylabel = (myrow == 0) ? 1.f : 0.f;
                const float gmine = sgns_coeff(ftot, mylabel, p.alpha, p.max_grad);
                const bool myact = (myrow <= n) && (myrow == 0 || e[4 + myrow - 1] != ctok);
                if (p.compute_loss && myact && lane == lane_of_row(row_of_lane(lane))) {
                    loss += softplus_clipped(mylabel > 0.5f ? -ftot : ftot);
                    maxdot = fmaxf(maxdot, fabsf(ftot));
                }
#pragma unroll
                for (int r = 0; r < V2_RB; ++r) {
                    const float g = __shfl_sync(0xffffffffu, gmine, lane_of_row(r));
                    if (!ract[r]) continue;
                    const int k = rb + r;
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
                        if (!act[c]) continue;
                        float gu[VEC];
#pragma unroll
                        for (int el = 0; el < VEC; ++el) {
                            du[c][el] = fmaf(g, v[r][c][el], du[c][el]);
                            gu[el] = g * u[c][el];
                        }
                        sts_vec<VEC>(stage + (size_t)k * K + coff[c], gu);     // row k now holds g*u
                    }
                }
            }
            if (flags & 2) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    if (act[c]) sts_vec<VEC>(stage + (size_t)(n + 1) * K + coff[c], du[c]);
            }
            v2_fence_async();            // generic-proxy writes -> visible to the async proxy (TMA)
            __syncwarp();
            if (lane <= n) {
                const int row = (lane == 0) ? ctok : e[4 + lane - 1];
                if ((lane == 0 || row != ctok) && !(p.debug & 1))
                    v2_bulk_reduce_add(p.syn1 + (size_t)row * K, stage + (size_t)lane * K, row_bytes);
            } else if (lane == n + 1 && (flags & 2) && !(p.debug & 2)) {
                v2_bulk_reduce_add(p.syn0 + (size_t)wtok * K, stage + (size_t)(n + 1) * K, row_bytes);
            }
            v2_commit();
            ++done;
        }
    }
    v2_wait_all();                       // smem must outlive the outstanding TMA reduces

    if (blockIdx.x == 0 && threadIdx.x == 0) p.stats[3] = (float)T;
    loss = warp_sum(loss);                 // per-lane partial sums (one owner lane per row)
    maxdot = warp_max(maxdot);
    if (lane == 0 && pairs) {
        atomicAdd(p.stats + 0, (float)pairs);
        if (p.compute_loss) {
            atomicAdd(p.stats + 1, loss);
            atomicMax(reinterpret_cast<int*>(p.stats + 2), __float_as_int(maxdot));
        }
    }
}

// ------------------------------------------------------------------ host side

struct PipeLayout { int warps, stages, stage_floats, warp_bytes; size_t total; };

static PipeLayout pipe_layout(int K, int negatives, size_t smem_budget) {
    PipeLayout best{0, 0, 0, 0, 0};
    const int stage_floats = (negatives + 2) * K;
    const size_t stage_bytes = (size_t)stage_floats * 4;
    const size_t fixed = 64 + (size_t)V2_RING * V2_ENTRY * 4;
    long best_score = -1;
    const int max_warps = (K > 256) ? 8 : 16;      // kernels with CHUNKS >= 3 are compiled for <= 256 threads
    for (int warps = max_warps; warps >= 2; --warps) {
        size_t per_warp = (smem_budget / warps) & ~(size_t)127;
        if (per_warp <= fixed + stage_bytes) continue;
        int stages = (int)((per_warp - fixed) / stage_bytes);
        if (stages > 8) stages = 8;
        if (stages < 2) continue;
        long score = (long)warps * (stages > 4 ? 4 : stages) * 16 + warps;
        if (score > best_score) {
            best_score = score;
            size_t wb = (fixed + (size_t)stages * stage_bytes + 127) & ~(size_t)127;
            best = PipeLayout{warps, stages, stage_floats, (int)wb, wb * warps};
        }
    }
    return best;
}

bool sgns_pipe_supported(int K, int window, int negatives) {
    if (negatives < 1 || negatives > V2_MAXNEG) return false;
    if (2 * window + 1 > 32 || 2 * window > V2_RING - 8) return false;
    if (K % 4 != 0 || K > 1024) return false;
    PipeLayout l = pipe_layout(K, negatives, 220 * 1024);
    return l.warps >= 2;
}

#define GW2V_PIPE_DISPATCH(K, CALL)                                \
    do {                                                           \
        if ((K) <= 64) { CALL(2, 1); }                             \
        else if ((K) <= 128) { CALL(4, 1); }                       \
        else if ((K) <= 256) { CALL(4, 2); }                       \
        else if ((K) <= 384) { CALL(4, 3); }                       \
        else if ((K) <= 512) { CALL(4, 4); }                       \
        else if ((K) <= 768) { CALL(4, 6); }                       \
        else { CALL(4, 8); }                                       \
    } while (0)

int sgns_pipe_grid(int K, int negatives, int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    PipeLayout l = pipe_layout(K, negatives, 220 * 1024);
    int occ = 1;
#define CALL(V, C)                                                                                          \
    do {                                                                                                    \
        cudaFuncSetAttribute(sgns_fused_pipe_kernel<V, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                             (int)l.total);                                                                 \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sgns_fused_pipe_kernel<V, C>, l.warps * 32,     \
                                                      l.total);                                             \
    } while (0)
    GW2V_PIPE_DISPATCH(K, CALL);
#undef CALL
    if (occ < 1) occ = 1;
    return sms * occ;
}

void launch_sgns_pipe(const SgnsParams& p, int grid, cudaStream_t stream) {
    PipeLayout l = pipe_layout(p.K, p.negatives, 220 * 1024);
#define CALL(V, C)                                                                                          \
    do {                                                                                                    \
        cudaFuncSetAttribute(sgns_fused_pipe_kernel<V, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                             (int)l.total);                                                                 \
        sgns_fused_pipe_kernel<V, C><<<grid, l.warps * 32, l.total, stream>>>(p, l.stages, l.stage_floats,  \
                                                                             l.warp_bytes);                \
    } while (0)
    GW2V_PIPE_DISPATCH(p.K, CALL);
#undef CALL
}

}  // namespace gw2v
