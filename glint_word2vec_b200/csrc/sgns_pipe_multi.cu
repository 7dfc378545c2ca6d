// sgns_fused_pipe_multi: column-sharded SGNS step with the partial-dot all-reduce fused into the
// per-warp TMA pipeline (world > 1).  This is the hot path of the reference -- dotprod fan-out,
// client-side sum of the shards' partial dot products, sigmoid, adjust fan-out (MLLIB:417-429,
// Glint server ops [G]) -- as ONE kernel per rank with no NCCL call:
//
//   issue   TMA bulk copies (cp.async.bulk, UBLKCP) of the pair's rows into this warp's smem stage
//   dots    partial dot products over this rank's K columns  -> smem
//   push    every few pairs ("batch"): lane j stores the batch's partials straight into rank j's
//           symmetric exchange slot over NVLink (st.global.v4 on a peer-mapped address) and then
//           publishes the batch sequence number with st.release.sys
//   ...     the warp keeps issuing loads / computing dots of LATER pairs while the batch is in flight
//           (rows stay resident in shared memory; nothing is re-gathered)
//   update  ld.acquire.sys on the peers' flags, sum the S partials in fixed rank order (all ranks
//           obtain bit-identical coefficients, so the reference's coefficient broadcast disappears),
//           g = (label - sigmoid(f)) * alpha, du += g v, rows overwritten in place with g*u,
//           cp.reduce.async.bulk.add.f32 (TMA reduce) back to this rank's shard.
//
// The schedule (when to push, when to update) is a pure function of per-warp counters, never of
// timing, so the S ranks' warps stay in lock-step on batch boundaries without any negotiation.
// Sequence numbers persist in device memory across launches and are never reset.
#include "pipe_common.cuh"
#include "sgns_params.h"
#include <cstdio>

namespace gw2v {

constexpr int M_RING = 64;           // pair descriptors per warp (must exceed lag + stages + 2*window)
constexpr int M_ENTRY = 12;          // ints per descriptor: wtok, ctok, flags, pad, negs[<=7]
constexpr int M_MAXNEG = 7;          // 1 + n <= 8 floats per pair in an exchange slot
constexpr int M_FP = 8;              // floats per pair in exchange slots
constexpr int M_G = 4;               // pairs per batch (slot = 32 floats = 128 B)
constexpr int M_RB = 8;

__device__ __forceinline__ uint32_t m_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void m_bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(m_smem(sdst)), "l"(gsrc), "r"(bytes), "r"(m_smem(bar)) : "memory");
}
__device__ __forceinline__ void m_bulk_reduce_add(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                 ::"l"(gdst), "r"(m_smem(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void m_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void m_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void m_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void m_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void m_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void m_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "M_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni M_DONE;\n\t"
        "bra.uni M_WAIT;\n\t"
        "M_DONE:\n\t"
        "}\n" ::"r"(m_smem(bar)), "r"(parity) : "memory");
}
template <int VEC>
__device__ __forceinline__ void m_lds(const float* p, float (&out)[VEC]) {
    if constexpr (VEC == 4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else {
        float2 v = *reinterpret_cast<const float2*>(p);
        out[0] = v.x; out[1] = v.y;
    }
}
template <int VEC>
__device__ __forceinline__ void m_sts(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}

struct MultiPipeArgs {
    int nstage;          // smem stages per warp
    int lag;             // pairs between the dots pass (A) and the update pass (B) of the same pair
    int stage_floats;    // (n + 2) * K
    int warp_bytes;      // smem per warp
    int nslot;           // exchange slots per (warp, source)
    uint32_t* warp_seq;  // [grid * warps] running batch sequence per warp (device memory, never reset)
};

// per-warp shared memory:  stages | mbarriers (128 B) | fpart/ftot[M_RING][8] | xsum[8][32]
//                          | item queue (M_IQ bytes) | ring[M_RING][M_ENTRY]
constexpr int M_IQ = 64;
__host__ __device__ inline size_t m_fixed_bytes() {
    return 128 + (size_t)M_RING * M_FP * 4 + 8 * 32 * 4 + M_IQ + (size_t)M_RING * M_ENTRY * 4;
}

// Every pair passes through the stage ring twice:
//   A item: TMA-load its rows, partial dots, (every M_G pairs) push the batch to all peers
//   B item: `lag` pairs later TMA-load the rows again (L2 hits), reduce the S partials, update
// The A/B interleaving is decided at issue time from per-warp counters only, recorded in a small
// item queue and replayed at compute time, so all ranks follow the identical schedule.
template <int VEC, int CHUNKS>
__global__ void __launch_bounds__((CHUNKS >= 3) ? 256 : 512)
sgns_fused_pipe_multi_kernel(const SgnsParams p, const MultiPipeArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp_cta = blockDim.x >> 5;
    const int nstage = a.nstage;
    unsigned char* wbase = smem_raw + (size_t)warp * a.warp_bytes;
    float* stages = reinterpret_cast<float*>(wbase);
    unsigned char* q = wbase + (size_t)nstage * a.stage_floats * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(q);            q += 128;
    float* fpart = reinterpret_cast<float*>(q);                 q += (size_t)M_RING * M_FP * 4;
    float* ftot = fpart;                                        // the reduced dots replace the partials in place
    float* xsum = reinterpret_cast<float*>(q);                  q += 8 * 32 * 4;
    unsigned char* iq = q;                                      q += M_IQ;
    int* ring = reinterpret_cast<int*>(q);

    const int K = p.K;
    const int n = p.negatives;
    const int ncalls = (n + 1) >> 1;
    const uint32_t row_bytes = (uint32_t)K * 4u;
    const int T = *p.n_tokens;
    const int maxctx = 2 * p.window;
    const int S = p.world;
    const int rank = p.rank;
    const int L = a.lag;
    const uint32_t sw_neg = stream_word(STREAM_NEG, p.iteration);
    const int gwarp = blockIdx.x * nwarp_cta + warp;            // identical on every rank
    const int n_warps = gridDim.x * nwarp_cta;

    if (lane == 0) {
        for (int s = 0; s < nstage; ++s) m_mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    bool act[CHUNKS];
    int coff[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) { coff[c] = (c * 32 + lane) * VEC; act[c] = coff[c] < K; }

    const size_t slot_stride = (size_t)M_G * M_FP;                               // floats per (slot, src)
    const size_t warp_x_base = (size_t)gwarp * a.nslot * S * slot_stride;
    uint32_t* my_flags = p.flags[rank] + (size_t)gwarp * S;
    uint32_t seq = a.warp_seq[gwarp];                                            // next batch to push
    const uint32_t seq0 = seq;

    int gen_i = gwarp;
    int head = 0;                         // pairs generated
    int ia = 0, ib = 0;                   // A / B items issued (pair indices)
    int ca = 0, cb = 0;                   // A / B items computed
    int items_issued = 0, items_done = 0; // positions in the item sequence
    int pushed = 0;                       // pairs whose partials have been pushed
    int nb_pushed = 0, nb_recv = 0, recv_end = 0;
    bool last_was_b = false;              // previous occupant kind of each stage is tracked via iq as well
    float ud[CHUNKS][VEC], uu[CHUNKS][VEC], du[CHUNKS][VEC];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
        for (int e = 0; e < VEC; ++e) { ud[c][e] = 0.f; uu[c][e] = 0.f; du[c][e] = 0.f; }
    float loss = 0.f, maxdot = 0.f;
    unsigned pairs = 0;
    unsigned long long wait_ns = 0;
    (void)last_was_b;

    // push pairs [pushed, ca) as one batch: lane j stores them into rank j's slot, then release-publishes seq+1
    auto push_batch = [&]() {
        __syncwarp();
        const int slot = (int)(seq % (uint32_t)a.nslot);
        if (lane < S && lane != rank) {
            float4* dst = reinterpret_cast<float4*>(p.xbuf[lane] + warp_x_base +
                                                    ((size_t)slot * S + rank) * slot_stride);
            const int cnt = ca - pushed;
            for (int g = 0; g < cnt; ++g) {
                const float4* src = reinterpret_cast<const float4*>(fpart + ((pushed + g) % M_RING) * M_FP);
                dst[g * 2 + 0] = src[0];
                dst[g * 2 + 1] = src[1];
            }
            st_release_sys(p.flags[lane] + (size_t)gwarp * S + rank, seq + 1u);
        }
        __syncwarp();
        pushed = ca;
        ++nb_pushed;
        ++seq;
    };

    while (true) {
        // ---------------------------------------------------------------- (1) generate descriptors
        while (gen_i < T && (M_RING - (head - cb)) >= maxctx) {
            const int i = gen_i;
            gen_i += n_warps;
            uint4 rw = rand4(p.seed_lo, p.seed_hi, stream_word(STREAM_WINDOW, p.iteration),
                             p.pos0 + (unsigned long long)i, 0u);
            int b = (int)__umulhi(rw.x, (uint32_t)p.window), lo, hi;
            if (p.window_mode == 0) { lo = -b; hi = b - 1; } else { int rad = p.window - b; lo = -rad; hi = rad; }
            lo = max(lo, -i);
            hi = min(hi, T - 1 - i);
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
            const int npair = __popc(mask);
            if (npair == 0) continue;
            if (valid) {
                const int rk = __popc(mask & ((1u << lane) - 1u));
                int* e = ring + ((head + rk) % M_RING) * M_ENTRY;
                e[0] = wi; e[1] = ctok;
                e[2] = (rk == 0 ? 1 : 0) | (rk == npair - 1 ? 2 : 0);
            }
            const int total = span * ncalls;
            const unsigned long long pos = p.pos0 + (unsigned long long)i;
            for (int item = lane; item < total; item += 32) {
                const int qq = item / ncalls, c = item - qq * ncalls;
                if ((mask >> qq) & 1u) {
                    const int slot = lo + qq + p.window;
                    uint4 r = rand4(p.seed_lo, p.seed_hi, sw_neg, pos, (uint32_t)(slot * ncalls + c));
                    const int rq = __popc(mask & ((1u << qq) - 1u));
                    int* e = ring + ((head + rq) % M_RING) * M_ENTRY;
                    e[4 + 2 * c] = alias_sample(p.alias, (uint32_t)p.vocab, r.x, r.y);
                    if (2 * c + 1 < n) e[4 + 2 * c + 1] = alias_sample(p.alias, (uint32_t)p.vocab, r.z, r.w);
                }
            }
            head += npair;
        }
        __syncwarp();
        const bool gen_done = gen_i >= T;

        // ---------------------------------------------------------------- (2) issue items (TMA loads)
        while (items_issued - items_done < nstage) {
            // deterministic choice of the next item: A while the dots pass is less than `lag` ahead, else B
            int kind;                                     // 0 = A, 1 = B
            if (ia < head && ia - ib < L) kind = 0;
            else if (ib < ia && (ia - ib >= L || (gen_done && ia == head))) kind = 1;
            else break;                                   // nothing issuable right now
            const int pidx = kind ? ib : ia;
            const int s = items_issued % nstage;
            const int* e = ring + (pidx % M_RING) * M_ENTRY;
            float* stage = stages + (size_t)s * a.stage_floats;
            if (items_issued >= nstage && iq[(items_issued - nstage) % M_IQ]) m_wait_read0();   // previous occupant was a B item
            const int wtok = e[0], ctok = e[1], flags = e[2];
            if (lane == 0) {
                int nact = 1 + (flags & 1);
                for (int k = 0; k < n; ++k) nact += (e[4 + k] != ctok) ? 1 : 0;
                m_mbar_expect_tx(bars + s, (uint32_t)nact * row_bytes);
                iq[items_issued % M_IQ] = (unsigned char)kind;
            }
            __syncwarp();
            if (lane <= n) {
                const int row = (lane == 0) ? ctok : e[4 + lane - 1];
                if (lane == 0 || row != ctok)
                    m_bulk_load(stage + (size_t)lane * K, p.syn1 + (size_t)row * K, row_bytes, bars + s);
            } else if (lane == n + 1 && (flags & 1)) {
                m_bulk_load(stage + (size_t)(n + 1) * K, p.syn0 + (size_t)wtok * K, row_bytes, bars + s);
            }
            if (kind) ++ib; else ++ia;
            ++items_issued;
        }

        // the stream ended after the last dots were computed: close the open batch (same decision on every rank)
        if (gen_done && ia == head && ca == head && ca > pushed) push_batch();

        if (items_done == items_issued) {
            if (gen_done && cb == head) break;
            continue;
        }

        // ---------------------------------------------------------------- (3) compute the next item
        const int s = items_done % nstage;
        const int kind = iq[items_done % M_IQ];
        float* stage = stages + (size_t)s * a.stage_floats;
        m_mbar_wait(bars + s, (uint32_t)((items_done / nstage) & 1));
        if (kind == 0) {
            // ---------------- A: partial dots of pair `ca`
            const int* e = ring + (ca % M_RING) * M_ENTRY;
            const int ctok = e[1], flags = e[2];
            if (flags & 1) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                    for (int el = 0; el < VEC; ++el) ud[c][el] = 0.f;
                    if (act[c]) m_lds<VEC>(stage + (size_t)(n + 1) * K + coff[c], ud[c]);
                }
            }
            float f[M_RB];
#pragma unroll
            for (int r = 0; r < M_RB; ++r) {
                float sacc = 0.f;
                const bool ra = (r <= n) && (r == 0 || e[4 + r - 1] != ctok);
                if (ra) {
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
                        if (!act[c]) continue;
                        float v[VEC];
                        m_lds<VEC>(stage + (size_t)r * K + coff[c], v);
#pragma unroll
                        for (int el = 0; el < VEC; ++el) sacc = fmaf(ud[c][el], v[el], sacc);
                    }
                }
                f[r] = sacc;
            }
            {
                const float tot = reduce8_transposed(f, lane);          // lane_of_row(r) holds the total of row r
                const int myrow = row_of_lane(lane);
                if (lane == lane_of_row(myrow)) fpart[(ca % M_RING) * M_FP + myrow] = tot;
            }
            ++ca;
            ++items_done;
            // push when a batch is full, or when the stream ends with an open batch
            if (ca - pushed == M_G || (gen_done && ca == head)) push_batch();
            continue;
        }
        // ---------------- B: reduce + update pair `cb`
        if (cb >= recv_end) {
            // receive the batch that starts at pair `cb`: pairs [cb, min(cb + M_G, pushed-at-that-time))
            // batch boundaries are multiples of M_G except for the final partial batch
            const int b_lo = cb;
            const uint32_t bseq = seq0 + (uint32_t)nb_recv;
            const int slot = (int)(bseq % (uint32_t)a.nslot);
            if (lane < S && lane != rank) {
                unsigned long long t0 = p.timing ? globaltimer_ns() : 0ull;
                uint32_t spins = 0;
                volatile uint32_t* fl = my_flags + lane;
                while ((int32_t)(*fl - (bseq + 1u)) < 0) {
                    if ((++spins & 0x3FFFu) == 0) {
                        if (t0 == 0ull) t0 = globaltimer_ns();
                        if (globaltimer_ns() - t0 > 20000000000ull) {
                            printf("[gw2v] rank %d warp %d: timeout waiting for rank %d batch %u (flag %u)\n",
                                   rank, gwarp, lane, bseq + 1u, *fl);
                            atomicExch(p.error_flag, 1);
                            __trap();
                        }
                    }
                }
                (void)ld_acquire_sys(my_flags + lane);          // acquire: the peer's data stores are visible
                if (p.timing) wait_ns += globaltimer_ns() - t0;
                const float4* src = reinterpret_cast<const float4*>(
                    p.xbuf[rank] + warp_x_base + ((size_t)slot * S + lane) * slot_stride);
                float4 got[M_G * M_FP / 4];
#pragma unroll
                for (int v4 = 0; v4 < M_G * M_FP / 4; ++v4) got[v4] = __ldcg(src + v4);
#pragma unroll
                for (int v4 = 0; v4 < M_G * M_FP / 4; ++v4)
                    reinterpret_cast<float4*>(xsum + lane * 32)[v4] = got[v4];
            }
            const int g = lane >> 3, vi = lane & 7;
            // the batch holds min(M_G, pairs pushed beyond b_lo) pairs; a partial batch only occurs at the stream end
            const int b_cnt = min(M_G, pushed - b_lo);
            xsum[rank * 32 + lane] = (g < b_cnt) ? fpart[((b_lo + g) % M_RING) * M_FP + vi] : 0.f;
            __syncwarp();
            float tot = 0.f;
            for (int r = 0; r < S; ++r) tot += xsum[r * 32 + lane];          // fixed order: bit-identical on all ranks
            if (g < b_cnt) ftot[((b_lo + g) % M_RING) * M_FP + vi] = tot;
            __syncwarp();
            recv_end = b_lo + b_cnt;
            ++nb_recv;
        }
        {
            const int* e = ring + (cb % M_RING) * M_ENTRY;
            const int wtok = e[0], ctok = e[1], flags = e[2];
            if (flags & 1) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                    for (int el = 0; el < VEC; ++el) { uu[c][el] = 0.f; du[c][el] = 0.f; }
                    if (act[c]) m_lds<VEC>(stage + (size_t)(n + 1) * K + coff[c], uu[c]);
                }
            }
            ++pairs;
            const float* ft = ftot + (cb % M_RING) * M_FP;
            // lane r (< 8) turns the reduced dot of row r into its coefficient / loss exactly once
            float gmine = 0.f;
            {
                const int myrow = lane & 7;
                const float f = ft[myrow];
                const float label = (myrow == 0) ? 1.f : 0.f;
                const bool myact = (myrow <= n) && (myrow == 0 || e[4 + myrow - 1] != ctok);
                gmine = myact ? sgns_coeff(f, label, p.alpha, p.max_grad) : 0.f;
                if (p.compute_loss && myact && lane < 8) {
                    loss += softplus_clipped(label > 0.5f ? -f : f);
                    maxdot = fmaxf(maxdot, fabsf(f));
                }
            }
            for (int r = 0; r <= n; ++r) {
                const float g = __shfl_sync(0xffffffffu, gmine, r);
                if (r > 0 && e[4 + r - 1] == ctok) continue;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    if (!act[c]) continue;
                    float v[VEC], gu[VEC];
                    m_lds<VEC>(stage + (size_t)r * K + coff[c], v);
#pragma unroll
                    for (int el = 0; el < VEC; ++el) {
                        du[c][el] = fmaf(g, v[el], du[c][el]);
                        gu[el] = g * uu[c][el];
                    }
                    m_sts<VEC>(stage + (size_t)r * K + coff[c], gu);
                }
            }
            if (flags & 2) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    if (act[c]) m_sts<VEC>(stage + (size_t)(n + 1) * K + coff[c], du[c]);
            }
            m_fence_async();
            __syncwarp();
            if (lane <= n) {
                const int row = (lane == 0) ? ctok : e[4 + lane - 1];
                if ((lane == 0 || row != ctok) && !(p.debug & 1))
                    m_bulk_reduce_add(p.syn1 + (size_t)row * K, stage + (size_t)lane * K, row_bytes);
            } else if (lane == n + 1 && (flags & 2) && !(p.debug & 2)) {
                m_bulk_reduce_add(p.syn0 + (size_t)wtok * K, stage + (size_t)(n + 1) * K, row_bytes);
            }
            m_commit();
            ++cb;
            ++items_done;
        }
    }
    m_wait_all();
    if (lane == 0) a.warp_seq[gwarp] = seq;
    (void)nb_pushed;

    if (blockIdx.x == 0 && threadIdx.x == 0) p.stats[3] = (float)T;
    loss = warp_sum(loss);
    maxdot = warp_max(maxdot);
    if (lane == 0 && pairs) {
        atomicAdd(p.stats + 0, (float)pairs);
        if (p.compute_loss) {
            atomicAdd(p.stats + 1, loss);
            atomicMax(reinterpret_cast<int*>(p.stats + 2), __float_as_int(maxdot));
        }
    }
    if (p.timing) {
        unsigned long long w = 0;
        for (int o = 0; o < 32; ++o) { unsigned long long x = __shfl_sync(0xffffffffu, wait_ns, o); w = x > w ? x : w; }
        if (lane == 0 && w) atomicAdd(p.timing + 0, w);
    }
}

// ------------------------------------------------------------------ host side

struct MultiLayout { int warps, stages, lag, stage_floats, warp_bytes, nslot; size_t total; };

static MultiLayout multi_layout(int K, int negatives) {
    MultiLayout best{0, 0, 0, 0, 0, 0, 0};
    const size_t budget = 220 * 1024;
    const int stage_floats = (negatives + 2) * K;
    const size_t stage_bytes = (size_t)stage_floats * 4;
    const int max_warps = (K > 256) ? 8 : 16;
    const int lag = 16;                                  // pairs between dots and update = 4 batches in flight
    long best_score = -1;
    for (int warps = max_warps; warps >= 2; --warps) {
        size_t per_warp = (budget / warps) & ~(size_t)127;
        int stages = 0;
        for (int st = 8; st >= 3; --st)
            if (m_fixed_bytes() + (size_t)st * stage_bytes <= per_warp) { stages = st; break; }
        if (stages == 0) continue;
        long score = (long)warps * (stages > 4 ? 4 : stages) * 16 + warps;
        if (score > best_score) {
            best_score = score;
            size_t wb = (m_fixed_bytes() + (size_t)stages * stage_bytes + 127) & ~(size_t)127;
            // batches in flight per warp <= lag / M_G + 2; slots must cover twice that (see the reuse argument)
            best = MultiLayout{warps, stages, lag, stage_floats, (int)wb, 2 * (lag / M_G + 3), wb * warps};
        }
    }
    return best;
}

bool sgns_pipe_multi_supported(int K, int window, int negatives) {
    if (negatives < 1 || negatives > M_MAXNEG) return false;
    if (2 * window + 1 > 32 || 2 * window > M_RING - 8) return false;
    if (K % 4 != 0 || K > 1024) return false;
    return multi_layout(K, negatives).warps >= 2;
}

#define GW2V_MP_DISPATCH(K, CALL)                                  \
    do {                                                           \
        if ((K) <= 64) { CALL(2, 1); }                             \
        else if ((K) <= 128) { CALL(4, 1); }                       \
        else if ((K) <= 256) { CALL(4, 2); }                       \
        else if ((K) <= 384) { CALL(4, 3); }                       \
        else if ((K) <= 512) { CALL(4, 4); }                       \
        else if ((K) <= 768) { CALL(4, 6); }                       \
        else { CALL(4, 8); }                                       \
    } while (0)

// geometry the host needs to size the symmetric buffers: {grid, warps, nslot, floats per (warp, slot, src)}
void sgns_pipe_multi_geometry(int K, int negatives, int device, int* grid, int* warps, int* nslot, int* slot_floats) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    MultiLayout l = multi_layout(K, negatives);
    int occ = 1;
#define CALL(V, C)                                                                                               \
    do {                                                                                                         \
        cudaFuncSetAttribute(sgns_fused_pipe_multi_kernel<V, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                             (int)l.total);                                                                      \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sgns_fused_pipe_multi_kernel<V, C>, l.warps * 32,    \
                                                      l.total);                                                  \
    } while (0)
    GW2V_MP_DISPATCH(K, CALL);
#undef CALL
    if (occ < 1) occ = 1;
    *grid = sms * occ;              // every CTA co-resident: required by the in-kernel flag protocol
    *warps = l.warps;
    *nslot = l.nslot;
    *slot_floats = M_G * M_FP;
}

void launch_sgns_pipe_multi(const SgnsParams& p, int grid, uint32_t* warp_seq, cudaStream_t stream) {
    MultiLayout l = multi_layout(p.K, p.negatives);
    MultiPipeArgs a{l.stages, l.lag, l.stage_floats, l.warp_bytes, l.nslot, warp_seq};
#define CALL(V, C)                                                                                               \
    do {                                                                                                         \
        cudaFuncSetAttribute(sgns_fused_pipe_multi_kernel<V, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                             (int)l.total);                                                                      \
        sgns_fused_pipe_multi_kernel<V, C><<<grid, l.warps * 32, l.total, stream>>>(p, a);                       \
    } while (0)
    GW2V_MP_DISPATCH(p.K, CALL);
#undef CALL
}

}  // namespace gw2v
