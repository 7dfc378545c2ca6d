// sgns_fused_pipe_multi: column-sharded SGNS step with the partial-dot all-reduce fused into the
// per-warp TMA pipeline (world > 1).  This is the hot path of the reference -- dotprod fan-out,
// client-side sum of the shards' partial dot products, sigmoid, adjust fan-out (MLLIB:417-429,
// Glint server ops [G]) -- as ONE kernel per rank with no NCCL call.
//
// Every pair passes through the warp's shared-memory stage ring twice:
//   A step  TMA-load the rows of P pairs (cp.async.bulk), partial dots over this rank's K columns;
//           every 4 pairs ("batch") lane j stores the batch's partials straight into rank j's
//           symmetric exchange slot over NVLink (st.global.v4 on a peer-mapped address) and then
//           publishes the batch sequence number with st.release.sys;
//   B step  `lag` pairs later: TMA-load the rows again (L2 hits), spin (ld.volatile + one
//           ld.acquire.sys) on the peers' flags, sum the S partials in fixed rank order (all ranks
//           obtain bit-identical coefficients, so the reference's coefficient broadcast disappears),
//           g = (label - sigmoid(f)) * alpha, rows overwritten in place with g*u, centre slot with
//           du, cp.reduce.async.bulk.add.f32 (TMA reduce) back to this rank's shard.
// Between A and B of a pair the warp keeps streaming other pairs, so the NVLink round trip is
// hidden behind useful work; only 32 bytes per pair wait in shared memory.
//
// The A/B interleaving is decided at issue time from per-warp counters only (never from timing),
// recorded in a small item queue and replayed at compute time, so the S ranks' warps follow the
// identical schedule and agree on batch boundaries without negotiation.  Batch sequence numbers
// persist in device memory across launches and are never reset.
#include "pipe_common.cuh"
#include <cstdio>

namespace gw2v {

constexpr int M_GEN = 4;             // centres per generation round
constexpr int M_RING = 96;           // pair descriptors per warp
constexpr int M_RF = 64;             // partial/total dot slots per warp (pairs alive between A and B)
constexpr int M_FP = 8;              // floats per pair in exchange slots (1 + n <= 8)
constexpr int M_G = 4;               // pairs per batch (slot = 32 floats = 128 B)
constexpr int M_MAXNEG = 7;
constexpr int M_IQ = 64;             // item queue entries (bytes)

struct MultiPipeArgs {
    int nstage;          // smem stages per warp
    int lag;             // pairs between the dots pass (A) and the update pass (B) of the same pair
    int stage_floats;    // P * (n + 2) * K
    int warp_bytes;      // smem per warp
    int nslot;           // exchange slots per (warp, source)
    uint32_t* warp_seq;  // [grid * warps] running batch sequence per warp (device memory, never reset)
};

// per-warp shared memory:  stages | mbarriers (128 B) | fdot[M_RF][8] | xsum[8][32] | item queue | ring
__host__ __device__ inline size_t m_fixed_bytes() {
    return 128 + (size_t)M_RF * M_FP * 4 + 8 * 32 * 4 + M_IQ + (size_t)M_RING * PIPE_ENTRY * 4;
}

template <int G, int CHUNKS>
__global__ void __launch_bounds__((CHUNKS >= 3) ? 256 : 512)
sgns_fused_pipe_multi_kernel(const SgnsParams p, const MultiPipeArgs a) {
    constexpr int P = 32 / G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp_cta = blockDim.x >> 5;
    const int nstage = a.nstage;
    unsigned char* wbase = smem_raw + (size_t)warp * a.warp_bytes;
    float* stages = reinterpret_cast<float*>(wbase);
    unsigned char* q = wbase + (size_t)nstage * a.stage_floats * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(q);            q += 128;
    float* fdot = reinterpret_cast<float*>(q);                  q += (size_t)M_RF * M_FP * 4;   // partials, then totals
    float* xsum = reinterpret_cast<float*>(q);                  q += 8 * 32 * 4;
    unsigned char* iq = q;                                      q += M_IQ;
    int* ring = reinterpret_cast<int*>(q);

    const int K = p.K;
    const int n = p.negatives;
    const int R = n + 2;
    const uint32_t row_bytes = (uint32_t)K * 4u;
    const int T = *p.n_tokens;
    const int maxgen = M_GEN * 2 * p.window;
    const int S = p.world;
    const int rank = p.rank;
    const int L = a.lag;
    const int gwarp = blockIdx.x * nwarp_cta + warp;            // identical on every rank
    const int n_warps = gridDim.x * nwarp_cta;

    if (lane == 0) {
        for (int s = 0; s < nstage; ++s) pp_mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int grp = lane / G, lg = lane % G;
    bool act[CHUNKS];
    int coff[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) { coff[c] = (c * G + lg) * 4; act[c] = coff[c] < K; }
    const int tma_pair = lane / R, tma_row = lane % R;
    const bool tma_lane = lane < P * R;

    const size_t slot_stride = (size_t)M_G * M_FP;                               // floats per (slot, src)
    const size_t warp_x_base = (size_t)gwarp * a.nslot * S * slot_stride;
    uint32_t* my_flags = p.flags[rank] + (size_t)gwarp * S;
    uint32_t seq = a.warp_seq[gwarp];                                            // next batch to push
    const uint32_t seq0 = seq;

    int gen_i = gwarp;
    int head = 0;                         // pairs generated
    int ia = 0, ib = 0;                   // pairs issued to the A / B pass
    int ca = 0, cb = 0;                   // pairs computed by the A / B pass
    int items_issued = 0, items_done = 0; // positions in the item sequence
    int pushed = 0;                       // pairs whose partials have been pushed
    int nb_recv = 0, recv_end = 0;
    float loss = 0.f, maxdot = 0.f;
    unsigned pairs = 0;
    unsigned long long wait_ns = 0;

    // push pairs [pushed, ca) as one batch: lane j stores them into rank j's slot, then release-publishes seq+1
    auto push_batch = [&]() {
        __syncwarp();
        const int slot = (int)(seq % (uint32_t)a.nslot);
        if (lane < S && lane != rank) {
            float4* dst = reinterpret_cast<float4*>(p.xbuf[lane] + warp_x_base +
                                                    ((size_t)slot * S + rank) * slot_stride);
            const int cnt = ca - pushed;
            for (int g = 0; g < cnt; ++g) {
                const float4* src = reinterpret_cast<const float4*>(fdot + ((pushed + g) % M_RF) * M_FP);
                dst[g * 2 + 0] = src[0];
                dst[g * 2 + 1] = src[1];
            }
            st_release_sys(p.flags[lane] + (size_t)gwarp * S + rank, seq + 1u);
        }
        __syncwarp();
        pushed = ca;
        ++seq;
    };

    // receive the next batch (pairs [recv_end, recv_end + cnt)): wait for every peer, reduce in rank order
    auto recv_batch = [&]() {
        const int b_lo = recv_end;
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
            for (int v4 = 0; v4 < M_G * M_FP / 4; ++v4) reinterpret_cast<float4*>(xsum + lane * 32)[v4] = got[v4];
        }
        const int g = lane >> 3, vi = lane & 7;
        const int b_cnt = min(M_G, pushed - b_lo);          // a short batch only occurs at the end of the stream
        xsum[rank * 32 + lane] = (g < b_cnt) ? fdot[((b_lo + g) % M_RF) * M_FP + vi] : 0.f;
        __syncwarp();
        float tot = 0.f;
        for (int r = 0; r < S; ++r) tot += xsum[r * 32 + lane];          // fixed order: bit-identical on all ranks
        if (g < b_cnt) fdot[((b_lo + g) % M_RF) * M_FP + vi] = tot;
        __syncwarp();
        recv_end = b_lo + b_cnt;
        ++nb_recv;
    };

    while (true) {
        // ---------------------------------------------------------------- (1) generate descriptors
        while (gen_i < T && (M_RING - (head - cb)) >= maxgen)
            head += generate_pairs<M_GEN, M_RING>(p, T, gen_i, n_warps, ring, head, lane);
        const bool gen_done = gen_i >= T;

        // ---------------------------------------------------------------- (2) issue steps (TMA loads)
        while (items_issued - items_done < nstage) {
            int kind, cnt;                                // 0 = A, 1 = B
            if ((head - ia >= P || (gen_done && head > ia)) && ia - ib < L) { kind = 0; cnt = min(P, head - ia); }
            else if (ib < ia && (ia - ib >= L || (gen_done && ia == head))) { kind = 1; cnt = min(P, ia - ib); }
            else break;
            const int first = kind ? ib : ia;
            const int s = items_issued % nstage;
            float* stage = stages + (size_t)s * a.stage_floats;
            if (items_issued >= nstage && (iq[(items_issued - nstage) % M_IQ] & 1)) pp_wait_read0();  // a B step's reduces read this stage
            bool mine = false;
            const float* src = nullptr;
            if (tma_lane && tma_pair < cnt) {
                const int* e = ring + ((first + tma_pair) % M_RING) * PIPE_ENTRY;
                const int ctok = e[1];
                if (tma_row == 0) { mine = true; src = p.syn1 + (size_t)ctok * K; }
                else if (tma_row <= n) { const int ng = e[4 + tma_row - 1]; mine = ng != ctok; src = p.syn1 + (size_t)ng * K; }
                else { mine = true; src = p.syn0 + (size_t)e[0] * K; }
            }
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (lane == 0) {
                pp_mbar_expect_tx(bars + s, (uint32_t)__popc(m) * row_bytes);
                iq[items_issued % M_IQ] = (unsigned char)(kind | (cnt << 1));
            }
            __syncwarp();
            if (mine) pp_bulk_load(stage + (size_t)(tma_pair * R + tma_row) * K, src, row_bytes, bars + s);
            if (kind) ib += cnt; else ia += cnt;
            ++items_issued;
        }

        // the stream ended after the last dots were computed: close the open batch (same decision on every rank)
        if (gen_done && ia == head && ca == head && ca > pushed) push_batch();

        if (items_done == items_issued) {
            if (gen_done && cb == head) break;
            continue;
        }

        // ---------------------------------------------------------------- (3) compute the next step
        const int s = items_done % nstage;
        const int item = iq[items_done % M_IQ];
        const int kind = item & 1, cnt = item >> 1;
        const bool gvalid = grp < cnt;
        float* stage = stages + (size_t)s * a.stage_floats + (size_t)grp * R * K;
        pp_mbar_wait(bars + s, (uint32_t)((items_done / nstage) & 1));
        if (kind == 0) {
            // ---------------- A: partial dots of pairs [ca, ca + cnt)
            const int pidx = ca + (gvalid ? grp : 0);
            const int* e = ring + (pidx % M_RING) * PIPE_ENTRY;
            const int ctok = e[1];
            float u[CHUNKS][4];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                for (int el = 0; el < 4; ++el) u[c][el] = 0.f;
                if (gvalid && act[c]) pp_lds4(stage + (size_t)(n + 1) * K + coff[c], u[c]);
            }
            float f[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float sacc = 0.f;
                if (gvalid && r <= n && (r == 0 || e[4 + r - 1] != ctok)) {
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
                        if (!act[c]) continue;
                        float v[4];
                        pp_lds4(stage + (size_t)r * K + coff[c], v);
#pragma unroll
                        for (int el = 0; el < 4; ++el) sacc = fmaf(u[c][el], v[el], sacc);
                    }
                }
                f[r] = sacc;
            }
            const float tot = group_reduce8<G>(f, lane);
            const int myrow = row_of_lane<G>(lane);
            if (gvalid && lg == lane_of_row<G>(myrow)) fdot[(pidx % M_RF) * M_FP + myrow] = tot;
            ca += cnt;
            ++items_done;
            if (ca - pushed >= M_G || (gen_done && ca == head)) push_batch();
            continue;
        }
        // ---------------- B: reduce + update pairs [cb, cb + cnt)
        while (recv_end < cb + cnt) recv_batch();
        {
            const int pidx = cb + (gvalid ? grp : 0);
            const int* e = ring + (pidx % M_RING) * PIPE_ENTRY;
            const int ctok = e[1];
            float u[CHUNKS][4], du[CHUNKS][4];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                for (int el = 0; el < 4; ++el) { u[c][el] = 0.f; du[c][el] = 0.f; }
                if (gvalid && act[c]) pp_lds4(stage + (size_t)(n + 1) * K + coff[c], u[c]);
            }
            if (lg == 0 && gvalid) ++pairs;
            // the lane owning row r turns the reduced dot into its coefficient / loss exactly once
            const int myrow = row_of_lane<G>(lane);
            const bool myact = gvalid && (myrow <= n) && (myrow == 0 || e[4 + myrow - 1] != ctok);
            const float fm = fdot[(pidx % M_RF) * M_FP + myrow];
            const float mylabel = (myrow == 0) ? 1.f : 0.f;
            const float gmine = myact ? sgns_coeff(fm, mylabel, p.alpha, p.max_grad, p.exp_table) : 0.f;
            if (p.compute_loss && myact && lg == lane_of_row<G>(myrow)) {
                loss += softplus_clipped(mylabel > 0.5f ? -fm : fm);
                maxdot = fmaxf(maxdot, fabsf(fm));
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float g = __shfl_sync(0xffffffffu, gmine, lane_of_row<G>(r), G);
                if (!(gvalid && r <= n && (r == 0 || e[4 + r - 1] != ctok))) continue;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    if (!act[c]) continue;
                    float v[4], gu[4];
                    pp_lds4(stage + (size_t)r * K + coff[c], v);
#pragma unroll
                    for (int el = 0; el < 4; ++el) {
                        du[c][el] = fmaf(g, v[el], du[c][el]);
                        gu[el] = g * u[c][el];
                    }
                    pp_sts4(stage + (size_t)r * K + coff[c], gu);
                }
            }
            if (gvalid) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    if (act[c]) pp_sts4(stage + (size_t)(n + 1) * K + coff[c], du[c]);
            }
            pp_fence_async();
            __syncwarp();
            if (tma_lane && tma_pair < cnt) {
                const int* e2 = ring + ((cb + tma_pair) % M_RING) * PIPE_ENTRY;
                const int ct = e2[1];
                float* sst = stages + (size_t)s * a.stage_floats + (size_t)(tma_pair * R + tma_row) * K;
                if (tma_row == 0) { if (!(p.debug & 1)) pp_bulk_reduce_add(p.syn1 + (size_t)ct * K, sst, row_bytes); }
                else if (tma_row <= n) {
                    const int ng = e2[4 + tma_row - 1];
                    if (ng != ct && !(p.debug & 1)) pp_bulk_reduce_add(p.syn1 + (size_t)ng * K, sst, row_bytes);
                } else if (!(p.debug & 2)) pp_bulk_reduce_add(p.syn0 + (size_t)e2[0] * K, sst, row_bytes);
            }
            pp_commit();
            cb += cnt;
            ++items_done;
        }
    }
    pp_wait_all();
    if (lane == 0) a.warp_seq[gwarp] = seq;

    if (blockIdx.x == 0 && threadIdx.x == 0) p.stats[3] = (float)T;
    loss = warp_sum(loss);
    maxdot = warp_max(maxdot);
    const float pf = warp_sum((float)pairs);
    if (lane == 0 && pf > 0.f) {
        atomicAdd(p.stats + 0, pf);
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

struct MultiLayout { int G, chunks, warps, stages, lag, stage_floats, warp_bytes, nslot; size_t total; };

static void mp_group(int K, int* G, int* chunks) {
    if (K <= 32) { *G = 8; *chunks = 1; }
    else if (K <= 64) { *G = 16; *chunks = 1; }
    else { *G = 32; *chunks = (K + 127) / 128; }
}

static MultiLayout multi_layout(int K, int negatives) {
    MultiLayout best{0, 0, 0, 0, 0, 0, 0, 0, 0};
    const size_t budget = 220 * 1024;
    int G, chunks;
    mp_group(K, &G, &chunks);
    const int P = 32 / G;
    const int stage_floats = P * (negatives + 2) * K;
    const size_t stage_bytes = (size_t)stage_floats * 4;
    const int max_warps = (chunks >= 3) ? 8 : 16;
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
            // batches in flight per warp <= (lag + stages * P) / M_G + 2; slots cover twice that (reuse argument)
            const int nslot = 2 * ((lag + stages * P) / M_G + 3);
            best = MultiLayout{G, chunks, warps, stages, lag, stage_floats, (int)wb, nslot, wb * warps};
        }
    }
    return best;
}

bool sgns_pipe_multi_supported(int K, int window, int negatives) {
    if (negatives < 1 || negatives > M_MAXNEG) return false;
    if (2 * window + 1 > 32) return false;
    if (M_GEN * 2 * window + 16 + 8 * 4 + 4 > M_RING) return false;     // generation round + pairs alive between A and B
    if (K % 4 != 0 || K > 1024) return false;
    int G, chunks;
    mp_group(K, &G, &chunks);
    if ((32 / G) * (negatives + 2) > 32) return false;
    return multi_layout(K, negatives).warps >= 2;
}

#define GW2V_MP_DISPATCH(L, CALL)                                            \
    do {                                                                     \
        if ((L).G == 8) { CALL(8, 1); }                                      \
        else if ((L).G == 16) { CALL(16, 1); }                               \
        else if ((L).chunks == 1) { CALL(32, 1); }                           \
        else if ((L).chunks == 2) { CALL(32, 2); }                           \
        else if ((L).chunks == 3) { CALL(32, 3); }                           \
        else if ((L).chunks == 4) { CALL(32, 4); }                           \
        else if ((L).chunks <= 6) { CALL(32, 6); }                           \
        else { CALL(32, 8); }                                                \
    } while (0)

// geometry the host needs to size the symmetric buffers: {grid, warps, nslot, floats per (warp, slot, src)}
void sgns_pipe_multi_geometry(int K, int negatives, int device, int* grid, int* warps, int* nslot, int* slot_floats) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    MultiLayout l = multi_layout(K, negatives);
    int occ = 1;
#define CALL(GG, C)                                                                                              \
    do {                                                                                                         \
        cudaFuncSetAttribute(sgns_fused_pipe_multi_kernel<GG, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)l.total);                                                                      \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sgns_fused_pipe_multi_kernel<GG, C>, l.warps * 32,   \
                                                      l.total);                                                  \
    } while (0)
    GW2V_MP_DISPATCH(l, CALL);
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
#define CALL(GG, C)                                                                                              \
    do {                                                                                                         \
        cudaFuncSetAttribute(sgns_fused_pipe_multi_kernel<GG, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)l.total);                                                                      \
        sgns_fused_pipe_multi_kernel<GG, C><<<grid, l.warps * 32, l.total, stream>>>(p, a);                      \
    } while (0)
    GW2V_MP_DISPATCH(l, CALL);
#undef CALL
}

}  // namespace gw2v
