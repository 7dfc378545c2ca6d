// sgns_fused_group_multi: column-sharded SGNS step, lane-group register path, with the partial-dot
// all-reduce fused in (world > 1).  One kernel per rank, no NCCL call on this path.
//
// Reference hot path: dotprod fan-out -> per-shard partial dots -> client sums -> sigmoid ->
// adjust fan-out (MLLIB:417-429, Glint server ops [G]).  Here, per warp:
//   A step  the G-lane groups of the warp compute the partial dots of P pairs over this rank's K
//           columns; every 4 pairs ("batch") lane j stores the batch straight into rank j's
//           symmetric exchange slot over NVLink (st.global.v4 on a peer-mapped address) and
//           publishes the batch sequence number with st.release.sys;
//   B step  `lag` pairs later: poll the peers' flags (ld.volatile + one ld.acquire.sys), sum the S
//           partials in fixed rank order (bit-identical coefficients on all ranks, so the
//           reference's coefficient broadcast disappears), re-read the rows (L2 hits) and apply
//           the updates with RED.128.
// The A/B schedule is a pure function of per-warp counters, so the S ranks agree on batch
// boundaries without negotiation; sequence numbers persist across launches and are never reset.
// With 24-40 resident warps per SM the NVLink round trip is covered by other warps' work.
#include "pipe_common.cuh"
#include <cstdio>

namespace gw2v {

constexpr int GM_GEN = 4;
constexpr int GM_RING = 64;          // pair descriptors per warp
constexpr int GM_RF = 32;            // dot slots per warp (pairs alive between A and B)
constexpr int GM_FP = 8;             // floats per pair in exchange slots (1 + n <= 8)
constexpr int GM_G = 4;              // pairs per batch (slot = 32 floats = 128 B)
constexpr int GM_LAG = 16;           // pairs between the dots pass and the update pass
constexpr int GM_NSLOT = 2 * (GM_LAG / GM_G + 3);
constexpr int GM_THREADS = 256;
constexpr int GM_MAXNEG = 7;

__device__ __forceinline__ void gm_ld4(const float* p, float (&o)[4]) {
    float4 v = __ldcg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void gm_red4(float* p, const float (&v)[4]) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}

struct GmWarpSmem {
    int ring[GM_RING * PIPE_ENTRY];
    float fdot[GM_RF * GM_FP];
    float xsum[8 * 32];
};

template <int G, int CHUNKS>
__global__ void __launch_bounds__(GM_THREADS, (CHUNKS == 1) ? 3 : ((CHUNKS == 2) ? 2 : 1))
sgns_fused_group_multi_kernel(const SgnsParams p, uint32_t* warp_seq) {
    constexpr int P = 32 / G;
    __shared__ __align__(16) GmWarpSmem wsm[GM_THREADS / 32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int* ring = wsm[warp].ring;
    float* fdot = wsm[warp].fdot;
    float* xsum = wsm[warp].xsum;

    const int K = p.K;
    const int n = p.negatives;
    const int T = *p.n_tokens;
    const int maxgen = GM_GEN * 2 * p.window;
    const int S = p.world;
    const int rank = p.rank;
    const int grp = lane / G, lg = lane % G;
    bool act[CHUNKS];
    int coff[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) { coff[c] = (c * G + lg) * 4; act[c] = coff[c] < K; }

    const int gwarp = blockIdx.x * (GM_THREADS / 32) + warp;          // identical on every rank
    const int n_warps = gridDim.x * (GM_THREADS / 32);
    const size_t slot_stride = (size_t)GM_G * GM_FP;
    const size_t warp_x_base = (size_t)gwarp * GM_NSLOT * S * slot_stride;
    uint32_t* my_flags = p.flags[rank] + (size_t)gwarp * S;
    uint32_t seq = warp_seq[gwarp];
    const uint32_t seq0 = seq;

    int gen_i = gwarp;
    int head = 0, ia = 0, ib = 0, pushed = 0, nb_recv = 0, recv_end = 0;
    float loss = 0.f, maxdot = 0.f;
    unsigned pairs = 0;
    unsigned long long wait_ns = 0;

    auto push_batch = [&]() {
        __syncwarp();
        const int slot = (int)(seq % (uint32_t)GM_NSLOT);
        if (lane < S && lane != rank) {
            float4* dst = reinterpret_cast<float4*>(p.xbuf[lane] + warp_x_base +
                                                    ((size_t)slot * S + rank) * slot_stride);
            const int cnt = ia - pushed;
            for (int g = 0; g < cnt; ++g) {
                const float4* src = reinterpret_cast<const float4*>(fdot + ((pushed + g) % GM_RF) * GM_FP);
                dst[g * 2 + 0] = src[0];
                dst[g * 2 + 1] = src[1];
            }
            st_release_sys(p.flags[lane] + (size_t)gwarp * S + rank, seq + 1u);
        }
        __syncwarp();
        pushed = ia;
        ++seq;
    };

    auto recv_batch = [&]() {
        const int b_lo = recv_end;
        const uint32_t bseq = seq0 + (uint32_t)nb_recv;
        const int slot = (int)(bseq % (uint32_t)GM_NSLOT);
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
            float4 got[GM_G * GM_FP / 4];
#pragma unroll
            for (int v4 = 0; v4 < GM_G * GM_FP / 4; ++v4) got[v4] = __ldcg(src + v4);
#pragma unroll
            for (int v4 = 0; v4 < GM_G * GM_FP / 4; ++v4) reinterpret_cast<float4*>(xsum + lane * 32)[v4] = got[v4];
        }
        const int g = lane >> 3, vi = lane & 7;
        const int b_cnt = min(GM_G, pushed - b_lo);         // a short batch only occurs at the end of the stream
        xsum[rank * 32 + lane] = (g < b_cnt) ? fdot[((b_lo + g) % GM_RF) * GM_FP + vi] : 0.f;
        __syncwarp();
        float tot = 0.f;
        for (int r = 0; r < S; ++r) tot += xsum[r * 32 + lane];          // fixed order: bit-identical on all ranks
        if (g < b_cnt) fdot[((b_lo + g) % GM_RF) * GM_FP + vi] = tot;
        __syncwarp();
        recv_end = b_lo + b_cnt;
        ++nb_recv;
    };

    while (true) {
        while (gen_i < T && (GM_RING - (head - ib)) >= maxgen)
            head += generate_pairs<GM_GEN, GM_RING>(p, T, gen_i, n_warps, ring, head, lane);
        const bool gen_done = gen_i >= T;

        if ((head - ia >= P || (gen_done && head > ia)) && ia - ib < GM_LAG) {
            // ---------------------------------------------------------------- A: partial dots
            const int cnt = min(P, head - ia);
            const bool gvalid = grp < cnt;
            const int pidx = ia + (gvalid ? grp : 0);
            const int* e = ring + (pidx % GM_RING) * PIPE_ENTRY;
            const int ctok = e[1];
            float u[CHUNKS][4];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                for (int el = 0; el < 4; ++el) u[c][el] = 0.f;
                if (gvalid && act[c]) gm_ld4(p.syn0 + (size_t)e[0] * K + coff[c], u[c]);
            }
            float v[8][CHUNKS][4];
            bool ract[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int row = (r == 0 || r > n) ? ctok : e[4 + r - 1];
                ract[r] = gvalid && r <= n && (r == 0 || row != ctok);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                    for (int el = 0; el < 4; ++el) v[r][c][el] = 0.f;
                    if (ract[r] && act[c]) gm_ld4(p.syn1 + (size_t)row * K + coff[c], v[r][c]);
                }
            }
            float f[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float sacc = 0.f;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
                    for (int el = 0; el < 4; ++el) sacc = fmaf(u[c][el], v[r][c][el], sacc);
                f[r] = sacc;
            }
            const float tot = group_reduce8<G>(f, lane);
            const int myrow = row_of_lane<G>(lane);
            if (gvalid && lg == lane_of_row<G>(myrow)) fdot[(pidx % GM_RF) * GM_FP + myrow] = tot;
            ia += cnt;
            if (ia - pushed >= GM_G || (gen_done && ia == head)) push_batch();
            continue;
        }
        if (gen_done && ia == head && ia > pushed) push_batch();          // stream ended with an open batch
        if (ib < ia && (ia - ib >= GM_LAG || (gen_done && ia == head))) {
            // ---------------------------------------------------------------- B: reduce + update
            const int cnt = min(P, ia - ib);
            while (recv_end < ib + cnt) recv_batch();
            const bool gvalid = grp < cnt;
            const int pidx = ib + (gvalid ? grp : 0);
            const int* e = ring + (pidx % GM_RING) * PIPE_ENTRY;
            const int ctok = e[1];
            float* urow = p.syn0 + (size_t)e[0] * K;
            float u[CHUNKS][4], du[CHUNKS][4];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                for (int el = 0; el < 4; ++el) { u[c][el] = 0.f; du[c][el] = 0.f; }
                if (gvalid && act[c]) gm_ld4(urow + coff[c], u[c]);
            }
            if (lg == 0 && gvalid) ++pairs;
            const int myrow = row_of_lane<G>(lane);
            const bool myact = gvalid && (myrow <= n) && (myrow == 0 || e[4 + myrow - 1] != ctok);
            const float fm = fdot[(pidx % GM_RF) * GM_FP + myrow];
            const float mylabel = (myrow == 0) ? 1.f : 0.f;
            const float gmine = myact ? sgns_coeff(fm, mylabel, p.alpha, p.max_grad, p.exp_table) : 0.f;
            if (p.compute_loss && myact && lg == lane_of_row<G>(myrow)) {
                loss += softplus_clipped(mylabel > 0.5f ? -fm : fm);
                maxdot = fmaxf(maxdot, fabsf(fm));
            }
            float v[8][CHUNKS][4];
            int rows[8];
            bool ract[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                rows[r] = (r == 0 || r > n) ? ctok : e[4 + r - 1];
                ract[r] = gvalid && r <= n && (r == 0 || rows[r] != ctok);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                    for (int el = 0; el < 4; ++el) v[r][c][el] = 0.f;
                    if (ract[r] && act[c]) gm_ld4(p.syn1 + (size_t)rows[r] * K + coff[c], v[r][c]);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float g = __shfl_sync(0xffffffffu, gmine, lane_of_row<G>(r), G);
                if (!ract[r]) continue;
                float* vrow = p.syn1 + (size_t)rows[r] * K;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    if (!act[c]) continue;
                    float gu[4];
#pragma unroll
                    for (int el = 0; el < 4; ++el) {
                        du[c][el] = fmaf(g, v[r][c][el], du[c][el]);
                        gu[el] = g * u[c][el];
                    }
                    if (!(p.debug & 1)) gm_red4(vrow + coff[c], gu);
                }
            }
            if (gvalid && !(p.debug & 2)) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    if (act[c]) gm_red4(urow + coff[c], du[c]);
            }
            ib += cnt;
            continue;
        }
        if (gen_done && ib == head) break;
    }
    if (lane == 0) warp_seq[gwarp] = seq;

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

static void gm_group(int K, int* G, int* chunks) {
    if (K <= 32) { *G = 8; *chunks = 1; }
    else if (K <= 64) { *G = 16; *chunks = 1; }
    else { *G = 32; *chunks = (K + 127) / 128; }
}

bool sgns_group_multi_supported(int K, int window, int negatives) {
    if (negatives < 1 || negatives > GM_MAXNEG) return false;
    if (2 * window + 1 > 32) return false;
    if (GM_GEN * 2 * window + GM_LAG + 8 > GM_RING) return false;
    return K % 4 == 0 && K <= 1024;
}

#define GW2V_GM_DISPATCH(K, CALL)                                            \
    do {                                                                     \
        int G_, ch_;                                                         \
        gm_group((K), &G_, &ch_);                                            \
        if (G_ == 8) { CALL(8, 1); }                                         \
        else if (G_ == 16) { CALL(16, 1); }                                  \
        else if (ch_ == 1) { CALL(32, 1); }                                  \
        else if (ch_ == 2) { CALL(32, 2); }                                  \
        else if (ch_ == 3) { CALL(32, 3); }                                  \
        else if (ch_ == 4) { CALL(32, 4); }                                  \
        else if (ch_ <= 6) { CALL(32, 6); }                                  \
        else { CALL(32, 8); }                                                \
    } while (0)

void sgns_group_multi_geometry(int K, int device, int* grid, int* warps, int* nslot, int* slot_floats) {
    int sms = 148, occ = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
#define CALL(GG, C) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sgns_fused_group_multi_kernel<GG, C>, GM_THREADS, 0)
    GW2V_GM_DISPATCH(K, CALL);
#undef CALL
    if (occ < 1) occ = 1;
    *grid = sms * occ;              // every CTA co-resident: required by the in-kernel flag protocol
    *warps = GM_THREADS / 32;
    *nslot = GM_NSLOT;
    *slot_floats = GM_G * GM_FP;
}

void launch_sgns_group_multi(const SgnsParams& p, int grid, uint32_t* warp_seq, cudaStream_t stream) {
#define CALL(GG, C) sgns_fused_group_multi_kernel<GG, C><<<grid, GM_THREADS, 0, stream>>>(p, warp_seq)
    GW2V_GM_DISPATCH(p.K, CALL);
#undef CALL
}

}  // namespace gw2v
