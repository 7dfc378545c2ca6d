// sgns_fused_pipe: the single-shard SGNS step as a per-warp TMA pipeline.
//
// v1 (sgns_kernels.cu) stages rows in registers (189 regs/thread at K=512, 8 warps/SM) and spends
// ~1300 warp instructions per pair; ncu showed it bound by its own dependent instruction stream,
// not by HBM (DRAM 44 %, issue 28 %, warps active 11 %).  This kernel
//   * never moves rows through registers on the way in or out: cp.async.bulk (TMA, UBLKCP) brings
//     the n+2 rows of a pair into the warp's shared-memory stage, and after the update the same
//     stage (each row overwritten IN PLACE with g*u, the centre slot with du) goes back with
//     cp.reduce.async.bulk.add.f32 (TMA reduce, UBLKRED) -- the scatter-add of Glint's `adjust`
//     (MLLIB:425) without a single per-lane RED instruction;
//   * lets G lanes (8/16/32) own a pair so a warp processes 32/G pairs per step with no divergence;
//   * generates windows / negatives for 8 centres at a time with all lanes sharing the Philox work;
//   * reduces the 8 dots of a pair with one transposed butterfly and evaluates each sigmoid once.
// Semantics are those of the oracle (models/sgns.py): per-pair private negatives, reference or
// word2vec.c window, +-6 clip, negatives equal to the context skipped.
#include "pipe_common.cuh"

namespace gw2v {

template <int G, int CHUNKS>
__global__ void __launch_bounds__((CHUNKS >= 3) ? 256 : 512)
sgns_fused_pipe_kernel(const SgnsParams p, const int nstage, const int stage_floats, const int warp_bytes) {
    constexpr int P = 32 / G;                       // pairs per step
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
    const int R = n + 2;                            // rows per pair: context, n negatives, centre
    const uint32_t row_bytes = (uint32_t)K * 4u;
    const int T = *p.n_tokens;
    const int maxgen = PIPE_GEN * 2 * p.window;     // worst-case pairs of one generation round

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
    // this lane's (pair, row) assignment for TMA issue: lane t -> pair t / R, row t % R
    const int tma_pair = lane / R, tma_row = lane % R;
    const bool tma_lane = lane < P * R;

    const int n_warps = gridDim.x * nwarp_cta;
    int gen_i = blockIdx.x * nwarp_cta + warp;
    int head = 0;                                   // pairs generated
    int issued = 0, done = 0;                       // steps
    int issued_pairs = 0, done_pairs = 0;
    float loss = 0.f, maxdot = 0.f;
    unsigned pairs = 0;

    while (true) {
        // ------------------------------------------------------------ (1) generate pair descriptors
        while (gen_i < T && (PIPE_RING - (head - done_pairs)) >= maxgen)
            head += generate_pairs(p, T, gen_i, n_warps, ring, head, lane);
        const bool gen_done = gen_i >= T;

        // ------------------------------------------------------------ (2) issue steps (TMA loads)
        while (issued - done < nstage && (head - issued_pairs >= P || (gen_done && head > issued_pairs))) {
            const int cnt = min(P, head - issued_pairs);
            const int s = issued % nstage;
            float* stage = stages + (size_t)s * stage_floats;
            if (issued >= nstage) pp_wait_read0();           // the TMA reduces that read this stage have drained it
            bool mine = false;
            const float* src = nullptr;
            if (tma_lane && tma_pair < cnt) {
                const int* e = ring + ((issued_pairs + tma_pair) % PIPE_RING) * PIPE_ENTRY;
                const int ctok = e[1];
                if (tma_row == 0) { mine = true; src = p.syn1 + (size_t)ctok * K; }
                else if (tma_row <= n) { const int ng = e[4 + tma_row - 1]; mine = ng != ctok; src = p.syn1 + (size_t)ng * K; }
                else { mine = true; src = p.syn0 + (size_t)e[0] * K; }
            }
            if (p.debug & 4) mine = false;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (lane == 0) pp_mbar_expect_tx(bars + s, (uint32_t)__popc(m) * row_bytes);
            __syncwarp();
            if (mine) pp_bulk_load(stage + (size_t)(tma_pair * R + tma_row) * K, src, row_bytes, bars + s);
            issued_pairs += cnt;
            ++issued;
        }
        if (done == issued) {
            if (gen_done && issued_pairs == head) break;
            continue;
        }

        // ------------------------------------------------------------ (3) compute step `done`
        {
            const int s = done % nstage;
            // the issue side packed min(P, generated - issued) pairs into this step; steps are only
            // short at the very end of the stream, so the same expression reproduces the count
            const int cnt = (done + 1 < issued) ? P : (issued_pairs - done_pairs);
            float* stage = stages + (size_t)s * stage_floats + (size_t)grp * R * K;
            const bool gvalid = grp < cnt;
            const int* e = ring + ((done_pairs + (gvalid ? grp : 0)) % PIPE_RING) * PIPE_ENTRY;
            const int ctok = e[1];
            pp_mbar_wait(bars + s, (uint32_t)((done / nstage) & 1));
            float u[CHUNKS][4], du[CHUNKS][4];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                for (int el = 0; el < 4; ++el) { u[c][el] = 0.f; du[c][el] = 0.f; }
                if (gvalid && act[c] && !(p.debug & 4)) pp_lds4(stage + (size_t)(n + 1) * K + coff[c], u[c]);
            }
            if (lg == 0 && gvalid) ++pairs;
            for (int rb = 0; rb <= n; rb += 8) {
                bool ract[8];
                float v[8][CHUNKS][4];
                float f[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int k = rb + r;
                    ract[r] = gvalid && (k <= n) && (k == 0 || e[4 + k - 1] != ctok);
                    float sacc = 0.f;
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                        for (int el = 0; el < 4; ++el) v[r][c][el] = 0.f;
                        if (ract[r] && act[c] && !(p.debug & 4)) {
                            pp_lds4(stage + (size_t)k * K + coff[c], v[r][c]);
#pragma unroll
                            for (int el = 0; el < 4; ++el) sacc = fmaf(u[c][el], v[r][c][el], sacc);
                        }
                    }
                    f[r] = sacc;
                }
                const float ftot = group_reduce8<G>(f, lane);            // total of row rb + row_of_lane
                const int myrow = rb + row_of_lane<G>(lane);
                const float mylabel = (myrow == 0) ? 1.f : 0.f;
                const bool myact = gvalid && (myrow <= n) && (myrow == 0 || e[4 + myrow - 1] != ctok);
                const float gmine = myact ? sgns_coeff(ftot, mylabel, p.alpha, p.max_grad, p.exp_table) : 0.f;
                if (p.compute_loss && myact && lg == lane_of_row<G>(row_of_lane<G>(lane))) {
                    loss += softplus_clipped(mylabel > 0.5f ? -ftot : ftot);
                    maxdot = fmaxf(maxdot, fabsf(ftot));
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float g = __shfl_sync(0xffffffffu, gmine, lane_of_row<G>(r), G);
                    if (!ract[r]) continue;
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) {
                        if (!act[c]) continue;
                        float gu[4];
#pragma unroll
                        for (int el = 0; el < 4; ++el) {
                            du[c][el] = fmaf(g, v[r][c][el], du[c][el]);
                            gu[el] = g * u[c][el];
                        }
                        pp_sts4(stage + (size_t)(rb + r) * K + coff[c], gu);     // row now holds g*u
                    }
                }
            }
            if (gvalid) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    if (act[c]) pp_sts4(stage + (size_t)(n + 1) * K + coff[c], du[c]);
            }
            pp_fence_async();                                // generic-proxy writes -> visible to the async proxy
            __syncwarp();
            if (tma_lane && tma_pair < cnt) {
                const int* e2 = ring + ((done_pairs + tma_pair) % PIPE_RING) * PIPE_ENTRY;
                const int ct = e2[1];
                float* sst = stages + (size_t)s * stage_floats + (size_t)(tma_pair * R + tma_row) * K;
                if (tma_row == 0) { if (!(p.debug & 1)) pp_bulk_reduce_add(p.syn1 + (size_t)ct * K, sst, row_bytes); }
                else if (tma_row <= n) {
                    const int ng = e2[4 + tma_row - 1];
                    if (ng != ct && !(p.debug & 1)) pp_bulk_reduce_add(p.syn1 + (size_t)ng * K, sst, row_bytes);
                } else if (!(p.debug & 2)) pp_bulk_reduce_add(p.syn0 + (size_t)e2[0] * K, sst, row_bytes);
            }
            pp_commit();
            done_pairs += cnt;
            ++done;
        }
    }
    pp_wait_all();                                           // smem must outlive the outstanding TMA reduces

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
}

// ------------------------------------------------------------------ host side

struct PipeLayout { int G, chunks, warps, stages, stage_floats, warp_bytes; size_t total; };

static void pipe_group(int K, int* G, int* chunks) {
    if (K <= 32) { *G = 8; *chunks = 1; }
    else if (K <= 64) { *G = 16; *chunks = 1; }
    else { *G = 32; *chunks = (K + 127) / 128; }
}

static PipeLayout pipe_layout(int K, int negatives, size_t smem_budget) {
    PipeLayout best{0, 0, 0, 0, 0, 0, 0};
    int G, chunks;
    pipe_group(K, &G, &chunks);
    const int P = 32 / G;
    const int stage_floats = P * (negatives + 2) * K;
    const size_t stage_bytes = (size_t)stage_floats * 4;
    const size_t fixed = 64 + (size_t)PIPE_RING * PIPE_ENTRY * 4;
    const int max_warps = (chunks >= 3) ? 8 : 16;
    long best_score = -1;
    for (int warps = max_warps; warps >= 2; --warps) {
        size_t per_warp = (smem_budget / warps) & ~(size_t)127;
        if (per_warp <= fixed + 2 * stage_bytes) continue;
        int stages = (int)((per_warp - fixed) / stage_bytes);
        if (stages > 8) stages = 8;
        long score = (long)warps * (stages > 4 ? 4 : stages) * 16 + warps;
        if (score > best_score) {
            best_score = score;
            size_t wb = (fixed + (size_t)stages * stage_bytes + 127) & ~(size_t)127;
            best = PipeLayout{G, chunks, warps, stages, stage_floats, (int)wb, wb * warps};
        }
    }
    return best;
}

bool sgns_pipe_supported(int K, int window, int negatives) {
    if (negatives < 1 || negatives > PIPE_MAXNEG) return false;
    if (2 * window + 1 > 32) return false;
    if (PIPE_GEN * 2 * window + 32 > PIPE_RING) return false;
    if (K % 4 != 0 || K > 1024) return false;
    int G, chunks;
    pipe_group(K, &G, &chunks);
    if ((32 / G) * (negatives + 2) > 32) return false;        // one TMA lane per (pair, row)
    PipeLayout l = pipe_layout(K, negatives, 220 * 1024);
    return l.warps >= 2;
}

#define GW2V_PIPE_DISPATCH(L, CALL)                                          \
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

int sgns_pipe_grid(int K, int negatives, int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    PipeLayout l = pipe_layout(K, negatives, 220 * 1024);
    int occ = 1;
#define CALL(GG, C)                                                                                          \
    do {                                                                                                     \
        cudaFuncSetAttribute(sgns_fused_pipe_kernel<GG, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                             (int)l.total);                                                                  \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sgns_fused_pipe_kernel<GG, C>, l.warps * 32,     \
                                                      l.total);                                              \
    } while (0)
    GW2V_PIPE_DISPATCH(l, CALL);
#undef CALL
    if (occ < 1) occ = 1;
    return sms * occ;
}

void launch_sgns_pipe(const SgnsParams& p, int grid, cudaStream_t stream) {
    PipeLayout l = pipe_layout(p.K, p.negatives, 220 * 1024);
#define CALL(GG, C)                                                                                          \
    do {                                                                                                     \
        cudaFuncSetAttribute(sgns_fused_pipe_kernel<GG, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                             (int)l.total);                                                                  \
        sgns_fused_pipe_kernel<GG, C><<<grid, l.warps * 32, l.total, stream>>>(p, l.stages, l.stage_floats,  \
                                                                              l.warp_bytes);                \
    } while (0)
    GW2V_PIPE_DISPATCH(l, CALL);
#undef CALL
}

}  // namespace gw2v
