// sgns_fused_group: single-shard SGNS step, register path with lane groups.
//
// Measured on B200 (profiles/): cp.async.bulk row copies cost ~25 SM cycles of TMA issue each, so for
// short rows (K <= 128, i.e. the per-GPU slice of a column-sharded model) the TMA pipeline
// (sgns_pipe.cu) is bound by TMA operations, while the original warp-per-centre kernel
// (sgns_kernels.cu) is bound by its ~600-1300 instructions per pair.  This kernel keeps the plain
// LDG / RED.128 memory path (no shared-memory staging, so 24-48 warps per SM hide latency) and
// removes the instruction overhead:
//   * a pair (1 context + n negative rows + the centre row) is owned by a GROUP of G = 8/16/32
//     lanes, so one warp instruction serves 32/G pairs;
//   * windows / negatives are generated for several centres at once, all lanes sharing the Philox
//     and alias-table work (pipe_common.cuh);
//   * the 8 dots of a pair are reduced by one transposed butterfly, each sigmoid is evaluated once.
// Semantics: per-pair private negatives; u is re-read and du applied per pair (word2vec.c order).
#include "pipe_common.cuh"
#include <cstdlib>

namespace gw2v {

constexpr int GK_GEN = 4;
constexpr int GK_RING = 64;
constexpr int GK_THREADS = 256;

__device__ __forceinline__ void gk_ld4(const float* p, float (&o)[4]) {
    float4 v = __ldcg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void gk_red4(float* p, const float (&v)[4]) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}

// MINB = resident CTAs per SM the register allocation is capped for (short rows: 3 -> 80 registers,
// 4 -> 64 registers with a few spills; long rows need the registers for the row data)
template <int G, int CHUNKS, int MINB>
__global__ void __launch_bounds__(GK_THREADS, MINB)
sgns_fused_group_kernel(const SgnsParams p) {
    constexpr int P = 32 / G;
    __shared__ int ring_all[(GK_THREADS / 32) * GK_RING * PIPE_ENTRY];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int* ring = ring_all + warp * GK_RING * PIPE_ENTRY;

    const int K = p.K;
    const int n = p.negatives;
    const int T = *p.n_tokens;
    const int maxgen = GK_GEN * 2 * p.window;
    const int grp = lane / G, lg = lane % G;
    bool act[CHUNKS];
    int coff[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) { coff[c] = (c * G + lg) * 4; act[c] = coff[c] < K; }

    const int n_warps = gridDim.x * (GK_THREADS / 32);
    int gen_i = blockIdx.x * (GK_THREADS / 32) + warp;
    int head = 0, done = 0;
    float loss = 0.f, maxdot = 0.f;
    unsigned pairs = 0;

    while (true) {
        while (gen_i < T && (GK_RING - (head - done)) >= maxgen)
            head += generate_pairs<GK_GEN, GK_RING>(p, T, gen_i, n_warps, ring, head, lane);
        if (head == done) {
            if (gen_i >= T) break;
            continue;
        }
        const int cnt = min(P, head - done);
        const bool gvalid = grp < cnt;
        const int* e = ring + ((done + (gvalid ? grp : 0)) % GK_RING) * PIPE_ENTRY;
        const int wtok = e[0], ctok = e[1];
        float* urow = p.syn0 + (size_t)wtok * K;
        float u[CHUNKS][4], du[CHUNKS][4];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
            for (int el = 0; el < 4; ++el) { u[c][el] = 0.f; du[c][el] = 0.f; }
            if (gvalid && act[c]) gk_ld4(urow + coff[c], u[c]);
        }
        if (lg == 0 && gvalid) ++pairs;
        for (int rb = 0; rb <= n; rb += 8) {
            bool ract[8];
            int rows[8];
            float v[8][CHUNKS][4];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int k = rb + r;
                rows[r] = (k == 0 || k > n) ? ctok : e[4 + k - 1];
                ract[r] = gvalid && (k <= n) && (k == 0 || rows[r] != ctok);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                    for (int el = 0; el < 4; ++el) v[r][c][el] = 0.f;
                    if (ract[r] && act[c] && !(p.debug & 4)) gk_ld4(p.syn1 + (size_t)rows[r] * K + coff[c], v[r][c]);
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
            const float ftot = group_reduce8<G>(f, lane);
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
                    if (!(p.debug & 1)) gk_red4(vrow + coff[c], gu);
                }
            }
        }
        if (gvalid && !(p.debug & 2)) {
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c)
                if (act[c]) gk_red4(urow + coff[c], du[c]);
        }
        done += cnt;
    }

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

static void gk_group(int K, int* G, int* chunks) {
    if (K <= 32) { *G = 8; *chunks = 1; }
    else if (K <= 64) { *G = 16; *chunks = 1; }
    else { *G = 32; *chunks = (K + 127) / 128; }
}

bool sgns_group_supported(int K, int window, int negatives) {
    if (negatives < 1 || negatives > PIPE_MAXNEG) return false;
    if (2 * window + 1 > 32) return false;
    if (GK_GEN * 2 * window + 8 > GK_RING) return false;
    return K % 4 == 0 && K <= 1024;
}

static int gk_occ() {
    static int occ = -1;
    if (occ < 0) { const char* e = getenv("GW2V_GROUP_OCC"); occ = e ? atoi(e) : 3; }
    return occ;
}

#define GW2V_GK_DISPATCH(K, CALL)                                            \
    do {                                                                     \
        int G_, ch_;                                                         \
        gk_group((K), &G_, &ch_);                                            \
        const bool o4 = gk_occ() >= 4;                                       \
        if (G_ == 8) { if (o4) CALL(8, 1, 4); else CALL(8, 1, 3); }          \
        else if (G_ == 16) { if (o4) CALL(16, 1, 4); else CALL(16, 1, 3); }  \
        else if (ch_ == 1) { if (o4) CALL(32, 1, 4); else CALL(32, 1, 3); }  \
        else if (ch_ == 2) { CALL(32, 2, 2); }                               \
        else if (ch_ == 3) { CALL(32, 3, 1); }                               \
        else if (ch_ == 4) { CALL(32, 4, 1); }                               \
        else if (ch_ <= 6) { CALL(32, 6, 1); }                               \
        else { CALL(32, 8, 1); }                                             \
    } while (0)

int sgns_group_grid(int K, int device) {
    int sms = 148, occ = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
#define CALL(GG, C, MB) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sgns_fused_group_kernel<GG, C, MB>, GK_THREADS, 0)
    GW2V_GK_DISPATCH(K, CALL);
#undef CALL
    if (occ < 1) occ = 1;
    return sms * occ;
}

void launch_sgns_group(const SgnsParams& p, int grid, cudaStream_t stream) {
#define CALL(GG, C, MB) sgns_fused_group_kernel<GG, C, MB><<<grid, GK_THREADS, 0, stream>>>(p)
    GW2V_GK_DISPATCH(p.K, CALL);
#undef CALL
}

}  // namespace gw2v
