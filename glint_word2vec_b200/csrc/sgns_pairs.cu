// sgns_pairs: the SGNS training step over a pre-generated pair-descriptor array (pairgen.cu).
//
// This is the production hot path: dotprod -> (all-reduce of the partial dots across the column
// shards) -> sigmoid / learning rate -> adjust of the reference (MLLIB:417-429 + the Glint server
// ops [G]), as ONE kernel per rank.
//
//   * purely pair-parallel: descriptor p = {centre word, context word, n negatives}; a GROUP of
//     G = 8/16/32 lanes owns a pair (G * 4 * CHUNKS >= K columns), a warp handles 32/G pairs per
//     step and walks the steps warp-strided -- no generation, no rings, few registers, so 24-48
//     warps per SM hide the latency of the random row gathers;
//   * rows move with 16-byte LDG.cg / RED.128 (L2 is the coherence point of the atomic updates);
//   * the 8 dots of a pair are reduced by one transposed butterfly, each sigmoid evaluated once;
//   * world > 1: every pair is visited twice.  Pass A computes the partial dots over this rank's
//     columns and, every 4 pairs, lane j stores the batch straight into rank j's symmetric exchange
//     slot over NVLink (st.global.v4 on a peer-mapped address) followed by st.release.sys of the
//     batch sequence number.  Pass B runs 16 pairs behind: it polls the peers' flags (ld.volatile +
//     one ld.acquire.sys), sums the S partials in fixed rank order (bit-identical coefficients on
//     every rank -- the reference's coefficient broadcast disappears), re-reads the rows (L2 hits)
//     and applies the updates.  Between A and B of a batch the warp keeps working on other pairs
//     and 24+ other warps are resident, so the NVLink round trip is off the critical path.  The
//     schedule is static per warp; sequence numbers persist across launches.  No NCCL on this path.
#include "common.cuh"
#include "sgns_params.h"
#include <cstdio>
#include <cstdlib>

namespace gw2v {

constexpr int PK_THREADS = 256;
constexpr int PK_FP = 8;              // floats per pair in exchange slots (1 + n <= 8)
constexpr int PK_BATCH = 4;           // pairs per batch: slot = 32 floats = 128 B
constexpr int PK_LAG = 16;            // max pairs between pass A and pass B (template parameter LAG: 4 or 16)
constexpr int PK_NSLOT = 14;          // >= 2 * (batches in flight + 1)
constexpr int PK_XCHUNKS = 11;        // 16-byte chunks {f, f, f, tag} carrying one batch (4 pairs x 8 dots = 32 floats)
constexpr int PK_XSTRIDE = 64;        // floats per (slot, source) region: 16 chunks, 256 B (the tagged-word format needs 16)
// Exchange formats (debug bit 32 selects the second):
//   fast  {f, f, f, tag} per 16-byte st.v4.f32 / ld.v4.u32.  Relies on an aligned 16-byte access being performed as ONE
//         transaction end to end (SM -> NVLink -> L2 -> SM); that is how the hardware behaves, but the PTX memory model
//         only promises single-copy atomicity up to 64 bits, so the engine runs xchg_selftest_kernel over every peer
//         pair at start-up and switches to the tagged-word format if it ever observes a torn chunk.
//   safe  {(f, tag), (f, tag)} per 16-byte st.v2.b64 / ld.v2.b64: every 64-bit element carries its own tag and is
//         single-copy atomic by the memory model; 16 chunks per batch instead of 11.
constexpr int PK_XCHUNKS_SAFE = 16;

template <int G> __device__ __forceinline__ int pk_row_of_lane(int lane) {
    const int lg = lane & (G - 1);
    return ((lg / (G / 2)) & 1) * 4 + ((lg / (G / 4)) & 1) * 2 + ((lg / (G / 8)) & 1);
}
template <int G> __device__ __forceinline__ int pk_lane_of_row(int r) {
    return ((r >> 2) & 1) * (G / 2) + ((r >> 1) & 1) * (G / 4) + (r & 1) * (G / 8);
}
// transposed butterfly over a group of G lanes: 8 values in, the lane keeps the total of row
// pk_row_of_lane<G>(lane); same pairing order as a plain xor butterfly (bit-identical totals)
template <int G>
__device__ __forceinline__ float pk_reduce8(const float (&f)[8], int lane) {
    float a[4];
    const bool h1 = lane & (G / 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float send = h1 ? f[j] : f[j + 4];
        const float keep = h1 ? f[j + 4] : f[j];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, G / 2);
    }
    float b[2];
    const bool h2 = lane & (G / 4);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float send = h2 ? a[j] : a[j + 2];
        const float keep = h2 ? a[j + 2] : a[j];
        b[j] = keep + __shfl_xor_sync(0xffffffffu, send, G / 4);
    }
    const bool h3 = lane & (G / 8);
    const float send = h3 ? b[0] : b[1];
    const float keep = h3 ? b[1] : b[0];
    float c = keep + __shfl_xor_sync(0xffffffffu, send, G / 8);
#pragma unroll
    for (int o = G / 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

__device__ __forceinline__ void pk_ld4(const float* p, float (&o)[4]) {
    float4 v = __ldcg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void pk_red4(float* p, const float (&v)[4]) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}

// descriptor words 0..11 -> row index of row r is word r + 1 (row 0 = context = word 1)
struct PairDesc { int w[12]; };
__device__ __forceinline__ void pk_load_desc(const int* __restrict__ desc, long long pair, int pd, PairDesc& d) {
    const int4* e = reinterpret_cast<const int4*>(desc + (size_t)pair * pd);
    const int4 a = __ldg(e), b = __ldg(e + 1);
    d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w;
    d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w;
    if (pd > 8) { const int4 c = __ldg(e + 2); d.w[8] = c.x; d.w[9] = c.y; d.w[10] = c.z; d.w[11] = c.w; }
    else { d.w[8] = d.w[9] = d.w[10] = d.w[11] = -1; }             // -1 = unused negative slot
}

// ------------------------------------------------------------------------------------------ single shard
template <int G, int CHUNKS, int MINB>
__global__ void __launch_bounds__(PK_THREADS, MINB)
sgns_pairs_kernel(const SgnsParams p, const int* __restrict__ desc, const int* __restrict__ n_pairs_ptr, const int pd,
                  const int splits) {
    constexpr int P = 32 / G;
    const int lane = threadIdx.x & 31;
    const int K = p.K;
    const int n = p.negatives;
    const int grp = lane / G, lg = lane % G;
    bool act[CHUNKS];
    int coff[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) { coff[c] = (c * G + lg) * 4; act[c] = coff[c] < K; }
    const long long n_real_pairs = *n_pairs_ptr;
    const long long n_pairs = n_real_pairs * splits;              // descriptors (n > 7: several per pair, pairgen.cu)
    const long long nsteps = (n_pairs + P - 1) / P;
    const long long warp_global = ((long long)blockIdx.x * PK_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * PK_THREADS) >> 5;
    float loss = 0.f, maxdot = 0.f;
    const int myrow = pk_row_of_lane<G>(lane);
    const bool owner = lg == pk_lane_of_row<G>(myrow);

    for (long long step = warp_global; step < nsteps; step += n_warps) {
        const long long pair = step * P + grp;
        const bool gvalid = pair < n_pairs;
        PairDesc d;
        pk_load_desc(desc, gvalid ? pair : 0, pd, d);
        const int ctok = d.w[1] & 0x7fffffff;                     // bit 31: context row inactive (continuation descriptor)
        const bool ctx_on = d.w[1] >= 0;
        d.w[1] = ctok;
        float* urow = p.syn0 + (size_t)d.w[0] * K;
        float u[CHUNKS][4], du[CHUNKS][4];
        float v[8][CHUNKS][4];
        bool ract[8];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
            for (int el = 0; el < 4; ++el) { u[c][el] = 0.f; du[c][el] = 0.f; }
            if (gvalid && act[c]) pk_ld4(urow + coff[c], u[c]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            ract[r] = gvalid && (r == 0 ? ctx_on : (d.w[r + 1] >= 0 && d.w[r + 1] != ctok));   // unused slots hold -1
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                for (int el = 0; el < 4; ++el) v[r][c][el] = 0.f;
                if (ract[r] && act[c] && !(p.debug & 4)) pk_ld4(p.syn1 + (size_t)d.w[r + 1] * K + coff[c], v[r][c]);
            }
        }
        float f[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
                for (int el = 0; el < 4; ++el) s = fmaf(u[c][el], v[r][c][el], s);
            f[r] = s;
        }
        const float ftot = pk_reduce8<G>(f, lane);
        unsigned actmask = 0;                       // bit r: row r takes part (no runtime indexing of d.w)
#pragma unroll
        for (int r = 0; r < 8; ++r) actmask |= ract[r] ? (1u << r) : 0u;
        const bool myact = (actmask >> myrow) & 1u;
        const float mylabel = (myrow == 0) ? 1.f : 0.f;
        const float gmine = myact ? sgns_coeff(ftot, mylabel, p.alpha, p.max_grad, p.exp_table) : 0.f;
        if (p.compute_loss && myact && owner) {
            loss += softplus_clipped(mylabel > 0.5f ? -ftot : ftot);
            maxdot = fmaxf(maxdot, fabsf(ftot));
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float g = __shfl_sync(0xffffffffu, gmine, pk_lane_of_row<G>(r), G);
            if (!ract[r]) continue;
            float* vrow = p.syn1 + (size_t)d.w[r + 1] * K;
            const float gs = g * row_scale(p.row_scale1, p.hot_rows, d.w[r + 1]);       // hot-row damping of this v row
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                if (!act[c]) continue;
                float gu[4];
#pragma unroll
                for (int el = 0; el < 4; ++el) {
                    du[c][el] = fmaf(g, v[r][c][el], du[c][el]);
                    gu[el] = gs * u[c][el];
                }
                if (!(p.debug & 1)) pk_red4(vrow + coff[c], gu);
            }
        }
        if (gvalid && !(p.debug & 2)) {
            const float su = row_scale(p.row_scale0, p.hot_rows, d.w[0]);
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                if (!act[c]) continue;
#pragma unroll
                for (int el = 0; el < 4; ++el) du[c][el] *= su;
                pk_red4(urow + coff[c], du[c]);
            }
        }
    }

    if (blockIdx.x == 0 && threadIdx.x == 0) { p.stats[3] = (float)(*p.n_tokens); p.stats[0] = (float)n_real_pairs; }
    if (p.compute_loss) {
        loss = warp_sum(loss);
        maxdot = warp_max(maxdot);
        if (lane == 0) {
            if (loss != 0.f) atomicAdd(p.stats + 1, loss);
            atomicMax(reinterpret_cast<int*>(p.stats + 2), __float_as_int(maxdot));
        }
    }
}

// ------------------------------------------------------------------------------------------ column shards
struct PkWarpSmem {
    float fdot[2 * PK_LAG * PK_FP];      // partial (then total) dots of the pairs between pass A and pass B
    float xsum[8 * 32];                  // per-source copy of one batch for the ordered reduction
    int dstash[2 * PK_LAG * 12];         // pair descriptors between pass A and pass B (saves an L2 round trip)
};

// LAG = pairs between pass A and pass B of a warp.  The rows touched in between must still be in L2 when
// pass B re-gathers them: in-flight footprint = LAG x warps x (2+n) rows.  LAG 16 at K=256 is 265 MB (> the
// 126 MB L2, pass B then re-reads DRAM: measured 1.21 ms vs 0.62 ms for the exchange-free kernel); one batch
// (LAG 4, ~35 us of work per warp) already covers the ~2-3 us NVLink round trip many times over.
// (Issuing pass B's row re-gather BEFORE pass A of the same iteration - to overlap its L2 latency with pass A's
// DRAM gathers - was measured and is slower at every occupancy: K=64 0.332 -> 0.398 ms, K=128 0.507 -> 0.694 ms.)
template <int G, int CHUNKS, int MINB, int LAG>
__global__ void __launch_bounds__(PK_THREADS, MINB)
sgns_pairs_multi_kernel(const SgnsParams p, const int* __restrict__ desc, const int* __restrict__ n_pairs_ptr,
                        const int pd, uint32_t* warp_seq, const int splits) {
    constexpr int P = 32 / G;
    constexpr int BS = PK_BATCH / P;           // steps per batch
    static_assert(LAG % PK_BATCH == 0 && LAG <= PK_LAG, "lag must be whole batches");
    constexpr int LAGS = LAG / P;              // lag in steps
    constexpr int RFS = 2 * LAGS;              // dot ring in steps (multiple of BS)
    __shared__ __align__(16) PkWarpSmem wsm[PK_THREADS / 32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* fdot = wsm[warp].fdot;
    float* xsum = wsm[warp].xsum;
    int* dstash = wsm[warp].dstash;
    const int K = p.K;
    const int n = p.negatives;
    const int S = p.world;
    const int rank = p.rank;
    const int grp = lane / G, lg = lane % G;
    bool act[CHUNKS];
    int coff[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) { coff[c] = (c * G + lg) * 4; act[c] = coff[c] < K; }
    const long long n_real_pairs = *n_pairs_ptr;
    const long long n_pairs = n_real_pairs * splits;              // descriptors (n > 7: several per pair, pairgen.cu)
    const long long nsteps = (n_pairs + P - 1) / P;
    const int gwarp = blockIdx.x * (PK_THREADS / 32) + warp;            // identical on every rank
    const int n_warps = gridDim.x * (PK_THREADS / 32);
    const int nk = (gwarp < nsteps) ? (int)((nsteps - gwarp + n_warps - 1) / n_warps) : 0;   // this warp's steps
    const int myrow = pk_row_of_lane<G>(lane);
    const bool owner = lg == pk_lane_of_row<G>(myrow);

    const size_t slot_stride = PK_XSTRIDE;                                       // floats per (slot, source)
    const size_t warp_x_base = (size_t)gwarp * PK_NSLOT * S * slot_stride;
    const uint32_t seq0 = warp_seq[gwarp];
    float loss = 0.f, maxdot = 0.f;
    unsigned long long wait_ns = 0;

    for (int k = 0; k < nk + LAGS; ++k) {
        const int kb = k - LAGS;
        // ---- pass B operands (descriptor from the shared-memory stash, rows re-gathered from L2)
        PairDesc dB;
        bool gvalidB = false;
        float uB[CHUNKS][4];
        float vB[8][CHUNKS][4];
        bool ractB[8];
        auto load_b = [&]() {
            const long long pairB = ((long long)gwarp + (long long)kb * n_warps) * P + grp;
            gvalidB = pairB < n_pairs;
            const int4* st = reinterpret_cast<const int4*>(dstash + ((kb % RFS) * P + grp) * 12);
            const int4 a = st[0], b2 = st[1], c2 = st[2];
            dB.w[0] = a.x; dB.w[1] = a.y; dB.w[2] = a.z; dB.w[3] = a.w;
            dB.w[4] = b2.x; dB.w[5] = b2.y; dB.w[6] = b2.z; dB.w[7] = b2.w;
            dB.w[8] = c2.x; dB.w[9] = c2.y; dB.w[10] = c2.z; dB.w[11] = c2.w;
            const int ctokB = dB.w[1] & 0x7fffffff;
            const bool ctxB = dB.w[1] >= 0;
            dB.w[1] = ctokB;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                for (int el = 0; el < 4; ++el) uB[c][el] = 0.f;
                if (gvalidB && act[c]) pk_ld4(p.syn0 + (size_t)dB.w[0] * K + coff[c], uB[c]);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                ractB[r] = gvalidB && (r == 0 ? ctxB : (dB.w[r + 1] >= 0 && dB.w[r + 1] != ctokB));
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                    for (int el = 0; el < 4; ++el) vB[r][c][el] = 0.f;
                    if (ractB[r] && act[c]) pk_ld4(p.syn1 + (size_t)dB.w[r + 1] * K + coff[c], vB[r][c]);
                }
            }
        };
        if (k < nk) {
            // ------------------------------------------------------------ pass A: partial dots of step k
            const long long pair = ((long long)gwarp + (long long)k * n_warps) * P + grp;
            const bool gvalid = pair < n_pairs;
            PairDesc d;
            pk_load_desc(desc, gvalid ? pair : 0, pd, d);
            const int ctok = d.w[1] & 0x7fffffff;                 // the stash keeps the raw word (flag included) for pass B
            const bool ctx_on = d.w[1] >= 0;
            if (lg == 0) {
                int4* st = reinterpret_cast<int4*>(dstash + ((k % RFS) * P + grp) * 12);
                st[0] = make_int4(d.w[0], d.w[1], d.w[2], d.w[3]);
                st[1] = make_int4(d.w[4], d.w[5], d.w[6], d.w[7]);
                st[2] = make_int4(d.w[8], d.w[9], d.w[10], d.w[11]);
            }
            float u[CHUNKS][4];
            float v[8][CHUNKS][4];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                for (int el = 0; el < 4; ++el) u[c][el] = 0.f;
                if (gvalid && act[c]) pk_ld4(p.syn0 + (size_t)d.w[0] * K + coff[c], u[c]);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const bool ra = gvalid && (r == 0 ? ctx_on : (d.w[r + 1] >= 0 && d.w[r + 1] != ctok));
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                    for (int el = 0; el < 4; ++el) v[r][c][el] = 0.f;
                    if (ra && act[c]) pk_ld4(p.syn1 + (size_t)(r == 0 ? ctok : d.w[r + 1]) * K + coff[c], v[r][c]);
                }
            }
            float f[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
                    for (int el = 0; el < 4; ++el) s = fmaf(u[c][el], v[r][c][el], s);
                f[r] = s;
            }
            const float tot = pk_reduce8<G>(f, lane);
            if (owner) fdot[((k % RFS) * P + grp) * PK_FP + myrow] = gvalid ? tot : 0.f;
            if ((k + 1) % BS == 0 || k == nk - 1) {
                // ---- push the batch.  No fence and no separate flag: every 16-byte chunk carries its own
                // sequence tag {f, f, f, tag} (one st.v4 = one NVLink write), the receiver polls the data itself.
                // (A st.release.sys flag costs a membar.sys round trip per batch: it was the top stall reason,
                // 6.3 stalled warps per issue in profiles/r1_run20_multi_ncu.md.)
                __syncwarp();
                const int b = k / BS;
                const uint32_t bseq = seq0 + (uint32_t)b;
                const int slot = (int)(bseq % (uint32_t)PK_NSLOT);
                if (p.debug & 16) {
                    // protocol stress test: pseudo-random delay (0-16 us) per warp, batch and rank before the push,
                    // so ranks drift by many batches and the slot ring / tags are exercised out of lock-step
                    const uint32_t h = ((uint32_t)gwarp * 2654435761u) ^ (bseq * 40503u) ^ ((uint32_t)rank * 0x9E3779B9u);
                    __nanosleep((h >> 9) & 0x3FFFu);
                }
                if (p.debug & 32) {
                    if (lane < PK_XCHUNKS_SAFE) {
                        const float* src = fdot + (size_t)((b * BS) % RFS) * P * PK_FP + 2 * lane;
                        const unsigned long long tg = (unsigned long long)(bseq + 1u) << 32;
                        const unsigned long long w0 = tg | (unsigned long long)__float_as_uint(src[0]);
                        const unsigned long long w1 = tg | (unsigned long long)__float_as_uint(src[1]);
                        for (int r = 0; r < S; ++r) {
                            if (r == rank) continue;
                            const int srcidx = (p.debug & 8) ? r : rank;
                            float* dst = p.xbuf[r] + warp_x_base + ((size_t)slot * S + srcidx) * slot_stride + 4 * lane;
                            asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
                        }
                    }
                } else if (lane < PK_XCHUNKS) {
                    const float* src = fdot + (size_t)((b * BS) % RFS) * P * PK_FP + 3 * lane;
                    const float x0 = src[0], x1 = src[1], x2 = (3 * lane + 2 < 32) ? src[2] : 0.f;
                    const uint32_t tag = bseq + 1u;
                    if (p.xbuf_mc != nullptr) {                          // (fast format only; the engine never combines it with bit 32)
                        // NVLS: one multicast store lands in every rank's slot (switch-replicated)
                        float* dst = p.xbuf_mc + warp_x_base + ((size_t)slot * S + rank) * slot_stride + 4 * lane;
                        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                                     ::"l"(dst), "f"(x0), "f"(x1), "f"(x2), "f"(__uint_as_float(tag)) : "memory");
                    } else {
                        for (int r = 0; r < S; ++r) {
                            if (r == rank) continue;
                            // debug bit 3 (profiling on ONE GPU): every xbuf pointer is this GPU's own buffer and
                            // the store to "rank r" plays the message FROM rank r, so the kernel feeds itself
                            const int srcidx = (p.debug & 8) ? r : rank;
                            float* dst = p.xbuf[r] + warp_x_base + ((size_t)slot * S + srcidx) * slot_stride + 4 * lane;
                            asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                                         ::"l"(dst), "f"(x0), "f"(x1), "f"(x2), "f"(__uint_as_float(tag)) : "memory");
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (kb >= 0) {
            // ------------------------------------------------------------ pass B: reduce + update step kb
            if (kb % BS == 0) {
                const int b = kb / BS;
                const uint32_t bseq = seq0 + (uint32_t)b;
                const int slot = (int)(bseq % (uint32_t)PK_NSLOT);
                float* mine = fdot + (size_t)((b * BS) % RFS) * P * PK_FP;       // 32 floats of this batch
                {
                    const uint32_t tag = bseq + 1u;
                    unsigned long long t0 = p.timing ? globaltimer_ns() : 0ull;
                    if (p.debug & 32) {
                        // tagged-word format: (S-1) sources x 16 chunks of two (value, tag) words
                        for (int it = lane; it < (S - 1) * PK_XCHUNKS_SAFE; it += 32) {
                            const int q = it / PK_XCHUNKS_SAFE, c = it - q * PK_XCHUNKS_SAFE;
                            const int srcr = q + (q >= rank ? 1 : 0);
                            const float* ptr = p.xbuf[rank] + warp_x_base + ((size_t)slot * S + srcr) * slot_stride + 4 * c;
                            unsigned long long w0, w1;
                            uint32_t spins = 0;
                            unsigned long long tw = 0ull;
                            while (true) {
                                asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(ptr) : "memory");
                                if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag) break;
                                if ((++spins & 0x3FFFu) == 0) {
                                    if (tw == 0ull) tw = globaltimer_ns();
                                    if (globaltimer_ns() - tw > 20000000000ull) {
                                        printf("[gw2v] rank %d warp %d: timeout waiting for rank %d batch %u\n", rank, gwarp, srcr, tag);
                                        atomicExch(p.error_flag, 1);
                                        __trap();
                                    }
                                }
                            }
                            xsum[srcr * 32 + 2 * c] = __uint_as_float((uint32_t)w0);
                            xsum[srcr * 32 + 2 * c + 1] = __uint_as_float((uint32_t)w1);
                        }
                    } else
                    // (S-1) sources x 11 chunks, strided over the lanes; each chunk is polled until its tag matches
                    for (int it = lane; it < (S - 1) * PK_XCHUNKS; it += 32) {
                        const int q = it / PK_XCHUNKS, c = it - q * PK_XCHUNKS;
                        const int srcr = q + (q >= rank ? 1 : 0);
                        const float* ptr = p.xbuf[rank] + warp_x_base + ((size_t)slot * S + srcr) * slot_stride + 4 * c;
                        uint32_t g0, g1, g2, g3, spins = 0;
                        unsigned long long tw = 0ull;
                        while (true) {
                            asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(g0), "=r"(g1), "=r"(g2), "=r"(g3) : "l"(ptr) : "memory");
                            if (g3 == tag) break;
                            if ((++spins & 0x3FFFu) == 0) {
                                if (tw == 0ull) tw = globaltimer_ns();
                                if (globaltimer_ns() - tw > 20000000000ull) {
                                    printf("[gw2v] rank %d warp %d: timeout waiting for rank %d batch %u (tag %u)\n",
                                           rank, gwarp, srcr, tag, g3);
                                    atomicExch(p.error_flag, 1);
                                    __trap();
                                }
                            }
                        }
                        float* xs = xsum + srcr * 32 + 3 * c;
                        xs[0] = __uint_as_float(g0);
                        xs[1] = __uint_as_float(g1);
                        if (3 * c + 2 < 32) xs[2] = __uint_as_float(g2);
                    }
                    if (p.timing) wait_ns += globaltimer_ns() - t0;
                }
                xsum[rank * 32 + lane] = mine[lane];
                __syncwarp();
                float tot = 0.f;
                for (int r = 0; r < S; ++r) tot += xsum[r * 32 + lane];      // fixed order: bit-identical on all ranks
                mine[lane] = tot;
                __syncwarp();
            }
            load_b();
            const PairDesc& d = dB;
            const bool gvalid = gvalidB;
            float* urow = p.syn0 + (size_t)d.w[0] * K;
            float du[CHUNKS][4];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c)
#pragma unroll
                for (int el = 0; el < 4; ++el) du[c][el] = 0.f;
            auto& u = uB;
            auto& v = vB;
            auto& ract = ractB;
            const float fm = fdot[((kb % RFS) * P + grp) * PK_FP + myrow];
            unsigned actmask = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) actmask |= ract[r] ? (1u << r) : 0u;
            const bool myact = (actmask >> myrow) & 1u;
            const float mylabel = (myrow == 0) ? 1.f : 0.f;
            const float gmine = myact ? sgns_coeff(fm, mylabel, p.alpha, p.max_grad, p.exp_table) : 0.f;
            if (p.compute_loss && myact && owner) {
                loss += softplus_clipped(mylabel > 0.5f ? -fm : fm);
                maxdot = fmaxf(maxdot, fabsf(fm));
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float g = __shfl_sync(0xffffffffu, gmine, pk_lane_of_row<G>(r), G);
                if (!ract[r]) continue;
                float* vrow = p.syn1 + (size_t)d.w[r + 1] * K;
                const float gs = g * row_scale(p.row_scale1, p.hot_rows, d.w[r + 1]);   // hot-row damping of this v row
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    if (!act[c]) continue;
                    float gu[4];
#pragma unroll
                    for (int el = 0; el < 4; ++el) {
                        du[c][el] = fmaf(g, v[r][c][el], du[c][el]);
                        gu[el] = gs * u[c][el];
                    }
                    if (!(p.debug & 1)) pk_red4(vrow + coff[c], gu);
                }
            }
            if (gvalid && !(p.debug & 2)) {
                const float su = row_scale(p.row_scale0, p.hot_rows, d.w[0]);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    if (!act[c]) continue;
#pragma unroll
                    for (int el = 0; el < 4; ++el) du[c][el] *= su;
                    pk_red4(urow + coff[c], du[c]);
                }
            }
        }
    }
    if (lane == 0) warp_seq[gwarp] = seq0 + (uint32_t)((nk + BS - 1) / BS);

    if (blockIdx.x == 0 && threadIdx.x == 0) { p.stats[3] = (float)(*p.n_tokens); p.stats[0] = (float)n_real_pairs; }
    if (p.compute_loss) {
        loss = warp_sum(loss);
        maxdot = warp_max(maxdot);
        if (lane == 0) {
            if (loss != 0.f) atomicAdd(p.stats + 1, loss);
            atomicMax(reinterpret_cast<int*>(p.stats + 2), __float_as_int(maxdot));
        }
    }
    if (p.timing) {
        unsigned long long w = 0;
        for (int o = 0; o < 32; ++o) { unsigned long long x = __shfl_sync(0xffffffffu, wait_ns, o); w = x > w ? x : w; }
        if (lane == 0 && w) atomicAdd(p.timing + 0, w);
    }
}

// ------------------------------------------------------------------------------------------ host side
static void pk_group(int K, int* G, int* chunks) {
    if (K <= 32) { *G = 8; *chunks = 1; }
    else if (K <= 64) { *G = 16; *chunks = 1; }
    else { *G = 32; *chunks = (K + 127) / 128; }
}

bool sgns_pairs_supported(int K, int window, int negatives) {
    // more than 7 negatives per pair travel as several descriptors (pairgen.cu): up to 21
    return negatives >= 1 && negatives <= 21 && window >= 1 && window <= 11 && K % 4 == 0 && K <= 1024;
}

static int pk_occ() {
    static int occ = -1;
    if (occ < 0) { const char* e = getenv("GW2V_PAIRS_OCC"); occ = e ? atoi(e) : 3; }
    return occ;
}
// pairs between pass A and pass B of the column-shard kernel (4 = one batch, 16 = four batches)
static int pk_lag() {
    static int lag = -1;
    if (lag < 0) { const char* e = getenv("GW2V_PAIRS_LAG"); lag = (e && atoi(e) >= 16) ? 16 : 4; }
    return lag;
}

// kernel variant selection for the column-shard kernel: LAG in {4, 16}
template <int GG, int C, int MB, typename F>
static void pk_multi_select(F&& f) {
    if (pk_lag() >= 16) f(sgns_pairs_multi_kernel<GG, C, MB, 16>);
    else f(sgns_pairs_multi_kernel<GG, C, MB, 4>);
}

#define GW2V_PK_DISPATCH(K, CALL)                                            \
    do {                                                                     \
        int G_, ch_;                                                         \
        pk_group((K), &G_, &ch_);                                            \
        const int o_ = pk_occ();                                             \
        if (G_ == 8) { if (o_ >= 4) CALL(8, 1, 4); else if (o_ == 3) CALL(8, 1, 3); else CALL(8, 1, 2); }        \
        else if (G_ == 16) { if (o_ >= 4) CALL(16, 1, 4); else if (o_ == 3) CALL(16, 1, 3); else CALL(16, 1, 2); } \
        else if (ch_ == 1) { if (o_ >= 4) CALL(32, 1, 4); else if (o_ == 3) CALL(32, 1, 3); else CALL(32, 1, 2); } \
        else if (ch_ == 2) { CALL(32, 2, 2); }                               \
        else if (ch_ == 3) { CALL(32, 3, 1); }                               \
        else if (ch_ == 4) { CALL(32, 4, 1); }                               \
        else if (ch_ <= 6) { CALL(32, 6, 1); }                               \
        else { CALL(32, 8, 1); }                                             \
    } while (0)

int sgns_pairs_grid(int K, int device, bool multi) {
    int sms = 148, occ = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (multi) {
#define CALL(GG, C, MB) pk_multi_select<GG, C, MB>([&](auto kern) {                      \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PK_THREADS, 0); })
        GW2V_PK_DISPATCH(K, CALL);
#undef CALL
    } else {
#define CALL(GG, C, MB) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sgns_pairs_kernel<GG, C, MB>, PK_THREADS, 0)
        GW2V_PK_DISPATCH(K, CALL);
#undef CALL
    }
    if (occ < 1) occ = 1;
    return sms * occ;               // multi: every CTA co-resident, required by the in-kernel flag protocol
}

// Start-up self-test of the fast exchange format (see PK_XSTRIDE): block 0 of every rank streams `iters` chunks
// {i, i, i, i} into each peer's test slot with the same st.relaxed.sys.v4 the training kernel uses, blocks 1.. poll
// their own slots (one warp per source) with the same ld.relaxed.sys.v4 and count every observation whose four words
// differ.  All ranks run it concurrently; result[0] = torn observations, result[1] = chunks observed.
__global__ void xchg_selftest_kernel(PeerTest t) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0) {
        // sender: warp w -> peer w; lanes write 32 different chunks so that many 16-byte stores are in flight
        for (int r = warp; r < t.world; r += blockDim.x / 32) {
            if (r == t.rank) continue;
            uint32_t* dst = t.buf[r] + ((size_t)t.rank * 32 + lane) * 4;
            for (uint32_t i = 1; i <= (uint32_t)t.iters; ++i)
                asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(dst), "r"(i) : "memory");
        }
    } else {
        for (int src = warp; src < t.world; src += blockDim.x / 32) {
            if (src == t.rank) continue;
            const uint32_t* ptr = t.buf[t.rank] + ((size_t)src * 32 + lane) * 4;
            unsigned long long torn = 0, seen = 0;
            const unsigned long long t0 = globaltimer_ns();
            while (true) {
                uint32_t a, b, c, d;
                asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(ptr) : "memory");
                ++seen;
                if (a != b || b != c || c != d) ++torn;
                if (a == (uint32_t)t.iters && b == a && c == a && d == a) break;
                if ((seen & 0xFFFu) == 0 && globaltimer_ns() - t0 > 5000000000ull) { torn |= 1ull << 62; break; }   // peer never finished
            }
            atomicAdd(t.result, torn);
            atomicAdd(t.result + 1, seen);
        }
    }
}

void launch_xchg_selftest(const PeerTest& t, cudaStream_t stream) {
    xchg_selftest_kernel<<<2, 256, 0, stream>>>(t);
}

void sgns_pairs_multi_geometry(int* warps_per_cta, int* nslot, int* slot_floats) {
    *warps_per_cta = PK_THREADS / 32;
    *nslot = PK_NSLOT;
    *slot_floats = PK_XSTRIDE;
}

void launch_sgns_pairs(const SgnsParams& p, const int* desc, const int* n_pairs, int pd, int grid, cudaStream_t stream) {
    const int splits = p.negatives <= 7 ? 1 : (p.negatives + 6) / 7;
#define CALL(GG, C, MB) sgns_pairs_kernel<GG, C, MB><<<grid, PK_THREADS, 0, stream>>>(p, desc, n_pairs, pd, splits)
    GW2V_PK_DISPATCH(p.K, CALL);
#undef CALL
}

void launch_sgns_pairs_multi(const SgnsParams& p, const int* desc, const int* n_pairs, int pd, int grid,
                             uint32_t* warp_seq, cudaStream_t stream) {
    const int splits = p.negatives <= 7 ? 1 : (p.negatives + 6) / 7;
#define CALL(GG, C, MB) pk_multi_select<GG, C, MB>([&](auto kern) {                      \
        kern<<<grid, PK_THREADS, 0, stream>>>(p, desc, n_pairs, pd, warp_seq, splits); })
    GW2V_PK_DISPATCH(p.K, CALL);
#undef CALL
}

}  // namespace gw2v
