// sgns_tile: the SGNS training step on the 5th-generation tensor cores (neg_sharing = "tile").
//
// Reference hot path: dotprod -> (sum of the partial dots over the column shards) -> sigmoid / learning rate ->
// adjust (MLLIB:417-429 + the Glint server ops [G]).  With the reference's per-pair private negatives every dot is
// a distinct row pair and there is no GEMM; with the negatives shared by a TILE of T = 128 consecutive centres
// (models/sgns.py, neg_sharing="tile": NN negatives per tile, negative term of centre i weighted m_i * n / NN) the
// step of a tile is GEMM-shaped over gathered rows:
//
//     U = syn0[centres]              [128 x K]      (K = this shard's columns)
//     V = syn1[contexts ; negatives] [R   x K]      R = 160 context rows (128 + 2 x 16 halo) + NN
//     S     = U . V^T                [128 x R]      pass A  (tcgen05, tf32, accumulated in TMEM over K chunks)
//     G     = mask * (label - sigmoid(S)) * alpha * weight     epilogue 1: band of the context block + NN block
//     dUneg = Gneg . Vneg [128 x K], dVneg = Gneg^T . U [NN x K]   pass B (tcgen05)
//     dUpos, dVctx : the <= 2(window-1) positive terms per centre  pass B (CUDA cores, rows read from the same
//                                                                  shared-memory stage) -> red.global.add.v4.f32
//
// Data path per CTA (persistent, one CTA per SM, warp specialised):
//   warps 9-12  producers  per 32-column chunk: (128 + R) / 4 x cp.async.bulk.tensor ... tile::gather4 (TMA) pull the
//                          U and V rows by vocabulary index into a ring stage (one copy per lane, four warps issue); the ring interleaves pass B of tile t with pass A of tile t+1.  Pass A uses SWIZZLE_128B tensor maps
//                          (K-major tf32 operands); pass B re-gathers the chunk (L2 hits) through
//                          SWIZZLE_128B_ATOM_32B maps because MN-major tf32 operands exist only in that layout
//                          (benchmarks/probe_umma*.py pins every descriptor used here against numpy on the GPU)
//   warp 8      MMA        one elected lane issues tcgen05.mma kind::tf32; tcgen05.commit releases stages /
//                          publishes accumulators through mbarriers
//   warps 0-7   epilogue   two groups of four warps (TMEM lane quadrant = warp % 4).  tcgen05.ld S -> coefficients ->
//                          Gneg (both layouts) + band coefficients in shared memory; the pass-B chunks alternate
//                          between the groups: tcgen05.ld dUneg / dVneg + positive terms -> 16-byte RED into syn0 / syn1
// Duplicate words inside a tile are separate rows whose updates are summed from PRE-update values: exactly the
// reference's mini-batch semantics with batchSize = 128 (MLLIB:417-425).  Across tiles the updates are asynchronous
// (Hogwild), as between the reference's partitions (MLLIB:392).
#include "common.cuh"
#include "tc_common.cuh"
#include "sgns_params.h"
#include "sgns_tile.h"

namespace gw2v {

using namespace tc;

namespace {

constexpr int TL_T = 128;                  // centres per tile (UMMA M)
constexpr int TL_HALO = 16;                // context halo rows on each side of the tile
constexpr int TL_CTX = TL_T + 2 * TL_HALO; // 160 context rows
constexpr int TL_BK = 32;                  // floats per K chunk: one 128-byte swizzle row
constexpr int TL_EPI_THREADS = 256;        // two epilogue groups of 4 warps (TMEM lane quadrant = warp % 4)
constexpr int TL_GROUP_THREADS = 128;
constexpr int TL_MMA_WARP = 8;
constexpr int TL_PROD_WARP0 = 9;           // producer warps 9..12
constexpr int TL_NPROD = 4;
constexpr int TL_THREADS = 13 * 32;
constexpr int TL_BLOCK_BYTES = TL_T * 128;   // one [128 x 32 floats] swizzled block: 16 KB
constexpr int TL_ACC_COL0 = 256;           // TMEM: S in columns [0, R), pass-B accumulators from column 256
constexpr int TL_ACC_STRIDE = 64;          // dUneg 32 | dVneg 32
constexpr int TL_GB_STRIDE = 25;           // floats per row in the band-coefficient array (<= 23 slots + pad)
constexpr int TL_BAND_BYTES = 17 * 1024;   // [160][25] coefficients by context row + [160] masks, padded to 1 KB
constexpr int TL_MAXSTAGE = 4;
constexpr int TL_XSLOTS = 4;               // exchange slots per CTA (>= 2 with the one-tile lead of the push)

template <int R> struct TileCfg {
    static constexpr int NN = R - TL_CTX;                      // shared negatives per tile
    static constexpr int NB = NN / 32;                         // 32-column blocks of Gneg
    static constexpr int GK_BYTES = NB * TL_BLOCK_BYTES;       // Gneg, K-major SWIZZLE_128B        (A of dUneg)
    static constexpr int G32_BYTES = NB * TL_BLOCK_BYTES;      // Gneg, SWIZZLE_128B_BASE32B        (A^T of dVneg)
    static constexpr int V_STAGE_BYTES = R * 128;
    static constexpr int STAGE_BYTES = TL_BLOCK_BYTES + V_STAGE_BYTES;
    static constexpr int NSTAGE = (NN <= 32) ? 4 : 3;
    static constexpr int NGROUPS = (TL_T + R) / 4;             // gather4 copies per stage
    static constexpr int META_INTS = TL_T + R + TL_T;          // utok | vtok | cinfo
    // byte offsets from the 1024-aligned base
    static constexpr int GK_OFF = 0;
    static constexpr int G32_OFF = GK_OFF + GK_BYTES;
    static constexpr int BAND_OFF = G32_OFF + G32_BYTES;
    static constexpr int MASK_OFF = BAND_OFF + TL_CTX * TL_GB_STRIDE * 4;
    static constexpr int STAGE_OFF = BAND_OFF + TL_BAND_BYTES;
    static constexpr int META_OFF = STAGE_OFF + NSTAGE * STAGE_BYTES;
    static constexpr int BAR_OFF = META_OFF + 3 * META_INTS * 4;
    static constexpr int SMEM_BYTES = 1024 + BAR_OFF + 512;
    static_assert(NGROUPS % TL_NPROD == 0, "gather groups must split evenly over the producer warps");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct TileArgs {
    SgnsParams p;
    const uint32_t* cinfo;        // [T] window masks from pair_count_kernel
    const int* tile_negs;         // [ntiles_max, NN]
    const int* n_pairs;           // device scalar (pair_tile_scan_kernel)
    float* dbg;                   // optional debug dump of tile 0: S [128 x R] (band window + negatives)
};

__device__ __forceinline__ void red_add_v4(float* p, float4 v) { atomicAdd(reinterpret_cast<float4*>(p), v); }

// plain C++ accesses through a pointer that is still known to be shared memory (base = smem_raw + padding), so the
// compiler emits LDS / STS and is free to batch and hoist them
__device__ __forceinline__ float4 lds4(const uint8_t* sm, uint32_t off) { return *reinterpret_cast<const float4*>(sm + off); }
__device__ __forceinline__ float lds1(const uint8_t* sm, uint32_t off) { return *reinterpret_cast<const float*>(sm + off); }
__device__ __forceinline__ uint32_t ldsu(const uint8_t* sm, uint32_t off) { return *reinterpret_cast<const uint32_t*>(sm + off); }
__device__ __forceinline__ void sts4(uint8_t* sm, uint32_t off, float4 v) { *reinterpret_cast<float4*>(sm + off) = v; }
__device__ __forceinline__ void sts1(uint8_t* sm, uint32_t off, float v) { *reinterpret_cast<float*>(sm + off) = v; }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TL_EPI_THREADS) : "memory"); }
__device__ __forceinline__ void fma4(float (&acc)[32], int e, float g, float4 x) {
    acc[4 * e + 0] = fmaf(g, x.x, acc[4 * e + 0]); acc[4 * e + 1] = fmaf(g, x.y, acc[4 * e + 1]);
    acc[4 * e + 2] = fmaf(g, x.z, acc[4 * e + 2]); acc[4 * e + 3] = fmaf(g, x.w, acc[4 * e + 3]);
}

}  // namespace

// SLP = 0: single shard.  SLP = 12 / 24: column shards -- band slots exchanged per centre (offsets -5..5 / -11..11)
template <int R, int SLP>
__global__ void __launch_bounds__(TL_THREADS, 1)
sgns_tile_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                 const __grid_constant__ CUtensorMap tm0s, const __grid_constant__ CUtensorMap tm1s, const TileArgs a) {
    using C = TileCfg<R>;
    constexpr int NN = C::NN;
    constexpr bool MULTI = SLP > 0;
    constexpr int WM = MULTI ? SLP / 2 - 1 : 0;                  // band slot s of the exchange payload <-> offset s - WM
    constexpr int PAYF = SLP + NN;                               // exchanged floats per centre
    const SgnsParams& p = a.p;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // 1024-aligned, still a shared pointer
    const uint32_t sm_addr = smem_u32(sm);
    // GK   : NB blocks [128 centres x 32 negatives], SWIZZLE_128B          G32 : the same values, SWIZZLE_128B_BASE32B
    // BAND : bandT [160][25]: coefficient of the pair (centre i, context row r = i + 16 + off) at [r][win - off]
    // MASK : maskT [160]: live slots of a context row   (dV: row r sums g * u over its centres; dU: centre i walks its mask)
    // STAGE: NSTAGE x { U block 16 KB | V rows R x 128 B }                 META: 3 x { utok 128 | vtok R | cinfo 128 }
    int* smMeta = reinterpret_cast<int*>(sm + C::META_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::BAR_OFF);
    uint64_t* full = bars;            // [4] TMA bytes landed (one expect_tx per producer warp)
    uint64_t* empty = bars + 4;       // [4] MMAs reading the stage retired
    uint64_t* epi_done = bars + 8;    // [4] epilogue finished reading the stage (pass B uses only)
    uint64_t* s_full = bars + 12;     // S accumulator complete
    uint64_t* g_ready = bars + 13;    // coefficients written to shared memory (256 epilogue threads)
    uint64_t* acc_full = bars + 14;   // [2] pass-B accumulators of a chunk complete
    uint64_t* acc_empty = bars + 16;  // [2] ... drained by the epilogue group that owns the buffer (128)
    uint64_t* meta_full = bars + 18;  // [3] tile meta loaded (three buffers: the producers run up to two tiles ahead)
    uint64_t* tile_done = bars + 21;  // [3] epilogue finished the tile (256): meta buffer reusable
    uint64_t* b_full = bars + 24;     // [2] rows of the pass-B chunk that owns accumulator buffer x have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int K = p.K;
    const int NC = (K + TL_BK - 1) / TL_BK;
    const int T = *p.n_tokens;
    const int ntiles = (T + TL_T - 1) / TL_T;

    // ---------------------------------------------------------------- one-time setup
    if (warp == TL_PROD_WARP0 && elect_one()) {
        tma_prefetch_desc(&tm0); tma_prefetch_desc(&tm1); tma_prefetch_desc(&tm0s); tma_prefetch_desc(&tm1s);
    }
    if (warp == TL_MMA_WARP) {
        if (elect_one()) {
            for (int s = 0; s < TL_MAXSTAGE; ++s) {
                mbar_init(full + s, TL_NPROD); mbar_init(empty + s, 1); mbar_init(epi_done + s, TL_GROUP_THREADS);
            }
            mbar_init(s_full, 1);
            mbar_init(g_ready, TL_EPI_THREADS);
            for (int x = 0; x < 2; ++x) {
                mbar_init(acc_full + x, 1); mbar_init(acc_empty + x, TL_GROUP_THREADS); mbar_init(b_full + x, 1);
            }
            for (int x = 0; x < 3; ++x) { mbar_init(meta_full + x, 1); mbar_init(tile_done + x, TL_EPI_THREADS); }
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp >= TL_PROD_WARP0) {
        // =========================================================== producers: tile meta + TMA row gathers
        const int pw = warp - TL_PROD_WARP0;
        constexpr int MYG = C::NGROUPS / TL_NPROD;                   // gather4 copies of this warp per stage
        int stage = 0; uint32_t phase = 0;
        uint32_t b_par = 0;                                          // bit s: parity of the number of pass-B uses of stage s
        uint32_t last_b = 0;                                         // bit s: the current occupant of stage s is a pass-B chunk
        const int g = pw + TL_NPROD * lane;                          // this lane's gather4 copy; lanes [0, MYG) are active
        const bool is_u = g < TL_T / 4, is_ctx = !is_u && g < (TL_T + TL_CTX) / 4;
        // pass A: everything SWIZZLE_128B (K-major); pass B: U and the negatives through the 32-byte-atom swizzle
        const CUtensorMap* tmA = is_u ? &tm0 : &tm1;
        const CUtensorMap* tmB = is_u ? &tm0s : (is_ctx ? &tm1 : &tm1s);
        int4 idc = make_int4(0, 0, 0, 0), idn = make_int4(0, 0, 0, 0);     // row ids of the current / next tile
        auto load_meta = [&](int tile, int it) {                     // producer warp 0: stage the tile's indices
            const int mb = it % 3;
            int* meta = smMeta + mb * C::META_INTS;
            if (pw == 0) {
                if (it >= 3) mbar_wait(tile_done + mb, ((it / 3) - 1) & 1, 10);
                const int t0 = tile * TL_T;
                for (int r = lane; r < TL_T; r += 32) {
                    const int pos = t0 + r;
                    meta[r] = pos < T ? __ldg(p.tokens + pos) : 0;
                    meta[TL_T + R + r] = pos < T ? (int)__ldg(a.cinfo + pos) : 0;
                }
                for (int r = lane; r < TL_CTX; r += 32) {
                    const int pos = t0 - TL_HALO + r;
                    meta[TL_T + r] = (pos >= 0 && pos < T) ? __ldg(p.tokens + pos) : 0;
                }
                for (int r = lane; r < NN; r += 32) meta[TL_T + TL_CTX + r] = __ldg(a.tile_negs + (size_t)tile * NN + r);
                __syncwarp();
                if (lane == 0) mbar_arrive(meta_full + mb);
            }
            mbar_wait(meta_full + mb, (it / 3) & 1, 13);
            int4 ids = make_int4(0, 0, 0, 0);
            if (lane < MYG) ids = *reinterpret_cast<const int4*>(meta + 4 * g);       // utok (32), contexts (40), negatives
            return ids;
        };
        auto fill = [&](const CUtensorMap* tm, const int4& ids, int c, bool pass_b) {
            mbar_wait(empty + stage, phase ^ 1, 11);
            if ((last_b >> stage) & 1u) mbar_wait(epi_done + stage, ((b_par >> stage) & 1u) ^ 1u, 12);
            uint8_t* st = sm + C::STAGE_OFF + stage * C::STAGE_BYTES;
            if (lane == 0) mbar_expect_tx(full + stage, (uint32_t)(MYG * 512));
            __syncwarp();
            if (lane < MYG) tma_gather4(st + g * 512, tm, c * TL_BK, ids.x, ids.y, ids.z, ids.w, full + stage);
            if (pass_b) { last_b |= 1u << stage; b_par ^= 1u << stage; } else { last_b &= ~(1u << stage); }
            if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
        };
        // Ring order: A(first tile), then per tile t: B(t, 0), A(t+1, 0), B(t, 1), A(t+1, 1), ... -- the S GEMM of the
        // next tile is fed while the epilogue works on this tile's updates.
        int it = 0;
        if ((int)blockIdx.x < ntiles) {
            idc = load_meta(blockIdx.x, 0);
            for (int c = 0; c < NC; ++c) fill(tmA, idc, c, false);
        }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int next = tile + gridDim.x;
            const bool has_next = next < ntiles;
            if (has_next) idn = load_meta(next, it + 1);
            for (int c = 0; c < NC; ++c) {
                fill(tmB, idc, c, true);
                if (has_next) fill(tmA, idn, c, false);
            }
            idc = idn;
        }
    } else if (warp == TL_MMA_WARP) {
        // =========================================================== MMA issuer
        const uint32_t idescS = make_idesc_tf32(TL_T, R, 0, 0);          // S     = U . V^T       (K-major, K-major)
        const uint32_t idescU = make_idesc_tf32(TL_T, TL_BK, 0, 1);      // dUneg = Gneg . Vneg   (K-major, MN-major)
        const uint32_t idescV = make_idesc_tf32(TL_T, TL_BK, 1, 1);      // dVneg = Gneg^T . U    (MN-major, MN-major)
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        uint32_t gc = 0;                                                 // running pass-B chunk counter
        const uint32_t gk_addr = sm_addr + C::GK_OFF, g32_addr = sm_addr + C::G32_OFF;
        auto mma_a = [&](int c) {                                       // S += U chunk . V chunk^T
            mbar_wait(full + stage, phase, 20);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t st = sm_addr + C::STAGE_OFF + stage * C::STAGE_BYTES;
                const uint64_t ad = make_sw128_desc(st, 16, 1024);
                const uint64_t bd = make_sw128_desc(st + TL_BLOCK_BYTES, 16, 1024);
#pragma unroll
                for (int k = 0; k < TL_BK / 8; ++k)
                    umma_tf32(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idescS, (c > 0 || k > 0) ? 1u : 0u);
                umma_commit(empty + stage);
                if (c == NC - 1) umma_commit(s_full);
            }
            __syncwarp();
            if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
        };
        if ((int)blockIdx.x < ntiles)
            for (int c = 0; c < NC; ++c) mma_a(c);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const bool has_next = tile + (int)gridDim.x < ntiles;
            // ---- coefficients of this tile are in shared memory (and S has been drained: the next tile may overwrite it)
            mbar_wait(g_ready, it & 1, 21);
            tc_fence_after();
            for (int c = 0; c < NC; ++c, ++gc) {
                const int acc = gc & 1;
                mbar_wait(full + stage, phase, 22);
                if (lane == 0) mbar_arrive(b_full + acc);               // the chunk's rows have landed: its epilogue group may read them
                mbar_wait(acc_empty + acc, ((gc >> 1) & 1) ^ 1, 23);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t st = sm_addr + C::STAGE_OFF + stage * C::STAGE_BYTES;
                    const uint32_t vneg = st + TL_BLOCK_BYTES + TL_CTX * 128;
                    const uint32_t d0 = tmem + TL_ACC_COL0 + acc * TL_ACC_STRIDE;
                    // dUneg chunk [128 x 32] = Gneg [128 x NN] . Vneg chunk [NN x 32]
#pragma unroll
                    for (int kk = 0; kk < NN / 8; ++kk) {
                        const uint64_t ad = make_sw128_desc(gk_addr + (kk >> 2) * TL_BLOCK_BYTES + (kk & 3) * 32, 16, 1024);
                        const uint64_t bd = make_smem_desc(vneg + kk * 1024, 1024, 512, 1);
                        umma_tf32(d0, ad, bd, idescU, kk > 0 ? 1u : 0u);
                    }
                    // dVneg chunk [NN (of M = 128) x 32] = Gneg^T . U chunk; rows >= NN read past Gneg and are ignored
#pragma unroll 4
                    for (int kk = 0; kk < TL_T / 8; ++kk) {
                        const uint64_t ad = make_smem_desc(g32_addr + kk * 1024, TL_BLOCK_BYTES, 512, 1);
                        const uint64_t bd = make_smem_desc(st + kk * 1024, 1024, 512, 1);
                        umma_tf32(d0 + 32, ad, bd, idescV, kk > 0 ? 1u : 0u);
                    }
                    umma_commit(empty + stage);
                    umma_commit(acc_full + acc);
                }
                __syncwarp();
                if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
                if (has_next) mma_a(c);                                  // pass A of the next tile, same chunk index
            }
        }
    } else {
        // =========================================================== epilogue: group 0 = warps 0-3, group 1 = warps 4-7
        // Epilogue 1 is split by columns (each group takes half of the band window and half of the shared negatives); the
        // pass-B chunks alternate between the groups with the accumulator buffer (chunk gc belongs to group gc & 1).
        const int grp = warp >> 2;
        const int q = warp & 3;                                          // TMEM lane quadrant
        const int row = q * 32 + lane;                                   // centre index in the tile / TMEM lane
        const int etid = threadIdx.x;                                    // 0..255
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const float nratio = (float)p.negatives / (float)NN;
        const int win = (p.window_mode == 0) ? p.window - 1 : p.window;  // farthest context offset
        float loss = 0.f, maxdot = 0.f;
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        uint32_t gc = 0;
        const uint32_t seq0 = MULTI ? p.cta_seq[blockIdx.x] : 0u;       // tile sequence of this CTA, persists across launches
        // Column shards, phase-A epilogue of tile iteration `pit` (one group, thread = centre): the used entries of the
        // partial S -- SLP band slots + NN negatives per centre -- go from TMEM straight into every rank's exchange slot
        // (st.global on peer-mapped addresses over NVLink), then ONE release per destination publishes the tile.
        auto exchange_push = [&](int pit, int bar_id) {
            if constexpr (MULTI) {
                mbar_wait(s_full, pit & 1, 34);
                tc_fence_after();
                const uint32_t tag = seq0 + (uint32_t)pit + 1u;
                const uint32_t slot = (seq0 + (uint32_t)pit) % TL_XSLOTS;
                const size_t slot_base = ((size_t)blockIdx.x * TL_XSLOTS + slot) * (size_t)p.world;
                const bool loop = (p.debug & 8) != 0;                    // single-GPU loopback of the protocol (tests, ncu)
                const float sc = loop ? 1.f / (float)p.world : 1.f;      // loopback: every "rank" contributes 1/S of the dots
                auto push4 = [&](int f0, float4 v) {
                    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
                    for (int r = 0; r < p.world; ++r) {
                        const int src = loop ? r : p.rank;
                        // payload layout [float4 index][centre][4]: the 32 lanes of a warp store 512 contiguous bytes
                        // (one coalesced NVLink write instead of 32 scattered 16-byte ones -- the row-major layout made
                        // the push of an 8-shard run 7x slower than the compute of the tile, profiles/r2_scale.md)
                        float* dst = p.xbuf[r] + (slot_base + src) * TL_T * PAYF + ((size_t)(f0 >> 2) * TL_T + (size_t)row) * 4;
                        asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                                     ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    }
                };
                {   // band: column 16 + lane + off of the window [32q, 32q + 64) for off = s - WM
                    float xw[64];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        uint32_t x[16];
                        tmem_ld16(lane_addr + (uint32_t)(32 * q + 16 * h), x);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) xw[16 * h + j] = __uint_as_float(x[j]);
                    }
#pragma unroll
                    for (int s4 = 0; s4 < SLP / 4; ++s4) {
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c0 = TL_HALO + 4 * s4 + e - WM;    // column for lane 0; lane l reads c0 + l
                            float t = 0.f;
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (c0 + k >= 0 && c0 + k < 64) t = (lane == k) ? xw[c0 + k] : t;
                            v[e] = t;
                        }
                        push4(4 * s4, make_float4(v[0], v[1], v[2], v[3]));
                    }
                }
#pragma unroll 1
                for (int h = 0; h < NN / 16; ++h) {
                    uint32_t x[16];
                    tmem_ld16(lane_addr + (uint32_t)(TL_CTX + 16 * h), x);
                    tmem_ld_wait();
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4)
                        push4(SLP + 16 * h + 4 * j4, make_float4(__uint_as_float(x[4 * j4]), __uint_as_float(x[4 * j4 + 1]),
                                                                 __uint_as_float(x[4 * j4 + 2]), __uint_as_float(x[4 * j4 + 3])));
                }
                tc_fence_before();
                // all 128 payload rows are written (by other threads): barrier, then one release per destination
                if (bar_id == 2) asm volatile("bar.sync 2, %0;" ::"n"(TL_GROUP_THREADS) : "memory");
                else asm volatile("bar.sync 3, %0;" ::"n"(TL_GROUP_THREADS) : "memory");
                if (row < p.world) {
                    const int src = loop ? row : p.rank;
                    __threadfence_system();
                    st_release_sys(p.flags[row] + slot_base + src, tag);
                }
            }
        };
        if (MULTI && (int)blockIdx.x < ntiles && grp == 0) exchange_push(0, 2);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int mb = it % 3;
            const uint32_t meta = C::META_OFF + mb * C::META_INTS * 4;
            mbar_wait(meta_full + mb, (it / 3) & 1, 30);
            const uint32_t info = ldsu(sm, meta + (TL_T + R + row) * 4);
            const uint32_t mask = info & 0xFFFFFFu;
            const int lo = -(int)(info >> 24);
            const int m = __popc(mask);
            const float wneg = (float)m * nratio;                            // weight in the loss
            const float wupd = wneg * p.tile_neg_weight;                     // weight in the updates (stability knob, engine.py)
            const bool has_next = tile + (int)gridDim.x < ntiles;
            if (it == 0)                                                 // the first tile's pass-A stage uses go by
                for (int c = 0; c < NC; ++c)
                    if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
            // every epilogue thread is done with the previous tile's coefficients before they are overwritten
            epi_bar();
            if (etid < TL_CTX) *reinterpret_cast<uint32_t*>(sm + C::MASK_OFF + etid * 4) = 0u;
            epi_bar();
            // one (centre, context) candidate: coefficient into bandT / maskT
            auto band_emit = [&](int off, float f, bool known_valid = false) {
                const int bit = off - lo;
                if (known_valid || (bit >= 0 && bit < 24 && ((mask >> bit) & 1u))) {
                    float l = 0.f;
                    const float g = sgns_coeff_loss(f, 1.f, p.alpha, p.max_grad, p.exp_table, p.compute_loss != 0, l);
                    const int cr = row + TL_HALO + off;                  // context row of the pair, slot = win - off
                    sts1(sm, C::BAND_OFF + (uint32_t)(cr * TL_GB_STRIDE + win - off) * 4, g);
                    atomicOr(reinterpret_cast<unsigned int*>(sm + C::MASK_OFF + cr * 4), 1u << (win - off));
                    if (p.compute_loss) { loss += l; maxdot = fmaxf(maxdot, fabsf(f)); }
                }
            };
            // 16 (centre, shared negative) dots: coefficients into Gneg in both operand layouts
            auto negs_emit = [&](int h, const float (&f)[16]) {
                float g[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float l = 0.f;
                    const float c = sgns_coeff_loss(f[j], 0.f, p.alpha, p.max_grad, p.exp_table, p.compute_loss != 0, l);
                    g[j] = m > 0 ? wupd * c : 0.f;
                    if (m > 0 && p.compute_loss) { loss += wneg * l; maxdot = fmaxf(maxdot, fabsf(f[j])); }
                }
                const uint32_t rowoff = (uint32_t)(h >> 1) * TL_BLOCK_BYTES + (uint32_t)row * 128;
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {
                    const int c16 = (h & 1) * 4 + cq;                    // 16-byte chunk of the 32-float block row
                    const float4 gv = make_float4(g[4 * cq], g[4 * cq + 1], g[4 * cq + 2], g[4 * cq + 3]);
                    sts4(sm, C::GK_OFF + rowoff + (uint32_t)((c16 ^ (row & 7)) << 4), gv);
                    const int c32 = (c16 >> 1) ^ (row & 3);              // 32-byte chunk, 4-row period
                    sts4(sm, C::G32_OFF + rowoff + (uint32_t)(c32 * 32 + (c16 & 1) * 16), gv);
                }
            };
            if constexpr (!MULTI) {
                mbar_wait(s_full, it & 1, 31);
                tc_fence_after();
                // Both groups take half of each part (measured: with "group 0 = band, group 1 = negatives" the band group
                // idled at the barrier below for ~12 % of the epilogue time -- 32 sigmoids against <= 10).
                {
                    // band of the context block: columns [32q, 32q + 64) of S hold every context of centres 32q..32q+31
#pragma unroll 1
                    // window-column mask of this centre: bit c set <=> column c of [32q, 32q + 64) is one of its contexts
                    // (mask bit b <-> offset lo + b <-> column 16 + lane + lo + b; lo >= -11, so the shift is >= 5)
                    const unsigned long long colmask = (unsigned long long)mask << (TL_HALO + lane + lo);
                    for (int h = 2 * grp; h < 2 * grp + 2; ++h) {
                        uint32_t x[16];
                        tmem_ld16(lane_addr + (uint32_t)(32 * q + 16 * h), x);
                        tmem_ld_wait();
                        const uint32_t cm = (uint32_t)(colmask >> (16 * h)) & 0xFFFFu;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if ((cm >> j) & 1u) band_emit(16 * h + j - TL_HALO - lane, __uint_as_float(x[j]), true);
                            if (a.dbg != nullptr && tile == 0) a.dbg[(size_t)row * R + 32 * q + 16 * h + j] = __uint_as_float(x[j]);
                        }
                    }
                }
                {
                    // shared negatives: columns [160, 160 + NN)
#pragma unroll 1
                    for (int h = grp * (NN / 32); h < (grp + 1) * (NN / 32); ++h) {
                        uint32_t x[16];
                        tmem_ld16(lane_addr + (uint32_t)(TL_CTX + 16 * h), x);
                        tmem_ld_wait();
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            f[j] = __uint_as_float(x[j]);
                            if (a.dbg != nullptr && tile == 0) a.dbg[(size_t)row * R + TL_CTX + 16 * h + j] = f[j];
                        }
                        negs_emit(h, f);
                    }
                }
            } else {
                // ---- column shards: the partial dots of this tile were pushed by every rank (exchange_push below, one
                // tile ago); wait for all of them, sum in rank order (bit-identical coefficients on every rank: the
                // reference's coefficient broadcast disappears, MLLIB:423-425) and continue as on a single shard
                const uint32_t tag = seq0 + (uint32_t)it + 1u;
                const uint32_t slot = (seq0 + (uint32_t)it) % TL_XSLOTS;
                const size_t slot_base = ((size_t)blockIdx.x * TL_XSLOTS + slot) * (size_t)p.world;
                if (etid < p.world) {
                    const uint32_t* fl = p.flags[p.rank] + slot_base + etid;
                    const unsigned long long t0 = tc_globaltimer_ns();
                    while (ld_acquire_sys(fl) != tag) {
                        if (tc_globaltimer_ns() - t0 > 10000000000ull) {
                            printf("[gw2v] tile exchange: rank %d cta %d timed out waiting for rank %d (tag %u)\n", p.rank,
                                   (int)blockIdx.x, etid, tag);
                            atomicExch(p.error_flag, 1);
                            __trap();
                        }
                    }
                    if (p.timing != nullptr) atomicAdd(p.timing, tc_globaltimer_ns() - t0);
                }
                epi_bar();
                const float* xin = p.xbuf[p.rank] + slot_base * TL_T * PAYF + (size_t)row * 4;    // + src * 128 * PAYF + j * 128 * 4
                auto ld_sys = [](const float* ptr) {
                    float4 v;
                    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr) : "memory");
                    return v;
                };
                {
                    // both groups take half of the band slots and half of the negatives (see the single-shard branch)
                    constexpr int S4 = SLP / 4;
                    constexpr int S4H = (S4 + 1) / 2;
#pragma unroll
                    for (int s4 = 0; s4 < S4; ++s4) {
                        if ((s4 < S4H) != (grp == 0)) continue;
                        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int src = 0; src < p.world; ++src) {
                            const float4 v = ld_sys(xin + (size_t)src * TL_T * PAYF + (size_t)s4 * TL_T * 4);
                            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
                        }
                        band_emit(4 * s4 + 0 - WM, t.x); band_emit(4 * s4 + 1 - WM, t.y);
                        band_emit(4 * s4 + 2 - WM, t.z); band_emit(4 * s4 + 3 - WM, t.w);
                    }
                }
                {
#pragma unroll 1
                    for (int h = grp * (NN / 32); h < (grp + 1) * (NN / 32); ++h) {
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = 0.f;
                        for (int src = 0; src < p.world; ++src) {
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 v = ld_sys(xin + (size_t)src * TL_T * PAYF + (size_t)(SLP / 4 + 4 * h + j4) * TL_T * 4);
                                f[4 * j4] += v.x; f[4 * j4 + 1] += v.y; f[4 * j4 + 2] += v.z; f[4 * j4 + 3] += v.w;
                            }
                        }
                        negs_emit(h, f);
                    }
                }
            }
            fence_proxy_async_smem();          // st.shared (generic proxy) -> tcgen05.mma operand reads (async proxy)
            tc_fence_before();
            mbar_arrive(g_ready);
            epi_bar();                         // both groups read each other's coefficients below

            // ---- rows this thread updates in the chunks of its group
            const int utok = (int)ldsu(sm, meta + row * 4);
            const bool u_on = m > 0;
            const float su = row_scale(p.row_scale0, p.hot_rows, utok);          // hot-row damping (sgns_params.h)
            const int ntok = row < NN ? (int)ldsu(sm, meta + (TL_T + TL_CTX + row) * 4) : 0;
            const float sn = row_scale(p.row_scale1, p.hot_rows, ntok) * p.tile_neg_scale;   // in-tile event cap
            // context rows owned by this thread: `row` and, in warp 3 of the group, also row 128 + lane
            uint32_t cm[2]; int ctok[2]; float cs[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const int rr = x == 0 ? row : ((q == 3) ? TL_T + lane : TL_CTX);
                cm[x] = 0; ctok[x] = 0; cs[x] = 1.f;
                if (rr < TL_CTX) {
                    cm[x] = ldsu(sm, C::MASK_OFF + rr * 4);
                    ctok[x] = (int)ldsu(sm, meta + (TL_T + rr) * 4);
                    cs[x] = row_scale(p.row_scale1, p.hot_rows, ctok[x]);
                }
            }
            // ---- pass B: accumulators + positive terms -> 16-byte atomics
            for (int c = 0; c < NC; ++c, ++gc) {
                if ((int)(gc & 1) != grp) {                              // the other group's chunk (+ the next tile's pass-A use)
                    if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
                    if (has_next) { if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; } }
                    continue;
                }
                const int acc = gc & 1;
                const int col0 = c * TL_BK;
                const uint32_t stU = C::STAGE_OFF + stage * C::STAGE_BYTES;      // byte offsets from sm
                const uint32_t stV = stU + TL_BLOCK_BYTES;
                // The stage's own `full` barrier cannot be used here: a group only handles every other chunk, so it would
                // skip phases of that barrier and a parity wait that skips a phase can pass early (seen with the odd ring of
                // NN = 64: stale U rows).  The MMA warp observes every stage use in order and forwards "landed" on a
                // barrier that belongs to the accumulator buffer, i.e. to this group alone.
                mbar_wait(b_full + acc, (gc >> 1) & 1, 32);
                // dV of one context row: sum over the centres ci = rr - 16 - win + k that have it as a context of
                // g * u_ci
                auto band_dv = [&](int rr, uint32_t mm, float (&av)[32]) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) av[e] = 0.f;
                    while (mm) {
                        const int k = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const int ci = rr - TL_HALO - win + k;
                        const float g = lds1(sm, C::BAND_OFF + (uint32_t)(rr * TL_GB_STRIDE + k) * 4);
                        float4 xv[8];
#pragma unroll
                        for (int c32 = 0; c32 < 4; ++c32) {                          // U row ci, SWIZZLE_128B_ATOM_32B
                            const uint32_t ph = stU + (uint32_t)ci * 128 + (uint32_t)((c32 ^ (ci & 3)) * 32);
                            xv[2 * c32] = lds4(sm, ph); xv[2 * c32 + 1] = lds4(sm, ph + 16);
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) fma4(av, e, g, xv[e]);
                    }
                };
                // Row updates leave as COALESCED 16-byte atomics: the 32 rows of a warp are transposed through 4 KB of
                // shared memory (the warp's quarter of this stage's U block, dead once its MMAs and the band reads are
                // done) so that one RED instruction covers 4 rows x 128 contiguous bytes instead of 32 scattered
                // 16-byte pieces -- the scattered form was measured LSU-bound (profiles/r2_tile_kernel.md).
                const uint32_t scratch = stU + (uint32_t)q * 4096;
                auto red_rows = [&](const float (&v)[32], float* mat, int tok, bool on, float scale) {
#pragma unroll
                    for (int cq = 0; cq < 8; ++cq)
                        sts4(sm, scratch + (uint32_t)lane * 128 + (uint32_t)((cq ^ (lane & 7)) << 4),
                             make_float4(scale * v[4 * cq], scale * v[4 * cq + 1], scale * v[4 * cq + 2], scale * v[4 * cq + 3]));
                    __syncwarp();
                    const int tk = on ? tok : -1;
                    const int cq = lane & 7;
                    float4 xr[8]; int rt[8];
#pragma unroll
                    for (int ps = 0; ps < 8; ++ps) {
                        const int r = ps * 4 + (lane >> 3);
                        rt[ps] = __shfl_sync(0xffffffffu, tk, r);
                        xr[ps] = lds4(sm, scratch + (uint32_t)r * 128 + (uint32_t)((cq ^ (r & 7)) << 4));
                    }
                    if (col0 + 4 * cq < K && !(p.debug & 1)) {
#pragma unroll
                        for (int ps = 0; ps < 8; ++ps)
                            if (rt[ps] >= 0) red_add_v4(mat + (size_t)rt[ps] * K + col0 + 4 * cq, xr[ps]);
                    }
                    __syncwarp();
                };
                // (1a) context rows 128..159 (warp 3 of the group): few rows, direct scattered atomics
                if (q == 3 && cm[1] != 0) {
                    float av[32];
                    band_dv(TL_T + lane, cm[1], av);
                    if (!(p.debug & 1)) {
                        float* vrow = p.syn1 + (size_t)ctok[1] * K + col0;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (col0 + 4 * j < K)
                                red_add_v4(vrow + 4 * j, make_float4(cs[1] * av[4 * j], cs[1] * av[4 * j + 1], cs[1] * av[4 * j + 2],
                                                                     cs[1] * av[4 * j + 3]));
                    }
                }
                // (1b) context rows 0..127 into registers
                float av0[32];
                band_dv(row, cm[0], av0);
                // the chunk's MMAs have retired and every thread of the group is done reading U: the U block is scratch now
                mbar_wait(acc_full + acc, (gc >> 1) & 1, 33);
                tc_fence_after();
                if (grp == 0) asm volatile("bar.sync 2, %0;" ::"n"(TL_GROUP_THREADS) : "memory");
                else asm volatile("bar.sync 3, %0;" ::"n"(TL_GROUP_THREADS) : "memory");
                red_rows(av0, p.syn1, ctok[0], cm[0] != 0, cs[0]);
                // (2) centre row: dU = dUneg (tensor cores) + sum over its contexts g * v
                const uint32_t d0 = lane_addr + TL_ACC_COL0 + acc * TL_ACC_STRIDE;
                {
                    uint32_t x[32];
                    tmem_ld32(d0, x);                                    // warp-collective: never predicated
                    tmem_ld_wait();
                    float au[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) au[e] = __uint_as_float(x[e]);
                    uint32_t mm = mask;
                    while (mm) {
                        const int bit = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const int j = row + TL_HALO + lo + bit;          // context row of this pair
                        const float g = lds1(sm, C::BAND_OFF + (uint32_t)(j * TL_GB_STRIDE + win - lo - bit) * 4);
                        float4 xv[8];
#pragma unroll
                        for (int cq = 0; cq < 8; ++cq)                   // V row j, SWIZZLE_128B
                            xv[cq] = lds4(sm, stV + (uint32_t)j * 128 + (uint32_t)((cq ^ (j & 7)) << 4));
#pragma unroll
                        for (int e = 0; e < 8; ++e) fma4(au, e, g, xv[e]);
                    }
                    red_rows(au, p.syn0, utok, u_on, su);
                }
                // (3) shared negatives: lanes [0, NN) of the dVneg accumulator (warp-uniform condition)
                if (q < NN / 32) {
                    uint32_t x[32];
                    tmem_ld32(d0 + 32, x);
                    tmem_ld_wait();
                    float an[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) an[e] = __uint_as_float(x[e]);
                    red_rows(an, p.syn1, ntok, true, sn);
                }
                fence_proxy_async_smem();      // scratch writes (generic proxy) before the next TMA fill of this stage
                tc_fence_before();
                mbar_arrive(acc_empty + acc);
                mbar_arrive(epi_done + stage);
                if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
                if (has_next) { if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; } }
            }
            // column shards: the group that did not own the last chunk pushes the next tile's partial dots while the other
            // group is still busy with its updates -- the NVLink latency hides behind the tail of pass B
            if (MULTI && has_next && grp == (int)(gc & 1)) exchange_push(it + 1, 2 + grp);
            mbar_arrive(tile_done + mb);
        }
        if (MULTI && blockIdx.x < gridDim.x && etid == 0) p.cta_seq[blockIdx.x] = seq0 + (uint32_t)it;
        if (p.compute_loss) {
            loss = warp_sum(loss);
            maxdot = warp_max(maxdot);
            if (lane == 0) {
                if (loss != 0.f) atomicAdd(p.stats + 1, loss);
                atomicMax(reinterpret_cast<int*>(p.stats + 2), __float_as_int(maxdot));
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { p.stats[3] = (float)T; p.stats[0] = (float)(*a.n_pairs); }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TL_MMA_WARP) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------ shared negatives of the tiles
// negs[tile, 2c], negs[tile, 2c+1] = alias samples of philox(seed, NEG | iteration, pos0 + tile * 128, c): bit-identical
// to models/sgns.py::tile_negatives, identical on every shard (no index ever crosses NVLink)
__global__ void tile_negs_kernel(const int* __restrict__ n_tokens, const int2* __restrict__ alias, int vocab,
                                 uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration, unsigned long long pos0, int nn,
                                 int* __restrict__ out) {
    const int T = *n_tokens;
    const int ntiles = (T + TL_T - 1) / TL_T;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = nn / 2;
    const int tile = gid / half, c = gid - tile * half;
    if (tile >= ntiles) return;
    const uint4 r = rand4(seed_lo, seed_hi, stream_word(STREAM_NEG, iteration), pos0 + (unsigned long long)tile * TL_T, (uint32_t)c);
    out[(size_t)tile * nn + 2 * c] = alias_sample(alias, (uint32_t)vocab, r.x, r.y);
    out[(size_t)tile * nn + 2 * c + 1] = alias_sample(alias, (uint32_t)vocab, r.z, r.w);
}

// ------------------------------------------------------------------------------------------ host side
bool sgns_tile_supported(int K, int window, int negatives, int tile_centres, int tile_negatives) {
    return K % 4 == 0 && K >= 4 && K <= 4096 && window >= 1 && window <= 11 && negatives >= 1 && tile_centres == TL_T &&
           (tile_negatives == 32 || tile_negatives == 64);
}

int sgns_tile_max_tiles(int max_tokens) { return (max_tokens + TL_T - 1) / TL_T; }

template <int R, int SLP>
static int launch_tile_r(const SgnsParams& p, const TileLaunch& l, cudaStream_t stream) {
    using C = TileCfg<R>;
    CUtensorMap tm0, tm1, tm0s, tm1s;
    const uint64_t V = (uint64_t)p.vocab, K = (uint64_t)p.K;
    if (!make_tensormap_f32(&tm0, p.syn0, V, K, K, TL_BK, 1, false)) return 2;
    if (!make_tensormap_f32(&tm1, p.syn1, V, K, K, TL_BK, 1, false)) return 2;
    if (!make_tensormap_f32(&tm0s, p.syn0, V, K, K, TL_BK, 1, true)) return 2;
    if (!make_tensormap_f32(&tm1s, p.syn1, V, K, K, TL_BK, 1, true)) return 2;
    TileArgs a{};
    a.p = p;
    a.cinfo = l.cinfo; a.tile_negs = l.tile_negs; a.n_pairs = l.n_pairs;
    a.dbg = l.dbg;
    auto kern = sgns_tile_kernel<R, SLP>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    kern<<<l.grid, TL_THREADS, C::SMEM_BYTES, stream>>>(tm0, tm1, tm0s, tm1s, a);
    return 0;
}

// band slots exchanged per centre over column shards: offsets -5..5 (12 floats) or -11..11 (24 floats)
int sgns_tile_band_slots(int window, int window_mode) {
    const int win = window_mode == 0 ? window - 1 : window;
    return win <= 5 ? 12 : 24;
}

void sgns_tile_exchange_geometry(int window, int window_mode, int tile_negatives, int* slots, int* floats_per_slot_src) {
    *slots = TL_XSLOTS;
    *floats_per_slot_src = TL_T * (sgns_tile_band_slots(window, window_mode) + tile_negatives);
}

int launch_sgns_tile(const SgnsParams& p, const TileLaunch& l, cudaStream_t stream) {
    if (l.max_tokens <= 0) return 0;
    const int nn = l.tile_negatives;
    if (nn != 32 && nn != 64) return 1;
    const int max_tiles = sgns_tile_max_tiles(l.max_tokens);
    const int total = max_tiles * (nn / 2);
    tile_negs_kernel<<<(total + 127) / 128, 128, 0, stream>>>(p.n_tokens, p.alias, p.vocab, p.seed_lo, p.seed_hi, p.iteration,
                                                              p.pos0, nn, l.tile_negs);
    const int slp = p.world > 1 ? sgns_tile_band_slots(p.window, p.window_mode) : 0;
    if (nn == 32) {
        if (slp == 0) return launch_tile_r<192, 0>(p, l, stream);
        if (slp == 12) return launch_tile_r<192, 12>(p, l, stream);
        return launch_tile_r<192, 24>(p, l, stream);
    }
    if (slp == 0) return launch_tile_r<224, 0>(p, l, stream);
    if (slp == 12) return launch_tile_r<224, 12>(p, l, stream);
    return launch_tile_r<224, 24>(p, l, stream);
}

}  // namespace gw2v
