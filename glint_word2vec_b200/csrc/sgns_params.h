// Launch parameters of the fused SGNS step (plain C struct shared by the
// kernels and the torch bindings; no torch headers here so kernel edits
// recompile in seconds).
#pragma once
#include <stdint.h>

namespace gw2v {

constexpr int MAX_WORLD = 8;

struct SgnsParams {
    float* syn0;                 // [V, K] input vectors  (this rank's column slice)
    float* syn1;                 // [V, K] output vectors
    const int* tokens;           // [T] compacted step tokens
    const int* sent_id;          // [T]
    const int* n_tokens;         // device scalar: T after sub-sampling
    const int2* alias;           // [V] {thresh, alias}
    float* stats;                // [4] pairs, loss, max|f|, spare
    unsigned long long pos0;     // stream position of tokens[0]
    uint32_t seed_lo, seed_hi, iteration;
    int vocab;
    int K;                       // row stride in floats (multiple of 4)
    int window, negatives, window_mode;   // window_mode 0 = reference (Q2), 1 = word2vec.c
    float alpha, max_grad;
    const float* exp_table;      // 1000-entry sigma table on [-6,6] (MLLIB:281-302 parity mode) or null = exact
    // Hot-row damping: the update of row r of syn0 / syn1 is multiplied by row_scale{0,1}[r] for r < hot_rows (rows
    // beyond are 1).  A device step applies thousands of stale, summed updates to the rows of very frequent words --
    // the exploding-gradient hazard README.md:17-19 warns about, 10^3 times stronger here -- so their effective
    // learning rate is capped (models/engine.py::row_scales).  null / 0 = off.
    const float* row_scale0;
    const float* row_scale1;
    int hot_rows;
    float tile_neg_scale;        // tile kernel: scale of the shared negatives' row updates (engine.tile_neg_scale)
    float tile_neg_weight;       // tile kernel: weight of the whole negative term, both dU and dV side (1 = the reference's n)
    int compute_loss;
    int debug;                   // bit0 skip syn1 atomics, bit1 skip syn0 atomics, bit2 skip row loads (profiling);
                                 // bit3 single-GPU loopback of the exchange (profiling); bit4 random push delays (stress test)
    // ---- cross-shard exchange (world > 1)
    int world, rank;
    int tile_centers;            // centres per CTA tile
    int slot_floats;             // floats per (cta, slot, source) region
    float* xbuf[MAX_WORLD];      // peer-mapped exchange buffers, xbuf[r] lives on rank r
    uint32_t* flags[MAX_WORLD];  // peer-mapped flag arrays [grid * world]
    float* xbuf_mc;              // multicast alias of xbuf (NVLS), or null
    uint32_t* cta_seq;           // [grid] running tile sequence number per CTA (local)
    int* error_flag;             // set by the spin watchdog
    unsigned long long* timing;  // optional [grid*2]: accumulated wait ns, tiles (exposed all-reduce time)
};

// start-up self-test of the 16-byte exchange chunks (sgns_pairs.cu)
struct PeerTest {
    uint32_t* buf[MAX_WORLD];     // per rank: [world][32] chunks of 16 bytes (symmetric)
    unsigned long long* result;   // [2] torn observations, observations
    int world, rank, iters;
};
void launch_xchg_selftest(const PeerTest& t, cudaStream_t stream);

// sgns_pairs.cu: production step over pre-generated pair descriptors (pairgen.cu)
bool sgns_pairs_supported(int K, int window, int negatives);
int sgns_pairs_grid(int K, int device, bool multi);
void sgns_pairs_multi_geometry(int* warps_per_cta, int* nslot, int* slot_floats);
void launch_sgns_pairs(const SgnsParams& p, const int* desc, const int* n_pairs, int pd, int grid, cudaStream_t stream);
void launch_sgns_pairs_multi(const SgnsParams& p, const int* desc, const int* n_pairs, int pd, int grid,
                             uint32_t* warp_seq, cudaStream_t stream);

}  // namespace gw2v
