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

// grid/block/smem helpers live in sgns_kernels.cu
void launch_sgns_single(const SgnsParams& p, int grid, cudaStream_t stream);
void launch_sgns_multi(const SgnsParams& p, int grid, cudaStream_t stream);
int sgns_multi_max_grid(int K, int window, int negatives, int tile_centers, int device);
int sgns_single_grid(int K, int device);
size_t sgns_multi_smem_bytes(int window, int negatives, int tile_centers);

// sgns_pipe.cu: per-warp TMA pipeline variant of the single-shard step
bool sgns_pipe_supported(int K, int window, int negatives);
int sgns_pipe_grid(int K, int negatives, int device);
void launch_sgns_pipe(const SgnsParams& p, int grid, cudaStream_t stream);

// sgns_pairs.cu: production step over pre-generated pair descriptors (pairgen.cu)
bool sgns_pairs_supported(int K, int window, int negatives);
int sgns_pairs_grid(int K, int device, bool multi);
void sgns_pairs_multi_geometry(int* warps_per_cta, int* nslot, int* slot_floats);
void launch_sgns_pairs(const SgnsParams& p, const int* desc, const int* n_pairs, int pd, int grid, cudaStream_t stream);
void launch_sgns_pairs_multi(const SgnsParams& p, const int* desc, const int* n_pairs, int pd, int grid,
                             uint32_t* warp_seq, cudaStream_t stream);

// sgns_group.cu: register-path kernel with lane groups (short rows / high occupancy)
bool sgns_group_supported(int K, int window, int negatives);
int sgns_group_grid(int K, int device);
void launch_sgns_group(const SgnsParams& p, int grid, cudaStream_t stream);

// sgns_group_multi.cu: lane-group register path with the in-kernel NVLink all-reduce (world > 1)
bool sgns_group_multi_supported(int K, int window, int negatives);
void sgns_group_multi_geometry(int K, int device, int* grid, int* warps, int* nslot, int* slot_floats);
void launch_sgns_group_multi(const SgnsParams& p, int grid, uint32_t* warp_seq, cudaStream_t stream);

// sgns_pipe_multi.cu: the pipeline with the in-kernel NVLink all-reduce (world > 1)
bool sgns_pipe_multi_supported(int K, int window, int negatives);
void sgns_pipe_multi_geometry(int K, int negatives, int device, int* grid, int* warps, int* nslot, int* slot_floats);
void launch_sgns_pipe_multi(const SgnsParams& p, int grid, uint32_t* warp_seq, cudaStream_t stream);

}  // namespace gw2v
