// Tensor-core SGNS step (neg_sharing = "tile"), see sgns_tile.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sgns_params.h"

namespace gw2v {

struct TileLaunch {
    const uint32_t* cinfo;        // [max_tokens] window masks (pair_count_kernel)
    int* tile_negs;               // [sgns_tile_max_tiles(max_tokens), tile_negatives] workspace
    const int* n_pairs;           // device scalar written by the pair-count scan
    float* dbg;                   // optional [128 * R] dump of tile 0's dot products (tests)
    int max_tokens;               // host-side upper bound of the step's token count
    int tile_negatives;           // 32 or 64
    int grid;                     // persistent CTAs (<= number of SMs)
};

bool sgns_tile_supported(int K, int window, int negatives, int tile_centres, int tile_negatives);
int sgns_tile_max_tiles(int max_tokens);
// column shards: exchange ring geometry (slots per CTA, floats per (slot, source rank))
void sgns_tile_exchange_geometry(int window, int window_mode, int tile_negatives, int* slots, int* floats_per_slot_src);
// pair_count + scan must have run on the same stream (launch_paircount); returns 0, 1 = unsupported, 2 = tensor map
int launch_sgns_tile(const SgnsParams& p, const TileLaunch& l, cudaStream_t stream);

// pairgen.cu: window masks + pair count only (no descriptors)
void launch_paircount(const int* tokens, const int* sent_id, const int* n_tokens, int max_tokens, uint32_t seed_lo,
                      uint32_t seed_hi, uint32_t iteration, unsigned long long pos0, int window, int window_mode,
                      uint32_t* cinfo, int* pair_off, int* n_pairs, int* tile_ws, float* stats, cudaStream_t stream);

}  // namespace gw2v
