// Host-callable launchers of the sm_100a kernels (no torch headers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gw2v {

// prep_kernels.cu
void launch_subsample_compact(const int* tok_in, const int* sid_in, int T, const uint32_t* keep_thresh,
                              uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration,
                              unsigned long long raw_pos0, int* tok_out, int* sid_out, int* count_out,
                              int* tile_ws,
                              cudaStream_t stream);
int subsample_max_blocks(int max_tokens);
int subsample_max_tokens();
void launch_zipf_stream(const int2* alias, int vocab, uint32_t seed_lo, uint32_t seed_hi,
                        unsigned long long pos0, int n, int* out, cudaStream_t stream);
void launch_init_syn0(float* syn0, long long vocab, int K, int col_start, int vector_size, uint32_t seed_lo,
                      uint32_t seed_hi, cudaStream_t stream);

// pairgen.cu
int pairgen_max_blocks(int max_tokens);
int pairgen_desc_ints(int negatives);
int pairgen_splits(int negatives);       // descriptors per pair: 1 up to 7 negatives, then one per 7
int pairgen_max_negatives();
int pairgen_max_tokens();
void launch_pairgen(const int* tokens, const int* sent_id, const int* n_tokens, int max_tokens, const int2* alias,
                    int vocab, uint32_t seed_lo, uint32_t seed_hi, uint32_t iteration, unsigned long long pos0,
                    int window, int window_mode, int negatives, int share_centre, uint32_t* cinfo, int* pair_off,
                    int* n_pairs, int* desc, int* tile_ws, float* stats,
                    cudaStream_t stream);   // stats (4 floats, may be null) is zeroed by the scan kernel

// infer_kernels.cu
void launch_gather_rows(const float* syn0, const long long* rows, int R, int K, float* out, cudaStream_t s);
void launch_segment_mean_rows(const float* syn0, const long long* rows, const long long* offsets, int NS, int K,
                              float* out, cudaStream_t s);
void launch_row_sqnorm(const float* syn0, long long V, int K, float* out, int sms, cudaStream_t s);
void launch_scores_rows(const float* syn0, long long V, int K, const float* qs, int Q, float* out, int sms,
                        cudaStream_t s);
void launch_topk_merge(float* cand_v, const long long* cand_i, int ncand, int Q, int k, float* out_v,
                       long long* out_i, cudaStream_t s);

// umma_probe.cu: descriptor probes (tests only)
int launch_umma_probe(const uint8_t* image, int image_bytes, const unsigned long long* ops, int n_ops, float* out,
                      int ncols, cudaStream_t stream);
int launch_gather4_probe(const float* table, long long rows, int cols, const int* row_idx, int n4, int col, int box_cols,
                         int bytes_per_op, int swizzle32, uint8_t* out, cudaStream_t stream);

}  // namespace gw2v
