// Plain-C parameter structs + launcher declarations of the fused serving kernels (serve_fused.cu, nn_tc.cu);
// no CUDA device code here so the torch bindings (host compiler) can include it.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "sgns_params.h"

namespace gw2v {

struct PeerPtrs { float* p[MAX_WORLD]; };
struct PeerIdx { long long* p[MAX_WORLD]; };

struct ServeSync {
    uint32_t* flags[MAX_WORLD];   // flags[r] = rank r's flag array [world]; entry [src] is written by rank src
    unsigned int* done;           // local CTA arrival counter (zero between kernels)
    int* error_flag;
    int world, rank;
    uint32_t seq;                 // this operation's sequence number (monotone, identical on all ranks)
};

void launch_serve_wait(const uint32_t* flags_local, int world, uint32_t seq, int* error_flag, cudaStream_t st);
void launch_serve_barrier(const ServeSync& s, cudaStream_t st);
void launch_gather_rows_push(const float* syn0, const long long* rows, int R, int K, const PeerPtrs& out, int ldo,
                             const ServeSync& s, cudaStream_t st);
void launch_segment_mean_push(const float* syn0, const long long* rows, const long long* offsets, int NS, int K,
                              const PeerPtrs& out, int ldo, const ServeSync& s, cudaStream_t st);
void launch_row_sqnorm_push(const float* syn0, long long V, int K, const PeerPtrs& slab, long long vown,
                            const ServeSync& s, int sms, cudaStream_t st);
void launch_reduce_finish_push(const float* slab_local, int nsrc, long long vown, long long nvalid, int take_sqrt,
                               const PeerPtrs& full, const ServeSync& s, cudaStream_t st);
void launch_scores_rows_push(const float* syn0, long long V, int K, const float* qs, int Q, const PeerPtrs& slab,
                             long long vown, const ServeSync& s, int sms, cudaStream_t st);
int topk_owned_num_chunks(long long nvalid);
void launch_topk_owned_push(const float* slab_local, int nsrc, int Q, long long vown, long long nvalid,
                            const float* norms_owned, long long row_base, int k, float* cand_v, long long* cand_i,
                            const PeerPtrs& out_v, const PeerIdx& out_i, const ServeSync& s, cudaStream_t st);
int topk_sample_chunks(long long nvalid);
void launch_topk_select(const float* slab_local, int nsrc, int Q, long long vown, long long nvalid,
                        const float* norms_owned, long long row_base, int k, float* cand_s_v, long long* cand_s_i,
                        float* tau_v, long long* tau_i, float* cand_f_v, long long* cand_f_i, int* counts, int cap,
                        cudaStream_t st);
void launch_topk_merge_push(float* cand_v, const long long* cand_i, int ncand, int Q, int k, const PeerPtrs& out_v,
                            const PeerIdx& out_i, const ServeSync& s, cudaStream_t st);
void launch_topk_owned_stage1(const float* slab_local, int nsrc, int Q, long long vown, long long nvalid,
                              const float* norms_owned, long long row_base, int k, float* cand_v, long long* cand_i,
                              cudaStream_t st);
void launch_push_block(const float* src, long long n, const PeerPtrs& dst, const ServeSync& s, cudaStream_t st);
// nn_tc.cu: tcgen05 score GEMM whose epilogue stores each tile into the owner rank's slab (reduce-scatter)
int launch_scores_tc_push(const float* syn0, long long V, int K, const float* qpad, int Q, const PeerPtrs& slab,
                          long long vown, const ServeSync& s, int sms, cudaStream_t stream);

}  // namespace gw2v
