// tcgen05 / TMEM nearest-neighbour score GEMM (see nn_tc.cu).
#pragma once
#include <cuda_runtime.h>
namespace gw2v {
bool scores_tc_supported(int K, int Q);
int scores_tc_padded_queries(int Q);
// out[q, v] = sum_k syn0[v, k] * qs[q, k]; qs must be zero padded to scores_tc_padded_queries(Q) rows.
// returns 0 on success, 1 if the shape is unsupported, 2 if the TMA descriptors could not be encoded
int launch_scores_tc(const float* syn0, long long V, int K, const float* qs, int Q, float* out, int sms,
                     cudaStream_t stream);

// nn_select.cu: score GEMM with the cosine + threshold selection fused into the tcgen05 epilogue
struct PeerPtrs;
struct ServeSync;
bool nn_select_supported(int K, int Q);
int launch_nn_select(const float* mat, long long rows, int K, long long pitch, const float* inv_norm, long long inv_stride,
                     const float* qpad, int Q, int dense, float* out, long long ld, const float* thr, int* cand, int* count,
                     int cap, int sms, cudaStream_t stream);
void launch_nn_rerank(const float* mat, int K, const float* inv_norm, const float* qpad, int Q, const int* cand,
                      const int* count, int cap, long long row_base, float* out_v, long long* out_i, cudaStream_t stream);
void launch_rowshard_push(const float* syn0, long long V, int K, const PeerPtrs& dst, long long vown, int ldr, int col0,
                          const ServeSync& sync, int sms, cudaStream_t stream);
}  // namespace gw2v
