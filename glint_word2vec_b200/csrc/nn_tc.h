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
}  // namespace gw2v
