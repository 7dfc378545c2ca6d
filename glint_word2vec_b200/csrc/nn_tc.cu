// placeholder until the tcgen05 kernel lands (keeps the module linkable)
#include "nn_tc.h"
namespace gw2v {
bool scores_tc_supported(int, int) { return false; }
int launch_scores_tc(const float*, long long, int, const float*, int, float*, int, cudaStream_t) { return 1; }
}  // namespace gw2v
