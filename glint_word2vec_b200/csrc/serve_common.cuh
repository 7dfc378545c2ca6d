// Cross-GPU plumbing of the serving kernels (pull / pullAverage / norms / multiply / top-k over column shards).
//
// Reference: every serving call is an Akka fan-out to the S servers followed by a client-side concat or sum
// (`BigWord2VecMatrix.pull/pullAverage/norms/multiply` [G], call sites MLLIB:486,514,598, ML:353,453).
// Here the shard kernels write their results straight into the peers' symmetric buffers over NVLink
// (st.global to peer-mapped pointers) and publish a monotone sequence number with st.release.sys when the
// LAST CTA of the grid has finished; the consumer is a 1-warp wait kernel (ld.acquire.sys) on the same stream.
// No NCCL call on the serving path.
#pragma once
#include "common.cuh"
#include "serve_params.h"

namespace gw2v {

// Call once per CTA after its last peer store (every thread of the CTA must reach it).  The last CTA to
// arrive publishes `seq` to every rank (including this one).
__device__ __forceinline__ void serve_cta_done(const ServeSync& s) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned int prev = atomicAdd(s.done, 1u);
        if (prev == total - 1) {
            atomicExch(s.done, 0u);
            __threadfence_system();
            for (int r = 0; r < s.world; ++r) st_release_sys(s.flags[r] + s.rank, s.seq);
        }
    }
}

}  // namespace gw2v
