// Descriptor probes for the tensor-core training kernel (tests/test_gpu_tile.py::test_umma_probe_*).
//
// There is no GPU on the development box, so every layout assumption of sgns_tile.cu is pinned down by two tiny
// kernels that are driven from Python on the GPU box:
//   umma_probe     runs an arbitrary list of tcgen05.mma (kind::tf32) instructions over a caller-supplied shared
//                  memory image with caller-supplied matrix / instruction descriptors and returns the TMEM
//                  accumulator -- the K-major and MN-major SWIZZLE_128B views, LBO/SBO, N = 32 and the overlapping
//                  M blocks used by G^T.U are all checked against numpy this way;
//   gather4_probe  issues cp.async.bulk.tensor.2d ... tile::gather4 and returns the shared-memory bytes, which
//                  proves the row order, the swizzle and the transaction byte count of the gather.
#include "tc_common.cuh"
#include "launchers.h"

namespace gw2v {

using namespace tc;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ image, int image_bytes, const unsigned long long* __restrict__ ops,
                  int n_ops, float* __restrict__ out, int ncols) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(base + i) = *reinterpret_cast<const uint4*>(image + i);
    fence_proxy_async_smem();
    if (warp == 0) {
        if (lane == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
        __syncwarp();
        tmem_alloc<512>(&tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && elect_one()) {
        const uint64_t rel = (uint64_t)(smem_u32(base) >> 4);
        for (int i = 0; i < n_ops; ++i) {
            const uint64_t a = ops[3 * i] + rel, b = ops[3 * i + 1] + rel, c = ops[3 * i + 2];
            umma_tf32(tmem + (uint32_t)((c >> 40) & 0x1FF), a, b, (uint32_t)c, (uint32_t)((c >> 32) & 1));
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c0 + j < ncols) out[(size_t)(warp * 32 + lane) * ncols + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

__global__ void __launch_bounds__(128, 1)
gather4_probe_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ rows, int n4, int col,
                     int bytes_per_op, uint8_t* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int total = n4 * 512;
    for (int i = threadIdx.x * 16; i < total; i += blockDim.x * 16) *reinterpret_cast<uint4*>(base + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)(n4 * bytes_per_op));
        for (int g = 0; g < n4; ++g)
            tma_gather4(base + g * 512, &map, col, rows[4 * g], rows[4 * g + 1], rows[4 * g + 2], rows[4 * g + 3], &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x * 16; i < total; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(base + i);
}

int launch_umma_probe(const uint8_t* image, int image_bytes, const unsigned long long* ops, int n_ops, float* out,
                      int ncols, cudaStream_t stream) {
    if (image_bytes % 16 || image_bytes > 200 * 1024 || ncols < 1 || ncols > 512) return 1;
    const int smem = image_bytes + 1024;
    cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_probe_kernel<<<1, 128, smem, stream>>>(image, image_bytes, ops, n_ops, out, ncols);
    return 0;
}

int launch_gather4_probe(const float* table, long long rows, int cols, const int* row_idx, int n4, int col, int box_cols,
                         int bytes_per_op, int swizzle32, uint8_t* out, cudaStream_t stream) {
    if (n4 < 1 || n4 * 512 > 200 * 1024) return 1;
    CUtensorMap map;
    if (!make_tensormap_f32(&map, table, (uint64_t)rows, (uint64_t)cols, (uint64_t)cols, (uint32_t)box_cols, 1, swizzle32 != 0)) return 2;
    const int smem = n4 * 512 + 1024;
    cudaFuncSetAttribute(gather4_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gather4_probe_kernel<<<1, 128, smem, stream>>>(map, row_idx, n4, col, bytes_per_op, out);
    return 0;
}

}  // namespace gw2v
