// Shared device helpers for the sm_100a kernels.
//
// * Philox4x32-10 -- bit-identical replica of utils/philox.py (tests/test_gpu_ops.py
//   checks it), so every rank regenerates the same windows / negatives /
//   sub-sampling decisions from (seed, stream, position, sub) with zero traffic
//   (reference: the seed-only dotprod request, MLLIB:420-421).
// * alias-table sampling (unigram^0.75 noise distribution).
// * vector load/store/atomic helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gw2v {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

enum Stream : uint32_t { STREAM_SUBSAMPLE = 1, STREAM_WINDOW = 2, STREAM_NEG = 3, STREAM_ZIPF = 4, STREAM_INIT = 5 };

__host__ __device__ __forceinline__ uint32_t stream_word(uint32_t stream, uint32_t iteration) {
    return (stream & 0xFFu) | ((iteration & 0xFFFFFFu) << 8);
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// random 4x32 for (position, sub) on a stream
__device__ __forceinline__ uint4 rand4(uint32_t seed_lo, uint32_t seed_hi, uint32_t stream_w,
                                       unsigned long long pos, uint32_t sub) {
    return philox4x32_10((uint32_t)pos, (uint32_t)(pos >> 32), sub, stream_w, seed_lo, seed_hi);
}

// {thresh (uint32 bits), alias}
__device__ __forceinline__ int alias_sample(const int2* __restrict__ table, uint32_t vocab,
                                            uint32_t r0, uint32_t r1) {
    uint32_t b = __umulhi(r0, vocab);
    int2 e = __ldg(table + b);
    return (r1 < (uint32_t)e.x) ? (int)b : e.y;
}

__device__ __forceinline__ float u32_to_unit_float(uint32_t r) {
    return (float)(r >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------- sigmoid / coefficient
constexpr float MAX_EXP = 6.0f;

// g = (label - sigmoid(f)) * alpha with the reference's hard clip at +-6 (MLLIB:292-302).
// table != null: the reference's 1000-entry lookup with its index scale 83.0 (integer 1000/6, MLLIB:296).
__device__ __forceinline__ float sgns_coeff(float f, float label, float alpha, float max_grad,
                                            const float* __restrict__ table = nullptr) {
    float sig;
    if (table != nullptr) {
        int ind = (int)((f + MAX_EXP) * 83.0f);
        ind = min(max(ind, 0), 999);
        sig = __ldg(table + ind);
    } else {
        sig = 1.0f / (1.0f + __expf(-f));
    }
    float g = label - sig;
    if (f > MAX_EXP) g = label - 1.0f;
    if (f < -MAX_EXP) g = label;
    g *= alpha;
    if (max_grad > 0.0f) g = fminf(fmaxf(g, -max_grad), max_grad);
    return g;
}

// coefficient and loss term of one dot product from ONE exponential (tile kernel epilogue: ~70 of these per thread and
// tile).  label 1: loss = log(1 + e^-f);  label 0: loss = log(1 + e^f) = f + log(1 + e^-f);  f clipped to +-MAX_EXP in
// the loss exactly like softplus_clipped; the reciprocal is the approximate one (2 ulp).
__device__ __forceinline__ float sgns_coeff_loss(float f, float label, float alpha, float max_grad,
                                                 const float* __restrict__ table, bool want_loss, float& loss) {
    const float fc = fminf(fmaxf(f, -MAX_EXP), MAX_EXP);
    const float e = __expf(-fc);
    float sig;
    if (table != nullptr) {
        int ind = (int)((f + MAX_EXP) * 83.0f);
        ind = min(max(ind, 0), 999);
        sig = __ldg(table + ind);
    } else {
        sig = __fdividef(1.0f, 1.0f + e);
    }
    float g = label - sig;
    if (f > MAX_EXP) g = label - 1.0f;
    if (f < -MAX_EXP) g = label;
    g *= alpha;
    if (max_grad > 0.0f) g = fminf(fmaxf(g, -max_grad), max_grad);
    if (want_loss) {
        const float l1 = __logf(1.0f + e);
        loss = label > 0.5f ? l1 : l1 + fc;
    }
    return g;
}

// per-row update scale of the hot-row damping (1 beyond the first `hot_rows` rows)
__device__ __forceinline__ float row_scale(const float* __restrict__ tab, int hot_rows, int row) {
    return (tab != nullptr && row < hot_rows) ? __ldg(tab + row) : 1.0f;
}

// softplus on the clipped dot: -log sigma(f) = softplus(-f)
__device__ __forceinline__ float softplus_clipped(float x) {
    x = fminf(fmaxf(x, -MAX_EXP), MAX_EXP);
    return (x > 0.0f ? x : 0.0f) + __logf(1.0f + __expf(-fabsf(x)));
}

// ---------------------------------------------------------------- system-scope flags (cross-GPU)
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

}  // namespace gw2v
