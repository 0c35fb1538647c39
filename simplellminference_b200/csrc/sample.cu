// csrc/sample.cu — temperature / top-k / top-p sampling of one logits vector on the device.
//
// The reference only has arg-max (source/op/argmax.cpp:7-17; LayerType::kLayerSoftmax is declared and unused, layer.h:17), so this op
// is additive and NOT part of the parity contract with the reference; its contract is the numpy restatement in
// tests/test_sample_gpu.py. One CTA of 1024 threads, everything deterministic for a given (seed, step):
//   z_i = logit_i / temperature,  w_i = exp(z_i - max z)
//   top-k : keep { i : z_i >= the k-th largest z }            (ties at the k-th value are all kept)
//   top-p : among those, keep { i : z_i >= tau } for the LARGEST tau whose kept mass is >= top_p * mass    (ties kept)
//   draw  : u = hash(seed, step) in [0,1); the first index (in index order) whose running kept mass exceeds u * kept mass
// The two thresholds are found by bisection over the order-preserving integer image of the floats (32 steps, each one
// block-wide count / sum) — no sort, no scratch memory. temperature <= 0 means arg-max (first maximum).
#include "common.cuh"

namespace sllm {

__device__ __forceinline__ uint32_t f2key(float f) {   // monotone: a < b  <=>  key(a) < key(b)
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline uint64_t sample_mix(uint64_t x) {   // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

constexpr int kSampleThreads = 1024;

__device__ __forceinline__ float block_sum_f(float v, float* red) {   // deterministic: fixed tree
    return block_sum(v, red);
}
__device__ __forceinline__ int block_sum_i(int v, int* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int t = red[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    const int r = red[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kSampleThreads) sample_kernel(const float* __restrict__ logits, int n, float temperature, int top_k, float top_p,
                                                                unsigned long long seed, unsigned long long step, int32_t* __restrict__ idx_out) {
    __shared__ float redf[33];
    __shared__ int redi[33];
    __shared__ float scan[kSampleThreads];
    __shared__ float s_max;
    const int tid = threadIdx.x;
    const float inv_t = 1.0f / temperature;
    // max of z (first maximum is irrelevant here)
    float m = -INFINITY;
    for (int i = tid; i < n; i += kSampleThreads) m = fmaxf(m, logits[i] * inv_t);
    m = warp_max(m);
    if ((tid & 31) == 0) redf[tid >> 5] = m;
    __syncthreads();
    if (tid < 32) {
        float t = redf[tid];
        t = warp_max(t);
        if (tid == 0) s_max = t;
    }
    __syncthreads();
    m = s_max;

    // ---- top-k threshold: the largest key T with count(key >= T) >= k  (bisection on the 32 key bits, high to low)
    uint32_t thr = 0;
    if (top_k > 0 && top_k < n) {
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = thr | (1u << bit);
            int c = 0;
            for (int i = tid; i < n; i += kSampleThreads) c += (f2key(logits[i] * inv_t) >= cand);
            if (block_sum_i(c, redi) >= top_k) thr = cand;
        }
    }
    // ---- top-p threshold over the top-k set: the largest key T' >= thr with mass(key >= T') >= top_p * mass(key >= thr)
    if (top_p > 0.f && top_p < 1.f) {
        float w = 0.f;
        for (int i = tid; i < n; i += kSampleThreads) {
            const float z = logits[i] * inv_t;
            if (f2key(z) >= thr) w += expf(z - m);
        }
        const float need = top_p * block_sum_f(w, redf);
        uint32_t t2 = 0;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = t2 | (1u << bit);
            float ws = 0.f;
            for (int i = tid; i < n; i += kSampleThreads) {
                const float z = logits[i] * inv_t;
                const uint32_t k = f2key(z);
                if (k >= cand && k >= thr) ws += expf(z - m);
            }
            if (block_sum_f(ws, redf) >= need) t2 = cand;
        }
        thr = max(thr, t2);
    }
    // ---- draw: contiguous chunk per thread, block-wide exclusive scan of the chunk masses, then a walk inside one chunk
    const int chunk = (n + kSampleThreads - 1) / kSampleThreads;
    const int i0 = min(n, tid * chunk), i1 = min(n, i0 + chunk);
    float mine = 0.f;
    for (int i = i0; i < i1; ++i) {
        const float z = logits[i] * inv_t;
        if (f2key(z) >= thr) mine += expf(z - m);
    }
    scan[tid] = mine;
    __syncthreads();
    if (tid == 0) {   // 1024 serial adds: fixed order, negligible next to the bisections
        float run = 0.f;
        for (int t = 0; t < kSampleThreads; ++t) { const float v = scan[t]; scan[t] = run; run += v; }
        redf[0] = run;
    }
    __syncthreads();
    const float total = redf[0];
    const float u = (float)(sample_mix(seed ^ sample_mix(step)) >> 40) * (1.0f / 16777216.0f);
    const float target = fminf(u * total, total * 0.99999994f);   // strictly below the total: exactly one thread owns it
    const float before = scan[tid], after = (tid + 1 < kSampleThreads) ? scan[tid + 1] : total;
    // exactly one thread owns the target: before <= target < after (empty chunks have before == after); the last kept index
    // catches target == total rounding
    if (target >= before && target < after) {
        float run = before;
        int pick = -1;
        for (int i = i0; i < i1; ++i) {
            const float z = logits[i] * inv_t;
            if (f2key(z) >= thr) {
                run += expf(z - m);
                pick = i;
                if (run > target) break;
            }
        }
        if (pick >= 0) *idx_out = pick;
    }
}

__global__ void __launch_bounds__(1024) sample_argmax_kernel(const float* __restrict__ logits, int n, int32_t* __restrict__ idx_out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float v = -INFINITY;
    int idx = 0x7fffffff;
    auto comb = [](float& a, int& ai, float b, int bi) { if (b > a || (b == a && bi < ai)) { a = b; ai = bi; } };
    for (int i = threadIdx.x; i < n; i += blockDim.x) comb(v, idx, logits[i], i);
    for (int o = 16; o > 0; o >>= 1) comb(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = sv[threadIdx.x]; idx = si[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) comb(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
        if (threadIdx.x == 0) *idx_out = (idx == 0x7fffffff) ? 0 : idx;
    }
}

}  // namespace sllm

using namespace sllm;

extern "C" int sllm_sample_f32(const float* logits, int32_t n, float temperature, int32_t top_k, float top_p, uint64_t seed, uint64_t step,
                               int32_t* idx_dev, sllm_stream_t stream) {
    SLLM_REQUIRE(logits && idx_dev && n > 0, SLLM_EINVAL, "sample: null pointer or n <= 0");
    SLLM_REQUIRE(top_k >= 0 && top_p >= 0.f && top_p <= 1.f, SLLM_EINVAL, "sample: top_k must be >= 0 and top_p in [0, 1]");
    if (temperature <= 0.f) sample_argmax_kernel<<<1, 1024, 0, as_stream(stream)>>>(logits, n, idx_dev);
    else sample_kernel<<<1, kSampleThreads, 0, as_stream(stream)>>>(logits, n, temperature, top_k, top_p, seed, step, idx_dev);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}
