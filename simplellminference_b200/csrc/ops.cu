// csrc/ops.cu — the small element-wise / reduction ops of the reference's op-by-op path, one launcher per
// kernel::*_cuda entry point (include/sllm_b200.h names each). These are the drop-in kernels behind the C++
// op layers; the fused decode step (decode_fused.cu) folds them into the GEMV kernels instead.
//
// Numerics follow the CPU reference's order of operations except for reduction order:
// IEEE sqrt/divide (no rsqrtf, no --use_fast_math), accurate expf, (x*inv)*w association in RMSNorm.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace sllm {

// ----------------------------------------------------------------------------------------------- add --
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 x = a4[i], y = b4[i];
        o4[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
    for (int i = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}

// ------------------------------------------------------------------------------------------ swiglu ----
__global__ void swiglu_kernel(const float* __restrict__ up, const float* __restrict__ gate, float* __restrict__ out, int n) {
    pdl_launch_dependents();
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float s = 1.0f / (1.0f + expf(-gate[i]));  // sigmoid(gate), swiglu_kernel.cpp:12
        out[i] = s * up[i];
    }
}

// ----------------------------------------------------------------------------------------- rmsnorm ----
// One CTA: d is at most a few thousand floats and the vector sits in L2; the fused path never launches this.
__global__ void __launch_bounds__(1024) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       float* __restrict__ y, int d, float eps) {
    __shared__ float red[33];
    pdl_launch_dependents();
    pdl_wait();
    float ss = 0.0f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) ss += x[i] * x[i];
    ss = block_sum(ss, red);
    const float inv = 1.0f / sqrtf(ss / (float)d + eps);
    for (int i = threadIdx.x; i < d; i += blockDim.x) y[i] = (x[i] * inv) * w[i];
}

// --------------------------------------------------------------------------------------- embedding ----
template <int WD>
__global__ void embedding_kernel(const int32_t* __restrict__ token_dev, int token, const void* __restrict__ table,
                                 const float* __restrict__ scales, int group, float* __restrict__ out, int vocab, int d) {
    pdl_launch_dependents();
    pdl_wait();
    int tok = token_dev ? *token_dev : token;
    tok = min(max(tok, 0), vocab - 1);
    const int64_t base = (int64_t)tok * d;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
        float v;
        if (WD == SLLM_F32) v = reinterpret_cast<const float*>(table)[base + i];
        else if (WD == SLLM_BF16) v = __uint_as_float((uint32_t) reinterpret_cast<const uint16_t*>(table)[base + i] << 16);
        else v = (float)reinterpret_cast<const int8_t*>(table)[base + i] * scales[(base + i) / group];
        out[i] = v;
    }
}

// -------------------------------------------------------------------------------------------- rope ----
// grid.x = max(q_heads, k_heads); thread j < head_dim/2 rotates pair (j, j+hd/2) of q (and of k while the
// head index is inside k). rope_kernel.cpp:27-38.
__global__ void rope_kernel(float* __restrict__ q, float* __restrict__ k, const int32_t* __restrict__ pos_dev, int pos,
                            const float* __restrict__ sin_t, const float* __restrict__ cos_t, int q_dim, int k_dim, int hd) {
    pdl_launch_dependents();
    pdl_wait();
    const int half = hd >> 1;
    const int p = pos_dev ? *pos_dev : pos;
    const int b = blockIdx.x * hd;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const float fci = sin_t[(int64_t)p * half + j], fcr = cos_t[(int64_t)p * half + j];
        if (b < q_dim) {
            const float v0 = q[b + j], v1 = q[b + j + half];
            q[b + j] = v0 * fcr - v1 * fci;
            q[b + j + half] = v1 * fcr + v0 * fci;
        }
        if (b < k_dim) {
            const float v0 = k[b + j], v1 = k[b + j + half];
            k[b + j] = v0 * fcr - v1 * fci;
            k[b + j + half] = v1 * fcr + v0 * fci;
        }
    }
}

// ------------------------------------------------------------------------------------------ argmax ----
// first maximum (std::max_element, argmax.cpp:11): ties resolve to the lowest index.
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ logits, int n, int32_t* __restrict__ idx_out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    pdl_launch_dependents();
    pdl_wait();
    float v = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) argmax_combine(v, idx, logits[i], i);
    for (int o = 16; o > 0; o >>= 1) argmax_combine(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sv[warp] = v; si[warp] = idx; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        v = lane < nw ? sv[lane] : -INFINITY;
        idx = lane < nw ? si[lane] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) argmax_combine(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
        if (lane == 0) *idx_out = (idx == 0x7fffffff) ? 0 : idx;
    }
}

// ------------------------------------------------------------------------------------ kv row store ----
__global__ void store_kv_kernel(const float* __restrict__ src, void* __restrict__ dst, int kv_dtype, int n) {
    pdl_launch_dependents();
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (kv_dtype == SLLM_BF16) reinterpret_cast<uint16_t*>(dst)[i] = f32_to_bf16_bits(src[i]);
        else reinterpret_cast<float*>(dst)[i] = src[i];
    }
}

static inline int blocks_for(int n, int threads) { return std::max(1, std::min((n + threads - 1) / threads, sm_count() * 4)); }

}  // namespace sllm

using namespace sllm;

extern "C" {

int sllm_add_f32(const float* a, const float* b, float* out, int32_t n, sllm_stream_t stream) {
    SLLM_REQUIRE(a && b && out && n > 0, SLLM_EINVAL, "add: null pointer or n <= 0");
    SLLM_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0, SLLM_EINVAL, "add: pointers must be 16-byte aligned");
    add_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(a, b, out, n);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

int sllm_swiglu_f32(const float* up, const float* gate, float* out, int32_t n, sllm_stream_t stream) {
    SLLM_REQUIRE(up && gate && out && n > 0, SLLM_EINVAL, "swiglu: null pointer or n <= 0");
    swiglu_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(up, gate, out, n);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

int sllm_rmsnorm_f32(const float* x, const float* w, float* y, int32_t d, float eps, sllm_stream_t stream) {
    SLLM_REQUIRE(x && w && y && d > 0, SLLM_EINVAL, "rmsnorm: null pointer or d <= 0");
    rmsnorm_kernel<<<1, 1024, 0, as_stream(stream)>>>(x, w, y, d, eps);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

int sllm_embedding(const int32_t* token_dev, int32_t token, const void* table, int32_t w_dtype, const float* scales,
                   int32_t group, float* out, int32_t vocab, int32_t d, sllm_stream_t stream) {
    SLLM_REQUIRE(table && out && vocab > 0 && d > 0, SLLM_EINVAL, "embedding: null pointer or bad size");
    SLLM_REQUIRE(token_dev || (token >= 0 && token < vocab), SLLM_EINVAL, "Token index %d is outside the vocabulary [0, %d).", token, vocab);
    cudaStream_t st = as_stream(stream);
    const int blocks = blocks_for(d, 256);
    switch (w_dtype) {
        case SLLM_F32: embedding_kernel<SLLM_F32><<<blocks, 256, 0, st>>>(token_dev, token, table, scales, group, out, vocab, d); break;
        case SLLM_BF16: embedding_kernel<SLLM_BF16><<<blocks, 256, 0, st>>>(token_dev, token, table, scales, group, out, vocab, d); break;
        case SLLM_INT8:
            SLLM_REQUIRE(scales && group > 0 && d % group == 0, SLLM_EINVAL, "embedding: int8 table needs scales and d %% group == 0");
            embedding_kernel<SLLM_INT8><<<blocks, 256, 0, st>>>(token_dev, token, table, scales, group, out, vocab, d);
            break;
        default: SLLM_REQUIRE(false, SLLM_EINVAL, "embedding: unknown weight dtype %d", w_dtype);
    }
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

int sllm_rope_tables(int32_t head_dim, int32_t max_len, float theta, float* sin_dev, float* cos_dev, sllm_stream_t stream) {
    SLLM_REQUIRE(sin_dev && cos_dev && head_dim >= 2 && head_dim % 2 == 0 && max_len > 0, SLLM_EINVAL, "rope_tables: bad arguments");
    const int half = head_dim / 2;
    std::vector<float> s((size_t)max_len * half), c((size_t)max_len * half);
    // Same libm calls, same fp32 rounding points as the reference CPU table builder (rope_kernel.cpp:8-17).
    for (int k = 0; k < half; ++k) {
        const float freq = 1.0f / powf(theta, (float)(2 * k) / (float)head_dim);
        for (int i = 0; i < max_len; ++i) {
            const float val = freq * (float)i;
            c[(size_t)i * half + k] = cosf(val);
            s[(size_t)i * half + k] = sinf(val);
        }
    }
    cudaStream_t st = as_stream(stream);
    SLLM_CUDA(cudaMemcpyAsync(sin_dev, s.data(), s.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    SLLM_CUDA(cudaMemcpyAsync(cos_dev, c.data(), c.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    SLLM_CUDA(cudaStreamSynchronize(st));
    return SLLM_OK;
}

int sllm_rope_f32(float* q, float* k, const int32_t* pos_dev, int32_t pos, const float* sin_tab, const float* cos_tab,
                  int32_t q_dim, int32_t k_dim, int32_t head_dim, sllm_stream_t stream) {
    SLLM_REQUIRE(q && k && sin_tab && cos_tab, SLLM_EINVAL, "rope: null pointer");
    SLLM_REQUIRE(head_dim >= 2 && head_dim % 2 == 0 && q_dim % head_dim == 0 && k_dim % head_dim == 0 && q_dim > 0 && k_dim >= 0,
                 SLLM_EINVAL, "rope: dims must be multiples of head_dim (q_dim=%d k_dim=%d head_dim=%d)", q_dim, k_dim, head_dim);
    SLLM_REQUIRE(pos_dev || pos >= 0, SLLM_EINVAL, "rope: negative position");
    const int heads = std::max(q_dim, k_dim) / head_dim;
    const int threads = std::min(1024, ((head_dim / 2 + 31) / 32) * 32);
    rope_kernel<<<heads, threads, 0, as_stream(stream)>>>(q, k, pos_dev, pos, sin_tab, cos_tab, q_dim, k_dim, head_dim);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

int sllm_argmax_f32(const float* logits, int32_t n, int32_t* idx_dev, sllm_stream_t stream) {
    SLLM_REQUIRE(logits && idx_dev && n > 0, SLLM_EINVAL, "argmax: null pointer or n <= 0");
    argmax_kernel<<<1, 1024, 0, as_stream(stream)>>>(logits, n, idx_dev);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

int sllm_store_kv_row(const float* src, void* cache_row, int32_t kv_dtype, int32_t n, sllm_stream_t stream) {
    SLLM_REQUIRE(src && cache_row && n > 0 && (kv_dtype == SLLM_F32 || kv_dtype == SLLM_BF16), SLLM_EINVAL, "store_kv_row: bad arguments");
    store_kv_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(src, cache_row, kv_dtype, n);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

}  // extern "C"
