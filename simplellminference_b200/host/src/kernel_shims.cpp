// host/src/kernel_shims.cpp — kernel::*_cuda over the extern-"C" launchers of include/sllm_b200.h.
#include <cuda_runtime_api.h>

#include "sllm/kernel.h"
#include "sllm_b200.h"

namespace kernel {

static void* g_stream = nullptr;
void set_stream(void* s) { g_stream = s; }
void* get_stream() { return g_stream; }

#define SLLM_DO(call)                                                              \
    do {                                                                           \
        if ((call) != 0) LOG(std::string(#call " failed: ") + sllm_last_error());  \
    } while (0)

static int wdtype(const mem::Tensor& w) {
    switch (w.data_type()) {
        case base::DataType::kBf16: return SLLM_BF16;
        case base::DataType::kInt8: return SLLM_INT8;
        default: return SLLM_F32;
    }
}
template <class T> static T* out_ptr(const mem::Tensor& t) { return const_cast<T*>(t.ptr<T>()); }

void add_kernel_cuda(const mem::Tensor& a, const mem::Tensor& b, const mem::Tensor& out, int32_t n) {
    SLLM_DO(sllm_add_f32(a.ptr<float>(), b.ptr<float>(), out_ptr<float>(out), n, g_stream));
}

void emb_kernel_cuda(const mem::Tensor& input, const mem::Tensor& weight, const mem::Tensor& output, int32_t vocab, int32_t d) {
    // the reference reads the token on the host (emb_kernel.cu:15); a device-resident token is accepted too
    if (input.device_type() == base::DeviceType::kDeviceCUDA)
        SLLM_DO(sllm_embedding(input.ptr<int32_t>(), 0, weight.ptr<void>(), wdtype(weight), nullptr, 64, out_ptr<float>(output), vocab, d, g_stream));
    else
        SLLM_DO(sllm_embedding(nullptr, *input.ptr<int32_t>(), weight.ptr<void>(), wdtype(weight), nullptr, 64, out_ptr<float>(output), vocab, d, g_stream));
}

void matmul_kernel_cuda(const mem::Tensor& input, const mem::Tensor& weight, const mem::Tensor& output, int32_t dim0, int32_t dim1,
                        float scale) {
    if (input.get_dim(0) != dim1) LOG("Tensor with Wrong Dim!");
    SLLM_DO(sllm_gemv(input.ptr<float>(), weight.ptr<void>(), wdtype(weight), nullptr, 64, out_ptr<float>(output), dim0, dim1, scale, g_stream));
}

// Scratch for the split-KV partials (the reference's `score` tensor is too small and not zero-initialised, so it is
// not reused): one lazily grown, zero-initialised device block per process.
static void* mha_workspace(size_t bytes) {
    static void* ws = nullptr;
    static size_t cap = 0;
    if (bytes > cap) {
        if (ws) cudaFree(ws);
        if (cudaMalloc(&ws, bytes) != cudaSuccess) LOG("mha workspace allocation failed");
        cudaMemset(ws, 0, bytes);
        cap = bytes;
    }
    return ws;
}

void mha_kernel_cuda(const mem::Tensor& query, const mem::Tensor& score, const mem::Tensor& key_cache, const mem::Tensor& value_cache,
                     const mem::Tensor& mha_out, int32_t layer_index, int32_t pos, int32_t max_seq_len, int32_t head_dim, int32_t hidden_dim,
                     int32_t kv_hidden_dim, int32_t att_kv_head_group, int32_t num_attention_heads, base::DeviceType) {
    (void)score; (void)hidden_dim; (void)att_kv_head_group;
    const int32_t kv_heads = kv_hidden_dim / head_dim;
    const int kvd = key_cache.data_type() == base::DataType::kBf16 ? SLLM_BF16 : SLLM_F32;
    void* ws = mha_workspace(sllm_mha_workspace_bytes(num_attention_heads, head_dim, max_seq_len));
    SLLM_DO(sllm_mha_decode(query.ptr<float>(), key_cache.ptr<void>(), value_cache.ptr<void>(), kvd, out_ptr<float>(mha_out), ws, layer_index,
                            nullptr, pos, max_seq_len, head_dim, num_attention_heads, kv_heads, g_stream));
}

void rmsnorm_kernel_cuda(const mem::Tensor& input, const mem::Tensor& weight, const mem::Tensor& output, int32_t d, float eps) {
    SLLM_DO(sllm_rmsnorm_f32(input.ptr<float>(), weight.ptr<float>(), out_ptr<float>(output), d, eps, g_stream));
}

void rope_cache_cal_cuda(int head_size, int max_seq_len, const mem::Tensor sin_cache, const mem::Tensor cos_cache, float theta) {
    SLLM_DO(sllm_rope_tables(head_size, max_seq_len, theta, out_ptr<float>(sin_cache), out_ptr<float>(cos_cache), g_stream));
}

void rope_kernel_cuda(const mem::Tensor& q, const mem::Tensor& k, const mem::Tensor& pos_now, const mem::Tensor& sin_cache,
                      const mem::Tensor& cos_cache, int32_t hidden_dim_size, int32_t head_dim) {
    // k is rotated over ITS OWN length (a kv_hidden-sized cache row), not over hidden_dim_size like the reference
    // does (GQA over-run).
    const int32_t k_dim = static_cast<int32_t>(k.size());
    if (pos_now.device_type() == base::DeviceType::kDeviceCUDA)
        SLLM_DO(sllm_rope_f32(out_ptr<float>(q), out_ptr<float>(k), pos_now.ptr<int32_t>(), 0, sin_cache.ptr<float>(), cos_cache.ptr<float>(),
                              hidden_dim_size, k_dim, head_dim, g_stream));
    else
        SLLM_DO(sllm_rope_f32(out_ptr<float>(q), out_ptr<float>(k), nullptr, *pos_now.ptr<int32_t>(0), sin_cache.ptr<float>(),
                              cos_cache.ptr<float>(), hidden_dim_size, k_dim, head_dim, g_stream));
}

void swiglu_kernel_cuda(const mem::Tensor& up, const mem::Tensor& gate, const mem::Tensor& output, int32_t n) {
    SLLM_DO(sllm_swiglu_f32(up.ptr<float>(), gate.ptr<float>(), out_ptr<float>(output), n, g_stream));
}

void argmax_kernel_cuda(const mem::Tensor& logits, const mem::Tensor& index_out, int32_t n) {
    SLLM_DO(sllm_argmax_f32(logits.ptr<float>(), n, out_ptr<int32_t>(index_out), g_stream));
}

}  // namespace kernel
