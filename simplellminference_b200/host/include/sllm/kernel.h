// sllm/kernel.h — namespace kernel: the reference's CUDA kernel entry points (include/kernel/cuda/<op>_kernel.cuh),
// same names, argument order and meaning. Each one is a thin shim over one extern-"C" launcher of
// include/sllm_b200.h; a non-zero status becomes the reference's error convention (LOG -> print + exit).
// Outputs are passed as const refs and written through, exactly like the reference.
// There are no *_cpu kernels in this build: the CPU path is the reference's own, kept as the parity oracle.
#pragma once
#include "memory.h"

namespace kernel {
void add_kernel_cuda(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& output, int32_t dim_size);
void emb_kernel_cuda(const mem::Tensor& input, const mem::Tensor& weight, const mem::Tensor& output, int32_t vocab_size,
                     int32_t hidden_dim_size);
void matmul_kernel_cuda(const mem::Tensor& input, const mem::Tensor& weight, const mem::Tensor& output, int32_t dim0,
                        int32_t dim1, float scale = 1.0f);
void mha_kernel_cuda(const mem::Tensor& query, const mem::Tensor& score, const mem::Tensor& key_cache,
                     const mem::Tensor& value_cache, const mem::Tensor& mha_out, int32_t layer_index, int32_t pos,
                     int32_t max_seq_len, int32_t head_dim, int32_t hidden_dim, int32_t kv_hidden_dim,
                     int32_t att_kv_head_group, int32_t num_attention_heads, base::DeviceType device_type);
void rmsnorm_kernel_cuda(const mem::Tensor& input, const mem::Tensor& weight, const mem::Tensor& output,
                         int32_t hidden_dim_size, float eps);
void rope_cache_cal_cuda(int head_size, int max_seq_len, const mem::Tensor sin_cache, const mem::Tensor cos_cache,
                         float rope_theta);
// k_dim < 0 keeps the reference call shape (k rotated over kv_hidden = its own length); see sllm_rope_f32
void rope_kernel_cuda(const mem::Tensor& input_q, const mem::Tensor& input_k, const mem::Tensor& pos_now,
                      const mem::Tensor& sin_cache, const mem::Tensor& cos_cache, int32_t hidden_dim_size, int32_t head_dim);
void swiglu_kernel_cuda(const mem::Tensor& up, const mem::Tensor& gate, const mem::Tensor& output, int32_t intermediate_size);
// new: device arg max (the reference's argmaxLayer is CPU-only, source/op/argmax.cpp:10-15)
void argmax_kernel_cuda(const mem::Tensor& logits, const mem::Tensor& index_out, int32_t n);

// the stream every shim launches on (default: the legacy default stream, like the reference)
void set_stream(void* cuda_stream);
void* get_stream();
}  // namespace kernel
