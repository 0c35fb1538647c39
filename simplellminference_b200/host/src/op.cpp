// host/src/op.cpp — the layer framework and the eight layers (contracts: reference source/op/*.cpp).
#include <cuda_runtime_api.h>

#include <numeric>

#include "sllm/kernel.h"
#include "sllm/op.h"
#include "sllm_b200.h"

namespace op {

static void require_cuda(base::DeviceType t, const char* layer) {
    if (t == base::DeviceType::kDeviceCUDA) return;
    if (t == base::DeviceType::kDeviceCPU)
        LOG(std::string(layer) + ": this build has no CPU kernels (the CPU path is the reference's own; construct the layer with kDeviceCUDA)");
    LOG("Device Type ERROR!");
}

// ---- Layer: slot plumbing ------------------------------------------------------------------------------
void Layer::set_weight(int32_t, const mem::Tensor&) { LOG("Function not Implementation!"); }
void Layer::set_weight(int32_t, const std::vector<int32_t>&, const void*, base::DeviceType) { LOG("Function not Implementation!"); }
void Layer::forward() { LOG("Function not Implementation!"); }

void Layer::to_cuda() {
    for (auto& t : inputs_) if (!t.is_empty()) t.to_cuda();
    for (auto& t : outputs_) if (!t.is_empty()) t.to_cuda();
}

void Layer::forward(const mem::Tensor& i1, const mem::Tensor& o1) {
    set_input(0, i1); set_output(0, o1);
    forward();
}
void Layer::forward(const mem::Tensor& i1, const mem::Tensor& i2, const mem::Tensor& o1) {
    set_input(0, i1); set_input(1, i2); set_output(0, o1);
    forward();
}
void Layer::forward(const mem::Tensor& i1, const mem::Tensor& i2, const mem::Tensor& i3, const mem::Tensor& o1) {
    set_input(0, i1); set_input(1, i2); set_input(2, i3); set_output(0, o1);
    forward();
}
void Layer::forward(const mem::Tensor& i1, const mem::Tensor& i2, const mem::Tensor& i3, const mem::Tensor& i4, const mem::Tensor& o1) {
    set_input(0, i1); set_input(1, i2); set_input(2, i3); set_input(3, i4); set_output(0, o1);
    forward();
}
void Layer::forward(const mem::Tensor& i1, const mem::Tensor& i2, const mem::Tensor& i3, const mem::Tensor& i4, const mem::Tensor& i5,
                    const mem::Tensor& o1) {
    set_input(0, i1); set_input(1, i2); set_input(2, i3); set_input(3, i4); set_input(4, i5); set_output(0, o1);
    forward();
}

// ---- LayerParam ------------------------------------------------------------------------------------------
void LayerParam::to_cuda() {
    Layer::to_cuda();
    for (auto& w : weights_) w.to_cuda();
}
void LayerParam::set_weight(int32_t idx, const mem::Tensor& weight) {
    if (weight.is_empty()) return;
    if (weight.device_type() != device_type_) LOG("Device not the same!");
    weights_.at(idx) = weight;
}
void LayerParam::set_weight(int32_t idx, const std::vector<int32_t>& dims, const void* weight_ptr, base::DeviceType device_type) {
    if (!weight_ptr) LOG("Ptr is empty!");
    mem::Tensor w(dims, false, nullptr, const_cast<void*>(weight_ptr));   // non-owning view of caller / mmapped memory
    if (device_type != base::DeviceType::kDeviceUnknown) w.set_device_type(device_type);
    weights_.at(idx) = w;
}

// ---- concrete layers ---------------------------------------------------------------------------------------
VecAddLayer::VecAddLayer(base::DeviceType t, int32_t n) : Layer(t, LayerType::kLayerAdd, "Add"), dim_size_(n) {
    reset_input_size(2); reset_output_size(1);
}
void VecAddLayer::forward() {
    require_cuda(device_type_, "VecAddLayer");
    kernel::add_kernel_cuda(get_input(0), get_input(1), get_output(0), dim_size_);
}

EmbeddingLayer::EmbeddingLayer(base::DeviceType t, int32_t vocab, int32_t d)
    : LayerParam(t, LayerType::kLayerEmbedding, "Embedding"), vocab_size_(vocab), hidden_dim_size_(d) {
    reset_weight_size(1); reset_input_size(1); reset_output_size(1);
}
void EmbeddingLayer::forward() {
    require_cuda(device_type_, "EmbeddingLayer");
    kernel::emb_kernel_cuda(get_input(0), get_weight(0), get_output(0), vocab_size_, hidden_dim_size_);
}

MatmulLayer::MatmulLayer(base::DeviceType t, int32_t dim0, int32_t dim1) : LayerParam(t, LayerType::kLayerMatmul, "Matmul"), dim0_(dim0), dim1_(dim1) {
    reset_input_size(1); reset_weight_size(1); reset_output_size(1);
}
void MatmulLayer::forward() {
    require_cuda(device_type_, "MatmulLayer");
    kernel::matmul_kernel_cuda(get_input(0), get_weight(0), get_output(0), dim0_, dim1_);
}
void MatmulLayer::quantize_weight_bf16() {
    mem::Tensor& w = get_weight(0);
    if (w.device_type() != base::DeviceType::kDeviceCUDA || w.data_type() != base::DataType::kFp32) LOG("quantize_weight_bf16 needs an fp32 weight on the device");
    mem::Tensor q({dim0_, dim1_}, true, mem::CUDADeviceAllocatorFactory::get_instance(), nullptr, base::DataType::kBf16);
    if (sllm_convert_weights(w.ptr<float>(), q.ptr<void>(), SLLM_BF16, nullptr, 64, dim0_, dim1_, kernel::get_stream()) != 0) LOG(sllm_last_error());
    cudaStreamSynchronize(static_cast<cudaStream_t>(kernel::get_stream()));
    w = q;   // the fp32 copy goes back to the pool
}

MultiHeadAttention::MultiHeadAttention(base::DeviceType t, int32_t max_seq_len, int32_t head_dim, int32_t n_heads, int32_t n_kv_heads)
    : Layer(t, LayerType::kLayerMHA, "MultiHeadAttention"), max_seq_len_(max_seq_len), head_dim_(head_dim),
      num_attention_heads_(n_heads), num_key_value_heads_(n_kv_heads) {
    reset_input_size(4); reset_output_size(1);
    hidden_dim_ = n_heads * head_dim;
    kv_hidden_dim_ = n_kv_heads * head_dim;
    att_kv_head_group_ = n_heads / n_kv_heads;
}
void MultiHeadAttention::forward() {
    require_cuda(device_type_, "MultiHeadAttention");
    kernel::mha_kernel_cuda(get_input(0), get_input(1), get_input(2), get_input(3), get_output(0), layer_index_, pos_, max_seq_len_,
                            head_dim_, hidden_dim_, kv_hidden_dim_, att_kv_head_group_, num_attention_heads_, device_type_);
}

RmsNormLayer::RmsNormLayer(base::DeviceType t, int32_t d, float eps) : LayerParam(t, LayerType::kLayerRMSNorm, "RMSNorm"), hidden_dim_size_(d), eps_(eps) {
    reset_input_size(1); reset_output_size(1); reset_weight_size(1);
}
void RmsNormLayer::forward() {
    require_cuda(device_type_, "RmsNormLayer");
    kernel::rmsnorm_kernel_cuda(get_input(0), get_weight(0), get_output(0), hidden_dim_size_, eps_);
}

RoPELayer::RoPELayer(base::DeviceType t, int32_t hidden, int32_t head_dim) : Layer(t, LayerType::kLayerRoPe, "RoPE"), hidden_dim_size_(hidden), head_dim_(head_dim) {
    reset_input_size(4); reset_output_size(1);
}
void RoPELayer::forward() {
    require_cuda(device_type_, "RoPELayer");
    kernel::rope_kernel_cuda(get_input(0), get_input(1), get_input(2), get_input(3), get_output(0), hidden_dim_size_, head_dim_);
}

SwigluLayer::SwigluLayer(base::DeviceType t, int32_t inter) : Layer(t, LayerType::kLayerSwiGLU, "SwiGLU"), intermediate_size_(inter) {
    reset_input_size(2); reset_output_size(1);
}
void SwigluLayer::forward() {
    require_cuda(device_type_, "SwigluLayer");
    kernel::swiglu_kernel_cuda(get_input(0), get_input(1), get_output(0), intermediate_size_);
}

argmaxLayer::argmaxLayer(base::DeviceType t, int32_t n) : device_type_(t), hidden_dim_size_(n) {}
void argmaxLayer::forward(const mem::Tensor& logits, const mem::Tensor& input_idx) {
    if (device_type_ != base::DeviceType::kDeviceCUDA) LOG("wrong device!\n");
    if (scratch_.is_empty()) scratch_ = mem::Tensor({1}, true, mem::CUDADeviceAllocatorFactory::get_instance(), nullptr, base::DataType::kInt32);
    kernel::argmax_kernel_cuda(logits, scratch_, hidden_dim_size_);
    // 4 bytes device -> host instead of the reference's vocab*4 bytes + CPU scan (model.cpp:175-179)
    mem::CPUDeviceAllocatorFactory::get_instance()->memcpy(scratch_.ptr<int32_t>(), const_cast<int32_t*>(input_idx.ptr<int32_t>()), sizeof(int32_t),
                                                           base::MemcpyKind::kMemcpyCUDA2CPU);
}

}  // namespace op
