// sllm/op.h — namespace op: the reference's layer framework and its eight layers
// (include/op/layer.h:22-150, add.h, embedding.h, matmul.h, mha.h, rmsnorm.h, rope.h, swiglu.h, argmax.h).
// Constructor signatures, slot arities and forward() overloads are the reference's, so model code written against
// it compiles unchanged. Differences: layers run on CUDA only (kDeviceCPU -> LOG error: the CPU path lives in the
// reference and is this project's oracle); argmaxLayer works on the device; MatmulLayer accepts bf16 / int8 weights.
#pragma once
#include <string>

#include "memory.h"

namespace op {

enum class LayerType : uint8_t {
    kLayerUnknown = 0, kLayerLinear = 1, kLayerEncode = 2, kLayerEmbedding = 3, kLayerRMSNorm = 4, kLayerMatmul = 5,
    kLayerRoPe = 6, kLayerMHA = 7, kLayerSoftmax = 8, kLayerAdd = 9, kLayerSwiGLU = 10,
};

class BaseLayer {
public:
    explicit BaseLayer(base::DeviceType device_type, LayerType layer_type, std::string layer_name = "")
        : layer_name_(std::move(layer_name)), layer_type_(layer_type), device_type_(device_type) {}
    virtual ~BaseLayer() = default;
    LayerType layer_type() const { return layer_type_; }
    const std::string& get_layer_name() const { return layer_name_; }
    void set_layer_name(const std::string& layer_name) { layer_name_ = layer_name; }
    base::DeviceType device_type() const { return device_type_; }
    void set_device_type(base::DeviceType device_type) { device_type_ = device_type; }

    virtual void forward() = 0;
    virtual void forward(const mem::Tensor& input1, const mem::Tensor& output1) = 0;
    virtual void forward(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& output1) = 0;
    virtual void forward(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& input3,
                         const mem::Tensor& output1) = 0;
    virtual void forward(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& input3,
                         const mem::Tensor& input4, const mem::Tensor& output1) = 0;
    virtual void forward(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& input3,
                         const mem::Tensor& input4, const mem::Tensor& input5, const mem::Tensor& output1) = 0;
    virtual void set_input(int32_t idx, const mem::Tensor& input) = 0;
    virtual void set_output(int32_t idx, const mem::Tensor& output) = 0;
    virtual size_t input_size() const = 0;
    virtual size_t output_size() const = 0;
    virtual mem::Tensor& get_input(int32_t idx) = 0;
    virtual mem::Tensor& get_output(int32_t idx) = 0;
    virtual const mem::Tensor& get_input(int32_t idx) const = 0;
    virtual const mem::Tensor& get_output(int32_t idx) const = 0;
    virtual void set_weight(int32_t idx, const mem::Tensor& weight) = 0;
    virtual void set_weight(int32_t idx, const std::vector<int32_t>& dims, const void* weight_ptr,
                            base::DeviceType device_type = base::DeviceType::kDeviceUnknown) = 0;

protected:
    std::string layer_name_;
    LayerType layer_type_ = LayerType::kLayerUnknown;
    base::DeviceType device_type_ = base::DeviceType::kDeviceUnknown;
};

class Layer : public BaseLayer {
public:
    using BaseLayer::BaseLayer;
    void set_input(int32_t idx, const mem::Tensor& input) override { inputs_.at(idx) = input; }
    void set_output(int32_t idx, const mem::Tensor& output) override { outputs_.at(idx) = output; }
    const mem::Tensor& get_input(int32_t idx) const override { return inputs_.at(idx); }
    const mem::Tensor& get_output(int32_t idx) const override { return outputs_.at(idx); }
    mem::Tensor& get_input(int32_t idx) override { return inputs_.at(idx); }
    mem::Tensor& get_output(int32_t idx) override { return outputs_.at(idx); }
    size_t input_size() const override { return inputs_.size(); }
    size_t output_size() const override { return outputs_.size(); }
    void reset_input_size(size_t size) { inputs_.resize(size); }
    void reset_output_size(size_t size) { outputs_.resize(size); }
    void set_weight(int32_t idx, const mem::Tensor& weight) override;
    void set_weight(int32_t idx, const std::vector<int32_t>& dims, const void* weight_ptr,
                    base::DeviceType device_type = base::DeviceType::kDeviceUnknown) override;
    virtual void to_cuda();

    void forward() override;
    void forward(const mem::Tensor& input1, const mem::Tensor& output1) override;
    void forward(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& output1) override;
    void forward(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& input3,
                 const mem::Tensor& output1) override;
    void forward(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& input3, const mem::Tensor& input4,
                 const mem::Tensor& output1) override;
    void forward(const mem::Tensor& input1, const mem::Tensor& input2, const mem::Tensor& input3, const mem::Tensor& input4,
                 const mem::Tensor& input5, const mem::Tensor& output1) override;

protected:
    std::vector<mem::Tensor> inputs_;
    std::vector<mem::Tensor> outputs_;
};

class LayerParam : public Layer {
public:
    using Layer::Layer;
    size_t weight_size() const { return weights_.size(); }
    void reset_weight_size(size_t size) { weights_.resize(size); }
    mem::Tensor& get_weight(int32_t idx) { return weights_.at(idx); }
    const mem::Tensor& get_weight(int32_t idx) const { return weights_.at(idx); }
    void to_cuda() override;
    void set_weight(int32_t idx, const mem::Tensor& weight) override;
    void set_weight(int32_t idx, const std::vector<int32_t>& dims, const void* weight_ptr,
                    base::DeviceType device_type = base::DeviceType::kDeviceUnknown) override;

protected:
    std::vector<mem::Tensor> weights_;
};

class VecAddLayer : public Layer {
public:
    explicit VecAddLayer(base::DeviceType device_type, int32_t dim_size);
    using Layer::forward;
    void forward() override;
private:
    int32_t dim_size_;
};

class EmbeddingLayer : public LayerParam {
public:
    explicit EmbeddingLayer(base::DeviceType device_type, int32_t vocab_size, int32_t hidden_dim_size);
    using Layer::forward;
    void forward() override;
private:
    int32_t vocab_size_ = 0, hidden_dim_size_ = 0;
};

class MatmulLayer : public LayerParam {
public:
    explicit MatmulLayer(base::DeviceType device_type, int32_t dim0, int32_t dim1);
    using Layer::forward;
    void forward() override;
    // extension: convert the (device, fp32) weight to bf16 storage in place; forward() then streams half the bytes
    void quantize_weight_bf16();
private:
    int32_t dim0_ = 0, dim1_ = 0;
};

class MultiHeadAttention : public Layer {
public:
    explicit MultiHeadAttention(base::DeviceType device_type, int32_t max_seq_len, int32_t head_dim,
                                int32_t num_attention_heads, int32_t num_key_value_heads);
    void set_pos(int32_t pos) { pos_ = pos; }
    void set_layer_index(int32_t index) { layer_index_ = index; }
    using Layer::forward;
    void forward() override;
private:
    int32_t layer_index_ = 0, pos_ = 0, max_seq_len_ = 0, head_dim_ = 0, hidden_dim_ = 0, kv_hidden_dim_ = 0;
    int32_t num_attention_heads_ = 0, num_key_value_heads_ = 0, att_kv_head_group_ = 0;
};

class RmsNormLayer : public LayerParam {
public:
    explicit RmsNormLayer(base::DeviceType device_type, int32_t hidden_dim_size, float eps);
    using Layer::forward;
    void forward() override;
private:
    int32_t hidden_dim_size_ = 0;
    float eps_ = 0;
};

class RoPELayer : public Layer {
public:
    explicit RoPELayer(base::DeviceType device_type, int32_t hidden_dim_size, int32_t head_dim);
    using Layer::forward;
    void forward() override;   // slots: in0 q, in1 k, in2 pos (CPU int32), in3 sin table, out0 cos table (rope.cpp:12-16)
private:
    int32_t hidden_dim_size_ = 0, head_dim_ = 0;
};

class SwigluLayer : public Layer {
public:
    explicit SwigluLayer(base::DeviceType device_type, int32_t intermediate_size);
    using Layer::forward;
    void forward() override;   // in0 = up, in1 = gate: out = sigmoid(gate) * up (swiglu.cpp:13-25)
private:
    int32_t intermediate_size_ = 0;
};

class argmaxLayer {
public:
    explicit argmaxLayer(base::DeviceType device_type, int32_t hidden_dim_size);
    // logits on the device (CUDA layer); input_idx is a CPU int32 tensor (the reference writes the next token into
    // the host-side input_token buffer, model.cpp:170). First maximum wins.
    void forward(const mem::Tensor& logits, const mem::Tensor& input_idx);
private:
    base::DeviceType device_type_ = base::DeviceType::kDeviceUnknown;
    int32_t hidden_dim_size_ = 0;
    mem::Tensor scratch_;   // device int32
};

}  // namespace op
