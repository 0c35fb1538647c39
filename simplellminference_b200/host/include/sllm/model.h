// sllm/model.h — namespace model: LlamaModelConfig, RawModelData, ModelBufferType, LlamaLayer, LlamaModel with the
// reference's public surface (include/model/model.h:14-89, config.h:5-17, weight_loader.h:7-17):
//   LlamaModel(tokenizer_path, model_path, device_type); init(); forward(); predict(...)
// Differences, all additive: the shape is settable (the reference hard-codes it), the weight source may be a file
// (headerless fp32 blob in the reference's tensor order, model.cpp:340-468) or caller memory, predict() takes token
// ids (the tokenizer is out of scope), and forward() runs either the reference-shaped op-by-op sequence through the
// op layers or the fused engine (persistent megakernel) — same results, selectable at run time.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "op.h"

struct sllm_engine;
struct sllm_batch;

namespace model {

struct LlamaModelConfig {   // field-for-field the reference's defaults (config.h:5-17)
    int vocab_size = 128256;
    int head_dim = 128;
    int hidden_size = 3072;
    int kv_hidden_size = 1024;
    int intermediate_size = 8192;
    int max_length = 1024;
    int num_hidden_layers = 28;
    int num_attention_heads = 24;
    int num_key_value_heads = 8;
    float rms_norm_eps = 1e-05f;
    float rope_theta = 100000.0f;
};

struct RawModelData {
    virtual ~RawModelData();
    int32_t fd = -1;
    size_t file_size = 0;
    void* data = nullptr;          // mapping owned by this object (munmap'ed on destruction)
    void* weight_data = nullptr;   // first weight (== data for a mapped file, or caller memory)
    virtual const void* weight(size_t offset) const = 0;
};
struct RawModelDataFp32 : RawModelData {
    const void* weight(size_t offset) const override { return static_cast<const float*>(weight_data) + offset; }
};

enum class ModelBufferType {
    input_token = 0, position = 1, key_cache = 2, value_cache = 3, emb_output = 4, rms_output = 5, query = 6, score = 7,
    mha_output = 8, att_output = 9, ffn_input = 10, up_output = 11, gate_output = 12, down_output = 13, swi_output = 14,
    ffn_output = 15, model_pred = 16, sin_cache = 17, cos_cache = 18,
};

struct LlamaLayer {
    std::shared_ptr<op::argmaxLayer> argmax_layer_;
    std::shared_ptr<op::Layer> add_layer_, rope_layer_, swiglu_layer_, mha_layer_, emb_layer_, cls_layer;
    std::vector<std::shared_ptr<op::Layer>> wq_layers_, wk_layers_, wv_layers_, wo_layers_;
    std::vector<std::shared_ptr<op::Layer>> up_layers_, gate_layers_, down_layers_, rmsnorm_layers_;
};

enum class ForwardMode { kEngine = 0 /* fused engine / megakernel */, kOpByOp = 1 /* the reference's 13 ops per layer */ };

class LlamaModel {
public:
    explicit LlamaModel(std::string tokenizer_path, std::string model_path, base::DeviceType device_type);
    ~LlamaModel();

    // ---- additive configuration (call before init()) ----
    void set_config(const LlamaModelConfig& config);
    void set_weights(const float* blob, size_t n_floats);     // caller memory instead of model_path
    void set_forward_mode(ForwardMode mode) { mode_ = mode; }
    void set_storage(base::DataType weights, base::DataType kv_cache) { w_dtype_ = weights; kv_dtype_ = kv_cache; }
    // predict(): run the prompt as ONE batched pass on the tensor cores (sllm_engine_prefill) instead of one forward per
    // prompt token (model.cpp:157-166). Engine mode, bf16 weights, head_dim 64/128; silently keeps the token-by-token
    // prompt where the engine cannot batch it. Results agree within the bf16-operand tolerance, not bit for bit.
    void set_batched_prefill(bool on) { batched_prefill_ = on; }
    bool batched_prefill_active() const;
    // Several prompts at once (additive; the reference's input_token{1} / position{1}, model.h:15-18, generalised): with a
    // capacity set before init(), the engine keeps its matrices row-major and owns a paged KV cache of n_pages pages of
    // page_len positions (0 = enough for max_seqs sequences of max_length), and predict_batch() decodes up to max_seqs
    // prompts per wave over ONE pass of the weights per step (sllm_batch_*). forward() / predict() keep working.
    void set_batch_capacity(int max_seqs, int page_len = 64, int n_pages = 0);
    // predict() for every prompt: element i = the max_length tokens that follow prompts[i][0]; each sequence is decoded
    // exactly as predict() would decode it alone (prompt echo, then first-max arg-max feedback, no EOS stop).
    std::vector<std::vector<int32_t>> predict_batch(const std::vector<std::vector<int32_t>>& prompts, int max_length);

    void init();
    void forward();   // one token at one position: reads input_token / position (CPU tensors), fills model_pred
    // greedy loop of the reference's predict (model.cpp:148-185) on token ids: returns the max_length tokens that
    // follow prompt[0] (prompt echo, then arg-max feedback); never stops at EOS, like the reference.
    std::vector<int32_t> predict(const std::vector<int32_t>& prompt_ids, int max_length);

    const mem::Tensor& get_buffer(ModelBufferType buffer_idx) { return buffers_.at(buffer_idx); }
    const LlamaModelConfig& config() const { return *config_; }

protected:
    void init_mem();
    void insert_buffer(ModelBufferType buffer_idx, const mem::Tensor& tensor);
    void read_model_file();
    void create_param_layers();
    void create_nonparam_layers();
    void forward_op_by_op();

    std::unique_ptr<LlamaModelConfig> config_;
    std::string tokenizer_path_, model_path_;
    std::map<ModelBufferType, mem::Tensor> buffers_;
    std::shared_ptr<RawModelData> raw_model_data_;
    base::DeviceType device_type_ = base::DeviceType::kDeviceUnknown;
    std::unique_ptr<LlamaLayer> Llama_layers_;
    ForwardMode mode_ = ForwardMode::kEngine;
    base::DataType w_dtype_ = base::DataType::kFp32, kv_dtype_ = base::DataType::kFp32;
    bool config_set_ = false, batched_prefill_ = false;
    sllm_engine* engine_ = nullptr;
    sllm_batch* batch_ = nullptr;
    int batch_max_seqs_ = 0, batch_page_len_ = 64, batch_n_pages_ = 0;
    size_t n_weight_floats_ = 0;
};

}  // namespace model
