// host/src/model.cpp — model::LlamaModel (contract: reference source/model/model.cpp). init() binds weights from a
// headerless fp32 blob in the reference's tensor order (model.cpp:340-468); forward() is one token at one
// position; predict() is the greedy loop on token ids.
#include <cuda_runtime_api.h>
#include <algorithm>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "sllm/kernel.h"
#include "sllm/model.h"
#include "sllm_b200.h"

namespace model {

RawModelData::~RawModelData() {
    if (data != nullptr && data != MAP_FAILED) munmap(data, file_size);
    if (fd != -1) close(fd);
}

LlamaModel::LlamaModel(std::string tokenizer_path, std::string model_path, base::DeviceType device_type)
    : tokenizer_path_(std::move(tokenizer_path)), model_path_(std::move(model_path)), device_type_(device_type) {}

LlamaModel::~LlamaModel() {
    if (batch_) sllm_batch_destroy(batch_);   // borrows the engine's weights: goes first
    if (engine_) sllm_engine_destroy(engine_);
}

void LlamaModel::set_batch_capacity(int max_seqs, int page_len, int n_pages) {
    if (max_seqs < 1 || page_len < 1 || n_pages < 0) LOG("set_batch_capacity: max_seqs and page_len must be positive");
    batch_max_seqs_ = max_seqs;
    batch_page_len_ = page_len;
    batch_n_pages_ = n_pages;
}

void LlamaModel::set_config(const LlamaModelConfig& config) {
    config_ = std::make_unique<LlamaModelConfig>(config);
    config_set_ = true;
}

void LlamaModel::set_weights(const float* blob, size_t n_floats) {
    auto raw = std::make_shared<RawModelDataFp32>();
    raw->weight_data = const_cast<float*>(blob);
    raw_model_data_ = raw;
    n_weight_floats_ = n_floats;
}

void LlamaModel::init() {
    if (device_type_ != base::DeviceType::kDeviceCUDA) LOG("LlamaModel: this build runs on CUDA only (the CPU path is the reference's own)");
    if (!config_set_) config_ = std::make_unique<LlamaModelConfig>();   // the reference's hard-coded shape
    if (!raw_model_data_) read_model_file();
    Llama_layers_ = std::make_unique<LlamaLayer>();
    if (mode_ == ForwardMode::kOpByOp) {
        create_param_layers();
        create_nonparam_layers();
    } else {
        Llama_layers_->argmax_layer_ = std::make_shared<op::argmaxLayer>(device_type_, config_->vocab_size);
    }
    init_mem();
}

void LlamaModel::read_model_file() {
    if (model_path_.empty()) LOG("No model weigth file!\n");
    const int32_t fd = open(model_path_.c_str(), O_RDONLY);
    if (fd == -1) LOG("Fail to open the weight file!\n");
    struct stat sb;
    if (fstat(fd, &sb) == -1) { close(fd); LOG("Failed to retrieve the file size information from the model file\n."); }
    auto raw = std::make_shared<RawModelDataFp32>();
    raw->fd = fd;
    raw->file_size = static_cast<size_t>(sb.st_size);
    raw->data = mmap(nullptr, raw->file_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (raw->data == MAP_FAILED) { raw->data = nullptr; LOG("Weight file wrong!\n"); }
    raw->weight_data = raw->data;
    raw_model_data_ = raw;
    n_weight_floats_ = raw->file_size / sizeof(float);
}

void LlamaModel::insert_buffer(ModelBufferType id, const mem::Tensor& tensor) {
    if (buffers_.count(id) > 0) LOG(std::to_string(int(id)) + " has exits in the buffers\n");
    if (tensor.is_empty()) LOG("The tensor is empty for inserting buffer.");
    buffers_.insert({id, tensor});
}

static mem::Tensor engine_view(sllm_engine* e, int id, std::vector<int32_t> dims) {
    void* p = nullptr;
    int64_t n = 0;
    int32_t dt = 0;
    if (sllm_engine_buffer(e, id, &p, &n, &dt) != 0) LOG(sllm_last_error());
    mem::Tensor t(std::move(dims), false, nullptr, p, dt == SLLM_BF16 ? base::DataType::kBf16 : base::DataType::kFp32);
    t.set_device_type(base::DeviceType::kDeviceCUDA);
    return t;
}

void LlamaModel::init_mem() {
    const LlamaModelConfig& c = *config_;
    auto cpu = mem::CPUDeviceAllocatorFactory::get_instance();
    // token and position live on the host, like the reference's (model.cpp:258-262)
    insert_buffer(ModelBufferType::input_token, mem::Tensor({1}, true, cpu, nullptr, base::DataType::kInt32));
    insert_buffer(ModelBufferType::position, mem::Tensor({1}, true, cpu, nullptr, base::DataType::kInt32));

    if (mode_ == ForwardMode::kEngine) {
        // ONE static device arena inside the engine (weights, KV cache, activations); the named buffers are views
        sllm_engine_config ec{};
        ec.shape = {c.vocab_size, c.head_dim, c.hidden_size, c.kv_hidden_size, c.intermediate_size, c.max_length, c.num_hidden_layers,
                    c.num_attention_heads, c.num_key_value_heads, c.rms_norm_eps, c.rope_theta};
        ec.w_dtype = w_dtype_ == base::DataType::kBf16 ? SLLM_BF16 : SLLM_F32;
        ec.kv_dtype = kv_dtype_ == base::DataType::kBf16 ? SLLM_BF16 : SLLM_F32;
        ec.group = 64;
        ec.tp_rank = 0;
        ec.tp_size = 1;
        ec.flags = batch_max_seqs_ > 0 ? 0 : SLLM_ENGINE_MEGAKERNEL;   // the batched kernels read row-major matrices
        if (sllm_engine_create(&ec, kernel::get_stream(), &engine_) != 0) LOG(sllm_last_error());
        if (sllm_engine_load_blob_f32(engine_, static_cast<const float*>(raw_model_data_->weight(0)), (int64_t)n_weight_floats_) != 0)
            LOG(sllm_last_error());
        if (batch_max_seqs_ > 0) {
            const int per_seq = (c.max_length + batch_page_len_ - 1) / batch_page_len_;
            const int n_pages = batch_n_pages_ > 0 ? batch_n_pages_ : batch_max_seqs_ * per_seq;
            if (sllm_batch_create(engine_, batch_max_seqs_, batch_page_len_, n_pages, ec.kv_dtype, &batch_) != 0) LOG(sllm_last_error());
        }
        const int L = c.num_hidden_layers, S = c.max_length, kv = c.kv_hidden_size, d = c.hidden_size, I = c.intermediate_size;
        insert_buffer(ModelBufferType::key_cache, engine_view(engine_, 2, {L, S, kv}));
        insert_buffer(ModelBufferType::value_cache, engine_view(engine_, 3, {L, S, kv}));
        insert_buffer(ModelBufferType::emb_output, engine_view(engine_, 4, {d}));
        insert_buffer(ModelBufferType::query, engine_view(engine_, 6, {d}));
        insert_buffer(ModelBufferType::ffn_input, engine_view(engine_, 10, {d}));
        insert_buffer(ModelBufferType::swi_output, engine_view(engine_, 14, {I}));
        insert_buffer(ModelBufferType::model_pred, engine_view(engine_, 16, {c.vocab_size}));
        insert_buffer(ModelBufferType::sin_cache, engine_view(engine_, 17, {S, c.head_dim / 2}));
        insert_buffer(ModelBufferType::cos_cache, engine_view(engine_, 18, {S, c.head_dim / 2}));
        return;
    }

    std::shared_ptr<mem::DeviceAllocator> alloc = mem::CUDADeviceAllocatorFactory::get_instance();
    auto dev = [&](std::vector<int32_t> dims) { return mem::Tensor(std::move(dims), true, alloc); };
    mem::Tensor kc = dev({c.num_hidden_layers, c.max_length, c.kv_hidden_size}), vc = dev({c.num_hidden_layers, c.max_length, c.kv_hidden_size});
    alloc->memset_zero(kc.ptr<float>(), kc.byte_size());
    alloc->memset_zero(vc.ptr<float>(), vc.byte_size());
    insert_buffer(ModelBufferType::key_cache, kc);
    insert_buffer(ModelBufferType::value_cache, vc);
    insert_buffer(ModelBufferType::emb_output, dev({c.hidden_size}));
    insert_buffer(ModelBufferType::rms_output, dev({c.hidden_size}));
    insert_buffer(ModelBufferType::query, dev({c.hidden_size}));
    insert_buffer(ModelBufferType::score, dev({std::max(c.head_dim, c.num_attention_heads), c.max_length}));
    insert_buffer(ModelBufferType::mha_output, dev({c.hidden_size}));
    insert_buffer(ModelBufferType::att_output, dev({c.hidden_size}));
    insert_buffer(ModelBufferType::ffn_input, dev({c.hidden_size}));
    insert_buffer(ModelBufferType::up_output, dev({c.intermediate_size}));
    insert_buffer(ModelBufferType::gate_output, dev({c.intermediate_size}));
    insert_buffer(ModelBufferType::down_output, dev({c.hidden_size}));
    insert_buffer(ModelBufferType::swi_output, dev({c.intermediate_size}));
    insert_buffer(ModelBufferType::ffn_output, dev({c.hidden_size}));
    insert_buffer(ModelBufferType::model_pred, dev({c.vocab_size}));
    mem::Tensor sin_cache = dev({c.max_length, c.head_dim / 2}), cos_cache = dev({c.max_length, c.head_dim / 2});
    kernel::rope_cache_cal_cuda(c.head_dim, c.max_length, sin_cache, cos_cache, c.rope_theta);
    insert_buffer(ModelBufferType::sin_cache, sin_cache);
    insert_buffer(ModelBufferType::cos_cache, cos_cache);
}

void LlamaModel::create_nonparam_layers() {
    const LlamaModelConfig& c = *config_;
    Llama_layers_->argmax_layer_ = std::make_shared<op::argmaxLayer>(device_type_, c.vocab_size);
    Llama_layers_->add_layer_ = std::make_shared<op::VecAddLayer>(device_type_, c.hidden_size);
    Llama_layers_->mha_layer_ = std::make_shared<op::MultiHeadAttention>(device_type_, c.max_length, c.head_dim, c.num_attention_heads, c.num_key_value_heads);
    Llama_layers_->rope_layer_ = std::make_shared<op::RoPELayer>(device_type_, c.hidden_size, c.head_dim);
    Llama_layers_->swiglu_layer_ = std::make_shared<op::SwigluLayer>(device_type_, c.intermediate_size);
}

void LlamaModel::create_param_layers() {
    const LlamaModelConfig& c = *config_;
    const size_t V = c.vocab_size, d = c.hidden_size, kv = c.kv_hidden_size, I = c.intermediate_size, L = c.num_hidden_layers;
    const size_t need = V * d + (2 * L + 1) * d + L * (2 * d * d + 2 * kv * d + 3 * I * d);
    if (n_weight_floats_ && n_weight_floats_ < need) LOG("weight blob is smaller than the configured shape needs");
    size_t off = 0;
    const bool bf16 = w_dtype_ == base::DataType::kBf16;
    auto upload = [&](const std::shared_ptr<op::Layer>& layer, std::vector<int32_t> dims, size_t count, bool matrix) {
        layer->set_weight(0, dims, raw_model_data_->weight(off), base::DeviceType::kDeviceCPU);
        layer->to_cuda();
        if (matrix && bf16) std::static_pointer_cast<op::MatmulLayer>(layer)->quantize_weight_bf16();
        off += count;
    };
    // tensor order of the blob: model.cpp:340-468
    Llama_layers_->emb_layer_ = std::make_shared<op::EmbeddingLayer>(device_type_, c.vocab_size, c.hidden_size);
    upload(Llama_layers_->emb_layer_, {c.vocab_size, c.hidden_size}, V * d, false);
    // the classifier is TIED to the embedding table (model.cpp:350-352); one device copy serves both (the reference
    // uploads it twice)
    auto cls = std::make_shared<op::MatmulLayer>(device_type_, c.vocab_size, c.hidden_size);
    cls->set_weight(0, std::static_pointer_cast<op::LayerParam>(Llama_layers_->emb_layer_)->get_weight(0));
    Llama_layers_->cls_layer = cls;
    for (size_t i = 0; i < 2 * L + 1; ++i) {
        Llama_layers_->rmsnorm_layers_.push_back(std::make_shared<op::RmsNormLayer>(device_type_, c.hidden_size, c.rms_norm_eps));
        upload(Llama_layers_->rmsnorm_layers_.back(), {c.hidden_size}, d, false);
    }
    auto stack = [&](std::vector<std::shared_ptr<op::Layer>>& dst, int rows, int cols) {
        for (size_t i = 0; i < L; ++i) {
            dst.push_back(std::make_shared<op::MatmulLayer>(device_type_, rows, cols));
            upload(dst.back(), {rows, cols}, (size_t)rows * cols, true);
        }
    };
    stack(Llama_layers_->wq_layers_, c.hidden_size, c.hidden_size);
    stack(Llama_layers_->wk_layers_, c.kv_hidden_size, c.hidden_size);
    stack(Llama_layers_->wv_layers_, c.kv_hidden_size, c.hidden_size);
    stack(Llama_layers_->wo_layers_, c.hidden_size, c.hidden_size);
    stack(Llama_layers_->up_layers_, c.intermediate_size, c.hidden_size);
    stack(Llama_layers_->gate_layers_, c.intermediate_size, c.hidden_size);
    stack(Llama_layers_->down_layers_, c.hidden_size, c.intermediate_size);
}

void LlamaModel::forward() {
    if (mode_ == ForwardMode::kOpByOp) { forward_op_by_op(); return; }
    const int32_t token = get_buffer(ModelBufferType::input_token).index<int32_t>(0);
    const int32_t pos = get_buffer(ModelBufferType::position).index<int32_t>(0);
    if (sllm_engine_forward(engine_, token, pos, nullptr, nullptr) != 0) LOG(sllm_last_error());
}

// The reference's layer loop, call for call (model.cpp:48-139), through the op layers.
void LlamaModel::forward_op_by_op() {
    const LlamaModelConfig& c = *config_;
    LlamaLayer& Ls = *Llama_layers_;
    auto B = [&](ModelBufferType t) -> const mem::Tensor& { return get_buffer(t); };
    const mem::Tensor& pos_tensor = B(ModelBufferType::position);
    const int pos = pos_tensor.index<int32_t>(0);
    Ls.emb_layer_->forward(B(ModelBufferType::input_token), B(ModelBufferType::emb_output));
    for (int l = 0; l < c.num_hidden_layers; ++l) {
        Ls.rmsnorm_layers_[2 * l]->forward(B(ModelBufferType::emb_output), B(ModelBufferType::rms_output));
        auto kvp = mem::slice_KV_cache(l, pos, c.max_length, c.kv_hidden_size, B(ModelBufferType::key_cache), B(ModelBufferType::value_cache));
        Ls.wq_layers_[l]->forward(B(ModelBufferType::rms_output), B(ModelBufferType::query));
        Ls.wk_layers_[l]->forward(B(ModelBufferType::rms_output), kvp.first);
        Ls.wv_layers_[l]->forward(B(ModelBufferType::rms_output), kvp.second);
        Ls.rope_layer_->forward(B(ModelBufferType::query), kvp.first, pos_tensor, B(ModelBufferType::sin_cache), B(ModelBufferType::cos_cache));
        auto* mha = static_cast<op::MultiHeadAttention*>(Ls.mha_layer_.get());
        mha->set_pos(pos);
        mha->set_layer_index(l);
        Ls.mha_layer_->forward(B(ModelBufferType::query), B(ModelBufferType::score), B(ModelBufferType::key_cache), B(ModelBufferType::value_cache),
                               B(ModelBufferType::mha_output));
        Ls.wo_layers_[l]->forward(B(ModelBufferType::mha_output), B(ModelBufferType::att_output));
        Ls.add_layer_->forward(B(ModelBufferType::emb_output), B(ModelBufferType::att_output), B(ModelBufferType::ffn_input));
        Ls.rmsnorm_layers_[2 * l + 1]->forward(B(ModelBufferType::ffn_input), B(ModelBufferType::rms_output));
        Ls.up_layers_[l]->forward(B(ModelBufferType::rms_output), B(ModelBufferType::up_output));
        Ls.gate_layers_[l]->forward(B(ModelBufferType::rms_output), B(ModelBufferType::gate_output));
        Ls.swiglu_layer_->forward(B(ModelBufferType::up_output), B(ModelBufferType::gate_output), B(ModelBufferType::swi_output));
        Ls.down_layers_[l]->forward(B(ModelBufferType::swi_output), B(ModelBufferType::ffn_output));
        Ls.add_layer_->forward(B(ModelBufferType::ffn_output), B(ModelBufferType::ffn_input), B(ModelBufferType::emb_output));
    }
    Ls.rmsnorm_layers_[2 * c.num_hidden_layers]->forward(B(ModelBufferType::emb_output), B(ModelBufferType::rms_output));
    Ls.cls_layer->forward(B(ModelBufferType::rms_output), B(ModelBufferType::model_pred));
}

bool LlamaModel::batched_prefill_active() const {
    return batched_prefill_ && engine_ != nullptr && sllm_engine_prefill_supported(engine_) == 1;
}

std::vector<std::vector<int32_t>> LlamaModel::predict_batch(const std::vector<std::vector<int32_t>>& prompts, int max_length) {
    if (!batch_) LOG("predict_batch: call set_batch_capacity() before init() (engine mode)");
    if (max_length < 1 || max_length >= config_->max_length) LOG("predict: max_length must be below the configured context (KV cache size)");
    std::vector<std::vector<int32_t>> out(prompts.size());
    for (size_t first = 0; first < prompts.size(); first += (size_t)batch_max_seqs_) {   // waves of up to max_seqs prompts
        const size_t n = std::min(prompts.size() - first, (size_t)batch_max_seqs_);
        std::vector<int32_t> slot(n);
        for (size_t i = 0; i < n; ++i) {
            const std::vector<int32_t>& p = prompts[first + i];
            if (p.empty()) LOG("predict: empty prompt");
            if (sllm_batch_add(batch_, p.data(), (int32_t)p.size(), &slot[i]) != 0) LOG(sllm_last_error());
        }
        if (sllm_batch_step(batch_, max_length) != 0) LOG(sllm_last_error());
        for (size_t i = 0; i < n; ++i) {
            out[first + i].resize((size_t)max_length);
            int32_t got = 0;
            if (sllm_batch_read(batch_, slot[i], out[first + i].data(), max_length, &got) != 0 || got != max_length) LOG(sllm_last_error());
            if (sllm_batch_remove(batch_, slot[i]) != 0) LOG(sllm_last_error());
        }
    }
    return out;
}

std::vector<int32_t> LlamaModel::predict(const std::vector<int32_t>& prompt_ids, int max_length) {
    if (prompt_ids.empty()) LOG("predict: empty prompt");
    if (max_length >= config_->max_length) LOG("predict: max_length must be below the configured context (KV cache size)");
    std::vector<int32_t> out;
    if (mode_ == ForwardMode::kEngine && batched_prefill_active() && (int)prompt_ids.size() <= max_length) {
        // prompt in one batched pass (fills the KV cache, leaves the first generated token), then device-resident decode
        out.resize(max_length);
        const int n = (int)prompt_ids.size();
        if (sllm_engine_prefill(engine_, prompt_ids.data(), n, 0) != 0) LOG(sllm_last_error());
        if (sllm_engine_enqueue_steps(engine_, max_length - n) != 0) LOG(sllm_last_error());
        if (sllm_engine_read_tokens(engine_, out.data(), max_length) != 0) LOG(sllm_last_error());
        return out;
    }
    if (mode_ == ForwardMode::kEngine) {   // whole loop on the device, one copy of the token list at the end
        out.resize(max_length);
        if (sllm_engine_greedy(engine_, prompt_ids.data(), (int32_t)prompt_ids.size(), max_length + 1, out.data()) != 0) LOG(sllm_last_error());
        return out;
    }
    mem::Tensor token = get_buffer(ModelBufferType::input_token), position = get_buffer(ModelBufferType::position);
    const int n_prompt = (int)prompt_ids.size();
    int pos = 0;
    token.index<int32_t>(0) = prompt_ids[0];
    while (pos < max_length) {   // model.cpp:157-185
        position.index<int32_t>(0) = pos;
        forward();
        pos++;
        if (pos < n_prompt) token.index<int32_t>(0) = prompt_ids[pos];
        else Llama_layers_->argmax_layer_->forward(get_buffer(ModelBufferType::model_pred), token);
        out.push_back(token.index<int32_t>(0));
    }
    return out;
}

}  // namespace model
