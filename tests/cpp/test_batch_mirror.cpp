// tests/cpp/test_batch_mirror.cpp — model::LlamaModel::predict_batch (the host mirror's face of sllm_batch_*, several
// prompts decoded together over a paged KV cache) against the CPU oracle run once per prompt: every sequence must get the
// tokens predict() / the reference would give it alone. Needs a GPU. Exit code 0 = all checks passed.
// (Separate from test_host_mirror.cpp on purpose: it is the newest code and runs last.)
#include <cstdio>
#include <vector>

#include <cuda_runtime_api.h>

#include "llama_oracle.h"
#include "sllm/model.h"

static int g_checks = 0, g_fail = 0;
#define CHECK(cond, ...)                                                  \
    do {                                                                  \
        ++g_checks;                                                       \
        if (!(cond)) { ++g_fail; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } \
    } while (0)

static model::LlamaModelConfig config_of(const syn_shape& s) {
    model::LlamaModelConfig c;
    c.vocab_size = s.vocab; c.head_dim = s.head_dim; c.hidden_size = s.hidden; c.kv_hidden_size = s.kv_hidden;
    c.intermediate_size = s.inter; c.max_length = s.max_len; c.num_hidden_layers = s.layers; c.num_attention_heads = s.heads;
    c.num_key_value_heads = s.kv_heads; c.rms_norm_eps = s.eps; c.rope_theta = s.theta;
    return c;
}

static void test_predict_batch(base::DataType wdt, const char* label) {
    syn_shape s{512, 32, 128, 64, 384, 48, 3, 4, 2, 1e-5f, 10000.f};   // tiny GQA model
    std::vector<float> blob((size_t)syn_blob_floats(&s));
    syn_fill_blob(&s, 1234, wdt == base::DataType::kBf16 ? SYN_BF16 : SYN_F32, 64, blob.data(), 4);
    model::LlamaModel m("", "", base::DeviceType::kDeviceCUDA);
    m.set_config(config_of(s));
    m.set_weights(blob.data(), blob.size());
    m.set_storage(wdt, base::DataType::kFp32);
    m.set_batch_capacity(3, 8);          // 3 slots: the 7 prompts below go through in waves of 3 + 3 + 1
    m.init();
    const std::vector<std::vector<int32_t>> prompts = {{1, 7, 300, 12, 44}, {5}, {9, 2}, {100, 200, 300}, {3}, {17, 18, 19, 20, 21, 22}, {511}};
    const int max_length = 40;
    const auto got = m.predict_batch(prompts, max_length);
    CHECK(got.size() == prompts.size(), "%s: %zu results for %zu prompts", label, got.size(), prompts.size());
    for (size_t i = 0; i < prompts.size() && i < got.size(); ++i) {
        std::vector<int32_t> ref(max_length);
        orc_model* o = orc_create(&s, blob.data());
        orc_greedy(o, prompts[i].data(), (int)prompts[i].size(), max_length + 1, ref.data(), nullptr);
        orc_destroy(o);
        CHECK(got[i] == ref, "%s: prompt %zu: predict_batch() tokens differ from the oracle", label, i);
    }
    // the single-sequence calls still work on the same model (per-kernel engine path), and agree
    const std::vector<int32_t> alone = m.predict(prompts[0], max_length);
    CHECK(!got.empty() && alone == got[0], "%s: predict() and predict_batch() disagree on prompt 0", label);
}

int main() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::printf("no CUDA device\n"); return 2; }
    test_predict_batch(base::DataType::kFp32, "fp32 weights");
    test_predict_batch(base::DataType::kBf16, "bf16 weights");
    std::printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "PASS", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
