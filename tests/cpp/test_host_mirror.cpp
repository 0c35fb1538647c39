// tests/cpp/test_host_mirror.cpp — exercises the C++ host mirror (simplellminference_b200/host) the way the
// reference's own code uses these classes: op::XLayer(kDeviceCUDA, ...)->forward(in..., out) on mem::Tensors and
// model::LlamaModel::{init, forward, predict}, checking every result against the CPU oracle
// (oracle/llama_oracle.c, pinned bit-for-bit to the reference). Needs a GPU. Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <cuda_runtime_api.h>
#include <cstring>

#include "llama_oracle.h"
#include "sllm/kernel.h"
#include "sllm/model.h"

static int g_checks = 0, g_fail = 0;
#define CHECK(cond, ...)                                                  \
    do {                                                                  \
        ++g_checks;                                                       \
        if (!(cond)) { ++g_fail; std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } \
    } while (0)

using base::DeviceType;
static const auto CUDA = DeviceType::kDeviceCUDA;
static std::mt19937 rng(12345);
static std::vector<float> randn(size_t n, float scale = 1.f) {
    std::normal_distribution<float> d(0.f, scale);
    std::vector<float> v(n);
    for (auto& x : v) x = d(rng);
    return v;
}
static mem::Tensor dev(const std::vector<float>& v, std::vector<int32_t> dims) {
    mem::Tensor t(std::move(dims), true, mem::CPUDeviceAllocatorFactory::get_instance());
    std::memcpy(t.ptr<float>(), v.data(), v.size() * 4);
    t.to_cuda();
    return t;
}
static mem::Tensor dev_empty(std::vector<int32_t> dims) { return mem::Tensor(std::move(dims), true, mem::CUDADeviceAllocatorFactory::get_instance()); }
static std::vector<float> host(const mem::Tensor& t) {
    mem::Tensor c = t;
    c.to_cpu();
    return std::vector<float>(c.ptr<float>(), c.ptr<float>() + c.size());
}
static float max_abs_diff(const std::vector<float>& a, const std::vector<float>& b) {
    float m = 0;
    for (size_t i = 0; i < a.size(); ++i) m = std::fmax(m, std::fabs(a[i] - b[i]));
    return m;
}
static mem::Tensor host_i32(int32_t v) {
    mem::Tensor t({1}, true, mem::CPUDeviceAllocatorFactory::get_instance(), nullptr, base::DataType::kInt32);
    t.index<int32_t>(0) = v;
    return t;
}

static void test_allocator_and_tensor() {
    auto alloc = mem::CUDADeviceAllocatorFactory::get_instance();
    CHECK(alloc->allocate(0) == nullptr, "allocate(0) must return nullptr");
    const size_t base_use = alloc->bytes_in_use();
    void* a = alloc->allocate(1000);
    void* b = alloc->allocate(3000);
    void* c = alloc->allocate(700);
    CHECK(a && b && c && a != b && b != c, "three distinct blocks");
    CHECK(alloc->bytes_in_use() == base_use + 1024 + 3072 + 1024, "512-byte rounding (round_size), got %zu", alloc->bytes_in_use() - base_use);
    alloc->release(b);
    void* b2 = alloc->allocate(2048);
    CHECK(b2 == b, "best fit reuses the freed block");
    alloc->release(a); alloc->release(b2); alloc->release(c);
    CHECK(alloc->bytes_in_use() == base_use, "everything returned");
    void* big = alloc->allocate(4608);
    CHECK(big == a, "released neighbours were coalesced (a+b+c region reused from its start)");
    alloc->release(big);
    alloc->release(reinterpret_cast<void*>(0x1234));   // unknown pointer is ignored, like the reference

    mem::Tensor t({3, 4, 5}, true, mem::CPUDeviceAllocatorFactory::get_instance());
    CHECK(t.size() == 60 && t.byte_size() == 240 && t.dims_size() == 3 && t.get_dim(1) == 4, "shape accounting");
    auto st = t.strides();
    CHECK(st.size() == 3 && st[0] == 20 && st[1] == 5 && st[2] == 1, "strides");
    for (int i = 0; i < 60; ++i) t.index<float>(i) = (float)i;
    mem::Tensor d = t.clone();
    d.to_cuda();
    CHECK(d.device_type() == CUDA && t.device_type() == DeviceType::kDeviceCPU, "to_cuda moves the clone only");
    d.to_cpu();
    CHECK(d.index<float>(59) == 59.f && d.ptr<float>() != t.ptr<float>(), "round trip keeps the data");
    mem::Tensor bf({8, 8}, true, mem::CUDADeviceAllocatorFactory::get_instance(), nullptr, base::DataType::kBf16);
    CHECK(bf.byte_size() == 128, "dtype-aware byte size (the reference hard-codes 4 bytes)");
    mem::Tensor view({4}, false, nullptr, t.ptr<float>(8));
    CHECK(view.get_buffer()->is_external() && view.ptr<float>()[0] == 8.f, "external view does not own memory");
    CHECK(!t.assign(nullptr), "assign(nullptr) fails softly");
}

static void test_ops() {
    const int d = 288, I = 768, hd = 48, H = 6, KVH = 2, S = 64, L = 2, pos = 41, layer = 1;
    {   // add
        auto a = randn(d), b = randn(d);
        op::VecAddLayer add(CUDA, d);
        mem::Tensor out = dev_empty({d});
        add.forward(dev(a, {d}), dev(b, {d}), out);
        std::vector<float> want(d);
        orc_add(a.data(), b.data(), want.data(), d);
        CHECK(max_abs_diff(host(out), want) == 0.f, "add must be exact");
    }
    {   // rmsnorm
        auto x = randn(d), w = randn(d, 0.02f);
        for (auto& v : w) v += 1.f;
        op::RmsNormLayer rms(CUDA, d, 1e-5f);
        mem::Tensor wt = dev(w, {d});
        rms.set_weight(0, wt);
        mem::Tensor out = dev_empty({d});
        rms.forward(dev(x, {d}), out);
        std::vector<float> want(d);
        orc_rmsnorm(x.data(), w.data(), want.data(), d, 1e-5f);
        CHECK(max_abs_diff(host(out), want) < 2e-5f, "rmsnorm %g", max_abs_diff(host(out), want));
    }
    {   // matmul: host weight via set_weight(dims, ptr, CPU) + to_cuda, as create_param_layers does; then bf16 storage
        const int rows = 77;
        auto x = randn(d), W = randn((size_t)rows * d, 0.2f);
        op::MatmulLayer mm(CUDA, rows, d);
        mm.set_weight(0, {rows, d}, W.data(), DeviceType::kDeviceCPU);
        mm.to_cuda();
        mem::Tensor out = dev_empty({rows});
        mm.forward(dev(x, {d}), out);
        std::vector<float> want(rows);
        orc_matmul(x.data(), W.data(), want.data(), rows, d, 1.0f);
        CHECK(max_abs_diff(host(out), want) < 1e-4f, "matmul fp32 %g", max_abs_diff(host(out), want));
        mm.quantize_weight_bf16();
        mm.forward(dev(x, {d}), out);
        std::vector<float> Wr(W.size());
        for (size_t i = 0; i < W.size(); ++i) Wr[i] = syn_round_bf16(W[i]);
        orc_matmul(x.data(), Wr.data(), want.data(), rows, d, 1.0f);
        CHECK(max_abs_diff(host(out), want) < 1e-4f, "matmul bf16 weights %g", max_abs_diff(host(out), want));
    }
    {   // "swiglu" = sigmoid(gate) * up
        auto up = randn(I), gate = randn(I, 3.f);
        op::SwigluLayer sw(CUDA, I);
        mem::Tensor out = dev_empty({I});
        sw.forward(dev(up, {I}), dev(gate, {I}), out);
        std::vector<float> want(I);
        orc_swiglu(up.data(), gate.data(), want.data(), I);
        CHECK(max_abs_diff(host(out), want) < 1e-6f, "swiglu %g", max_abs_diff(host(out), want));
    }
    {   // embedding: host-side token like the reference
        auto table = randn((size_t)100 * d);
        op::EmbeddingLayer emb(CUDA, 100, d);
        emb.set_weight(0, {100, d}, table.data(), DeviceType::kDeviceCPU);
        emb.to_cuda();
        mem::Tensor out = dev_empty({d});
        emb.forward(host_i32(37), out);
        CHECK(max_abs_diff(host(out), std::vector<float>(table.begin() + 37 * d, table.begin() + 38 * d)) == 0.f, "embedding row copy");
    }
    {   // rope tables (bit exact: host libm) + rope with a GQA-sized k
        mem::Tensor sin_t = dev_empty({S, hd / 2}), cos_t = dev_empty({S, hd / 2});
        kernel::rope_cache_cal_cuda(hd, S, sin_t, cos_t, 10000.f);
        std::vector<float> ws((size_t)S * hd / 2), wc((size_t)S * hd / 2);
        orc_rope_cache(hd, S, 10000.f, ws.data(), wc.data());
        CHECK(max_abs_diff(host(sin_t), ws) == 0.f && max_abs_diff(host(cos_t), wc) == 0.f, "rope tables bit exact");
        auto q = randn(H * hd), k = randn(KVH * hd);
        mem::Tensor qd = dev(q, {H * hd}), kd = dev(k, {KVH * hd});
        op::RoPELayer rope(CUDA, H * hd, hd);
        rope.forward(qd, kd, host_i32(pos), sin_t, cos_t);
        orc_rope(q.data(), k.data(), pos, ws.data(), wc.data(), H * hd, KVH * hd, hd);
        CHECK(max_abs_diff(host(qd), q) < 1e-6f && max_abs_diff(host(kd), k) < 1e-6f, "rope");
        // mha over a GQA cache
        auto kc = randn((size_t)L * S * KVH * hd), vc = randn((size_t)L * S * KVH * hd), qq = randn(H * hd);
        op::MultiHeadAttention mha(CUDA, S, hd, H, KVH);
        mha.set_pos(pos);
        mha.set_layer_index(layer);
        mem::Tensor out = dev_empty({H * hd});
        mha.forward(dev(qq, {H * hd}), dev_empty({std::max(hd, H), S}), dev(kc, {L, S, KVH * hd}), dev(vc, {L, S, KVH * hd}), out);
        std::vector<float> want(H * hd), score((size_t)std::max(hd, H) * S);
        orc_mha(qq.data(), score.data(), kc.data(), vc.data(), want.data(), layer, pos, S, hd, H, KVH);
        CHECK(max_abs_diff(host(out), want) < 2e-5f, "mha %g", max_abs_diff(host(out), want));
    }
    {   // argmax: first maximum, result lands in a HOST int32 tensor
        std::vector<float> lg(5000, 0.f);
        lg[4999] = lg[1234] = lg[77] = 1.f;
        op::argmaxLayer am(CUDA, 5000);
        mem::Tensor idx = host_i32(-1);
        am.forward(dev(lg, {5000}), idx);
        CHECK(idx.index<int32_t>(0) == 77, "argmax tie -> first, got %d", idx.index<int32_t>(0));
    }
}

static void test_model(model::ForwardMode mode, base::DataType wdt, const char* label) {
    syn_shape s{512, 32, 128, 64, 384, 48, 3, 4, 2, 1e-5f, 10000.f};   // tiny GQA model
    std::vector<float> blob((size_t)syn_blob_floats(&s));
    syn_fill_blob(&s, 1234, wdt == base::DataType::kBf16 ? SYN_BF16 : SYN_F32, 64, blob.data(), 4);
    model::LlamaModelConfig c;
    c.vocab_size = s.vocab; c.head_dim = s.head_dim; c.hidden_size = s.hidden; c.kv_hidden_size = s.kv_hidden;
    c.intermediate_size = s.inter; c.max_length = s.max_len; c.num_hidden_layers = s.layers; c.num_attention_heads = s.heads;
    c.num_key_value_heads = s.kv_heads; c.rms_norm_eps = s.eps; c.rope_theta = s.theta;
    model::LlamaModel m("", "", CUDA);
    m.set_config(c);
    m.set_weights(blob.data(), blob.size());
    m.set_forward_mode(mode);
    m.set_storage(wdt, base::DataType::kFp32);
    m.init();
    orc_model* o = orc_create(&s, blob.data());
    // forward(): token/position in the host tensors, logits in model_pred
    std::vector<float> want(s.vocab);
    int32_t tok = 7;
    float worst = 0;
    for (int pos = 0; pos < 12; ++pos) {
        const_cast<mem::Tensor&>(m.get_buffer(model::ModelBufferType::input_token)).index<int32_t>(0) = tok;
        const_cast<mem::Tensor&>(m.get_buffer(model::ModelBufferType::position)).index<int32_t>(0) = pos;
        m.forward();
        orc_forward(o, tok, pos, want.data());
        worst = std::fmax(worst, max_abs_diff(host(m.get_buffer(model::ModelBufferType::model_pred)), want));
        tok = orc_argmax(want.data(), s.vocab);
    }
    float scale = 1.f;
    for (float v : want) scale = std::fmax(scale, std::fabs(v));
    CHECK(worst <= 1e-4f * scale, "%s: forward logits max|d|=%g (scale %g)", label, worst, scale);
    // predict(): greedy loop
    std::vector<int32_t> prompt = {1, 7, 300, 12, 44};
    std::vector<int32_t> got = m.predict(prompt, 45), ref(45);
    orc_model* o2 = orc_create(&s, blob.data());
    orc_greedy(o2, prompt.data(), (int)prompt.size(), 46, ref.data(), nullptr);
    CHECK(got == ref, "%s: predict() tokens differ from the oracle", label);
    orc_destroy(o);
    orc_destroy(o2);
}

// predict() with the prompt as one batched tensor-core pass: prompt echo exact, the generated tokens equal to the
// token-by-token predict() of the same model (gain-1 weights: the bf16 operand rounding stays far below the logit margins)
static void test_batched_prefill() {
    syn_shape s{2048, 64, 512, 128, 1408, 320, 3, 8, 2, 1e-5f, 10000.f};
    std::vector<float> blob((size_t)syn_blob_floats(&s));
    syn_fill_blob(&s, 21, SYN_BF16, 64, blob.data(), 4);
    for (int seg = 2; seg < 9; ++seg) {   // projections scaled to gain 1 (exact in bf16)
        int64_t off, cnt, row; float mean, cc;
        syn_segment(&s, seg, &off, &cnt, &row, &mean, &cc);
        for (int64_t i = 0; i < cnt; ++i) blob[(size_t)(off + i)] *= 0.25f;
    }
    model::LlamaModelConfig c;
    c.vocab_size = s.vocab; c.head_dim = s.head_dim; c.hidden_size = s.hidden; c.kv_hidden_size = s.kv_hidden;
    c.intermediate_size = s.inter; c.max_length = s.max_len; c.num_hidden_layers = s.layers; c.num_attention_heads = s.heads;
    c.num_key_value_heads = s.kv_heads; c.rms_norm_eps = s.eps; c.rope_theta = s.theta;
    std::vector<int32_t> prompt(150);
    for (size_t i = 0; i < prompt.size(); ++i) prompt[i] = (int32_t)(1 + (i * 37 + 11) % (s.vocab - 1));
    std::vector<int32_t> got[2];
    for (int batched = 0; batched < 2; ++batched) {
        model::LlamaModel m("", "", CUDA);
        m.set_config(c);
        m.set_weights(blob.data(), blob.size());
        m.set_storage(base::DataType::kBf16, base::DataType::kBf16);
        m.set_batched_prefill(batched != 0);
        m.init();
        CHECK(m.batched_prefill_active() == (batched != 0), "batched prefill active = %d", (int)m.batched_prefill_active());
        got[batched] = m.predict(prompt, 170);
    }
    CHECK(got[0].size() == 170 && got[1].size() == 170, "predict sizes %zu %zu", got[0].size(), got[1].size());
    bool echo = true;
    for (size_t i = 0; i + 1 < prompt.size(); ++i) echo = echo && got[1][i] == prompt[i + 1];
    CHECK(echo, "batched prefill: prompt echo differs");
    CHECK(got[0] == got[1], "batched prefill: generated tokens differ from the token-by-token prompt");
}

int main() {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::printf("no CUDA device\n"); return 2; }
    test_allocator_and_tensor();
    test_ops();
    test_model(model::ForwardMode::kOpByOp, base::DataType::kFp32, "op-by-op fp32");
    test_model(model::ForwardMode::kOpByOp, base::DataType::kBf16, "op-by-op bf16 weights");
    test_model(model::ForwardMode::kEngine, base::DataType::kFp32, "engine fp32");
    test_model(model::ForwardMode::kEngine, base::DataType::kBf16, "engine bf16 weights");
    test_batched_prefill();
    std::printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "PASS", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
