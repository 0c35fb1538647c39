// csrc/batch.cu — sllm_batch_* / sllm_kvpages_*: batched multi-sequence greedy decode over a paged KV cache
// (SURVEY.md §8f rank 3). Additive: the reference decodes ONE sequence (include/model/model.h:15-18 input_token{1},
// position{1}; source/model/model.cpp:148-185); here up to 64 sequences, each at its own position, advance one token
// per step and share a single pass over the weights.
//
// A batch borrows the weights, RoPE tables and stream of an engine that stores its matrices row-major (any engine
// created without SLLM_ENGINE_MEGAKERNEL, one GPU). Per step and layer (same five-kernel split as decode_fused.cuh):
//   A  RMSNorm -> QKV GEMV for all slots -> RoPE -> q buffer + the K/V rows of each slot's position in ITS page
//   B  split-KV flash decoding per (slot, KV head) through the slot's block table (mha.cu, PAGED instantiation)
//   C  Wo GEMV + residual      D  RMSNorm -> up/gate GEMV -> sigmoid(gate)*up      E  Wdown GEMV + residual
// then RMSNorm -> tied classifier for all slots, and one CTA per slot for first-max arg-max + token/position feedback
// (prompt tokens are fed verbatim first, exactly as LlamaModel::predict does, model.cpp:157-166). Everything a step
// needs (tokens, positions, block tables) lives in device memory, so n steps are n launch sequences with no host
// round trip; the host only hands out pages ahead of the steps it enqueues.
//
// Options: a slot may sample instead of taking the arg-max (sllm_batch_set_sampling: one sllm_sample_f32 per live slot keyed by
// (seed, position)); development knobs sllm_tune 5 (replay the step as one CUDA graph per live-slot count) and 6 (GEMV body with
// four weight rows per warp), both off until measured.
//
// Pages: sllm_kvpages is plain host bookkeeping (free stack + per-sequence page lists), exported on its own so that it
// is testable without a GPU. The device pools are [pages][layers][kv_heads][page_len][head_dim] (paged_kv.cuh).
#include <algorithm>
#include <cstring>
#include <vector>

#include "batch_gemv.cuh"
#include "engine_view.cuh"
#include "prefill.cuh"

using namespace sllm;

// ---------------------------------------------------------------------------------------- page bookkeeping ---
struct sllm_kvpages {
    int n_pages = 0, page_len = 0, max_seqs = 0, max_pages = 0;
    std::vector<int32_t> free_;    // stack of free page ids; page 0 is handed out first
    std::vector<int32_t> table;    // [max_seqs][max_pages], -1 = none
    std::vector<int32_t> held;     // pages held by each sequence

    int pages_for(int n_positions) const { return (n_positions + page_len - 1) / page_len; }
    // pages a sequence would have to take to cover positions [0, n_positions); < 0: it cannot (beyond max_pages)
    int extra_needed(int seq, int n_positions) const {
        const int need = pages_for(n_positions);
        if (need > max_pages) return -1;
        return std::max(0, need - held[seq]);
    }
    void take(int seq, int extra) {
        for (int i = 0; i < extra; ++i) {
            table[(size_t)seq * max_pages + held[seq]++] = free_.back();
            free_.pop_back();
        }
    }
    void release(int seq) {
        for (int i = held[seq] - 1; i >= 0; --i) {   // so that the pages come back in the order they were handed out
            free_.push_back(table[(size_t)seq * max_pages + i]);
            table[(size_t)seq * max_pages + i] = -1;
        }
        held[seq] = 0;
    }
};

extern "C" {

sllm_kvpages* sllm_kvpages_create(int32_t n_pages, int32_t page_len, int32_t max_seqs, int32_t max_pages_per_seq) {
    if (n_pages < 1 || page_len < 1 || max_seqs < 1 || max_pages_per_seq < 1) {
        set_error("kvpages: n_pages=%d page_len=%d max_seqs=%d max_pages_per_seq=%d must all be positive", n_pages, page_len, max_seqs, max_pages_per_seq);
        return nullptr;
    }
    auto* kp = new sllm_kvpages();
    kp->n_pages = n_pages; kp->page_len = page_len; kp->max_seqs = max_seqs; kp->max_pages = max_pages_per_seq;
    kp->free_.resize(n_pages);
    for (int i = 0; i < n_pages; ++i) kp->free_[i] = n_pages - 1 - i;
    kp->table.assign((size_t)max_seqs * max_pages_per_seq, -1);
    kp->held.assign(max_seqs, 0);
    return kp;
}

void sllm_kvpages_destroy(sllm_kvpages* kp) { delete kp; }

int32_t sllm_kvpages_reserve(sllm_kvpages* kp, int32_t seq, int32_t n_positions) {
    SLLM_REQUIRE(kp && seq >= 0 && seq < kp->max_seqs && n_positions >= 0, SLLM_EINVAL, "kvpages_reserve: bad argument (seq %d, positions %d)", seq, n_positions);
    const int extra = kp->extra_needed(seq, n_positions);
    SLLM_REQUIRE(extra >= 0, SLLM_EINVAL, "kvpages_reserve: %d positions need more than %d pages of %d", n_positions, kp->max_pages, kp->page_len);
    SLLM_REQUIRE(extra <= (int)kp->free_.size(), SLLM_ENOMEM, "out of KV pages: sequence %d needs %d more, %d free", seq, extra, (int)kp->free_.size());
    kp->take(seq, extra);
    return extra;
}

int sllm_kvpages_release(sllm_kvpages* kp, int32_t seq) {
    SLLM_REQUIRE(kp && seq >= 0 && seq < kp->max_seqs, SLLM_EINVAL, "kvpages_release: bad sequence %d", seq);
    kp->release(seq);
    return SLLM_OK;
}

int32_t sllm_kvpages_free_count(const sllm_kvpages* kp) { return kp ? (int32_t)kp->free_.size() : 0; }
int32_t sllm_kvpages_held(const sllm_kvpages* kp, int32_t seq) { return (kp && seq >= 0 && seq < kp->max_seqs) ? kp->held[seq] : 0; }
const int32_t* sllm_kvpages_table(const sllm_kvpages* kp) { return kp ? kp->table.data() : nullptr; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ the batch ---
constexpr int kBatchMaxSeqs = 64;

struct sllm_batch {
    bool poisoned = false;   // a step failed between its launches (a CUDA error): the device-side positions no longer match the host's; only destroy is allowed
    EngineView ev{};
    cudaStream_t stream = nullptr;
    int max_seqs = 0, kv_dtype = SLLM_BF16, esz_kv = 2;
    int d = 0, hd = 0, L = 0, S = 0, V = 0, H = 0, KVH = 0, I = 0, q_dim = 0, kv_dim = 0;
    sllm_kvpages* pages = nullptr;
    // device arena (one cudaMalloc)
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0, arena_used = 0;
    void *k_pool = nullptr, *v_pool = nullptr;
    float *x = nullptr, *h = nullptr, *q = nullptr, *att = nullptr, *swi = nullptr, *logits = nullptr;
    int32_t *token = nullptr, *pos = nullptr, *n_prompt = nullptr, *next = nullptr, *block_table = nullptr, *prompt = nullptr, *history = nullptr;
    void* mha_ws = nullptr;
    float* part = nullptr;
    // host mirror
    std::vector<int> host_pos;     // position of the slot's next step; -1 = slot free
    int64_t total_launches = 0;
    // per-slot sampling instead of arg-max (sllm_batch_set_sampling); temperature <= 0 = arg-max
    std::vector<float> s_temp, s_top_p;
    std::vector<int32_t> s_top_k;
    std::vector<uint64_t> s_seed;
    int n_sampling = 0;            // live slots with temperature > 0
    // opt-in (sllm_tune key 5): one CUDA graph per live-slot count, captured the second time that count steps
    std::vector<cudaGraphExec_t> graphs;   // [max_seqs + 1]
    std::vector<int> graph_launches;       // kernel nodes of each graph
    std::vector<char> warmed;              // a direct step with this count has run (modules loaded, attributes set)
    // opt-in tensor-core step (sllm_batch_set_tensor_cores): the five GEMV groups become tcgen05 GEMMs over the live slots' rows
    bool tc = false;
    PfCache* tc_pf = nullptr;
    uint8_t* tc_ws = nullptr;
    uint16_t *tc_xn = nullptr, *tc_att = nullptr, *tc_s = nullptr;   // bf16 GEMM operands: normalised rows, attention output, sigmoid(gate)*up
    float *tc_qkv = nullptr, *tc_ug = nullptr;                       // fp32 GEMM results before RoPE / SwiGLU
};

template <class T>
static T* bcarve(sllm_batch* b, size_t bytes) {
    const size_t off = (b->arena_used + 255) / 256 * 256;
    b->arena_used = off + bytes;
    return b->arena ? reinterpret_cast<T*>(b->arena + off) : nullptr;
}

static void batch_layout(sllm_batch* b) {   // first pass (arena == nullptr) only measures
    b->arena_used = 0;
    const size_t ms = b->max_seqs;
    const size_t pool = (size_t)b->pages->n_pages * b->L * b->KVH * b->pages->page_len * b->hd * b->esz_kv;
    b->k_pool = bcarve<void>(b, pool);
    b->v_pool = bcarve<void>(b, pool);
    b->x = bcarve<float>(b, 4 * ms * b->d);
    b->h = bcarve<float>(b, 4 * ms * b->d);
    b->q = bcarve<float>(b, 4 * ms * b->q_dim);
    b->att = bcarve<float>(b, 4 * ms * b->q_dim);
    b->swi = bcarve<float>(b, 4 * ms * b->I);
    b->logits = bcarve<float>(b, 4 * ms * b->V);
    b->token = bcarve<int32_t>(b, 4 * ms);
    b->n_prompt = bcarve<int32_t>(b, 4 * ms);
    b->next = bcarve<int32_t>(b, 4 * ms);
    b->prompt = bcarve<int32_t>(b, 4 * ms * b->S);
    b->history = bcarve<int32_t>(b, 4 * ms * b->S);
    b->mha_ws = bcarve<void>(b, mha_paged_workspace_bytes(b->max_seqs, b->H, b->KVH, b->hd));
    b->part = bcarve<float>(b, 4 * 2 * ms * b->d);   // K-split partial sums of the down projection (sllm_tune key 7)
    // the two arrays that start at -1 (no position, no page) come last: one memset of 0xFF covers both
    b->pos = bcarve<int32_t>(b, 4 * ms);
    b->block_table = bcarve<int32_t>(b, 4 * ms * b->pages->max_pages);
}

static size_t wbytes(int dtype, int64_t n) { return dtype == SLLM_F32 ? 4 * (size_t)n : dtype == SLLM_BF16 ? 2 * (size_t)n : (size_t)n; }

// ------------------------------------------------------------------------------------------ small kernels ---
__global__ void batch_embed_kernel(const int32_t* __restrict__ token, const int32_t* __restrict__ pos, const void* __restrict__ table,
                                   int w_dtype, const float* __restrict__ scales, int group, float* __restrict__ x, int vocab, int d) {
    const int slot = blockIdx.y;
    const bool live = pos[slot] >= 0;
    const int tok = min(max(token[slot], 0), vocab - 1);
    const int64_t base = (int64_t)tok * d;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
        float v = 0.f;   // a slot that is not in use carries zeros through the step
        if (live) {
            if (w_dtype == SLLM_F32) v = reinterpret_cast<const float*>(table)[base + i];
            else if (w_dtype == SLLM_BF16) v = __uint_as_float((uint32_t) reinterpret_cast<const uint16_t*>(table)[base + i] << 16);
            else v = (float)reinterpret_cast<const int8_t*>(table)[base + i] * scales[(base + i) / group];
        }
        x[(size_t)slot * d + i] = v;
    }
}

__device__ __forceinline__ void first_max(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }   // first maximum, argmax.cpp:11
}

// One CTA per slot: arg-max of the slot's logits, then the predict loop's feedback (model.cpp:157-185): prompt tokens
// are fed verbatim while pos + 1 < n_prompt, afterwards the arg-max; history[p] = the token that followed position p.
constexpr int kArgmaxThreads = 1024;
__global__ void __launch_bounds__(kArgmaxThreads)
batch_argmax_feedback_kernel(const float* __restrict__ logits, int V, int32_t* __restrict__ token, int32_t* __restrict__ pos,
                             const int32_t* __restrict__ n_prompt, int32_t* __restrict__ next, const int32_t* __restrict__ prompt,
                             int32_t* __restrict__ history, int S) {
    __shared__ float sv[kArgmaxThreads / 32];
    __shared__ int si[kArgmaxThreads / 32];
    const int slot = blockIdx.x;
    const int p = pos[slot];
    if (p < 0) return;   // uniform over the CTA
    const float* lg = logits + (size_t)slot * V;
    float v = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < V; i += kArgmaxThreads) first_max(v, idx, lg[i], i);
    for (int o = 16; o > 0; o >>= 1) first_max(v, idx, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, idx, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sv[warp] = v; si[warp] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kArgmaxThreads / 32; ++w) first_max(v, idx, sv[w], si[w]);
        if (idx == 0x7fffffff) idx = 0;
        const int nxt = (p + 1 < n_prompt[slot]) ? prompt[(size_t)slot * S + p + 1] : idx;
        next[slot] = idx;
        history[(size_t)slot * S + p] = nxt;
        token[slot] = nxt;
        pos[slot] = p + 1;
    }
}

// Sampling mode: next[slot] was written by sllm_sample_f32 (a draw, or the arg-max for slots that do not sample); same feedback.
__global__ void batch_feedback_kernel(const int32_t* __restrict__ next, int32_t* __restrict__ token, int32_t* __restrict__ pos,
                                      const int32_t* __restrict__ n_prompt, const int32_t* __restrict__ prompt, int32_t* __restrict__ history, int S) {
    const int slot = blockIdx.x;
    if (threadIdx.x != 0) return;
    const int p = pos[slot];
    if (p < 0) return;
    const int nxt = (p + 1 < n_prompt[slot]) ? prompt[(size_t)slot * S + p + 1] : next[slot];
    history[(size_t)slot * S + p] = nxt;
    token[slot] = nxt;
    pos[slot] = p + 1;
}

// x = h + (p0 + p1): the two K halves of the split down projection (sllm_tune key 7)
__global__ void batch_add_partials_kernel(const float* __restrict__ resid, const float* __restrict__ p0, const float* __restrict__ p1,
                                          float* __restrict__ y, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = resid[i] + (p0[i] + p1[i]);
}

__global__ void batch_set_slot_kernel(int32_t* token, int32_t* pos, int32_t* n_prompt, int slot, int tok, int p, int np) {
    token[slot] = tok;
    pos[slot] = p;
    n_prompt[slot] = np;
}

// ------------------------------------------------------------------------------------------ GEMV launches ---
namespace sllm { extern int g_tune_batch_rows4, g_tune_batch_ksplit; }

template <int WD, int NB, bool FOUR, class Policy>
static int launch_bgemv_nb(sllm_batch* b, Policy& p, int units, int grid_y = 1) {
    void (*kernel)(Policy);
    if constexpr (FOUR) kernel = bgemv4_kernel<WD, NB, Policy>;
    else kernel = bgemv_kernel<WD, NB, Policy>;
    const size_t smem = bgemv_smem_bytes(p.cols_, p.nb_);
    static size_t configured = 0;   // per instantiation, FOUR included
    if (smem > 48 * 1024 && smem > configured) {
        SLLM_REQUIRE(smem <= (size_t)smem_optin_bytes(), SLLM_ENOTSUP, "%d activation vectors of %d floats do not fit shared memory", p.nb_, p.cols_);
        SLLM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int per_sm = (smem + 1024) * 2 <= (size_t)smem_optin_bytes() ? 2 : 1;   // CTAs of this size that fit one SM
    LaunchCfg lc(dim3(gemv_grid(FOUR ? (units + 1) / 2 : units, per_sm), grid_y), dim3(kGemvThreads), smem, b->stream, false);
    SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, kernel, p));
    g_launches++;
    b->total_launches++;
    return SLLM_OK;
}

template <int WD, class Policy>
static int launch_bgemv(sllm_batch* b, Policy& p, int units) {
    if (p.nb_ <= 1) return launch_bgemv_nb<WD, 1, false>(b, p, units);
    if (p.nb_ <= 2) return launch_bgemv_nb<WD, 2, false>(b, p, units);
    if (g_tune_batch_rows4) {   // experimental four-row body (sllm_tune key 6): where the shared-memory reads per FMA matter
        if (p.nb_ <= 4) return launch_bgemv_nb<WD, 4, true>(b, p, units);
        return launch_bgemv_nb<WD, kBatchMaxNb, true>(b, p, units);
    }
    if (p.nb_ <= 4) return launch_bgemv_nb<WD, 4, false>(b, p, units);
    return launch_bgemv_nb<WD, kBatchMaxNb, false>(b, p, units);
}

// activation vectors of `cols` floats one launch can stage
static int fit_nb(int cols) {
    int nb = kBatchMaxNb;
    while (nb > 1 && bgemv_smem_bytes(cols, nb) > (size_t)smem_optin_bytes()) --nb;
    return nb;
}

static const void* layer_w(const sllm_batch* b, const void* w, int64_t rows, int64_t cols, int l) {
    return reinterpret_cast<const uint8_t*>(w) + wbytes(b->ev.w_dtype, (int64_t)l * rows * cols);
}
static const float* layer_sc(const sllm_batch* b, const float* sc, int64_t rows, int64_t cols, int l) {
    return sc ? sc + (int64_t)l * rows * cols / b->ev.group : nullptr;
}

// one token for every live slot among the first `hi`
template <int WD>
static int enqueue_batch_step(sllm_batch* b, int hi) {
    const EngineView& ev = b->ev;
    const int d = b->d, L = b->L, I = b->I, V = b->V, q_dim = b->q_dim, kv_dim = b->kv_dim;
    PagedKv pk{};
    pk.block_table = b->block_table; pk.pos = b->pos; pk.max_pages = b->pages->max_pages; pk.page_len = b->pages->page_len;
    pk.layers = L; pk.q_stride = q_dim; pk.heads = b->H;
    const int nsplit = mha_paged_nsplit(b->KVH, hi, b->S);
    const int g_d = fit_nb(d), g_q = fit_nb(q_dim), g_i = fit_nb(I);

    {
        dim3 grid(std::max(1, std::min((d + 255) / 256, 64)), hi);
        batch_embed_kernel<<<grid, 256, 0, b->stream>>>(b->token, b->pos, ev.emb, ev.w_dtype, ev.emb_sc, ev.group, b->x, V, d);
        SLLM_LAUNCH_CHECK();
        g_launches++;
        b->total_launches++;
    }
    for (int l = 0; l < L; ++l) {
        for (int b0 = 0; b0 < hi; b0 += g_d) {   // A
            BQkvPolicy<WD> A{};
            A.W_ = layer_w(b, ev.wqkv, q_dim + 2 * kv_dim, d, l); A.sc_ = layer_sc(b, ev.wqkv_sc, q_dim + 2 * kv_dim, d, l);
            A.grp_ = ev.group; A.cols_ = d; A.nb_ = std::min(g_d, hi - b0); A.b0 = b0;
            A.x = b->x; A.norm_w = ev.norms + (int64_t)(2 * l) * d; A.eps = ev.shape.eps; A.sin_t = ev.sin_t; A.cos_t = ev.cos_t;
            A.q_out = b->q; A.k_pool = b->k_pool; A.v_pool = b->v_pool; A.pk = pk; A.layer = l; A.kv_dtype = b->kv_dtype;
            A.q_dim = q_dim; A.kv_dim = kv_dim; A.hd = b->hd;
            if (int rc = launch_bgemv<WD>(b, A, (q_dim + 2 * kv_dim) / 2)) return rc;
        }
        if (int rc = mha_paged_dispatch(b->q, b->k_pool, b->v_pool, b->kv_dtype, b->att, b->mha_ws, l, pk, hi, b->max_seqs, nsplit, b->hd,
                                        b->KVH, b->stream)) return rc;   // B
        b->total_launches++;
        for (int b0 = 0; b0 < hi; b0 += g_q) {   // C: h = x + Wo.att
            BResidualPolicy<WD> C{};
            C.W_ = layer_w(b, ev.wo, d, q_dim, l); C.sc_ = layer_sc(b, ev.wo_sc, d, q_dim, l);
            C.grp_ = ev.group; C.cols_ = q_dim; C.nb_ = std::min(g_q, hi - b0); C.b0 = b0;
            C.x = b->att; C.resid = b->x; C.y = b->h; C.nrows = d;
            if (int rc = launch_bgemv<WD>(b, C, (d + 1) / 2)) return rc;
        }
        for (int b0 = 0; b0 < hi; b0 += g_d) {   // D
            BGateUpPolicy<WD> D{};
            D.W_ = layer_w(b, ev.wug, 2 * (int64_t)I, d, l); D.sc_ = layer_sc(b, ev.wug_sc, 2 * (int64_t)I, d, l);
            D.grp_ = ev.group; D.cols_ = d; D.nb_ = std::min(g_d, hi - b0); D.b0 = b0;
            D.h = b->h; D.norm_w = ev.norms + (int64_t)(2 * l + 1) * d; D.eps = ev.shape.eps; D.s_out = b->swi; D.inter = I;
            if (int rc = launch_bgemv<WD>(b, D, I)) return rc;
        }
        const int E_w = WInfo<WD>::E, half = I / 2;
        const bool split_down = g_tune_batch_ksplit && g_tune_batch_rows4 && hi > g_i && hi > 2 && I % (2 * E_w) == 0 &&
                                (WD != SLLM_INT8 || half % ev.group == 0) && fit_nb(half) > g_i;
        if (split_down) {   // E' (experimental): K halves over grid.y, then x = h + (p0 + p1)
            const int g_h = fit_nb(half);
            for (int b0 = 0; b0 < hi; b0 += g_h) {
                BDownSplitPolicy<WD> E{};
                E.W_ = layer_w(b, ev.wdown, d, I, l); E.sc_ = layer_sc(b, ev.wdown_sc, d, I, l);
                E.grp_ = ev.group; E.cols_ = half; E.nb_ = std::min(g_h, hi - b0); E.b0 = b0;
                E.x = b->swi; E.part = b->part; E.nrows = d; E.row_cols_ = I; E.slots_total = b->max_seqs;
                int rc;
                if (E.nb_ <= 4) rc = launch_bgemv_nb<WD, 4, true>(b, E, (d + 1) / 2, 2);
                else rc = launch_bgemv_nb<WD, kBatchMaxNb, true>(b, E, (d + 1) / 2, 2);
                if (rc) return rc;
            }
            const int n = hi * d;
            batch_add_partials_kernel<<<std::min((n + 255) / 256, 4 * sm_count()), 256, 0, b->stream>>>(b->h, b->part, b->part + (size_t)b->max_seqs * d, b->x, n);
            SLLM_LAUNCH_CHECK();
            g_launches++;
            b->total_launches++;
        } else
        for (int b0 = 0; b0 < hi; b0 += g_i) {   // E: x = Wdown.s + h
            BResidualPolicy<WD> E{};
            E.W_ = layer_w(b, ev.wdown, d, I, l); E.sc_ = layer_sc(b, ev.wdown_sc, d, I, l);
            E.grp_ = ev.group; E.cols_ = I; E.nb_ = std::min(g_i, hi - b0); E.b0 = b0;
            E.x = b->swi; E.resid = b->h; E.y = b->x; E.nrows = d;
            if (int rc = launch_bgemv<WD>(b, E, (d + 1) / 2)) return rc;
        }
    }
    for (int b0 = 0; b0 < hi; b0 += g_d) {   // F
        BClsPolicy<WD> F{};
        F.W_ = ev.emb; F.sc_ = ev.emb_sc; F.grp_ = ev.group; F.cols_ = d; F.nb_ = std::min(g_d, hi - b0); F.b0 = b0;
        F.x = b->x; F.norm_w = ev.norms + (int64_t)(2 * L) * d; F.eps = ev.shape.eps; F.logits = b->logits; F.nrows = V;
        if (int rc = launch_bgemv<WD>(b, F, (V + 1) / 2)) return rc;
    }
    if (b->n_sampling == 0) {
        batch_argmax_feedback_kernel<<<hi, kArgmaxThreads, 0, b->stream>>>(b->logits, V, b->token, b->pos, b->n_prompt, b->next, b->prompt,
                                                                          b->history, b->S);
        SLLM_LAUNCH_CHECK();
        g_launches++;
        b->total_launches++;
        return SLLM_OK;
    }
    // some slot samples: one draw (or arg-max, temperature 0) per live slot with the op launcher of sample.cu, keyed by (seed, position)
    for (int s = 0; s < hi; ++s) {
        if (b->host_pos[s] < 0) continue;
        if (int rc = sllm_sample_f32(b->logits + (size_t)s * V, V, b->s_temp[s], b->s_top_k[s], b->s_top_p[s], b->s_seed[s],
                                     (uint64_t)b->host_pos[s], b->next + s, b->stream)) return rc;
        b->total_launches++;
    }
    batch_feedback_kernel<<<hi, 32, 0, b->stream>>>(b->next, b->token, b->pos, b->n_prompt, b->prompt, b->history, b->S);
    SLLM_LAUNCH_CHECK();
    g_launches++;
    b->total_launches++;
    return SLLM_OK;
}

namespace sllm { extern int g_tune_pf_pdl; }
// ------------------------------------------------------------------------------- tensor-core step (opt-in) ---
// With many live sequences the GEMV kernels above re-read the activations from shared memory once per weight row and stop scaling at
// ~4 sequences (profiles/r02_batch_decode.jsonl). Here a step's rows [hi][d] go through the prefill GEMM (prefill_gemm.cu: tcgen05.mma,
// TMEM accumulators, TMA operands; row-major bf16 weights): every weight byte is read once per step whatever the number of sequences.
// The price is the prefill's: GEMM operands are bf16 (activations rounded after RMSNorm / attention / SwiGLU), accumulation fp32 —
// per-sequence results agree with the reference within the bf16-operand tolerance (tests/test_zz_batch_gpu.py), not bit for bit.
// RoPE + paged K/V write of the fp32 q/k/v rows: the epilogue of BQkvPolicy as a kernel of its own (rope_kernel.cpp:36-37)
__global__ void __launch_bounds__(256) btc_rope_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ sin_t, const float* __restrict__ cos_t,
                                                         float* __restrict__ q_out, void* k_pool, void* v_pool, PagedKv pk, int layer, int kv_dtype,
                                                         int q_dim, int kv_dim, int hd) {
    const int slot = blockIdx.x;
    pdl_wait();                 // launched with programmatic stream serialization: resident early, the GEMM's results are complete from here on
    pdl_launch_dependents();
    const int pos = pk.pos[slot];
    if (pos < 0) return;
    const float* row = qkv + (size_t)slot * (q_dim + 2 * kv_dim);
    const int half = hd >> 1, rope_units = (q_dim + kv_dim) >> 1, units = (q_dim + 2 * kv_dim) >> 1, kv_heads = kv_dim / hd;
    const int pi = pos / pk.page_len, in_page = pos - pi * pk.page_len;
    const int page = pk.block_table[(size_t)slot * pk.max_pages + pi];
    auto store_kv = [&](void* pool, size_t idx, float v) {
        if (kv_dtype == SLLM_BF16) reinterpret_cast<uint16_t*>(pool)[idx] = f32_to_bf16_bits(v);
        else reinterpret_cast<float*>(pool)[idx] = v;
    };
    for (int u = blockIdx.y * blockDim.x + threadIdx.x; u < units; u += gridDim.y * blockDim.x) {   // grid.y CTAs share a slot's row
        if (u < rope_units) {
            const int head = u / half, j = u - head * half;
            const int r0 = head * hd + j;
            const float s0 = row[r0], s1 = row[r0 + half];
            const float fci = sin_t[(int64_t)pos * half + j], fcr = cos_t[(int64_t)pos * half + j];
            const float o0 = s0 * fcr - s1 * fci, o1 = s1 * fcr + s0 * fci;
            if (r0 < q_dim) {
                q_out[(size_t)slot * q_dim + r0] = o0;
                q_out[(size_t)slot * q_dim + r0 + half] = o1;
            } else {
                const int kvh = head - q_dim / hd;
                const size_t idx = paged_row_index(page, pk.layers, layer, kv_heads, kvh, pk.page_len, in_page, hd) + j;
                store_kv(k_pool, idx, o0);
                store_kv(k_pool, idx + half, o1);
            }
        } else {
            const int c = 2 * (u - rope_units);
            const int kvh = c / hd, j = c - kvh * hd;
            const size_t idx = paged_row_index(page, pk.layers, layer, kv_heads, kvh, pk.page_len, in_page, hd) + j;
            store_kv(v_pool, idx, row[q_dim + kv_dim + c]);
            store_kv(v_pool, idx + 1, row[q_dim + kv_dim + c + 1]);
        }
    }
}
__global__ void btc_to_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, size_t n4) {
    pdl_wait();
    pdl_launch_dependents();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        reinterpret_cast<uint2*>(dst)[i] = make_uint2((uint32_t)f32_to_bf16_bits(v.x) | ((uint32_t)f32_to_bf16_bits(v.y) << 16),
                                                      (uint32_t)f32_to_bf16_bits(v.z) | ((uint32_t)f32_to_bf16_bits(v.w) << 16));
    }
}
// ug = [rows][up (inter) | gate (inter)] fp32 -> s = sigmoid(gate) * up as bf16 (swiglu_kernel.cpp:12-13)
__global__ void btc_swiglu_kernel(const float* __restrict__ ug, uint16_t* __restrict__ s, int inter, size_t n) {
    pdl_wait();
    pdl_launch_dependents();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / inter, u = i - r * inter;
        const float up = ug[r * 2 * inter + u], gate = ug[r * 2 * inter + inter + u];
        s[i] = f32_to_bf16_bits((1.0f / (1.0f + expf(-gate))) * up);
    }
}

static int enqueue_batch_step_tc(sllm_batch* b, int hi) {
    const EngineView& ev = b->ev;
    const int d = b->d, L = b->L, I = b->I, V = b->V, q_dim = b->q_dim, kv_dim = b->kv_dim, nqkv = q_dim + 2 * kv_dim;
    cudaStream_t st = b->stream;
    PagedKv pk{};
    pk.block_table = b->block_table; pk.pos = b->pos; pk.max_pages = b->pages->max_pages; pk.page_len = b->pages->page_len;
    pk.layers = L; pk.q_stride = q_dim; pk.heads = b->H;
    const int nsplit = mha_paged_nsplit(b->KVH, hi, b->S);
    const int64_t before = g_launches;
    const bool pdl = g_tune_pf_pdl != 0;
    auto small_grid = [&](size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)4 * sm_count()); };
#define BT(call) do { if (int rc = (call)) return rc; } while (0)
    auto gemm = [&](const void* A, const void* W, int N, int K, int epilogue, float* out) -> int {
        PfGemmArgs a{};
        a.A = A; a.W = W; a.T = hi; a.N = N; a.K = K; a.tiled = 0; a.epilogue = epilogue; a.out = out; a.ld_out = N; a.n_valid = N;
        return pf_gemm(b->tc_pf, a, st);
    };
    {
        dim3 grid(std::max(1, std::min((d + 255) / 256, 64)), hi);
        batch_embed_kernel<<<grid, 256, 0, st>>>(b->token, b->pos, ev.emb, ev.w_dtype, ev.emb_sc, ev.group, b->x, V, d);
        SLLM_LAUNCH_CHECK();
        g_launches++;
    }
    for (int l = 0; l < L; ++l) {
        BT(pf_rmsnorm(b->x, nullptr, ev.norms + (int64_t)(2 * l) * d, b->tc_xn, hi, d, ev.shape.eps, st));                       // A
        BT(gemm(b->tc_xn, layer_w(b, ev.wqkv, nqkv, d, l), nqkv, d, PF_EPI_STORE, b->tc_qkv));
        {
            LaunchCfg lc(dim3(hi, std::max(1, std::min(16, (nqkv / 2 + 255) / 256))), dim3(256), 0, st, pdl);
            SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, btc_rope_kv_kernel, (const float*)b->tc_qkv, ev.sin_t, ev.cos_t, b->q, b->k_pool, b->v_pool, pk, l, b->kv_dtype, q_dim,
                                         kv_dim, b->hd));
        }
        BT(mha_paged_dispatch(b->q, b->k_pool, b->v_pool, b->kv_dtype, b->att, b->mha_ws, l, pk, hi, b->max_seqs, nsplit, b->hd, b->KVH, st));   // B
        {
            LaunchCfg lc(dim3(small_grid((size_t)hi * q_dim / 4)), dim3(256), 0, st, pdl);
            SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, btc_to_bf16_kernel, (const float*)b->att, b->tc_att, (size_t)hi * q_dim / 4));
        }
        BT(gemm(b->tc_att, layer_w(b, ev.wo, d, q_dim, l), d, q_dim, PF_EPI_RESID, b->x));                                          // C: x += Wo.att
        BT(pf_rmsnorm(b->x, nullptr, ev.norms + (int64_t)(2 * l + 1) * d, b->tc_xn, hi, d, ev.shape.eps, st));                   // D
        BT(gemm(b->tc_xn, layer_w(b, ev.wug, 2 * (int64_t)I, d, l), 2 * I, d, PF_EPI_STORE, b->tc_ug));
        {
            LaunchCfg lc(dim3(small_grid((size_t)hi * I)), dim3(256), 0, st, pdl);
            SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, btc_swiglu_kernel, (const float*)b->tc_ug, b->tc_s, I, (size_t)hi * I));
        }
        BT(gemm(b->tc_s, layer_w(b, ev.wdown, d, I, l), d, I, PF_EPI_RESID, b->x));                                                 // E: x += Wdown.s
        g_launches += 4;   // the three small kernels + attention
    }
    BT(pf_rmsnorm(b->x, nullptr, ev.norms + (int64_t)(2 * L) * d, b->tc_xn, hi, d, ev.shape.eps, st));                           // F
    BT(gemm(b->tc_xn, ev.emb, V, d, PF_EPI_STORE, b->logits));
#undef BT
    b->total_launches += g_launches - before;
    if (b->n_sampling == 0) {
        batch_argmax_feedback_kernel<<<hi, kArgmaxThreads, 0, st>>>(b->logits, V, b->token, b->pos, b->n_prompt, b->next, b->prompt, b->history, b->S);
        SLLM_LAUNCH_CHECK();
        g_launches++;
        b->total_launches++;
        return SLLM_OK;
    }
    for (int s = 0; s < hi; ++s) {
        if (b->host_pos[s] < 0) continue;
        if (int rc = sllm_sample_f32(b->logits + (size_t)s * V, V, b->s_temp[s], b->s_top_k[s], b->s_top_p[s], b->s_seed[s],
                                     (uint64_t)b->host_pos[s], b->next + s, st)) return rc;
        b->total_launches++;
    }
    batch_feedback_kernel<<<hi, 32, 0, st>>>(b->next, b->token, b->pos, b->n_prompt, b->prompt, b->history, b->S);
    SLLM_LAUNCH_CHECK();
    g_launches++;
    b->total_launches++;
    return SLLM_OK;
}

static int dispatch_batch_step(sllm_batch* b, int hi) {
    if (b->tc) return enqueue_batch_step_tc(b, hi);
    switch (b->ev.w_dtype) {
        case SLLM_F32: return enqueue_batch_step<SLLM_F32>(b, hi);
        case SLLM_BF16: return enqueue_batch_step<SLLM_BF16>(b, hi);
        default: return enqueue_batch_step<SLLM_INT8>(b, hi);
    }
}

namespace sllm { extern int g_tune_batch_graph; }

// One step. Default: the launch sequence itself. With sllm_tune(5, 1): the sequence is a function of `hi` alone (tokens, positions
// and block tables are read from device memory), so it is captured once per hi and replayed — 5L + 3 launches become one.
static int batch_step_once(sllm_batch* b, int hi) {
    if (!g_tune_batch_graph || b->n_sampling > 0) return dispatch_batch_step(b, hi);   // a draw's step index is a by-value argument
    if (!b->warmed[hi]) {   // first step with this count runs directly: it loads the kernels and sets their shared-memory attributes
        b->warmed[hi] = 1;
        return dispatch_batch_step(b, hi);
    }
    if (!b->graphs[hi]) {
        cudaGraph_t g = nullptr;
        const int64_t l0 = b->total_launches, gl0 = g_launches;
        SLLM_CUDA(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = dispatch_batch_step(b, hi);
        const cudaError_t ce = cudaStreamEndCapture(b->stream, &g);
        b->graph_launches[hi] = (int)(b->total_launches - l0);
        b->total_launches = l0;   // the captured pass launched nothing
        g_launches = gl0;
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        SLLM_CUDA(ce);
        const cudaError_t ci = cudaGraphInstantiate(&b->graphs[hi], g, 0);
        cudaGraphDestroy(g);
        SLLM_CUDA(ci);
    }
    SLLM_CUDA(cudaGraphLaunch(b->graphs[hi], b->stream));
    b->total_launches += b->graph_launches[hi];
    g_launches += b->graph_launches[hi];
    return SLLM_OK;
}

static int live_hi(const sllm_batch* b) {
    int hi = 0;
    for (int s = 0; s < b->max_seqs; ++s)
        if (b->host_pos[s] >= 0) hi = s + 1;
    return hi;
}

extern "C" {

int sllm_batch_create(sllm_engine* e, int32_t max_seqs, int32_t page_len, int32_t n_pages, int32_t kv_dtype, sllm_batch** out) {
    SLLM_REQUIRE(e && out, SLLM_EINVAL, "batch_create: null argument");
    SLLM_REQUIRE(max_seqs >= 1 && max_seqs <= kBatchMaxSeqs, SLLM_EINVAL, "batch_create: max_seqs=%d outside [1, %d]", max_seqs, kBatchMaxSeqs);
    SLLM_REQUIRE(page_len >= 1 && n_pages >= 1, SLLM_EINVAL, "batch_create: page_len=%d, n_pages=%d must be positive", page_len, n_pages);
    SLLM_REQUIRE(kv_dtype == SLLM_F32 || kv_dtype == SLLM_BF16, SLLM_EINVAL, "batch_create: bad kv dtype %d", kv_dtype);
    EngineView ev{};
    if (int rc = engine_view(e, &ev)) return rc;
    SLLM_REQUIRE(ev.weights_loaded, SLLM_ESTATE, "batch_create: the engine's weights are not loaded");
    SLLM_REQUIRE(ev.tp == 1, SLLM_ENOTSUP, "batched decode is single-GPU (tensor-parallel size %d)", ev.tp);
    SLLM_REQUIRE(!ev.mega, SLLM_ENOTSUP, "batched decode reads row-major weights: create the engine without SLLM_ENGINE_MEGAKERNEL (its matrices are tiled)");
    auto* b = new sllm_batch();
    b->ev = ev;
    b->stream = ev.stream;
    b->max_seqs = max_seqs;
    b->kv_dtype = kv_dtype;
    b->esz_kv = kv_dtype == SLLM_F32 ? 4 : 2;
    const sllm_shape& s = ev.shape;
    b->d = s.hidden; b->hd = s.head_dim; b->L = s.layers; b->S = s.max_len; b->V = s.vocab; b->H = s.heads; b->KVH = s.kv_heads; b->I = s.inter;
    b->q_dim = s.heads * s.head_dim; b->kv_dim = s.kv_heads * s.head_dim;
    b->pages = sllm_kvpages_create(n_pages, page_len, max_seqs, (s.max_len + page_len - 1) / page_len);
    if (!b->pages) { delete b; return SLLM_EINVAL; }
    b->host_pos.assign(max_seqs, -1);
    b->s_temp.assign(max_seqs, 0.f);
    b->s_top_p.assign(max_seqs, 0.f);
    b->s_top_k.assign(max_seqs, 0);
    b->s_seed.assign(max_seqs, 0);
    b->graphs.assign(max_seqs + 1, nullptr);
    b->graph_launches.assign(max_seqs + 1, 0);
    b->warmed.assign(max_seqs + 1, 0);
    batch_layout(b);   // measure
    b->arena_bytes = (b->arena_used + ((size_t)1 << 20) - 1) >> 20 << 20;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (b->arena_bytes > free_b) {
        set_error("batch needs %zu MiB of HBM (%d pages of %d positions), %zu MiB free", b->arena_bytes >> 20, n_pages, page_len, free_b >> 20);
        sllm_batch_destroy(b);
        return SLLM_ENOMEM;
    }
    if (cudaMalloc(&b->arena, b->arena_bytes) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc(%zu MiB) failed", b->arena_bytes >> 20);
        sllm_batch_destroy(b);
        return SLLM_ENOMEM;
    }
    batch_layout(b);   // assign
    uint8_t* minus1 = reinterpret_cast<uint8_t*>(b->pos);
    cudaError_t ce = cudaMemsetAsync(b->arena, 0, minus1 - b->arena, b->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(minus1, 0xFF, b->arena + b->arena_used - minus1, b->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(b->stream);
    if (ce != cudaSuccess) { cuda_fail(ce, "batch init", __FILE__, __LINE__); sllm_batch_destroy(b); return (int)ce; }
    *out = b;
    return SLLM_OK;
}

void sllm_batch_destroy(sllm_batch* b) {
    if (!b) return;
    if (b->arena) {
        cudaStreamSynchronize(b->stream);
        cudaFree(b->arena);
    }
    for (cudaGraphExec_t g : b->graphs)
        if (g) cudaGraphExecDestroy(g);
    if (b->tc_ws) cudaFree(b->tc_ws);
    if (b->tc_pf) pf_cache_destroy(b->tc_pf);
    sllm_kvpages_destroy(b->pages);
    delete b;
}

int sllm_batch_add(sllm_batch* b, const int32_t* prompt_host, int32_t n_prompt, int32_t* slot_out) {
    SLLM_REQUIRE(b && prompt_host && slot_out, SLLM_EINVAL, "batch_add: null argument");
    SLLM_REQUIRE(n_prompt >= 1 && n_prompt <= b->S, SLLM_EINVAL, "batch_add: prompt length %d outside [1, %d]", n_prompt, b->S);
    for (int i = 0; i < n_prompt; ++i)
        SLLM_REQUIRE(prompt_host[i] >= 0 && prompt_host[i] < b->V, SLLM_EINVAL, "Token index %d is outside the vocabulary [0, %d).", prompt_host[i], b->V);
    int slot = -1;
    for (int s = 0; s < b->max_seqs && slot < 0; ++s)
        if (b->host_pos[s] < 0) slot = s;
    SLLM_REQUIRE(slot >= 0, SLLM_ESTATE, "batch_add: all %d slots are in use", b->max_seqs);
    // pageable source: the runtime stages it before returning, so the caller's buffer is free afterwards
    SLLM_CUDA(cudaMemcpyAsync(b->prompt + (size_t)slot * b->S, prompt_host, sizeof(int32_t) * (size_t)n_prompt, cudaMemcpyHostToDevice, b->stream));
    batch_set_slot_kernel<<<1, 1, 0, b->stream>>>(b->token, b->pos, b->n_prompt, slot, prompt_host[0], 0, n_prompt);
    SLLM_LAUNCH_CHECK();
    g_launches++;
    b->total_launches++;
    b->host_pos[slot] = 0;
    *slot_out = slot;
    return SLLM_OK;
}

int sllm_batch_remove(sllm_batch* b, int32_t slot) {
    SLLM_REQUIRE(b && slot >= 0 && slot < b->max_seqs && b->host_pos[slot] >= 0, SLLM_EINVAL, "batch_remove: slot %d is not in use", slot);
    batch_set_slot_kernel<<<1, 1, 0, b->stream>>>(b->token, b->pos, b->n_prompt, slot, 0, -1, 0);
    SLLM_LAUNCH_CHECK();
    g_launches++;
    b->total_launches++;
    // the pages may be handed to another sequence right away: everything that touches them is ordered on the stream
    b->pages->release(slot);
    b->host_pos[slot] = -1;
    if (b->s_temp[slot] > 0.f) b->n_sampling--;
    b->s_temp[slot] = 0.f;
    return SLLM_OK;
}

int sllm_batch_set_sampling(sllm_batch* b, int32_t slot, float temperature, int32_t top_k, float top_p, uint64_t seed) {
    SLLM_REQUIRE(b && slot >= 0 && slot < b->max_seqs && b->host_pos[slot] >= 0, SLLM_EINVAL, "batch_set_sampling: slot %d is not in use", slot);
    SLLM_REQUIRE(top_k >= 0 && top_p >= 0.f && top_p <= 1.f, SLLM_EINVAL, "sample: top_k must be >= 0 and top_p in [0, 1]");
    const bool was = b->s_temp[slot] > 0.f, is = temperature > 0.f;
    b->n_sampling += (is ? 1 : 0) - (was ? 1 : 0);
    b->s_temp[slot] = is ? temperature : 0.f;
    b->s_top_k[slot] = top_k;
    b->s_top_p[slot] = top_p;
    b->s_seed[slot] = seed;
    return SLLM_OK;
}

/* Opt-in: run the steps of this batch on the tensor cores (see enqueue_batch_step_tc). bf16 weights only; SLLM_ENOTSUP otherwise. */
int sllm_batch_set_tensor_cores(sllm_batch* b, int32_t on) {
    SLLM_REQUIRE(b, SLLM_EINVAL, "null batch");
    if (!on) { b->tc = false; return SLLM_OK; }
    if (b->ev.w_dtype != SLLM_BF16) { set_error("tensor-core batch step: bf16 weights only"); return SLLM_ENOTSUP; }
    if (b->d % 8 || b->q_dim % 8 || b->I % 8 || b->d % 4 || b->q_dim % 4) { set_error("tensor-core batch step: hidden, q and intermediate sizes must be multiples of 8"); return SLLM_ENOTSUP; }
    if (!b->tc_ws) {
        const size_t ms = (size_t)b->max_seqs;
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 1023) / 1024 * 1024; return o; };
        const size_t o_xn = take(2 * ms * b->d), o_att = take(2 * ms * b->q_dim), o_s = take(2 * ms * b->I),
                     o_qkv = take(4 * ms * (b->q_dim + 2 * (size_t)b->kv_dim)), o_ug = take(4 * ms * 2 * (size_t)b->I);
        if (cudaMalloc(&b->tc_ws, off) != cudaSuccess) { cudaGetLastError(); set_error("tensor-core batch step: cudaMalloc(%zu MiB) failed", off >> 20); return SLLM_ENOMEM; }
        SLLM_CUDA(cudaMemsetAsync(b->tc_ws, 0, off, b->stream));
        b->tc_xn = reinterpret_cast<uint16_t*>(b->tc_ws + o_xn); b->tc_att = reinterpret_cast<uint16_t*>(b->tc_ws + o_att);
        b->tc_s = reinterpret_cast<uint16_t*>(b->tc_ws + o_s); b->tc_qkv = reinterpret_cast<float*>(b->tc_ws + o_qkv);
        b->tc_ug = reinterpret_cast<float*>(b->tc_ws + o_ug);
        b->tc_pf = pf_cache_create();
    }
    for (cudaGraphExec_t& g : b->graphs) if (g) { cudaGraphExecDestroy(g); g = nullptr; }   // graphs captured in the other mode
    std::fill(b->warmed.begin(), b->warmed.end(), 0);
    b->tc = true;
    return SLLM_OK;
}

int sllm_batch_step(sllm_batch* b, int32_t n_steps) {
    SLLM_REQUIRE(b && n_steps >= 0, SLLM_EINVAL, "batch_step: bad argument");
    SLLM_REQUIRE(!b->poisoned, SLLM_ESTATE, "batch_step: an earlier step of this batch failed between its launches; destroy the batch");
    const int hi = live_hi(b);
    if (hi == 0 || n_steps == 0) return SLLM_OK;
    // pages for every position the n steps will write, for all slots or for none
    int extra_total = 0;
    for (int s = 0; s < hi; ++s) {
        if (b->host_pos[s] < 0) continue;
        SLLM_REQUIRE(b->host_pos[s] + n_steps <= b->S, SLLM_EINVAL, "batch_step: slot %d at position %d cannot take %d more steps (max_len %d): remove it first",
                     s, b->host_pos[s], n_steps, b->S);
        extra_total += b->pages->extra_needed(s, b->host_pos[s] + n_steps);
    }
    SLLM_REQUIRE(extra_total <= sllm_kvpages_free_count(b->pages), SLLM_ENOMEM, "out of KV pages: %d steps need %d more pages, %d free", n_steps,
                 extra_total, sllm_kvpages_free_count(b->pages));
    for (int s = 0; s < hi; ++s) {
        if (b->host_pos[s] < 0) continue;
        const int extra = b->pages->extra_needed(s, b->host_pos[s] + n_steps);
        if (extra == 0) continue;
        b->pages->take(s, extra);
        const size_t row = (size_t)s * b->pages->max_pages;
        SLLM_CUDA(cudaMemcpyAsync(b->block_table + row, b->pages->table.data() + row, sizeof(int32_t) * (size_t)b->pages->max_pages,
                                  cudaMemcpyHostToDevice, b->stream));   // pageable source: staged before the call returns
    }
    for (int i = 0; i < n_steps; ++i) {
        if (int rc = batch_step_once(b, hi)) { b->poisoned = true; return rc; }   // pages already taken stay with their slots (released on remove)
        for (int s = 0; s < hi; ++s)
            if (b->host_pos[s] >= 0) b->host_pos[s] += 1;
    }
    return SLLM_OK;
}

int sllm_batch_read(sllm_batch* b, int32_t slot, int32_t* tokens_out_host, int32_t max_tokens, int32_t* n_out) {
    SLLM_REQUIRE(b && tokens_out_host && slot >= 0 && slot < b->max_seqs && b->host_pos[slot] >= 0 && max_tokens >= 0, SLLM_EINVAL,
                 "batch_read: bad argument (slot %d)", slot);
    const int n = std::min((int)max_tokens, b->host_pos[slot]);
    if (n > 0)
        SLLM_CUDA(cudaMemcpyAsync(tokens_out_host, b->history + (size_t)slot * b->S, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, b->stream));
    SLLM_CUDA(cudaStreamSynchronize(b->stream));
    if (n_out) *n_out = n;
    return SLLM_OK;
}

int sllm_batch_logits(sllm_batch* b, int32_t slot, float* logits_host) {
    SLLM_REQUIRE(b && logits_host && slot >= 0 && slot < b->max_seqs && b->host_pos[slot] > 0, SLLM_EINVAL, "batch_logits: slot %d has not stepped", slot);
    SLLM_CUDA(cudaMemcpyAsync(logits_host, b->logits + (size_t)slot * b->V, sizeof(float) * (size_t)b->V, cudaMemcpyDeviceToHost, b->stream));
    SLLM_CUDA(cudaStreamSynchronize(b->stream));
    return SLLM_OK;
}

int sllm_batch_buffer(sllm_batch* b, int32_t id, void** dev_ptr, int64_t* n_elems, int32_t* dtype) {
    SLLM_REQUIRE(b && dev_ptr && n_elems && dtype, SLLM_EINVAL, "batch_buffer: null argument");
    const int64_t ms = b->max_seqs;
    *dtype = SLLM_F32;
    switch (id) {
        case 2: *dev_ptr = b->k_pool; *dtype = b->kv_dtype; *n_elems = (int64_t)b->pages->n_pages * b->L * b->KVH * b->pages->page_len * b->hd; break;
        case 3: *dev_ptr = b->v_pool; *dtype = b->kv_dtype; *n_elems = (int64_t)b->pages->n_pages * b->L * b->KVH * b->pages->page_len * b->hd; break;
        case 4: *dev_ptr = b->x; *n_elems = ms * b->d; break;
        case 6: *dev_ptr = b->q; *n_elems = ms * b->q_dim; break;
        case 8: *dev_ptr = b->att; *n_elems = ms * b->q_dim; break;
        case 10: *dev_ptr = b->h; *n_elems = ms * b->d; break;
        case 14: *dev_ptr = b->swi; *n_elems = ms * b->I; break;
        case 16: *dev_ptr = b->logits; *n_elems = ms * b->V; break;
        default: SLLM_REQUIRE(false, SLLM_EINVAL, "batch_buffer: unknown buffer id %d", id);
    }
    SLLM_CUDA(cudaStreamSynchronize(b->stream));
    return SLLM_OK;
}

int32_t sllm_batch_free_pages(const sllm_batch* b) { return b ? sllm_kvpages_free_count(b->pages) : 0; }
int32_t sllm_batch_position(const sllm_batch* b, int32_t slot) { return (b && slot >= 0 && slot < b->max_seqs) ? b->host_pos[slot] : -1; }
int64_t sllm_batch_total_launches(const sllm_batch* b) { return b ? b->total_launches : 0; }

// Device bytes sllm_batch_create would take (page pools + per-slot buffers + attention workspace), as pure arithmetic: lets a
// host size n_pages against the free HBM (sllm_device_info) before creating anything. Same layout pass as the real one.
int64_t sllm_batch_arena_bytes(const sllm_shape* shape, int32_t max_seqs, int32_t page_len, int32_t n_pages, int32_t kv_dtype) {
    if (!shape || max_seqs < 1 || max_seqs > kBatchMaxSeqs || page_len < 1 || n_pages < 1 || (kv_dtype != SLLM_F32 && kv_dtype != SLLM_BF16) ||
        shape->head_dim < 1 || shape->heads < 1 || shape->kv_heads < 1 || shape->max_len < 1) {
        set_error("batch_arena_bytes: bad argument (max_seqs %d, page_len %d, n_pages %d, kv dtype %d)", max_seqs, page_len, n_pages, kv_dtype);
        return -1;
    }
    sllm_batch b;
    sllm_kvpages kp;
    kp.n_pages = n_pages; kp.page_len = page_len; kp.max_seqs = max_seqs; kp.max_pages = (shape->max_len + page_len - 1) / page_len;
    b.pages = &kp;
    b.max_seqs = max_seqs;
    b.kv_dtype = kv_dtype;
    b.esz_kv = kv_dtype == SLLM_F32 ? 4 : 2;
    b.d = shape->hidden; b.hd = shape->head_dim; b.L = shape->layers; b.S = shape->max_len; b.V = shape->vocab; b.H = shape->heads; b.KVH = shape->kv_heads;
    b.I = shape->inter; b.q_dim = shape->heads * shape->head_dim; b.kv_dim = shape->kv_heads * shape->head_dim;
    batch_layout(&b);   // arena == nullptr: measures only
    b.pages = nullptr;
    return (int64_t)((b.arena_used + ((size_t)1 << 20) - 1) >> 20 << 20);
}

int64_t sllm_batch_step_bytes(const sllm_batch* b) {
    if (!b) return 0;
    const double bw = b->ev.w_dtype == SLLM_F32 ? 4.0 : b->ev.w_dtype == SLLM_BF16 ? 2.0 : 1.0 + 4.0 / b->ev.group;
    const double d = b->d, kv = b->kv_dim, I = b->I, L = b->L, V = b->V;
    double bytes = bw * (V * d + L * (2 * d * d + 2 * kv * d + 3 * I * d)) + 4.0 * (2 * L + 1) * d;   // every weight once per STEP
    for (int s = 0; s < b->max_seqs; ++s) {
        if (b->host_pos[s] < 0) continue;
        bytes += bw * d + (double)b->esz_kv * 2 * L * kv * (b->host_pos[s] + 1) + (double)b->esz_kv * 2 * L * kv;   // per sequence
    }
    return (int64_t)bytes;
}

}  // extern "C"
