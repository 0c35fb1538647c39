// csrc/engine.cu — sllm_engine_*: device arena, KV cache, weight loading/sharding and the decode step
// (replaces model::LlamaModel::{init_mem, create_param_layers, forward} and the loop of predict, reference
// source/model/model.cpp:40-187, 247-469).
//
// Memory: ONE cudaMalloc per engine (a static arena, bump-allocated at 256-byte granularity): weights, KV
// cache [L][S][kv_local], RoPE tables, activations, workspaces, device-resident step state. Nothing is
// allocated or freed in the hot loop (the reference allocates per call, rms_kernel.cu:48-51).
// Decode step: 5 fused kernels per layer + embedding + classifier/argmax, chained with programmatic dependent
// launch and captured once into a CUDA graph; token and position live in device memory and are advanced by
// the last kernel of the step, so n tokens = n graph replays with no host round trip.
// Tensor parallelism: one engine per GPU/process; Q/K/V/gate/up split by output rows, O/down by input
// columns (repacked contiguous at load), classifier by vocab rows; NCCL all-reduce after O and down.
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "engine_view.cuh"
#include "megakernel.cuh"
#include "prefill.cuh"

namespace sllm {
int mha_decode_dispatch(const float* q, const void* kc, const void* vc, int kv_dtype, float* out, void* ws, int layer,
                        const int32_t* pos_dev, int pos, int max_len, int hd, int heads, int kv_heads, cudaStream_t st, bool pdl);
size_t mha_workspace_bytes(int heads, int head_dim, int max_len);
int check_gemv_args(const void* W, int w_dtype, const float* scales, int group, int rows, int cols);
extern int g_tune_mega_debug;   // runtime.cu: measurement aids (sllm_tune key 8)
}  // namespace sllm

using namespace sllm;

#define SLLM_NCCL(call)                                                                           \
    do {                                                                                          \
        ncclResult_t _r = (call);                                                                 \
        if (_r != ncclSuccess) {                                                                  \
            set_error("NCCL error %d (%s) at %s:%d in %s", (int)_r, ncclGetErrorString(_r), __FILE__, __LINE__, #call); \
            return SLLM_ECOMM;                                                                    \
        }                                                                                         \
    } while (0)

struct Matrix {          // a weight matrix (or stack of per-layer matrices) in storage dtype
    void* w = nullptr;   // final location (row-major; TILED in megakernel mode, see megakernel.cuh)
    void* rm = nullptr;  // megakernel mode, during loading only: temporary row-major staging copy
    float* sc = nullptr; // int8 group scales
    int64_t rows = 0, cols = 0;   // per layer
    int64_t layers = 1;
    int kind = 0;        // PH_* pairing rule of the tiled layout
};

struct sllm_engine {
    sllm_engine_config cfg{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool fused = true, use_graph = true, pdl = true;
    bool mega = false;            // persistent one-launch-per-token kernel (megakernel.cu / megakernel_ll.cu)
    bool mega_ll = false;         // barrier-free {value, epoch}-word version (also the tensor-parallel one)
    bool mega_fuse = false;       // experimental: down projection fused into the gate_up phase (SLLM_ENGINE_MEGA_FUSE_DOWN)
    bool mega2 = false;           // megakernel2.cu: two grid-wide dependency points per layer (SLLM_ENGINE_MEGA_V2)
    void* wo_t = nullptr;         // its column-block copy of every layer's Wo (megakernel.cuh "PH_WO_T"); wo stays for the batched prefill
    float* v2_bufs = nullptr;     // xbuf[2], hbuf[2]: d floats each, touched by that kernel only (zero-on-entry invariants)
    unsigned* v2_flags = nullptr; // per-kv-head dependency counters
    Mega2Params mega2_params{};
    std::vector<PhaseDesc> phases_host;   // what phases_dev holds (sllm_engine_calibrate rewrites the share tables in it)
    uint32_t* cum_dev = nullptr;           // share tables of the calibrated partition: up to kCumTables x (grid + 1) words
    std::vector<double> calib_tau;         // measured relative time per byte of every CTA (empty = never calibrated)
    void* wdown_t = nullptr;      // its transposed per-layer matrices (megakernel.cuh "PH_DOWN_T"); wdown stays for the batched prefill
    MegaLLPlan ll_plan{};
    MegaLLParams ll_params{};
    uint8_t* ll_block = nullptr;  // this rank's word area (own cudaMalloc: exported through CUDA IPC under TP)
    bool ll_ready = false;
    MegaPlan mega_plan_{};
    MegaParams mega_params{};
    PhaseDesc* phases_dev = nullptr;
    float* mega_att = nullptr;
    unsigned* bar_counter = nullptr;
    unsigned long long* trace = nullptr;   // development aid: per-CTA phase timeline of the megakernel
    // local (this rank's) dimensions
    int d = 0, hd = 0, L = 0, S = 0, V = 0, H = 0, KVH = 0, I = 0;
    int tp = 1, rank = 0;
    int q_loc = 0, kv_loc = 0, I_loc = 0, V_loc = 0, v0 = 0, H_loc = 0, KVH_loc = 0;
    int esz_w = 2, esz_kv = 2;
    // arena
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0, arena_used = 0;
    // weights
    Matrix emb, wqkv, wo, wug, wdown;
    float* norms = nullptr;
    bool weights_loaded = false;
    // caches, tables, activations (names follow ModelBufferType, model.h:14-34)
    void *key_cache = nullptr, *value_cache = nullptr;
    float *sin_t = nullptr, *cos_t = nullptr;
    float *x = nullptr /*emb_output*/, *xb = nullptr /*rms_output*/, *q = nullptr /*query*/, *att = nullptr /*mha_output*/,
          *att_out = nullptr /*att_output*/, *h = nullptr /*ffn_input*/, *up = nullptr, *gate = nullptr, *swi = nullptr,
          *ffn_out = nullptr, *logits = nullptr /*model_pred*/, *kv_tmp = nullptr;
    float *part_a = nullptr, *part_b = nullptr;   // TP partial sums (wo / down)
    void* mha_ws = nullptr;
    float* blk_val = nullptr;
    int32_t* blk_idx = nullptr;
    float* tp_pairs = nullptr;   // [tp][2] gathered (value, index-as-float-bits)
    StepState* state = nullptr;
    int32_t *prompt_dev = nullptr, *history_dev = nullptr;
    int cls_grid = 0;
    // pinned host staging
    int32_t* h_state = nullptr;   // 8 ints
    // graph
    cudaGraphExec_t graph_exec = nullptr;
    int step_launches = 0;
    int64_t total_launches = 0;
    // comm
    ncclComm_t comm = nullptr;
    bool p2p_mode = false;            // all-reduce over NVLink peer memory fused into the GEMV kernels (no NCCL call)
    bool p2p_ready = false;
    bool timing_only = false;         // sllm_engine_enqueue_kernel: single kernels must not wait for peers
    uint8_t* p2p_block = nullptr;     // this rank's receive area + flags (own cudaMalloc: exported through CUDA IPC)
    size_t p2p_recv_bytes = 0;
    void* p2p_peer[kMaxTp] = {};      // mapped peer blocks
    P2PComm* p2p_dev = nullptr;       // descriptor in device memory
    // batched prefill (prefill.cuh): tensor-map cache + a workspace allocated on first use, sized for pf_rows prompt rows
    PfCache* pf = nullptr;
    uint8_t* pf_ws = nullptr;
    int pf_rows = 0;
    // tensor-parallel prefill exchange over peer memory (prefill_tp.cu): this rank's block, the peers' mappings, call counter
    uint8_t* pfx_block = nullptr;
    void* pfx_peer[kMaxTp] = {};
    bool pfx_ready = false;
    uint32_t pfx_calls = 0, pfx_ctas = 0;
    bool pfx_dirty = false;
    float *pf_x = nullptr, *pf_part = nullptr;
    uint16_t *pf_xn = nullptr, *pf_q = nullptr, *pf_att = nullptr, *pf_s = nullptr;
    int32_t* pf_ids = nullptr;
    int64_t pf_gemm_launches = 0;
};

constexpr int kCumTables = 8;
// ------------------------------------------------------------------------------------------- helpers ---
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr int kPrefillBlock = 1024;   // prompt rows per pass of the batched prefill through the layers
static size_t wbytes(int dtype, int64_t n) { return dtype == SLLM_F32 ? 4 * (size_t)n : dtype == SLLM_BF16 ? 2 * (size_t)n : (size_t)n; }

template <class T>
static T* carve(sllm_engine* e, size_t bytes) {
    const size_t off = align_up(e->arena_used, 256);
    e->arena_used = off + bytes;
    return e->arena ? reinterpret_cast<T*>(e->arena + off) : nullptr;
}

static void carve_matrix(sllm_engine* e, Matrix& m, int64_t layers, int64_t rows, int64_t cols, int kind) {
    m.rows = rows;
    m.cols = cols;
    m.layers = layers;
    m.kind = kind;
    const size_t bytes = e->mega ? (size_t)layers * mega_matrix_bytes((int)rows, (int)cols, kind, e->cfg.w_dtype)
                                 : wbytes(e->cfg.w_dtype, layers * rows * cols);
    m.w = carve<void>(e, bytes);
    m.sc = (e->cfg.w_dtype == SLLM_INT8) ? carve<float>(e, sizeof(float) * (size_t)(layers * rows * cols / e->cfg.group)) : nullptr;
}

// Lays out every buffer; first pass (arena == nullptr) only measures.
static void layout(sllm_engine* e) {
    e->arena_used = 0;
    const int64_t d = e->d, L = e->L, S = e->S;
    carve_matrix(e, e->emb, 1, e->V, d, PH_CLS);
    e->norms = carve<float>(e, sizeof(float) * (size_t)((2 * L + 1) * d));
    carve_matrix(e, e->wqkv, L, e->q_loc + 2 * e->kv_loc, d, PH_QKV);
    carve_matrix(e, e->wo, L, d, e->q_loc, PH_WO);
    carve_matrix(e, e->wug, L, 2 * e->I_loc, d, PH_GATEUP);
    carve_matrix(e, e->wdown, L, d, e->I_loc, PH_DOWN);
    e->key_cache = carve<void>(e, (size_t)e->esz_kv * L * S * e->kv_loc);
    e->value_cache = carve<void>(e, (size_t)e->esz_kv * L * S * e->kv_loc);
    e->sin_t = carve<float>(e, sizeof(float) * (size_t)S * (e->hd / 2));
    e->cos_t = carve<float>(e, sizeof(float) * (size_t)S * (e->hd / 2));
    e->x = carve<float>(e, 4 * (size_t)d);
    e->xb = carve<float>(e, 4 * (size_t)d);
    e->q = carve<float>(e, 4 * (size_t)std::max(e->q_loc, (int)d));
    e->att = carve<float>(e, 4 * (size_t)std::max(e->q_loc, (int)d));
    e->att_out = carve<float>(e, 4 * (size_t)d);
    e->h = carve<float>(e, 4 * (size_t)d);
    e->up = carve<float>(e, 4 * (size_t)e->I_loc);
    e->gate = carve<float>(e, 4 * (size_t)e->I_loc);
    e->swi = carve<float>(e, 4 * (size_t)e->I_loc);
    e->ffn_out = carve<float>(e, 4 * (size_t)d);
    e->logits = carve<float>(e, 4 * (size_t)e->V_loc);
    e->kv_tmp = carve<float>(e, 4 * (size_t)2 * e->kv_loc);
    e->part_a = carve<float>(e, 4 * (size_t)d);
    e->part_b = carve<float>(e, 4 * (size_t)d);
    e->mha_ws = carve<void>(e, mha_workspace_bytes(e->H_loc, e->hd, e->S));
    e->cls_grid = gemv_grid((e->V_loc + 1) / 2, 2);
    const size_t nblk = (size_t)std::max(e->cls_grid, sm_count()) + 1;
    e->blk_val = carve<float>(e, 4 * nblk);
    e->blk_idx = carve<int32_t>(e, 4 * nblk);
    e->tp_pairs = carve<float>(e, 8 * (size_t)std::max(1, e->tp));
    e->state = carve<StepState>(e, sizeof(StepState));
    e->mega_att = carve<float>(e, 4 * std::max<size_t>(4, e->mega_plan_.att_part_floats));
    e->bar_counter = carve<unsigned>(e, 256);
    e->phases_dev = carve<PhaseDesc>(e, sizeof(PhaseDesc) * (size_t)(4 * L + 1));
    e->prompt_dev = carve<int32_t>(e, 4 * (size_t)S);
    e->history_dev = carve<int32_t>(e, 4 * (size_t)S);
    e->wdown_t = e->mega_fuse ? carve<void>(e, (size_t)L * mega_down_t_bytes(e->d, e->I_loc, e->cfg.w_dtype)) : nullptr;
    e->wo_t = e->mega2 ? carve<void>(e, (size_t)L * mega2_wot_bytes(e->d, e->hd, e->H_loc, e->KVH_loc, e->cfg.w_dtype)) : nullptr;
    e->v2_bufs = e->mega2 ? carve<float>(e, 4 * (size_t)4 * d) : nullptr;
    e->v2_flags = e->mega2 ? carve<unsigned>(e, (size_t)2 * e->KVH_loc * 128) : nullptr;
    e->cum_dev = (e->mega && !e->mega_ll) ? carve<uint32_t>(e, sizeof(uint32_t) * (size_t)kCumTables * (size_t)(e->mega_plan_.grid + 1)) : nullptr;
}

static int count_launch(sllm_engine* e) {
    e->total_launches++;
    return 0;
}

// ------------------------------------------------------------------------------------- fused launches --
template <int WD, class Policy>
static int launch_policy(sllm_engine* e, Policy& p, int units, int grid_override = 0) {
    const size_t smem = gemv_smem_bytes(p.cols_);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        SLLM_REQUIRE(smem <= (size_t)smem_optin_bytes(), SLLM_ENOTSUP, "input length %d does not fit shared memory", p.cols_);
        SLLM_CUDA(cudaFuncSetAttribute(fused_gemv_kernel<WD, Policy>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int grid = grid_override ? grid_override : gemv_grid(units, 2);
    LaunchCfg lc(dim3(grid), dim3(kGemvThreads), smem, e->stream, e->pdl);
    SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, fused_gemv_kernel<WD, Policy>, p));
    g_launches++;
    count_launch(e);
    return SLLM_OK;
}

static const void* mat_layer(const sllm_engine* e, const Matrix& m, int l) {
    return reinterpret_cast<const uint8_t*>(m.w) + wbytes(e->cfg.w_dtype, (int64_t)l * m.rows * m.cols);
}
static const float* sc_layer(const sllm_engine* e, const Matrix& m, int l) {
    return m.sc ? m.sc + (int64_t)l * m.rows * m.cols / e->cfg.group : nullptr;
}
static void* kv_layer(const sllm_engine* e, void* cache, int l) {
    return reinterpret_cast<uint8_t*>(cache) + (size_t)e->esz_kv * l * e->S * e->kv_loc;
}

__global__ void embed_state_kernel(const StepState* st, const void* table, int w_dtype, const float* scales, int group,
                                   float* out, int vocab, int d) {
    pdl_launch_dependents();
    pdl_wait();
    int tok = st->token;
    tok = min(max(tok, 0), vocab - 1);
    const int64_t base = (int64_t)tok * d;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
        float v;
        if (w_dtype == SLLM_F32) v = reinterpret_cast<const float*>(table)[base + i];
        else if (w_dtype == SLLM_BF16) v = __uint_as_float((uint32_t) reinterpret_cast<const uint16_t*>(table)[base + i] << 16);
        else v = (float)reinterpret_cast<const int8_t*>(table)[base + i] * scales[(base + i) / group];
        out[i] = v;
    }
}

// TP: merge the ranks' (best value, best index) pairs and advance the step state (every rank identically)
__global__ void tp_merge_kernel(const float* pairs, int tp, StepState* st, const int32_t* prompt, int32_t* history) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float v = -INFINITY;
    int idx = 0x7fffffff;
    for (int r = 0; r < tp; ++r) {
        const float ov = pairs[2 * r];
        const int oi = __float_as_int(pairs[2 * r + 1]);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    ClsPolicy<SLLM_F32>::step_feedback(st, prompt, history, idx == 0x7fffffff ? 0 : idx);
}
__global__ void tp_argmax_p2p_kernel(const float* blk_val, const int32_t* blk_idx, int slot, const P2PComm* c, int op, StepState* st,
                                     const int32_t* prompt, int32_t* history) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned e = p2p_epoch(*c, op);
    const float v = blk_val[slot];
    const int idx = blk_idx[slot];
    for (int dst = 0; dst < c->tp; ++dst) {
        uint2* s = p2p_slot(*c, dst, e, c->rank);
        p2p_send(s, v, e);
        p2p_send(s + 1, __int_as_float(idx), e);
    }
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int r = 0; r < c->tp; ++r) {
        const float2 pr = p2p_recv2(p2p_slot(*c, c->rank, e, r), e);
        const int oi = __float_as_int(pr.y);
        if (pr.x > best || (pr.x == best && oi < bi)) { best = pr.x; bi = oi; }
    }
    ClsPolicy<SLLM_F32>::step_feedback(st, prompt, history, bi == 0x7fffffff ? 0 : bi);
}
__global__ void tp_pack_kernel(const float* blk_val, const int32_t* blk_idx, int slot, float* pair) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        pair[0] = blk_val[slot];
        pair[1] = __int_as_float(blk_idx[slot]);
    }
}

// One fused kernel of the step. kind: 0 embed, 1 qkv(A), 2 mha(B), 3 wo(C), 4 gate_up(D), 5 down(E), 6 cls(F).
// Residual-stream bookkeeping under TP: after the all-reduce of a row-parallel GEMV the partial sum sits in
// part_*; the NEXT kernel's prologue adds it to the residual (and CTA 0 stores the sum).
enum { K_EMBED = 0, K_QKV = 1, K_MHA = 2, K_WO = 3, K_GATEUP = 4, K_DOWN = 5, K_CLS = 6 };

template <int WD>
static int enqueue_kernel(sllm_engine* e, int kind, int l) {
    const sllm_shape& s = e->cfg.shape;
    const int d = e->d, L = e->L;
    const bool tp = e->tp > 1;
    const int32_t* pos_dev = &e->state->pos;
    switch (kind) {
        case K_EMBED: {
            LaunchCfg lc(dim3(std::max(1, std::min((d + 255) / 256, 64))), dim3(256), 0, e->stream, e->pdl);
            SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, embed_state_kernel, (const StepState*)e->state, (const void*)e->emb.w, e->cfg.w_dtype,
                                         (const float*)e->emb.sc, e->cfg.group, e->x, e->V, d));
            g_launches++;
            count_launch(e);
            return SLLM_OK;
        }
        case K_QKV: {
            QkvPolicy<WD> A{};
            A.W_ = mat_layer(e, e->wqkv, l); A.sc_ = sc_layer(e, e->wqkv, l); A.grp_ = e->cfg.group; A.cols_ = d;
            if (tp && l > 0) { A.x = e->h; A.add = e->part_b; A.sum_out = e->x; } else { A.x = e->x; A.add = nullptr; A.sum_out = nullptr; }
            if (e->p2p_mode && !e->timing_only && l > 0) { A.add = nullptr; A.p2p = e->p2p_dev; A.p2p_op = 2 * (l - 1) + 1; }
            A.norm_w = e->norms + (int64_t)(2 * l) * d; A.eps = s.eps; A.pos_dev = pos_dev; A.sin_t = e->sin_t; A.cos_t = e->cos_t;
            A.q_out = e->q; A.k_cache = kv_layer(e, e->key_cache, l); A.v_cache = kv_layer(e, e->value_cache, l);
            A.kv_dtype = e->cfg.kv_dtype; A.q_dim = e->q_loc; A.kv_dim = e->kv_loc; A.hd = e->hd;
            return launch_policy<WD>(e, A, (e->q_loc + 2 * e->kv_loc) / 2);
        }
        case K_MHA: {
            if (int rc = mha_decode_dispatch(e->q, e->key_cache, e->value_cache, e->cfg.kv_dtype, e->att, e->mha_ws, l, pos_dev, 0, e->S,
                                             e->hd, e->H_loc, e->KVH_loc, e->stream, e->pdl)) return rc;
            count_launch(e);
            return SLLM_OK;
        }
        case K_WO: {
            ResidualPolicy<WD> C{};
            C.W_ = mat_layer(e, e->wo, l); C.sc_ = sc_layer(e, e->wo, l); C.grp_ = e->cfg.group; C.cols_ = e->q_loc;
            C.x = e->att; C.nrows = d;
            if (tp) { C.resid = nullptr; C.y = e->part_a; } else { C.resid = e->x; C.y = e->h; }
            if (e->p2p_mode && !e->timing_only) { C.p2p = e->p2p_dev; C.p2p_op = 2 * l; C.st = e->state; }
            return launch_policy<WD>(e, C, (d + 1) / 2);
        }
        case K_GATEUP: {
            GateUpPolicy<WD> D{};
            D.W_ = mat_layer(e, e->wug, l); D.sc_ = sc_layer(e, e->wug, l); D.grp_ = e->cfg.group; D.cols_ = d;
            // TP: the residual stream entering this layer is x (layer 0) or h+part_b, which kernel A stored in x
            if (tp) { D.h = e->x; D.add = e->part_a; D.sum_out = e->h; } else { D.h = e->h; D.add = nullptr; D.sum_out = nullptr; }
            if (e->p2p_mode && !e->timing_only) { D.add = nullptr; D.p2p = e->p2p_dev; D.p2p_op = 2 * l; }
            D.norm_w = e->norms + (int64_t)(2 * l + 1) * d; D.eps = s.eps; D.s_out = e->swi; D.inter = e->I_loc;
            return launch_policy<WD>(e, D, e->I_loc);
        }
        case K_DOWN: {
            ResidualPolicy<WD> E{};
            E.W_ = mat_layer(e, e->wdown, l); E.sc_ = sc_layer(e, e->wdown, l); E.grp_ = e->cfg.group; E.cols_ = e->I_loc;
            E.x = e->swi; E.nrows = d;
            if (tp) { E.resid = nullptr; E.y = e->part_b; } else { E.resid = e->h; E.y = e->x; }
            if (e->p2p_mode && !e->timing_only) { E.p2p = e->p2p_dev; E.p2p_op = 2 * l + 1; E.st = e->state; }
            return launch_policy<WD>(e, E, (d + 1) / 2);
        }
        case K_CLS: {
            ClsPolicy<WD> F{};
            F.W_ = reinterpret_cast<const uint8_t*>(e->emb.w) + wbytes(e->cfg.w_dtype, (int64_t)e->v0 * d);
            F.sc_ = e->emb.sc ? e->emb.sc + (int64_t)e->v0 * d / e->cfg.group : nullptr;
            F.grp_ = e->cfg.group; F.cols_ = d;
            if (tp) { F.x = e->h; F.add = e->part_b; F.sum_out = e->x; } else { F.x = e->x; F.add = nullptr; F.sum_out = nullptr; }
            if (e->p2p_mode && !e->timing_only) { F.add = nullptr; F.p2p = e->p2p_dev; F.p2p_op = 2 * (L - 1) + 1; }
            F.norm_w = e->norms + (int64_t)(2 * L) * d; F.eps = s.eps; F.logits = e->logits; F.nrows = e->V_loc; F.row0 = e->v0;
            F.blk_val = e->blk_val; F.blk_idx = e->blk_idx; F.st = e->state; F.prompt = e->prompt_dev; F.history = e->history_dev;
            F.single_rank = (tp || l < 0) ? 0 : 1;   // l < 0: timing-only launch, leave the step state alone
            return launch_policy<WD>(e, F, (e->V_loc + 1) / 2, e->cls_grid);
        }
        default: SLLM_REQUIRE(false, SLLM_EINVAL, "unknown kernel kind %d", kind);
    }
}

template <int WD>
static int enqueue_fused_step(sllm_engine* e) {
    const int d = e->d, L = e->L;
    const bool tp = e->tp > 1;
#define STEP(kind, layer) do { if (int rc = enqueue_kernel<WD>(e, kind, layer)) return rc; } while (0)
    STEP(K_EMBED, 0);
    for (int l = 0; l < L; ++l) {
        STEP(K_QKV, l);
        STEP(K_MHA, l);
        STEP(K_WO, l);
        if (tp && !e->p2p_mode) { SLLM_NCCL(ncclAllReduce(e->part_a, e->part_a, d, ncclFloat, ncclSum, e->comm, e->stream)); count_launch(e); }
        STEP(K_GATEUP, l);
        STEP(K_DOWN, l);
        if (tp && !e->p2p_mode) { SLLM_NCCL(ncclAllReduce(e->part_b, e->part_b, d, ncclFloat, ncclSum, e->comm, e->stream)); count_launch(e); }
    }
    STEP(K_CLS, 0);
#undef STEP
    if (tp && e->p2p_mode) {   // (value, index) exchange + first-max merge + step feedback in ONE tiny kernel over peer memory
        tp_argmax_p2p_kernel<<<1, 32, 0, e->stream>>>(e->blk_val, e->blk_idx, e->cls_grid, e->p2p_dev, 2 * L, e->state, e->prompt_dev, e->history_dev);
        count_launch(e);
        SLLM_LAUNCH_CHECK();
    } else if (tp) {
        tp_pack_kernel<<<1, 32, 0, e->stream>>>(e->blk_val, e->blk_idx, e->cls_grid, e->tp_pairs + 2 * e->rank);
        count_launch(e);
        SLLM_NCCL(ncclAllGather(e->tp_pairs + 2 * e->rank, e->tp_pairs, 2, ncclFloat, e->comm, e->stream));
        count_launch(e);
        tp_merge_kernel<<<1, 32, 0, e->stream>>>(e->tp_pairs, e->tp, e->state, e->prompt_dev, e->history_dev);
        count_launch(e);
        SLLM_LAUNCH_CHECK();
    }
    return SLLM_OK;
}

// ------------------------------------------------------------------------------ unfused (op-by-op) step --
// The reference's own 13-ops-per-layer sequence (model.cpp:48-139) through the public op launchers; position
// known on the host. Single rank only. Used for A/B against the fused path and as a second parity target.
__global__ void set_state_kernel(StepState* st, int token, int pos, int n_prompt) {
    st->token = token;
    st->pos = pos;
    st->n_prompt = n_prompt;
}
__global__ void feedback_kernel(StepState* st, const int32_t* idx, const int32_t* prompt, int32_t* history) {
    ClsPolicy<SLLM_F32>::step_feedback(st, prompt, history, *idx);
}

static int enqueue_unfused_step(sllm_engine* e, int pos) {
    SLLM_REQUIRE(e->tp == 1, SLLM_ENOTSUP, "the op-by-op path is single-GPU");
    const sllm_shape& s = e->cfg.shape;
    const int d = e->d, L = e->L, wd = e->cfg.w_dtype, grp = e->cfg.group;
    sllm_stream_t st = e->stream;
    const int64_t before = g_launches;
#define OP(call) do { if (int rc = (call)) return rc; } while (0)
    OP(sllm_embedding(&e->state->token, 0, e->emb.w, wd, e->emb.sc, grp, e->x, e->V, d, st));
    for (int l = 0; l < L; ++l) {
        const uint8_t* wqkv = reinterpret_cast<const uint8_t*>(mat_layer(e, e->wqkv, l));
        const float* sqkv = sc_layer(e, e->wqkv, l);
        auto row_ptr = [&](int64_t row) { return wqkv + wbytes(wd, row * d); };
        auto row_sc = [&](int64_t row) { return sqkv ? sqkv + row * d / grp : nullptr; };
        OP(sllm_rmsnorm_f32(e->x, e->norms + (int64_t)(2 * l) * d, e->xb, d, s.eps, st));
        float *krow, *vrow;
        uint8_t* kdst = reinterpret_cast<uint8_t*>(kv_layer(e, e->key_cache, l)) + (size_t)e->esz_kv * pos * e->kv_loc;
        uint8_t* vdst = reinterpret_cast<uint8_t*>(kv_layer(e, e->value_cache, l)) + (size_t)e->esz_kv * pos * e->kv_loc;
        if (e->cfg.kv_dtype == SLLM_F32) { krow = reinterpret_cast<float*>(kdst); vrow = reinterpret_cast<float*>(vdst); }
        else { krow = e->kv_tmp; vrow = e->kv_tmp + e->kv_loc; }
        OP(sllm_gemv(e->xb, row_ptr(0), wd, row_sc(0), grp, e->q, e->q_loc, d, 1.0f, st));
        OP(sllm_gemv(e->xb, row_ptr(e->q_loc), wd, row_sc(e->q_loc), grp, krow, e->kv_loc, d, 1.0f, st));
        OP(sllm_gemv(e->xb, row_ptr(e->q_loc + e->kv_loc), wd, row_sc(e->q_loc + e->kv_loc), grp, vrow, e->kv_loc, d, 1.0f, st));
        OP(sllm_rope_f32(e->q, krow, nullptr, pos, e->sin_t, e->cos_t, e->q_loc, e->kv_loc, e->hd, st));
        if (e->cfg.kv_dtype != SLLM_F32) {
            OP(sllm_store_kv_row(krow, kdst, e->cfg.kv_dtype, e->kv_loc, st));
            OP(sllm_store_kv_row(vrow, vdst, e->cfg.kv_dtype, e->kv_loc, st));
        }
        OP(sllm_mha_decode(e->q, e->key_cache, e->value_cache, e->cfg.kv_dtype, e->att, e->mha_ws, l, nullptr, pos, e->S, e->hd,
                           e->H_loc, e->KVH_loc, st));
        OP(sllm_gemv(e->att, mat_layer(e, e->wo, l), wd, sc_layer(e, e->wo, l), grp, e->att_out, d, e->q_loc, 1.0f, st));
        OP(sllm_add_f32(e->x, e->att_out, e->h, d, st));
        OP(sllm_rmsnorm_f32(e->h, e->norms + (int64_t)(2 * l + 1) * d, e->xb, d, s.eps, st));
        const uint8_t* wug = reinterpret_cast<const uint8_t*>(mat_layer(e, e->wug, l));
        const float* sug = sc_layer(e, e->wug, l);
        OP(sllm_gemv(e->xb, wug, wd, sug, grp, e->up, e->I_loc, d, 1.0f, st));
        OP(sllm_gemv(e->xb, wug + wbytes(wd, (int64_t)e->I_loc * d), wd, sug ? sug + (int64_t)e->I_loc * d / grp : nullptr, grp, e->gate,
                     e->I_loc, d, 1.0f, st));
        OP(sllm_swiglu_f32(e->up, e->gate, e->swi, e->I_loc, st));
        OP(sllm_gemv(e->swi, mat_layer(e, e->wdown, l), wd, sc_layer(e, e->wdown, l), grp, e->ffn_out, d, e->I_loc, 1.0f, st));
        OP(sllm_add_f32(e->ffn_out, e->h, e->x, d, st));
    }
    OP(sllm_rmsnorm_f32(e->x, e->norms + (int64_t)(2 * L) * d, e->xb, d, s.eps, st));
    OP(sllm_gemv(e->xb, e->emb.w, wd, e->emb.sc, grp, e->logits, e->V, d, 1.0f, st));
    OP(sllm_argmax_f32(e->logits, e->V, e->blk_idx, st));
#undef OP
    feedback_kernel<<<1, 1, 0, e->stream>>>(e->state, e->blk_idx, e->prompt_dev, e->history_dev);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    e->total_launches += g_launches - before;
    return SLLM_OK;
}

static int enqueue_fused_dispatch(sllm_engine* e) {
    switch (e->cfg.w_dtype) {
        case SLLM_F32: return enqueue_fused_step<SLLM_F32>(e);
        case SLLM_BF16: return enqueue_fused_step<SLLM_BF16>(e);
        default: return enqueue_fused_step<SLLM_INT8>(e);
    }
}

// capture one fused step into a graph (once)
static int build_graph(sllm_engine* e) {
    if (e->graph_exec) return SLLM_OK;
    // warm-up launch outside capture: sets the per-kernel shared-memory attributes and loads the modules
    const int64_t t0 = e->total_launches;
    if (int rc = enqueue_fused_dispatch(e)) return rc;
    e->step_launches = (int)(e->total_launches - t0);
    SLLM_CUDA(cudaStreamSynchronize(e->stream));
    cudaGraph_t g = nullptr;
    SLLM_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_fused_dispatch(e);
    cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
    e->total_launches -= e->step_launches;  // the captured pass launched nothing
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    SLLM_CUDA(ce);
    SLLM_CUDA(cudaGraphInstantiate(&e->graph_exec, g, 0));
    SLLM_CUDA(cudaGraphDestroy(g));
    return SLLM_OK;
}

// *done = steps actually enqueued (the host shadow of the position must follow the device, also when a launch fails half way)
static int enqueue_steps(sllm_engine* e, int n, int host_pos, int* done) {
    *done = 0;
    SLLM_REQUIRE(e->weights_loaded, SLLM_ESTATE, "weights not loaded");
    if (!e->fused) {
        for (int i = 0; i < n; ++i) {
            if (int rc = enqueue_unfused_step(e, host_pos + i)) return rc;
            ++*done;
        }
        return SLLM_OK;
    }
    if (e->mega) {
        SLLM_REQUIRE(!e->mega_ll || e->ll_ready, SLLM_ESTATE, "tensor-parallel megakernel: peer areas not exchanged yet (sllm_engine_p2p_import)");
        for (int i = 0; i < n; ++i) {
            const int rc = e->mega_ll ? mega_ll_launch(e->ll_params, e->H_loc / e->KVH_loc, e->ll_plan.grid, e->ll_plan.smem, e->stream)
                           : e->mega2 ? mega2_launch(e->mega2_params, e->H_loc / e->KVH_loc, e->mega_plan_.grid, e->mega_plan_.smem, e->stream, e->mega_fuse)
                                      : mega_launch(e->mega_params, e->H_loc / e->KVH_loc, e->mega_plan_.grid, e->mega_plan_.smem, e->stream, e->mega_fuse);
            if (rc) return rc;
            e->total_launches++;
            ++*done;
        }
        return SLLM_OK;
    }
    if (e->use_graph) {
        SLLM_REQUIRE(e->graph_exec, SLLM_ESTATE, "decode graph not built (weights or communicator missing)");
        for (int i = 0; i < n; ++i) {
            SLLM_CUDA(cudaGraphLaunch(e->graph_exec, e->stream));
            e->total_launches += e->step_launches;
            g_launches += e->step_launches;
            ++*done;
        }
        return SLLM_OK;
    }
    for (int i = 0; i < n; ++i) {
        const int64_t t0 = e->total_launches;
        if (int rc = enqueue_fused_dispatch(e)) return rc;   // a step that failed between its kernels leaves the device state undefined: the caller must set_state again
        e->step_launches = (int)(e->total_launches - t0);
        ++*done;
    }
    return SLLM_OK;
}

static int set_state(sllm_engine* e, int token, int pos, int n_prompt) {
    SLLM_REQUIRE(pos >= 0 && pos < e->S, SLLM_EINVAL, "position %d outside [0, %d)", pos, e->S);
    SLLM_REQUIRE(token >= 0 && token < e->V, SLLM_EINVAL, "Token index %d is outside the vocabulary [0, %d).", token, e->V);
    set_state_kernel<<<1, 1, 0, e->stream>>>(e->state, token, pos, n_prompt);
    g_launches++;
    e->total_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

static int set_state(sllm_engine* e, int token, int pos, int n_prompt);
// ------------------------------------------------------------------------------------------ weights ----
struct ShardSpec { int seg; int64_t first_row, n_rows, src_row_len, col0, row_len; void* dst; float* sc; };

static std::vector<ShardSpec> shard_plan(sllm_engine* e) {
    std::vector<ShardSpec> plan;
    const int64_t d = e->d, kv = e->cfg.shape.kv_hidden, I = e->I, L = e->L, r = e->rank;
    const int wd = e->cfg.w_dtype, grp = e->cfg.group;
    auto off_w = [&](const Matrix& m, int64_t elems) { return (void*)(reinterpret_cast<uint8_t*>(m.rm ? m.rm : m.w) + wbytes(wd, elems)); };
    auto off_s = [&](const Matrix& m, int64_t elems) { return m.sc ? m.sc + elems / grp : nullptr; };
    plan.push_back({0, 0, e->V, d, 0, d, off_w(e->emb, 0), e->emb.sc});
    plan.push_back({1, 0, 2 * L + 1, d, 0, d, e->norms, nullptr});
    for (int64_t l = 0; l < L; ++l) {
        const int64_t qkv_rows = e->q_loc + 2 * e->kv_loc;
        const int64_t base = l * qkv_rows * d;
        plan.push_back({2, l * d + r * e->q_loc, e->q_loc, d, 0, d, off_w(e->wqkv, base), off_s(e->wqkv, base)});
        plan.push_back({3, l * kv + r * e->kv_loc, e->kv_loc, d, 0, d, off_w(e->wqkv, base + (int64_t)e->q_loc * d), off_s(e->wqkv, base + (int64_t)e->q_loc * d)});
        plan.push_back({4, l * kv + r * e->kv_loc, e->kv_loc, d, 0, d, off_w(e->wqkv, base + (int64_t)(e->q_loc + e->kv_loc) * d), off_s(e->wqkv, base + (int64_t)(e->q_loc + e->kv_loc) * d)});
        plan.push_back({5, l * d, d, d, r * e->q_loc, e->q_loc, off_w(e->wo, l * d * e->q_loc), off_s(e->wo, l * d * e->q_loc)});
        const int64_t ug = l * 2 * e->I_loc * d;
        plan.push_back({6, l * I + r * e->I_loc, e->I_loc, d, 0, d, off_w(e->wug, ug), off_s(e->wug, ug)});
        plan.push_back({7, l * I + r * e->I_loc, e->I_loc, d, 0, d, off_w(e->wug, ug + (int64_t)e->I_loc * d), off_s(e->wug, ug + (int64_t)e->I_loc * d)});
        plan.push_back({8, l * d, d, I, r * e->I_loc, e->I_loc, off_w(e->wdown, l * d * e->I_loc), off_s(e->wdown, l * d * e->I_loc)});
    }
    return plan;
}

static int64_t segment_offset(const sllm_shape& s, int seg) {
    const int64_t V = s.vocab, d = s.hidden, kv = s.kv_hidden, I = s.inter, L = s.layers;
    const int64_t counts[9] = {V * d, (2 * L + 1) * d, L * d * d, L * kv * d, L * kv * d, L * d * d, L * I * d, L * I * d, L * d * I};
    int64_t off = 0;
    for (int k = 0; k < seg; ++k) off += counts[k];
    return off;
}

// The graph is built eagerly (when weights and, under TP, the communicator are both present), never lazily in
// the middle of a run: its warm-up pass executes one real step at (token 0, position 0).
static int maybe_build_graph(sllm_engine* e) {
    if (!e->fused || !e->use_graph || e->graph_exec || !e->weights_loaded) return SLLM_OK;
    if (e->tp > 1 && !(e->p2p_mode ? e->p2p_ready : e->comm != nullptr)) return SLLM_OK;
    if (int rc = set_state(e, 0, 0, 0)) return rc;
    if (int rc = build_graph(e)) return rc;
    if (int rc = set_state(e, 0, 0, 0)) return rc;
    SLLM_CUDA(cudaMemsetAsync(e->history_dev, 0, sizeof(int32_t) * (size_t)e->S, e->stream));
    SLLM_CUDA(cudaStreamSynchronize(e->stream));
    return SLLM_OK;
}

// megakernel mode: weights are generated / uploaded row-major into ONE temporary allocation, then repacked into the
// tiled arena layout matrix by matrix, and the temporary is freed.
static int mega_stage_begin(sllm_engine* e) {
    if (!e->mega) return SLLM_OK;
    Matrix* ms[5] = {&e->emb, &e->wqkv, &e->wo, &e->wug, &e->wdown};
    size_t total = 0;
    for (Matrix* m : ms) total += align_up(wbytes(e->cfg.w_dtype, m->layers * m->rows * m->cols), 256);
    uint8_t* tmp = nullptr;
    if (cudaMalloc(&tmp, total) != cudaSuccess) { cudaGetLastError(); set_error("megakernel staging: cudaMalloc(%zu MiB) failed", total >> 20); return SLLM_ENOMEM; }
    size_t off = 0;
    for (Matrix* m : ms) { m->rm = tmp + off; off += align_up(wbytes(e->cfg.w_dtype, m->layers * m->rows * m->cols), 256); }
    return SLLM_OK;
}
static int mega_stage_end(sllm_engine* e) {
    if (!e->mega || !e->emb.rm) return SLLM_OK;
    Matrix* ms[5] = {&e->emb, &e->wqkv, &e->wo, &e->wug, &e->wdown};
    int rc = SLLM_OK;
    for (Matrix* m : ms) {
        const size_t tiled = mega_matrix_bytes((int)m->rows, (int)m->cols, m->kind, e->cfg.w_dtype);
        for (int64_t l = 0; l < m->layers && rc == SLLM_OK; ++l)
            rc = mega_repack(reinterpret_cast<uint8_t*>(m->rm) + wbytes(e->cfg.w_dtype, l * m->rows * m->cols),
                             m->sc ? m->sc + l * m->rows * m->cols / e->cfg.group : nullptr,
                             reinterpret_cast<uint8_t*>(m->w) + (size_t)l * tiled, (int)m->rows, (int)m->cols, m->kind, e->cfg.w_dtype, e->hd,
                             e->q_loc, e->kv_loc, e->I_loc, e->stream);
    }
    if (e->mega_fuse) {   // the fused kernel's transposed copy of every layer's down matrix, from the same row-major staging copy
        const size_t per_layer = mega_down_t_bytes(e->d, e->I_loc, e->cfg.w_dtype);
        for (int64_t l = 0; l < e->wdown.layers && rc == SLLM_OK; ++l)
            rc = mega_repack_down_t(reinterpret_cast<uint8_t*>(e->wdown.rm) + wbytes(e->cfg.w_dtype, l * e->wdown.rows * e->wdown.cols),
                                    reinterpret_cast<uint8_t*>(e->wdown_t) + (size_t)l * per_layer, e->d, e->I_loc, e->cfg.w_dtype, e->stream);
    }
    if (e->mega2) {   // the column-block copy of every layer's Wo, from the same row-major staging copy
        const size_t per_layer = mega2_wot_bytes(e->d, e->hd, e->H_loc, e->KVH_loc, e->cfg.w_dtype);
        for (int64_t l = 0; l < e->wo.layers && rc == SLLM_OK; ++l)
            rc = mega2_repack_wot(reinterpret_cast<uint8_t*>(e->wo.rm) + wbytes(e->cfg.w_dtype, l * e->wo.rows * e->wo.cols),
                                  reinterpret_cast<uint8_t*>(e->wo_t) + (size_t)l * per_layer, e->d, e->hd, e->H_loc, e->KVH_loc, e->cfg.w_dtype, e->stream);
    }
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    cudaFree(e->emb.rm);
    for (Matrix* m : ms) m->rm = nullptr;
    if (rc) return rc;
    SLLM_CUDA(ce);
    return SLLM_OK;
}

// a load that failed half way: give the row-major staging copy back (the engine stays without weights)
static void mega_stage_abort(sllm_engine* e) {
    if (!e->emb.rm) return;
    cudaStreamSynchronize(e->stream);
    cudaFree(e->emb.rm);
    Matrix* ms[5] = {&e->emb, &e->wqkv, &e->wo, &e->wug, &e->wdown};
    for (Matrix* m : ms) m->rm = nullptr;
}

static int setup_mega(sllm_engine* e) {
    std::vector<PhaseDesc> host((size_t)4 * e->L + 1);
    // the classifier streams this rank's vocab rows [v0, v0+V_loc) of the (tiled, full) embedding matrix
    const TileGeom eg = mega_tile_geom(((e->V + 1) / 2) * 2, e->d, e->cfg.w_dtype);
    const uint8_t* cls = reinterpret_cast<const uint8_t*>(e->emb.w) + (size_t)(e->v0 / eg.R) * eg.KS * eg.tile_bytes;
    mega_fill_phases(host.data(), e->L, e->cfg.w_dtype, e->wqkv.w, e->wo.w, e->wug.w, e->wdown.w, cls, e->d, e->q_loc, e->kv_loc,
                     e->I_loc, e->V_loc);
    if (e->mega_fuse)
        for (int l = 0; l < e->L; ++l)
            mega_fill_down_t(host[(size_t)4 * l + 3], reinterpret_cast<const uint8_t*>(e->wdown_t) + (size_t)l * mega_down_t_bytes(e->d, e->I_loc, e->cfg.w_dtype),
                             e->d, e->I_loc, l, e->cfg.w_dtype);
    if (e->mega2)
        for (int l = 0; l < e->L; ++l)
            mega2_fill_wot(host[(size_t)4 * l + 1], reinterpret_cast<const uint8_t*>(e->wo_t) + (size_t)l * mega2_wot_bytes(e->d, e->hd, e->H_loc, e->KVH_loc, e->cfg.w_dtype),
                           e->d, e->hd, e->H_loc, e->KVH_loc, l, e->cfg.w_dtype);
    if (e->mega_ll) {
        MegaLLParams& q = e->ll_params;
        q.phases = e->phases_dev;
        q.d = e->d; q.hd = e->hd; q.L = e->L; q.S = e->S; q.V = e->V; q.V_loc = e->V_loc; q.v0 = e->v0; q.q_loc = e->q_loc; q.kv_loc = e->kv_loc;
        q.I_loc = e->I_loc; q.H_loc = e->H_loc; q.KVH_loc = e->KVH_loc; q.nsplit = e->ll_plan.nsplit;
        q.w_dtype = e->cfg.w_dtype; q.kv_dtype = e->cfg.kv_dtype; q.eps = e->cfg.shape.eps;
        q.emb = reinterpret_cast<const uint8_t*>(e->emb.w); q.norms = e->norms;
        q.kc = reinterpret_cast<uint8_t*>(e->key_cache); q.vc = reinterpret_cast<uint8_t*>(e->value_cache);
        q.sin_t = e->sin_t; q.cos_t = e->cos_t; q.logits = e->logits; q.x_out = e->x; q.blk_val = e->blk_val; q.blk_idx = e->blk_idx; q.st = e->state;
        q.prompt = e->prompt_dev; q.history = e->history_dev; q.tp = e->tp; q.rank = e->rank;
        q.area[e->rank] = reinterpret_cast<uint2*>(e->ll_block);
        q.off_wop = e->ll_plan.off_wop; q.off_dnp = e->ll_plan.off_dnp; q.off_qv = e->ll_plan.off_qv; q.off_kvn = e->ll_plan.off_kvn;
        q.off_att = e->ll_plan.off_att; q.off_swi = e->ll_plan.off_swi; q.off_arg = e->ll_plan.off_arg;
    }
    SLLM_CUDA(cudaMemcpyAsync(e->phases_dev, host.data(), sizeof(PhaseDesc) * host.size(), cudaMemcpyHostToDevice, e->stream));
    SLLM_CUDA(cudaStreamSynchronize(e->stream));
    e->phases_host = host;
    MegaParams& p = e->mega_params;
    p.phases = e->phases_dev;
    p.d = e->d; p.hd = e->hd; p.L = e->L; p.S = e->S; p.V = e->V; p.V_loc = e->V_loc; p.v0 = e->v0; p.q_loc = e->q_loc; p.kv_loc = e->kv_loc;
    p.I_loc = e->I_loc; p.H_loc = e->H_loc; p.KVH_loc = e->KVH_loc; p.nsplit = e->mega_plan_.nsplit;
    p.w_dtype = e->cfg.w_dtype; p.kv_dtype = e->cfg.kv_dtype; p.eps = e->cfg.shape.eps;
    p.emb = reinterpret_cast<const uint8_t*>(e->emb.w); p.norms = e->norms;
    p.kc = reinterpret_cast<uint8_t*>(e->key_cache); p.vc = reinterpret_cast<uint8_t*>(e->value_cache);
    p.sin_t = e->sin_t; p.cos_t = e->cos_t; p.x = e->x; p.h = e->h; p.q = e->q; p.swi = e->swi; p.logits = e->logits;
    p.att_part = e->mega_att; p.blk_val = e->blk_val; p.blk_idx = e->blk_idx; p.st = e->state; p.prompt = e->prompt_dev;
    p.history = e->history_dev; p.bar_counter = e->bar_counter; p.trace = nullptr;
    if (e->mega2) {
        Mega2Params& P = e->mega2_params;
        P.m = p;
        P.xbuf[0] = e->v2_bufs; P.xbuf[1] = e->v2_bufs + e->d; P.hbuf[0] = e->v2_bufs + 2 * (size_t)e->d; P.hbuf[1] = e->v2_bufs + 3 * (size_t)e->d;
        P.x_copy = e->x; P.h_copy = e->h; P.flags = e->v2_flags;
    }
    return SLLM_OK;
}

static int finish_weights(sllm_engine* e) {
    if (int rc = sllm_rope_tables(e->hd, e->S, e->cfg.shape.theta, e->sin_t, e->cos_t, e->stream)) return rc;
    SLLM_CUDA(cudaStreamSynchronize(e->stream));
    e->weights_loaded = true;
    if (e->mega) {
        if (int rc = mega_stage_end(e)) return rc;
        if (int rc = setup_mega(e)) return rc;
        e->step_launches = 1;
        return SLLM_OK;
    }
    return maybe_build_graph(e);
}

// -------------------------------------------------------------------------------------------- C ABI ----
namespace sllm {
int engine_view(const sllm_engine* e, EngineView* v) {
    SLLM_REQUIRE(e && v, SLLM_EINVAL, "engine_view: null argument");
    v->shape = e->cfg.shape;
    v->w_dtype = e->cfg.w_dtype;
    v->group = e->cfg.group;
    v->tp = e->tp;
    v->mega = e->mega ? 1 : 0;
    v->weights_loaded = e->weights_loaded ? 1 : 0;
    v->stream = e->stream;
    v->emb = e->emb.w; v->emb_sc = e->emb.sc;
    v->norms = e->norms;
    v->wqkv = e->wqkv.w; v->wqkv_sc = e->wqkv.sc;
    v->wo = e->wo.w; v->wo_sc = e->wo.sc;
    v->wug = e->wug.w; v->wug_sc = e->wug.sc;
    v->wdown = e->wdown.w; v->wdown_sc = e->wdown.sc;
    v->sin_t = e->sin_t; v->cos_t = e->cos_t;
    return SLLM_OK;
}
}  // namespace sllm

extern "C" {

int sllm_engine_create(const sllm_engine_config* cfg, sllm_stream_t stream, sllm_engine** out) {
    SLLM_REQUIRE(cfg && out, SLLM_EINVAL, "engine_create: null argument");
    const sllm_shape& s = cfg->shape;
    SLLM_REQUIRE(s.vocab > 0 && s.head_dim > 0 && s.hidden > 0 && s.kv_hidden > 0 && s.inter > 0 && s.max_len > 0 && s.layers > 0 &&
                 s.heads > 0 && s.kv_heads > 0, SLLM_EINVAL, "engine_create: non-positive dimension");
    SLLM_REQUIRE(s.heads * s.head_dim == s.hidden, SLLM_ENOTSUP, "heads*head_dim (%d) != hidden (%d): the reference's wq/wo are d x d (model.cpp:372-378)", s.heads * s.head_dim, s.hidden);
    SLLM_REQUIRE(s.kv_heads * s.head_dim == s.kv_hidden && s.heads % s.kv_heads == 0, SLLM_EINVAL, "inconsistent kv dims");
    const int tp = cfg->tp_size < 1 ? 1 : cfg->tp_size;
    SLLM_REQUIRE(cfg->tp_rank >= 0 && cfg->tp_rank < tp, SLLM_EINVAL, "tp_rank %d outside [0,%d)", cfg->tp_rank, tp);
    SLLM_REQUIRE(s.heads % tp == 0 && s.kv_heads % tp == 0 && s.inter % tp == 0 && s.vocab % tp == 0, SLLM_ENOTSUP,
                 "tensor parallel size %d must divide heads, kv_heads, intermediate and vocab", tp);
    SLLM_REQUIRE(cfg->w_dtype >= SLLM_F32 && cfg->w_dtype <= SLLM_INT8, SLLM_EINVAL, "bad weight dtype");
    SLLM_REQUIRE(cfg->kv_dtype == SLLM_F32 || cfg->kv_dtype == SLLM_BF16, SLLM_EINVAL, "bad kv dtype");
    int ndev = 0;
    SLLM_CUDA(cudaGetDeviceCount(&ndev));
    SLLM_REQUIRE(ndev > 0, SLLM_ESTATE, "no CUDA device: libsllm_b200 has no CPU fallback");

    auto* e = new sllm_engine();
    e->cfg = *cfg;
    e->cfg.tp_size = tp;
    if (e->cfg.w_dtype != SLLM_INT8) e->cfg.group = e->cfg.group > 0 ? e->cfg.group : 64;
    e->stream = as_stream(stream);
    if (!e->stream) {  // the legacy default stream cannot be captured into a graph: own a stream instead
        if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete e; set_error("cudaStreamCreate failed"); return SLLM_ENOMEM; }
        e->own_stream = true;
    }
    e->fused = !(cfg->flags & SLLM_ENGINE_UNFUSED);
    e->use_graph = e->fused && !(cfg->flags & SLLM_ENGINE_NO_GRAPH);
    e->pdl = e->fused && (cfg->flags & SLLM_ENGINE_PDL);
    e->p2p_mode = e->fused && tp > 1 && (cfg->flags & SLLM_ENGINE_P2P_ALLREDUCE);
    e->d = s.hidden; e->hd = s.head_dim; e->L = s.layers; e->S = s.max_len; e->V = s.vocab; e->H = s.heads; e->KVH = s.kv_heads; e->I = s.inter;
    e->tp = tp; e->rank = cfg->tp_rank;
    e->H_loc = s.heads / tp; e->KVH_loc = s.kv_heads / tp; e->q_loc = e->H_loc * s.head_dim; e->kv_loc = e->KVH_loc * s.head_dim;
    e->I_loc = s.inter / tp; e->V_loc = s.vocab / tp; e->v0 = e->rank * e->V_loc;
    e->esz_w = (int)wbytes(cfg->w_dtype, 1); e->esz_kv = cfg->kv_dtype == SLLM_F32 ? 4 : 2;
    auto fail = [&](int rc) { delete e; return rc; };
    const int E = cfg->w_dtype == SLLM_F32 ? 4 : cfg->w_dtype == SLLM_BF16 ? 8 : 16;
    if (e->d % E || e->q_loc % E || e->I_loc % E) { set_error("hidden/local dims must be multiples of %d for this weight type", E); return fail(SLLM_ENOTSUP); }
    if (cfg->w_dtype == SLLM_INT8 && (cfg->group < 16 || cfg->group % 16 || e->d % cfg->group || e->q_loc % cfg->group || e->I_loc % cfg->group)) {
        set_error("int8: group=%d must be a multiple of 16 dividing hidden=%d, local q dim=%d and local intermediate=%d", cfg->group, e->d, e->q_loc, e->I_loc);
        return fail(SLLM_ENOTSUP);
    }
    if (e->hd % 16 || e->hd > 256) { set_error("head_dim=%d must be a multiple of 16, <= 256", e->hd); return fail(SLLM_ENOTSUP); }
    if (gemv_smem_bytes(std::max(e->d, e->I_loc)) > (size_t)smem_optin_bytes()) { set_error("activation vector does not fit shared memory"); return fail(SLLM_ENOTSUP); }

    if ((cfg->flags & SLLM_ENGINE_MEGAKERNEL) && e->fused) {
        if (tp == 1 ? (cfg->flags & SLLM_ENGINE_MEGA_LL) != 0 : (cfg->flags & SLLM_ENGINE_P2P_ALLREDUCE) != 0) {
            e->ll_plan = mega_ll_plan(cfg->w_dtype, cfg->kv_dtype, e->d, e->hd, e->q_loc, e->kv_loc, e->I_loc, e->V_loc, e->v0, e->H_loc, e->KVH_loc, e->S, tp);
            if (cfg->w_dtype == SLLM_INT8 && cfg->group != 64) { e->ll_plan.ok = false; e->ll_plan.why = "int8 group size other than 64"; }   // the tile format carries one scale per 4 chunks
            e->mega_ll = e->mega = e->ll_plan.ok;
        }
        if (!e->mega && tp == 1) {
            e->mega_plan_ = mega_plan(cfg->w_dtype, e->cfg.group, cfg->kv_dtype, e->d, e->hd, e->q_loc, e->kv_loc, e->I_loc, e->V_loc, e->H_loc, e->KVH_loc, e->S);
            e->mega = e->mega_plan_.ok;
        }
        if (e->mega) e->p2p_mode = false;   // the megakernel carries its own in-kernel all-reduce
        const int g_q = e->H_loc / e->KVH_loc;
        e->mega_fuse = e->mega && !e->mega_ll && (cfg->flags & SLLM_ENGINE_MEGA_FUSE_DOWN) && mega_fuse_down_ok(cfg->w_dtype, e->d, e->I_loc) &&
                       (g_q == 1 || g_q == 2 || g_q == 4 || g_q == 8);
        e->mega2 = e->mega && !e->mega_ll && (cfg->flags & SLLM_ENGINE_MEGA_V2) && (g_q == 1 || g_q == 2 || g_q == 4 || g_q == 8) &&
                   mega2_ok(cfg->w_dtype, e->d, e->hd, e->H_loc, e->KVH_loc, e->mega_plan_.nsplit, e->mega_plan_.grid, nullptr) &&
                   e->mega_plan_.smem + 10 * 1024 <= (size_t)smem_optin_bytes();   // its static shared memory: phase table + publication list
    }
    layout(e);  // measure
    e->arena_bytes = align_up(e->arena_used, 1 << 20);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (e->arena_bytes > free_b) { set_error("engine needs %zu MiB of HBM, %zu MiB free", e->arena_bytes >> 20, free_b >> 20); return fail(SLLM_ENOMEM); }
    if (cudaMalloc(&e->arena, e->arena_bytes) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu MiB) failed", e->arena_bytes >> 20); return fail(SLLM_ENOMEM); }
    layout(e);  // assign
    if (e->mega_ll) {
        const size_t bytes = (size_t)e->ll_plan.area_words * sizeof(uint2);
        if (cudaMalloc(&e->ll_block, bytes) != cudaSuccess || cudaMemset(e->ll_block, 0, bytes) != cudaSuccess) {
            cudaGetLastError(); set_error("megakernel word area: allocation failed"); sllm_engine_destroy(e); return SLLM_ENOMEM;
        }
        e->ll_ready = (tp == 1);
    }
    if (e->p2p_mode) {
        if (tp > kMaxTp) { set_error("peer-memory all-reduce supports up to %d ranks", kMaxTp); sllm_engine_destroy(e); return SLLM_ENOTSUP; }
        e->p2p_recv_bytes = align_up((size_t)2 * tp * e->d * sizeof(uint2), 256);   // {value, epoch} words
        const size_t total = e->p2p_recv_bytes + 256 + sizeof(P2PComm) + 256;
        if (cudaMalloc(&e->p2p_block, total) != cudaSuccess || cudaMemset(e->p2p_block, 0, total) != cudaSuccess) {
            cudaGetLastError(); set_error("peer-memory all-reduce: allocation failed"); sllm_engine_destroy(e); return SLLM_ENOMEM;
        }
    }
    if ((tp == 2 || tp == 4 || tp == 8) && (cfg->flags & SLLM_ENGINE_P2P_ALLREDUCE) && e->mega && !pf_unsupported_reason(cfg->w_dtype, e->hd, e->d, e->q_loc, e->I_loc) && e->d % 8 == 0 && e->d <= 8192) {
        const size_t total = pfx_block_bytes(kPrefillBlock, e->d);
        if (cudaMalloc(&e->pfx_block, total) != cudaSuccess || cudaMemset(e->pfx_block, 0, total) != cudaSuccess) {
            cudaGetLastError(); set_error("prefill exchange block: allocation of %zu MiB failed", total >> 20); sllm_engine_destroy(e); return SLLM_ENOMEM;
        }
    }
    // zero everything that is read before it is written: KV cache, workspaces, state
    cudaError_t ce = cudaMemsetAsync(e->key_cache, 0, e->arena + e->arena_used - reinterpret_cast<uint8_t*>(e->key_cache), e->stream);
    if (ce == cudaSuccess) ce = cudaMallocHost(reinterpret_cast<void**>(&e->h_state), 64);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) { cuda_fail(ce, "engine init", __FILE__, __LINE__); sllm_engine_destroy(e); return (int)ce; }
    *out = e;
    return SLLM_OK;
}

void sllm_engine_destroy(sllm_engine* e) {
    if (!e) return;
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->comm) ncclCommDestroy(e->comm);
    if (e->h_state) cudaFreeHost(e->h_state);
    for (int r = 0; r < kMaxTp; ++r) if (e->p2p_peer[r]) cudaIpcCloseMemHandle(e->p2p_peer[r]);
    if (e->p2p_dev) cudaFree(e->p2p_dev);
    if (e->p2p_block) cudaFree(e->p2p_block);
    if (e->ll_block) cudaFree(e->ll_block);
    for (int r = 0; r < kMaxTp; ++r) if (e->pfx_peer[r]) cudaIpcCloseMemHandle(e->pfx_peer[r]);
    if (e->pfx_block) cudaFree(e->pfx_block);
    if (e->pf_ws) cudaFree(e->pf_ws);
    if (e->pf) pf_cache_destroy(e->pf);
    if (e->emb.rm) cudaFree(e->emb.rm);   // a weight load that failed half way
    if (e->arena) cudaFree(e->arena);
    if (e->trace) cudaFree(e->trace);
    if (e->own_stream) cudaStreamDestroy(e->stream);
    delete e;
}

int sllm_engine_load_synthetic(sllm_engine* e, uint64_t seed) {
    SLLM_REQUIRE(e, SLLM_EINVAL, "null engine");
    if (int rc = mega_stage_begin(e)) return rc;
    for (const ShardSpec& p : shard_plan(e))
        if (int rc = sllm_synth_fill(&e->cfg.shape, seed, p.seg, p.first_row, p.n_rows, p.src_row_len, p.col0, p.row_len, p.dst,
                                     e->cfg.w_dtype, p.sc, e->cfg.group, e->stream)) { mega_stage_abort(e); return rc; }
    return finish_weights(e);
}

int sllm_engine_load_blob_f32(sllm_engine* e, const float* blob, int64_t n_floats) {
    SLLM_REQUIRE(e && blob, SLLM_EINVAL, "null argument");
    const sllm_shape& s = e->cfg.shape;
    const int64_t need = segment_offset(s, 8) + (int64_t)s.layers * s.hidden * s.inter;
    SLLM_REQUIRE(n_floats >= need, SLLM_EINVAL, "weight blob has %lld floats, the shape needs %lld (model.cpp:340-468 layout)", (long long)n_floats, (long long)need);
    if (int rc0 = mega_stage_begin(e)) return rc0;
    const size_t stage_bytes = (size_t)64 << 20;
    float* stage = nullptr;
    if (cudaError_t ce = cudaMalloc(&stage, stage_bytes)) { mega_stage_abort(e); return cuda_fail(ce, "cudaMalloc(upload staging)", __FILE__, __LINE__); }
    int rc = SLLM_OK;
    for (const ShardSpec& p : shard_plan(e)) {
        const float* src = blob + segment_offset(s, p.seg);
        const int64_t rows_per_chunk = std::max<int64_t>(1, (int64_t)(stage_bytes / 4) / p.row_len);
        const int wd = (p.seg == 1) ? SLLM_F32 : e->cfg.w_dtype;
        for (int64_t r0 = 0; r0 < p.n_rows && rc == SLLM_OK; r0 += rows_per_chunk) {
            const int64_t nr = std::min(rows_per_chunk, p.n_rows - r0);
            cudaError_t ce = cudaMemcpy2DAsync(stage, (size_t)p.row_len * 4, src + (p.first_row + r0) * p.src_row_len + p.col0,
                                               (size_t)p.src_row_len * 4, (size_t)p.row_len * 4, (size_t)nr, cudaMemcpyHostToDevice, e->stream);
            if (ce != cudaSuccess) { rc = cuda_fail(ce, "cudaMemcpy2DAsync(weights)", __FILE__, __LINE__); break; }
            void* dst = reinterpret_cast<uint8_t*>(p.dst) + wbytes(wd, r0 * p.row_len);
            float* sc = p.sc ? p.sc + r0 * p.row_len / e->cfg.group : nullptr;
            rc = sllm_convert_weights(stage, dst, wd, sc, e->cfg.group, nr, p.row_len, e->stream);
            if (rc == SLLM_OK) { ce = cudaStreamSynchronize(e->stream); if (ce != cudaSuccess) rc = cuda_fail(ce, "sync", __FILE__, __LINE__); }
        }
        if (rc) break;
    }
    cudaFree(stage);
    if (rc) { mega_stage_abort(e); return rc; }
    return finish_weights(e);
}

int sllm_comm_unique_id(void* id_bytes_128) {
    SLLM_REQUIRE(id_bytes_128, SLLM_EINVAL, "null id buffer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    SLLM_NCCL(ncclGetUniqueId(&id));
    std::memcpy(id_bytes_128, &id, 128);
    return SLLM_OK;
}

int sllm_engine_init_comm(sllm_engine* e, const void* id_bytes_128) {
    SLLM_REQUIRE(e && id_bytes_128, SLLM_EINVAL, "null argument");
    SLLM_REQUIRE(e->tp > 1, SLLM_ESTATE, "init_comm on a single-rank engine");
    ncclUniqueId id;
    std::memcpy(&id, id_bytes_128, 128);
    SLLM_NCCL(ncclCommInitRank(&e->comm, e->tp, id, e->rank));
    return maybe_build_graph(e);
}

int sllm_engine_p2p_export(sllm_engine* e, void* handle_bytes_64) {
    SLLM_REQUIRE(e && handle_bytes_64, SLLM_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t h;
    if (e->mega_ll && e->tp > 1) {
        SLLM_CUDA(cudaIpcGetMemHandle(&h, e->ll_block));
        std::memcpy(handle_bytes_64, &h, 64);
        return SLLM_OK;
    }
    SLLM_REQUIRE(e->p2p_mode && e->p2p_block, SLLM_ESTATE, "engine was not created with SLLM_ENGINE_P2P_ALLREDUCE (tp_size > 1)");
    SLLM_CUDA(cudaIpcGetMemHandle(&h, e->p2p_block));
    std::memcpy(handle_bytes_64, &h, 64);
    return SLLM_OK;
}

int sllm_engine_p2p_import(sllm_engine* e, const void* all_handles) {
    SLLM_REQUIRE(e && all_handles, SLLM_EINVAL, "null argument");
    if (e->mega_ll && e->tp > 1) {
        for (int r = 0; r < e->tp; ++r) {
            if (r == e->rank) continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, reinterpret_cast<const uint8_t*>(all_handles) + (size_t)r * 64, 64);
            void* pp = nullptr;
            SLLM_CUDA(cudaIpcOpenMemHandle(&pp, h, cudaIpcMemLazyEnablePeerAccess));
            e->p2p_peer[r] = pp;
            e->ll_params.area[r] = reinterpret_cast<uint2*>(pp);
        }
        e->ll_ready = true;
        e->p2p_ready = true;
        return SLLM_OK;
    }
    SLLM_REQUIRE(e->p2p_mode && e->p2p_block, SLLM_ESTATE, "engine was not created with SLLM_ENGINE_P2P_ALLREDUCE (tp_size > 1)");
    P2PComm c{};
    c.tp = e->tp; c.rank = e->rank; c.n = e->d; c.ops_per_step = 2 * e->L + 1; c.step = &e->state->pad[1];
    for (int r = 0; r < e->tp; ++r) {
        uint8_t* base = e->p2p_block;
        if (r != e->rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, reinterpret_cast<const uint8_t*>(all_handles) + (size_t)r * 64, 64);
            void* p = nullptr;
            SLLM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            e->p2p_peer[r] = p;
            base = reinterpret_cast<uint8_t*>(p);
        }
        c.recv[r] = reinterpret_cast<uint2*>(base);
    }
    SLLM_CUDA(cudaMalloc(&e->p2p_dev, sizeof(P2PComm)));
    SLLM_CUDA(cudaMemcpyAsync(e->p2p_dev, &c, sizeof(P2PComm), cudaMemcpyHostToDevice, e->stream));
    SLLM_CUDA(cudaStreamSynchronize(e->stream));
    e->p2p_ready = true;
    return maybe_build_graph(e);
}

/* The prefill exchange block (prefill_tp.cu): same messenger protocol as sllm_engine_p2p_export / _import. SLLM_ENOTSUP when the engine has
   none (single rank, no SLLM_ENGINE_P2P_ALLREDUCE, or a shape the batched prefill does not take): prefill then all-reduces with NCCL. */
int sllm_engine_prefill_p2p_export(sllm_engine* e, void* handle_bytes_64) {
    SLLM_REQUIRE(e && handle_bytes_64, SLLM_EINVAL, "null argument");
    if (!e->pfx_block) { set_error("this engine has no prefill exchange block"); return SLLM_ENOTSUP; }
    cudaIpcMemHandle_t h;
    SLLM_CUDA(cudaIpcGetMemHandle(&h, e->pfx_block));
    std::memcpy(handle_bytes_64, &h, 64);
    return SLLM_OK;
}

int sllm_engine_prefill_p2p_import(sllm_engine* e, const void* all_handles) {
    SLLM_REQUIRE(e && all_handles, SLLM_EINVAL, "null argument");
    if (!e->pfx_block) { set_error("this engine has no prefill exchange block"); return SLLM_ENOTSUP; }
    for (int r = 0; r < e->tp; ++r) {
        if (r == e->rank || e->pfx_peer[r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, reinterpret_cast<const uint8_t*>(all_handles) + (size_t)r * 64, 64);
        void* pp = nullptr;
        SLLM_CUDA(cudaIpcOpenMemHandle(&pp, h, cudaIpcMemLazyEnablePeerAccess));
        e->pfx_peer[r] = pp;
    }
    e->pfx_ready = true;
    return SLLM_OK;
}

int sllm_engine_set_state(sllm_engine* e, int32_t token, int32_t pos) {
    SLLM_REQUIRE(e, SLLM_EINVAL, "null engine");
    e->h_state[7] = pos;  // host shadow of the position for the op-by-op path
    return set_state(e, token, pos, 0);
}

int sllm_engine_enqueue_steps(sllm_engine* e, int32_t n_steps) {
    SLLM_REQUIRE(e && n_steps >= 0, SLLM_EINVAL, "bad argument");
    SLLM_REQUIRE(e->tp == 1 || e->comm || e->p2p_ready, SLLM_ESTATE, "tensor-parallel engine without a communicator");
    SLLM_REQUIRE(e->h_state[7] + n_steps <= e->S, SLLM_EINVAL, "%d steps from position %d overrun max_len %d", n_steps, e->h_state[7], e->S);
    int done = 0;
    const int rc = enqueue_steps(e, n_steps, e->h_state[7], &done);
    e->h_state[7] += done;   // the host shadow follows what was really enqueued, also when a launch failed half way
    return rc;
}

int sllm_engine_read_tokens(sllm_engine* e, int32_t* out, int32_t n) {
    SLLM_REQUIRE(e && out && n >= 0 && n <= e->S, SLLM_EINVAL, "bad argument");
    const int first = e->h_state[7] - n;
    SLLM_REQUIRE(first >= 0, SLLM_EINVAL, "asked for %d tokens, only %d steps taken", n, e->h_state[7]);
    SLLM_CUDA(cudaMemcpyAsync(out, e->history_dev + first, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    SLLM_CUDA(cudaStreamSynchronize(e->stream));
    return SLLM_OK;
}

int sllm_engine_forward(sllm_engine* e, int32_t token, int32_t pos, float* logits_host, int32_t* next_token_host) {
    SLLM_REQUIRE(e, SLLM_EINVAL, "null engine");
    SLLM_REQUIRE(e->tp == 1 || e->comm || e->p2p_ready, SLLM_ESTATE, "tensor-parallel engine without a communicator");
    // host -> device: token and position travel through pinned memory like any other input
    e->h_state[0] = token; e->h_state[1] = pos; e->h_state[2] = 0; e->h_state[3] = 0;
    SLLM_REQUIRE(pos >= 0 && pos < e->S, SLLM_EINVAL, "position %d outside [0, %d)", pos, e->S);
    SLLM_REQUIRE(token >= 0 && token < e->V, SLLM_EINVAL, "Token index %d is outside the vocabulary [0, %d).", token, e->V);
    SLLM_CUDA(cudaMemcpyAsync(e->state, e->h_state, 16, cudaMemcpyHostToDevice, e->stream));
    e->h_state[7] = pos;
    int done = 0;
    if (int rc = enqueue_steps(e, 1, pos, &done)) return rc;
    e->h_state[7] = pos + 1;
    if (logits_host) SLLM_CUDA(cudaMemcpyAsync(logits_host, e->logits, sizeof(float) * (size_t)e->V_loc, cudaMemcpyDeviceToHost, e->stream));
    if (next_token_host) SLLM_CUDA(cudaMemcpyAsync(e->h_state + 4, &e->state->next, 4, cudaMemcpyDeviceToHost, e->stream));
    SLLM_CUDA(cudaStreamSynchronize(e->stream));
    if (next_token_host) *next_token_host = e->h_state[4];
    return SLLM_OK;
}

int sllm_engine_greedy(sllm_engine* e, const int32_t* prompt, int32_t n_prompt, int32_t n_total, int32_t* tokens_out) {
    SLLM_REQUIRE(e && prompt && tokens_out, SLLM_EINVAL, "null argument");
    SLLM_REQUIRE(n_prompt >= 1 && n_total >= n_prompt && n_total <= e->S, SLLM_EINVAL, "need 1 <= n_prompt <= n_total <= max_len (%d, %d, %d)", n_prompt, n_total, e->S);
    for (int i = 0; i < n_prompt; ++i) SLLM_REQUIRE(prompt[i] >= 0 && prompt[i] < e->V, SLLM_EINVAL, "Token index %d is outside the vocabulary [0, %d).", prompt[i], e->V);
    SLLM_CUDA(cudaMemcpyAsync(e->prompt_dev, prompt, sizeof(int32_t) * (size_t)n_prompt, cudaMemcpyHostToDevice, e->stream));
    if (int rc = set_state(e, prompt[0], 0, n_prompt)) return rc;
    e->h_state[7] = 0;
    if (int rc = sllm_engine_enqueue_steps(e, n_total - 1)) return rc;
    return sllm_engine_read_tokens(e, tokens_out, n_total - 1);
}

// ---- calibrated partition ---------------------------------------------------------------------------------------
// SMs do not stream from HBM at the same rate: a stable pattern of about +-5 % per GPU (position on the dies / distance to the memory
// partitions), and every phase of the persistent kernel ends with the slowest CTA. sllm_engine_calibrate times a few real steps with
// the kernel's own %globaltimer stamps, takes every CTA's time per byte over the big weight phases, and gives each CTA a share of
// every phase's tile rows in inverse proportion — an integer allocation that minimises the latest finisher, contiguous per CTA.
static int items_of(const sllm_engine* e, const PhaseDesc& ph) {   // the n the kernels pass to cta_cut for this phase (0 = not partitioned by it)
    if (ph.kind == PH_WO_T) return 0;
    if (e->mega_fuse && ph.kind == PH_GATEUP) return ph.nunits / kFuseJT;
    return ph.ntr;
}
static std::vector<int> host_cuts(const std::vector<uint32_t>* cum, int n, int ncta) {   // boundaries b[0..ncta] exactly as cta_cut computes them
    std::vector<int> b((size_t)ncta + 1);
    for (int c = 0; c <= ncta; ++c) b[c] = cum ? (int)(((uint64_t)n * (*cum)[c]) >> 24) : (int)(((int64_t)n * c) / ncta);
    return b;
}
static std::vector<uint32_t> build_cum(int n, const std::vector<double>& tau) {
    const int ncta = (int)tau.size();
    double lo = 0.0, hi = 0.0;
    for (double t : tau) hi = std::max(hi, t);
    hi *= (double)n + 1.0;
    auto total = [&](double T) { int64_t s = 0; for (double t : tau) s += (int64_t)std::floor(T / t); return s; };
    for (int it = 0; it < 80; ++it) { const double mid = 0.5 * (lo + hi); (total(mid) >= n ? hi : lo) = mid; }
    std::vector<int> cnt((size_t)ncta);
    int64_t sum = 0;
    for (int c = 0; c < ncta; ++c) { cnt[c] = (int)std::floor(hi / tau[c]); sum += cnt[c]; }
    while (sum > n) {   // give back the surplus where it hurts most
        int worst = 0;
        for (int c = 1; c < ncta; ++c) if (cnt[c] * tau[c] > cnt[worst] * tau[worst]) worst = c;
        cnt[worst]--; sum--;
    }
    std::vector<uint32_t> cum((size_t)ncta + 1);
    int64_t b = 0;
    for (int c = 0; c < ncta; ++c) { cum[c] = (uint32_t)((b * (1ll << 24) + n - 1) / n); b += cnt[c]; }
    cum[ncta] = 1u << 24;
    return cum;
}

extern "C" int sllm_engine_calibrate(sllm_engine* e, int32_t rounds) {
    SLLM_REQUIRE(e && e->weights_loaded, SLLM_ESTATE, "calibrate: engine without weights");
    SLLM_REQUIRE(e->mega && !e->mega_ll && e->tp == 1, SLLM_ENOTSUP, "calibrate: only the single-GPU grid-barrier megakernels partition their phases by a share table");
    if (rounds < 1) rounds = 2;
    const int ncta = e->mega_plan_.grid, L = e->L, nev = 5 * L + 1;
    SLLM_REQUIRE(nev <= 512 && e->S >= 24, SLLM_ENOTSUP, "calibrate: needs at most 102 layers and max_len >= 24");
    const size_t tb = (size_t)ncta * 512 * 8 * 8;
    if (!e->trace) { SLLM_CUDA(cudaMalloc(&e->trace, tb)); SLLM_CUDA(cudaMemset(e->trace, 0, tb)); }
    e->mega_params.trace = e->trace; e->mega2_params.m.trace = e->trace;
    std::vector<unsigned long long> tr(tb / 8);
    std::map<int, std::vector<uint32_t>> tables;   // by item count
    int rc = SLLM_OK;
    for (int round = 0; round < rounds && rc == SLLM_OK; ++round) {
        std::vector<double> T((size_t)ncta, 0.0), B((size_t)ncta, 0.0);
        if ((rc = set_state(e, 1, 0, 0))) break;
        e->h_state[7] = 0;
        int done = 0;
        if ((rc = enqueue_steps(e, 3, 0, &done))) break;
        for (int step = 0; step < 6 && rc == SLLM_OK; ++step) {
            if ((rc = enqueue_steps(e, 1, 3 + step, &done))) break;
            SLLM_CUDA(cudaMemcpyAsync(tr.data(), e->trace, tb, cudaMemcpyDeviceToHost, e->stream));
            SLLM_CUDA(cudaStreamSynchronize(e->stream));
            for (size_t wp = 0; wp < e->phases_host.size(); ++wp) {
                const PhaseDesc& ph = e->phases_host[wp];
                const int n = items_of(e, ph);
                if (n <= 0 || ph.kind == PH_WO) continue;          // the big streams: qkv, gate_up, down, classifier
                const int ev = (int)wp + ph.layer + ((ph.kind != PH_QKV && ph.kind != PH_CLS) ? 1 : 0);
                const auto it = tables.find(n);
                const std::vector<int> cut = host_cuts(it == tables.end() ? nullptr : &it->second, n, ncta);
                for (int c = 0; c < ncta; ++c) {
                    const unsigned long long t1 = tr[((size_t)c * 512 + ev) * 8 + 1], t3 = tr[((size_t)c * 512 + ev) * 8 + 3];
                    if (t3 > t1 && cut[c + 1] > cut[c]) { T[c] += (double)(t3 - t1); B[c] += (double)(cut[c + 1] - cut[c]) / (double)n * (double)ph.ntr * ph.KS * ph.tile_bytes; }
                }
            }
        }
        if (rc) break;
        // CTAs without a sample (a small model: fewer tile rows in every phase than CTAs) keep the mean share
        std::vector<double> tau((size_t)ncta, 0.0);
        double mean = 0.0;
        int cnt = 0;
        for (int c = 0; c < ncta; ++c) if (B[c] > 0.0) { tau[c] = T[c] / B[c]; mean += tau[c]; cnt++; }
        if (cnt == 0 || !(mean > 0.0)) { set_error("calibrate: the kernel's timeline is empty"); rc = SLLM_ESTATE; break; }
        mean /= cnt;
        for (double& t : tau) t = (t > 0.0) ? std::min(1.25, std::max(0.8, t / mean)) : 1.0;
        e->calib_tau = tau;
        // one table per distinct item count, then every phase points at its table
        tables.clear();
        std::vector<uint32_t> flat;
        std::map<int, size_t> slot;
        for (const PhaseDesc& ph : e->phases_host) {
            const int n = items_of(e, ph);
            if (n <= 0 || slot.count(n) || (int)slot.size() >= kCumTables) continue;
            slot[n] = slot.size();
            tables[n] = build_cum(n, tau);
            flat.insert(flat.end(), tables[n].begin(), tables[n].end());
        }
        SLLM_CUDA(cudaMemcpyAsync(e->cum_dev, flat.data(), flat.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
        for (PhaseDesc& ph : e->phases_host) {
            const int n = items_of(e, ph);
            ph.cum = (n > 0 && slot.count(n)) ? e->cum_dev + slot[n] * (size_t)(ncta + 1) : nullptr;
        }
        SLLM_CUDA(cudaMemcpyAsync(e->phases_dev, e->phases_host.data(), sizeof(PhaseDesc) * e->phases_host.size(), cudaMemcpyHostToDevice, e->stream));
        SLLM_CUDA(cudaStreamSynchronize(e->stream));
    }
    e->mega_params.trace = nullptr; e->mega2_params.m.trace = nullptr;
    SLLM_CUDA(cudaMemsetAsync(e->history_dev, 0, sizeof(int32_t) * (size_t)e->S, e->stream));
    if (int rc2 = set_state(e, 0, 0, 0)) return rc2;
    e->h_state[7] = 0;
    SLLM_CUDA(cudaStreamSynchronize(e->stream));
    return rc;
}

/* the measured relative time per byte of CTA `cta` (1.0 = the mean; 0 when the engine was never calibrated) */
extern "C" float sllm_engine_calibration(const sllm_engine* e, int32_t cta) {
    return (e && cta >= 0 && (size_t)cta < e->calib_tau.size()) ? (float)e->calib_tau[(size_t)cta] : 0.0f;
}

// ---- batched prefill ------------------------------------------------------------------------------------------
static const char* prefill_unsupported(const sllm_engine* e) {
    if (!e->mega) return "batched prefill reads the megakernel's tiled weights and head-major KV cache (create the engine with SLLM_ENGINE_MEGAKERNEL)";
    return pf_unsupported_reason(e->cfg.w_dtype, e->hd, e->d, e->q_loc, e->I_loc);
}

static int prefill_workspace(sllm_engine* e, int rows) {
    if (rows <= e->pf_rows) return SLLM_OK;
    if (e->pf_ws) { SLLM_CUDA(cudaStreamSynchronize(e->stream)); cudaFree(e->pf_ws); e->pf_ws = nullptr; e->pf_rows = 0; }
    if (!e->pf) e->pf = pf_cache_create();
    const size_t T = (size_t)rows;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 1024); return o; };
    const size_t o_x = take(4 * T * e->d), o_part = take(4 * T * e->d), o_xn = take(2 * T * e->d), o_q = take(2 * T * e->q_loc),
                 o_att = take(2 * T * e->q_loc), o_s = take(2 * T * e->I_loc), o_ids = take(4 * T);
    if (cudaMalloc(&e->pf_ws, off) != cudaSuccess) { cudaGetLastError(); set_error("prefill workspace: cudaMalloc(%zu MiB) failed", off >> 20); return SLLM_ENOMEM; }
    e->pf_x = reinterpret_cast<float*>(e->pf_ws + o_x); e->pf_part = reinterpret_cast<float*>(e->pf_ws + o_part);
    e->pf_xn = reinterpret_cast<uint16_t*>(e->pf_ws + o_xn); e->pf_q = reinterpret_cast<uint16_t*>(e->pf_ws + o_q);
    e->pf_att = reinterpret_cast<uint16_t*>(e->pf_ws + o_att); e->pf_s = reinterpret_cast<uint16_t*>(e->pf_ws + o_s);
    e->pf_ids = reinterpret_cast<int32_t*>(e->pf_ws + o_ids);
    if (e->pfx_block) {   // the partial sums and the GEMM operand rows live in the block the peers have mapped
        e->pf_part = reinterpret_cast<float*>(e->pfx_block + pfx_off_part(0, kPrefillBlock, e->d));
        e->pf_xn = reinterpret_cast<uint16_t*>(e->pfx_block + pfx_off_xn(kPrefillBlock, e->d));
    }
    e->pf_rows = rows;
    return SLLM_OK;
}

__global__ void pf_pack_pair_kernel(const float* logits, const int32_t* idx, int v0, float* pair) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        pair[0] = logits[*idx];
        pair[1] = __int_as_float(v0 + *idx);
    }
}

// One block of T prompt rows at positions pos0.. through all layers. Not the last block: the last layer stops after its
// K/V rows are in the cache (its attention/FFN output would only feed logits nobody reads, model.cpp:159-165). Last block:
// the last row goes on through the final norm and the classifier (a T = 1 GEMM over this rank's vocab rows), arg-max, and
// the step state advances exactly like after a decode step.
static int prefill_block(sllm_engine* e, int T, int pos0, bool last_block) {
    const sllm_shape& s = e->cfg.shape;
    const int d = e->d, L = e->L;
    const bool tp = e->tp > 1;
    cudaStream_t st = e->stream;
    const int64_t before = g_launches;
    auto tiled_layer = [&](const Matrix& m, int l) {
        return reinterpret_cast<const uint8_t*>(m.w) + (size_t)l * mega_matrix_bytes((int)m.rows, (int)m.cols, m.kind, e->cfg.w_dtype);
    };
    const size_t kv_layer_bytes = (size_t)e->esz_kv * e->S * e->kv_loc;
    // tensor parallel with the peers' exchange blocks mapped: sum over ranks + residual + RMSNorm + distribution of the bf16 rows in ONE kernel
    // over NVLink peer memory (prefill_tp.cu) instead of ncclAllReduce + RMSNorm; the fp32 residual stream is row-sharded from the first call on
    const bool pfx = tp && e->pfx_ready;
    PfxParams xp{};
    if (pfx) {
        xp.tp = e->tp; xp.rank = e->rank; xp.T = T; xp.d = d; xp.eps = s.eps; xp.x = e->pf_x;
        xp.off_xn = pfx_off_xn(kPrefillBlock, d);
        if (e->pfx_dirty) {   // a measurement run skipped the exchanges (and with them the zeroing of the partial-sum matrices)
            SLLM_CUDA(cudaMemsetAsync(e->pfx_block + pfx_off_part(0, kPrefillBlock, d), 0, pfx_off_xn(kPrefillBlock, d) - pfx_off_part(0, kPrefillBlock, d), st));
            e->pfx_dirty = false;
        }
        for (int r = 0; r < e->tp; ++r) xp.block[r] = (r == e->rank) ? e->pfx_block : reinterpret_cast<uint8_t*>(e->pfx_peer[r]);
    }
    // which = 0: the wo partial sums, 1: the down partial sums; the call zeroes the other matrix for the GEMM that accumulates into it next
    float* const part2[2] = {pfx ? reinterpret_cast<float*>(e->pfx_block + pfx_off_part(0, kPrefillBlock, d)) : nullptr,
                             pfx ? reinterpret_cast<float*>(e->pfx_block + pfx_off_part(1, kPrefillBlock, d)) : nullptr};
    auto exchange = [&](int which, const float* norm_w, int row_lo, int row_hi, bool xlast) -> int {
        if (g_tune_mega_debug & 16) { e->pfx_dirty = true; return SLLM_OK; }
        xp.off_part = pfx_off_part(which, kPrefillBlock, d); xp.zero = part2[which ^ 1];   // measurement aid: prefill WITHOUT its exchanges (garbage results) = the GEMM / attention time alone
        xp.norm_w = norm_w; xp.row_lo = row_lo; xp.row_hi = row_hi; xp.write_xlast = xlast ? 1 : 0;
        xp.epoch = ++e->pfx_calls;
        e->pfx_ctas += (uint32_t)pf_tp_exchange_grid(T, e->tp, sm_count());
        xp.done_target = e->pfx_ctas;
        return pf_tp_exchange(xp, sm_count(), st);
    };
#define PF(call) do { if (int rc = (call)) return rc; } while (0)
    PF(pf_embed(e->pf_ids, e->emb.w, e->V, d, e->pf_x, T, st));
    for (int l = 0; l < L; ++l) {
        uint8_t* kc = reinterpret_cast<uint8_t*>(e->key_cache) + (size_t)l * kv_layer_bytes;
        uint8_t* vc = reinterpret_cast<uint8_t*>(e->value_cache) + (size_t)l * kv_layer_bytes;
        // attention norm (+ under TP the all-reduced down partial of the previous layer)
        if (!(pfx && l > 0)) PF(pf_rmsnorm(e->pf_x, (tp && l > 0) ? e->pf_part : nullptr, e->norms + (size_t)(2 * l) * d, e->pf_xn, T, d, s.eps, st));
        PfGemmArgs a{};
        a.A = e->pf_xn; a.W = tiled_layer(e->wqkv, l); a.T = T; a.N = e->q_loc + 2 * e->kv_loc; a.K = d; a.tiled = 1; a.epilogue = PF_EPI_QKV;
        a.q_out = e->pf_q; a.k_cache = kc; a.v_cache = vc; a.kv_dtype = e->cfg.kv_dtype; a.q_loc = e->q_loc; a.kv_loc = e->kv_loc; a.hd = e->hd;
        a.S = e->S; a.pos0 = pos0; a.sin_t = e->sin_t; a.cos_t = e->cos_t;
        PF(pf_gemm(e->pf, a, st));
        if (l == L - 1 && !last_block) break;
        PF(pf_attention(e->pf_q, kc, vc, e->cfg.kv_dtype, e->pf_att, T, pos0, e->S, e->hd, e->H_loc, e->KVH_loc, st));
        PfGemmArgs c{};
        c.A = e->pf_att; c.W = tiled_layer(e->wo, l); c.T = T; c.N = ((d + 1) / 2) * 2; c.K = e->q_loc; c.tiled = 1;
        c.epilogue = pfx ? PF_EPI_ACCUM : tp ? PF_EPI_STORE : PF_EPI_RESID; c.out = pfx ? part2[0] : tp ? e->pf_part : e->pf_x; c.ld_out = d; c.n_valid = d;
        PF(pf_gemm(e->pf, c, st));
        if (pfx) {
            PF(exchange(0, e->norms + (size_t)(2 * l + 1) * d, 0, T, false));
        } else {
            if (tp) SLLM_NCCL(ncclAllReduce(e->pf_part, e->pf_part, (size_t)T * d, ncclFloat, ncclSum, e->comm, st));
            PF(pf_rmsnorm(e->pf_x, tp ? e->pf_part : nullptr, e->norms + (size_t)(2 * l + 1) * d, e->pf_xn, T, d, s.eps, st));
        }
        PfGemmArgs g{};
        g.A = e->pf_xn; g.W = tiled_layer(e->wug, l); g.T = T; g.N = 2 * e->I_loc; g.K = d; g.tiled = 1; g.epilogue = PF_EPI_GATEUP;
        g.s_out = e->pf_s; g.I_loc = e->I_loc;
        PF(pf_gemm(e->pf, g, st));
        PfGemmArgs dn = c;
        dn.A = e->pf_s; dn.W = tiled_layer(e->wdown, l); dn.K = e->I_loc;
        if (pfx) dn.out = part2[1];
        PF(pf_gemm(e->pf, dn, st));
        if (pfx) {   // sum over ranks + residual + the NEXT norm (the next layer's attention norm; after the last layer: the final norm of the last row)
            if (l + 1 < L) PF(exchange(1, e->norms + (size_t)(2 * l + 2) * d, 0, T, false));
            else PF(exchange(1, e->norms + (size_t)(2 * L) * d, T - 1, T, true));
        } else if (tp) {
            SLLM_NCCL(ncclAllReduce(e->pf_part, e->pf_part, (size_t)T * d, ncclFloat, ncclSum, e->comm, st));
        }
    }
    if (last_block) {
        const size_t last = (size_t)(T - 1) * d;
        if (pfx) {   // the exchange left the normalised last row in every rank's operand buffer and its fp32 residual row in every block
            SLLM_CUDA(cudaMemcpyAsync(e->x, e->pfx_block + kPfxXlast, sizeof(float) * (size_t)d, cudaMemcpyDeviceToDevice, st));
        } else {
            PF(pf_rmsnorm(e->pf_x + last, tp ? e->pf_part + last : nullptr, e->norms + (size_t)(2 * L) * d, e->pf_xn, 1, d, s.eps, st));
            SLLM_CUDA(cudaMemcpyAsync(e->x, e->pf_x + last, sizeof(float) * (size_t)d, cudaMemcpyDeviceToDevice, st));   // emb_output: final residual
        }
        const TileGeom eg = mega_tile_geom(((e->V + 1) / 2) * 2, d, e->cfg.w_dtype);
        PfGemmArgs c{};
        c.A = pfx ? e->pf_xn + last : e->pf_xn; c.W = reinterpret_cast<const uint8_t*>(e->emb.w) + (size_t)(e->v0 / eg.R) * eg.KS * eg.tile_bytes;
        c.T = 1; c.N = ((e->V_loc + 1) / 2) * 2; c.K = d; c.tiled = 1; c.epilogue = PF_EPI_STORE; c.out = e->logits; c.ld_out = e->V_loc; c.n_valid = e->V_loc;
        PF(pf_gemm(e->pf, c, st));
        PF(sllm_argmax_f32(e->logits, e->V_loc, e->blk_idx, st));
        if (pfx) {
            xp.epoch = ++e->pfx_calls;
            PF(pf_tp_argmax(xp, e->logits, e->blk_idx, e->v0, e->state, e->prompt_dev, e->history_dev, st));
        } else if (tp) {
            pf_pack_pair_kernel<<<1, 32, 0, st>>>(e->logits, e->blk_idx, e->v0, e->tp_pairs + 2 * e->rank);
            SLLM_NCCL(ncclAllGather(e->tp_pairs + 2 * e->rank, e->tp_pairs, 2, ncclFloat, e->comm, st));
            tp_merge_kernel<<<1, 32, 0, st>>>(e->tp_pairs, e->tp, e->state, e->prompt_dev, e->history_dev);
            g_launches += 2;
        } else {
            feedback_kernel<<<1, 1, 0, st>>>(e->state, e->blk_idx, e->prompt_dev, e->history_dev);
            g_launches++;
        }
        SLLM_LAUNCH_CHECK();
    }
#undef PF
    e->total_launches += g_launches - before;
    return SLLM_OK;
}

int sllm_engine_prefill_supported(const sllm_engine* e) {
    if (!e) return 0;
    const char* why = prefill_unsupported(e);
    if (why) { set_error("%s", why); return 0; }
    return 1;
}

int sllm_engine_prefill(sllm_engine* e, const int32_t* prompt, int32_t n, int32_t start_pos) {
    SLLM_REQUIRE(e && prompt && n >= 1, SLLM_EINVAL, "bad argument");
    SLLM_REQUIRE(e->weights_loaded, SLLM_ESTATE, "weights not loaded");
    SLLM_REQUIRE(start_pos >= 0 && start_pos + n <= e->S, SLLM_EINVAL, "prompt of %d tokens at position %d overruns max_len %d", n, start_pos, e->S);
    for (int i = 0; i < n; ++i) SLLM_REQUIRE(prompt[i] >= 0 && prompt[i] < e->V, SLLM_EINVAL, "Token index %d is outside the vocabulary [0, %d).", prompt[i], e->V);
    if (const char* why = prefill_unsupported(e)) { set_error("%s", why); return SLLM_ENOTSUP; }
    SLLM_REQUIRE(e->tp == 1 || e->comm || e->pfx_ready, SLLM_ESTATE,
                 "tensor-parallel prefill needs the peers' exchange blocks (sllm_engine_prefill_p2p_import) or the NCCL communicator (sllm_engine_init_comm)");
    SLLM_REQUIRE(e->tp == 1 || e->ll_ready, SLLM_ESTATE, "tensor-parallel engine: peer areas not exchanged yet");
    // blocks of up to kBlock rows through the tensor-core path; the last block also produces the last row's logits,
    // the arg-max and the step-state update (token = arg-max, position = start_pos + n, history)
    constexpr int kBlock = kPrefillBlock;
    if (int rc = prefill_workspace(e, std::min(n, kBlock))) return rc;
    if (n > 1) SLLM_CUDA(cudaMemcpyAsync(e->history_dev + start_pos, prompt + 1, sizeof(int32_t) * (size_t)(n - 1), cudaMemcpyHostToDevice, e->stream));
    if (int rc = set_state(e, prompt[n - 1], start_pos + n - 1, 0)) return rc;
    for (int b0 = 0; b0 < n; b0 += kBlock) {
        const int T = std::min(kBlock, n - b0);
        SLLM_CUDA(cudaMemcpyAsync(e->pf_ids, prompt + b0, sizeof(int32_t) * (size_t)T, cudaMemcpyHostToDevice, e->stream));
        if (int rc = prefill_block(e, T, start_pos + b0, b0 + T == n)) return rc;
    }
    e->h_state[7] = start_pos + n;
    return SLLM_OK;
}

int sllm_engine_buffer(sllm_engine* e, int32_t id, void** ptr, int64_t* n, int32_t* dtype) {
    SLLM_REQUIRE(e && ptr && n && dtype, SLLM_EINVAL, "null argument");
    const int64_t kvn = (int64_t)e->L * e->S * e->kv_loc;
    *dtype = SLLM_F32;
    switch (id) {
        case 2: *ptr = e->key_cache; *n = kvn; *dtype = e->cfg.kv_dtype; break;
        case 3: *ptr = e->value_cache; *n = kvn; *dtype = e->cfg.kv_dtype; break;
        case 4: *ptr = e->x; *n = e->d; break;
        case 5: *ptr = e->xb; *n = e->d; break;
        case 6: *ptr = e->q; *n = e->q_loc; break;
        case 8: *ptr = e->att; *n = e->q_loc; break;
        case 9: *ptr = e->att_out; *n = e->d; break;
        case 10: *ptr = e->h; *n = e->d; break;
        case 11: *ptr = e->up; *n = e->I_loc; break;
        case 12: *ptr = e->gate; *n = e->I_loc; break;
        case 14: *ptr = e->swi; *n = e->I_loc; break;
        case 15: *ptr = e->ffn_out; *n = e->d; break;
        case 16: *ptr = e->logits; *n = e->V_loc; break;
        case 17: *ptr = e->sin_t; *n = (int64_t)e->S * (e->hd / 2); break;
        case 18: *ptr = e->cos_t; *n = (int64_t)e->S * (e->hd / 2); break;
        case 100: *ptr = e->emb.w; *n = (int64_t)e->V * e->d; *dtype = e->cfg.w_dtype; break;
        case 101: *ptr = e->norms; *n = (int64_t)(2 * e->L + 1) * e->d; break;
        case 102: *ptr = e->wqkv.w; *n = (int64_t)e->L * e->wqkv.rows * e->wqkv.cols; *dtype = e->cfg.w_dtype; break;
        case 103: *ptr = e->wo.w; *n = (int64_t)e->L * e->wo.rows * e->wo.cols; *dtype = e->cfg.w_dtype; break;
        case 104: *ptr = e->wug.w; *n = (int64_t)e->L * e->wug.rows * e->wug.cols; *dtype = e->cfg.w_dtype; break;
        case 105: *ptr = e->wdown.w; *n = (int64_t)e->L * e->wdown.rows * e->wdown.cols; *dtype = e->cfg.w_dtype; break;
        case 200: {   // megakernel timeline: allocate on first request, stamps are written from the next step on
            SLLM_REQUIRE(e->mega, SLLM_ESTATE, "trace needs megakernel mode");
            const size_t nb = (size_t)(e->mega_ll ? e->ll_plan.grid : e->mega_plan_.grid) * 512 * 8 * 8;
            if (!e->trace) { SLLM_CUDA(cudaMalloc(&e->trace, nb)); SLLM_CUDA(cudaMemset(e->trace, 0, nb)); e->mega_params.trace = e->trace; e->mega2_params.m.trace = e->trace; e->ll_params.trace = e->trace; }
            *ptr = e->trace; *n = (int64_t)(nb / 4); *dtype = SLLM_F32; break;
        }
        case 201:   // timeline of the most recent prefill exchange (prefill_tp.cu): 2 x 8 u64 nanosecond stamps (last full call, last one-row call)
            SLLM_REQUIRE(e->pfx_block, SLLM_ESTATE, "no prefill exchange block");
            *ptr = e->pfx_block + kPfxTrace; *n = 32; *dtype = SLLM_F32; break;
        case 110: *ptr = e->emb.sc; *n = e->emb.sc ? (int64_t)e->V * e->d / e->cfg.group : 0; break;
        case 112: *ptr = e->wqkv.sc; *n = e->wqkv.sc ? (int64_t)e->L * e->wqkv.rows * e->wqkv.cols / e->cfg.group : 0; break;
        default: SLLM_REQUIRE(false, SLLM_EINVAL, "unknown buffer id %d", id);
    }
    return SLLM_OK;
}

int64_t sllm_engine_step_bytes(const sllm_engine* e, int32_t pos) {
    if (!e) return 0;
    const double bw = e->cfg.w_dtype == SLLM_F32 ? 4.0 : e->cfg.w_dtype == SLLM_BF16 ? 2.0 : 1.0 + 4.0 / e->cfg.group;
    const double d = e->d, L = e->L;
    const double mats = (double)e->V_loc * d + L * ((double)(e->q_loc + 2 * e->kv_loc) * d + d * e->q_loc + 3.0 * e->I_loc * d);
    const double bytes = bw * mats + 4.0 * (2 * L + 1) * d + bw * d + (double)e->esz_kv * 2 * L * e->kv_loc * (pos + 1) +
                         (double)e->esz_kv * 2 * L * e->kv_loc;
    return (int64_t)bytes;
}

int sllm_engine_enqueue_kernel(sllm_engine* e, int32_t kind, int32_t layer) {
    SLLM_REQUIRE(e && e->weights_loaded && e->fused, SLLM_ESTATE, "enqueue_kernel needs a fused engine with weights");
    SLLM_REQUIRE(!e->mega, SLLM_ESTATE, "enqueue_kernel is not available in megakernel mode (weights are tiled, the step is one kernel)");
    SLLM_REQUIRE(layer >= 0 && layer < e->L, SLLM_EINVAL, "layer %d outside [0,%d)", layer, e->L);
    const int l = (kind == K_CLS) ? -1 : layer;
    e->timing_only = true;
    int rc;
    switch (e->cfg.w_dtype) {
        case SLLM_F32: rc = enqueue_kernel<SLLM_F32>(e, kind, l); break;
        case SLLM_BF16: rc = enqueue_kernel<SLLM_BF16>(e, kind, l); break;
        default: rc = enqueue_kernel<SLLM_INT8>(e, kind, l); break;
    }
    e->timing_only = false;
    return rc;
}

int64_t sllm_engine_kernel_bytes(const sllm_engine* e, int32_t kind, int32_t pos) {
    if (!e) return 0;
    const double bw = e->cfg.w_dtype == SLLM_F32 ? 4.0 : e->cfg.w_dtype == SLLM_BF16 ? 2.0 : 1.0 + 4.0 / e->cfg.group;
    const double d = e->d;
    switch (kind) {
        case K_EMBED: return (int64_t)(bw * d);
        case K_QKV: return (int64_t)(bw * (e->q_loc + 2.0 * e->kv_loc) * d + 4.0 * d + e->esz_kv * 2.0 * e->kv_loc);
        case K_MHA: return (int64_t)(e->esz_kv * 2.0 * e->kv_loc * (pos + 1));
        case K_WO: return (int64_t)(bw * d * e->q_loc);
        case K_GATEUP: return (int64_t)(bw * 2.0 * e->I_loc * d + 4.0 * d);
        case K_DOWN: return (int64_t)(bw * d * e->I_loc);
        case K_CLS: return (int64_t)(bw * e->V_loc * d + 4.0 * d);
        default: return 0;
    }
}

const char* sllm_engine_mode(const sllm_engine* e) {
    if (!e) return "null";
    if (e->mega && e->mega2) return e->mega_fuse ? "megakernel(v2,fused-down)" : "megakernel(v2)";
    if (e->mega) return e->mega_ll ? "megakernel(ll)" : e->mega_fuse ? "megakernel(fused-down)" : "megakernel";
    if (!e->fused) return "unfused";
    return e->use_graph ? (e->pdl ? "fused+graph+pdl" : "fused+graph") : (e->pdl ? "fused+pdl" : "fused");
}

int32_t sllm_engine_kv_layout(const sllm_engine* e) { return (e && e->mega) ? 1 : 0; }

int32_t sllm_engine_step_launches(const sllm_engine* e) { return e ? e->step_launches : 0; }
int64_t sllm_engine_total_launches(const sllm_engine* e) { return e ? e->total_launches : 0; }

}  // extern "C"
