// csrc/decode_fused.cuh — the five fused kernels of one decode layer plus the classifier/argmax kernel.
//
// Per layer (reference model.cpp:50-129 runs 13 op calls; here 5 launches, HBM traffic = weights + KV only):
//   A  qkv      : RMSNorm(x) -> [Wq;Wk;Wv] GEMV -> RoPE(q,k) -> q buffer, K/V cache row `pos`   (ops :52-:66)
//   B  mha      : split-KV flash decoding (mha.cu)                                             (:70-:78)
//   C  wo       : Wo GEMV + residual  h = x + Wo.att                                           (:80-:90)
//   D  gate_up  : RMSNorm(h) -> [Wup;Wgate] GEMV -> sigmoid(gate)*up                           (:93-:115)
//   E  down     : Wdown GEMV + residual  x = Wdown.s + h                                       (:118-:128)
// and once per token:  embedding gather (:48);  F  RMSNorm -> tied classifier GEMV -> first-max argmax ->
// token/position feedback on the device (:131-:139 and the predict loop :157-:185).
// Under tensor parallelism C and E write partial sums instead (the all-reduce and the residual add follow).
#pragma once
#include "gemv_core.cuh"

namespace sllm {

struct StepState {          // device-resident decode state (one per engine)
    int32_t token;          // input token of the current step
    int32_t pos;            // position of the current step
    int32_t n_prompt;       // prompt length (tokens fed verbatim while pos+1 < n_prompt)
    int32_t ticket;         // last-block ticket of the classifier kernel (always returns to 0)
    int32_t next;           // argmax of the last step (this rank's view / global after TP reduction)
    int32_t pad[3];         // pad[0] = megakernel barrier epoch, pad[1] = monotonic step counter (peer all-reduce epochs)
};

// ---- one-shot all-reduce over NVLink peer memory, fused into the GEMV kernels -------------------------------
// Low-latency ("LL") protocol: every fp32 partial sum travels as ONE 8-byte word {value bits, epoch}, and an aligned
// 8-byte store is a single transaction, so value and flag arrive together: no fence, no separate flag, no ticket.
// Every rank owns a receive area [2 parities][tp][n] of such words; peers map it through CUDA IPC. The kernel that
// PRODUCES partial sums (row-parallel GEMV: wo / down) stores each value straight into slot [my rank] of every
// rank's area (NVLink stores; a local store for itself). The kernel that CONSUMES them (the next RMSNorm prologue)
// spins on each word until its epoch field equals e, and adds the tp partial vectors in rank order — the same
// order on every rank, so all ranks compute bit-identical activations. e = step * ops_per_step + op + 1 comes from
// the device-resident step counter, so the launch sequence is CUDA-graph capturable; areas alternate by epoch
// parity (dependencies keep every rank at most one op ahead of the slowest one).
constexpr int kMaxTp = 8;
struct P2PComm {
    uint2* recv[kMaxTp];        // recv[r] = rank r's receive area (recv[rank] is local memory)
    int tp, rank, n;            // n = words per partial vector
    int ops_per_step;
    const int32_t* step;        // device step counter (StepState::pad[1])
};
__device__ __forceinline__ unsigned p2p_epoch(const P2PComm& c, int op) { return (unsigned)(*c.step) * (unsigned)c.ops_per_step + (unsigned)op + 1u; }
__device__ __forceinline__ uint2* p2p_slot(const P2PComm& c, int dst_rank, unsigned epoch, int src_rank) {
    return c.recv[dst_rank] + ((size_t)(epoch & 1u) * c.tp + src_rank) * c.n;
}
__device__ __forceinline__ void p2p_send(uint2* p, float v, unsigned epoch) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
// two adjacent words (16 bytes); spins until both carry `epoch`
__device__ __forceinline__ float2 p2p_recv2(const uint2* p, unsigned epoch) {
    uint4 w;
    unsigned spins = 0;
    while (true) {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
        if (w.y == epoch && w.w == epoch) break;
        if (++spins > (1u << 26)) __trap();   // a peer died: fail instead of hanging the GPU
    }
    return make_float2(__uint_as_float(w.x), __uint_as_float(w.z));
}

// used by stage_x_rmsnorm (gemv_core.cuh): the tp slot addresses of op `op` in THIS rank's receive area + its epoch
__device__ __forceinline__ unsigned p2p_slots(const P2PComm* c, int op, const uint2** slots, int* tp) {
    const unsigned e = p2p_epoch(*c, op);
    for (int r = 0; r < c->tp; ++r) slots[r] = p2p_slot(*c, c->rank, e, r);
    *tp = c->tp;
    return e;
}

struct GemvBase {
    const void* W_;
    const float* sc_;
    int grp_, cols_;
    __device__ const void* W() const { return W_; }
    __device__ const float* scales() const { return sc_; }
    __device__ int group() const { return grp_; }
    __device__ int cols() const { return cols_; }
};

// ---- A: RMSNorm -> QKV -> RoPE -> cache ---------------------------------------------------------------
template <int WD>
struct QkvPolicy : GemvBase {
    const float* x;        // residual stream [d]
    const float* add;      // TP only: all-reduced partial to add to x first (else nullptr)
    float* sum_out;        // TP only: where CTA 0 stores x + add
    const P2PComm* p2p;    // TP over peer memory: add = sum of the tp partial vectors of op `p2p_op` (else nullptr)
    int p2p_op;
    const float* norm_w;   // [d]
    float eps;
    const int32_t* pos_dev;
    const float* sin_t;    // [S][hd/2]
    const float* cos_t;
    float* q_out;          // [q_dim]
    void* k_cache;         // layer base: [S][kv_dim] in kv dtype
    void* v_cache;
    int kv_dtype, q_dim, kv_dim, hd;
    __device__ int units() const { return (q_dim + 2 * kv_dim) >> 1; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const {
        const int half = hd >> 1;
        const int rope_units = (q_dim + kv_dim) >> 1;
        if (u < rope_units) {  // q or k head: RoPE partners (j, j + hd/2)
            const int head = u / half, j = u - head * half;
            r0 = (int64_t)head * hd + j;
            r1 = r0 + half;
        } else {  // v: two consecutive rows
            r0 = (int64_t)q_dim + kv_dim + 2 * (int64_t)(u - rope_units);
            r1 = r0 + 1;
        }
    }
    __device__ void stage(float* xs, float* red) const { stage_x_rmsnorm<WD>(xs, red, x, norm_w, cols_, eps, add, sum_out, p2p, p2p_op); }
    __device__ void store_kv(void* cache, int64_t idx, float v) const {
        if (kv_dtype == SLLM_BF16) reinterpret_cast<uint16_t*>(cache)[idx] = f32_to_bf16_bits(v);
        else reinterpret_cast<float*>(cache)[idx] = v;
    }
    __device__ void emit(int u, float s0, float s1) {
        const int half = hd >> 1;
        const int rope_units = (q_dim + kv_dim) >> 1;
        const int pos = *pos_dev;
        if (u < rope_units) {
            const int head = u / half, j = u - head * half;
            const float fci = sin_t[(int64_t)pos * half + j], fcr = cos_t[(int64_t)pos * half + j];
            const float o0 = s0 * fcr - s1 * fci;  // rope_kernel.cpp:36-37
            const float o1 = s1 * fcr + s0 * fci;
            const int r0 = head * hd + j;
            if (r0 < q_dim) {
                q_out[r0] = o0;
                q_out[r0 + half] = o1;
            } else {
                const int64_t base = (int64_t)pos * kv_dim + (r0 - q_dim);
                store_kv(k_cache, base, o0);
                store_kv(k_cache, base + half, o1);
            }
        } else {
            const int64_t base = (int64_t)pos * kv_dim + 2 * (int64_t)(u - rope_units);
            store_kv(v_cache, base, s0);
            store_kv(v_cache, base + 1, s1);
        }
    }
    __device__ void finish(float*) {}
};

// ---- C / E: GEMV + residual ---------------------------------------------------------------------------
template <int WD>
struct ResidualPolicy : GemvBase {
    const float* x;         // GEMV input [cols]
    const float* resid;     // [rows] or nullptr (tensor-parallel partial sums: no residual here)
    float* y;               // [rows]
    int nrows;
    const P2PComm* p2p;     // TP over peer memory: partial sums go straight into every rank's receive area
    int p2p_op;
    StepState* st;
    __device__ int units() const { return (nrows + 1) >> 1; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const { r0 = 2 * (int64_t)u; r1 = min(2 * u + 1, nrows - 1); }
    __device__ void stage(float* xs, float*) const { stage_x_plain<WD>(xs, x, cols_); }
    __device__ void emit(int u, float s0, float s1) {
        const int r = 2 * u;
        if (p2p) {
            const unsigned e = p2p_epoch(*p2p, p2p_op);
            for (int dst = 0; dst < p2p->tp; ++dst) {                 // NVLink stores (local for dst == rank)
                uint2* slot = p2p_slot(*p2p, dst, e, p2p->rank);
                p2p_send(slot + r, s0, e);
                if (r + 1 < nrows) p2p_send(slot + r + 1, s1, e);
            }
            return;
        }
        y[r] = resid ? resid[r] + s0 : s0;  // add_kernel.cpp:10-13: out = in1 + in2
        if (r + 1 < nrows) y[r + 1] = resid ? resid[r + 1] + s1 : s1;
    }
    __device__ void finish(float*) {}
};

// ---- D: RMSNorm -> up/gate -> sigmoid(gate)*up --------------------------------------------------------
template <int WD>
struct GateUpPolicy : GemvBase {
    const float* h;
    const float* add;      // TP only (see QkvPolicy)
    float* sum_out;
    const P2PComm* p2p;
    int p2p_op;
    const float* norm_w;
    float eps;
    float* s_out;  // [inter]
    int inter;     // local intermediate size; W = [up rows (inter)][gate rows (inter)]
    __device__ int units() const { return inter; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const { r0 = u; r1 = (int64_t)inter + u; }
    __device__ void stage(float* xs, float* red) const { stage_x_rmsnorm<WD>(xs, red, h, norm_w, cols_, eps, add, sum_out, p2p, p2p_op); }
    __device__ void emit(int u, float up, float gate) {
        const float sg = 1.0f / (1.0f + expf(-gate));  // swiglu_kernel.cpp:12-13
        s_out[u] = sg * up;
    }
    __device__ void finish(float*) {}
};

// ---- F: RMSNorm -> classifier -> argmax -> feedback ---------------------------------------------------

template <int WD>
struct ClsPolicy : GemvBase {
    const float* x;
    const float* add;       // TP only (see QkvPolicy)
    float* sum_out;
    const P2PComm* p2p;
    int p2p_op;
    const float* norm_w;
    float eps;
    float* logits;          // [nrows] local logits (model_pred)
    int nrows, row0;        // local vocab rows and the global index of the first one
    float* blk_val;         // [gridDim.x] per-CTA best value
    int32_t* blk_idx;       // [gridDim.x] per-CTA best (global) index
    StepState* st;
    const int32_t* prompt;  // [n_prompt] device copy of the prompt
    int32_t* history;       // [max_len] token that followed position p
    int single_rank;        // 1: finish the step here (token feedback); 0: TP — a later kernel merges ranks
    float best_v;
    int best_i;
    __device__ int units() const { return (nrows + 1) >> 1; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const { r0 = 2 * (int64_t)u; r1 = min(2 * u + 1, nrows - 1); }
    __device__ void stage(float* xs, float* red) {
        best_v = -INFINITY;
        best_i = 0x7fffffff;
        stage_x_rmsnorm<WD>(xs, red, x, norm_w, cols_, eps, add, sum_out, p2p, p2p_op);
    }
    __device__ static void better(float& v, int& i, float ov, int oi) {
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }  // first maximum, argmax.cpp:11
    }
    __device__ void emit(int u, float s0, float s1) {
        const int r = 2 * u;
        logits[r] = s0;
        better(best_v, best_i, s0, row0 + r);
        if (r + 1 < nrows) {
            logits[r + 1] = s1;
            better(best_v, best_i, s1, row0 + r + 1);
        }
    }
    // all threads of the CTA; lane 0 of each warp holds that warp's best
    __device__ void finish(float* red) {
        __shared__ float sv[kGemvWarps];
        __shared__ int si[kGemvWarps];
        __shared__ int s_last;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { sv[warp] = best_v; si[warp] = best_i; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = sv[0];
            int i = si[0];
            for (int w = 1; w < kGemvWarps; ++w) better(v, i, sv[w], si[w]);
            blk_val[blockIdx.x] = v;
            blk_idx[blockIdx.x] = i;
            __threadfence();
            s_last = (atomicAdd(&st->ticket, 1) == (int)gridDim.x - 1);
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        float v = -INFINITY;
        int i = 0x7fffffff;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) better(v, i, __ldcg(blk_val + b), __ldcg(blk_idx + b));
        for (int o = 16; o > 0; o >>= 1) better(v, i, __shfl_xor_sync(0xffffffffu, v, o), __shfl_xor_sync(0xffffffffu, i, o));
        if (lane == 0) { sv[warp] = v; si[warp] = i; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kGemvWarps; ++w) better(v, i, sv[w], si[w]);
            if (i == 0x7fffffff) i = 0;
            st->ticket = 0;
            blk_val[gridDim.x] = v;  // slot [grid] = this rank's best, for the TP merge
            blk_idx[gridDim.x] = i;
            if (single_rank) step_feedback(st, prompt, history, i);
        }
        (void)red;
    }
    // predict loop, model.cpp:157-185: prompt tokens are fed verbatim, then the argmax is fed back
    __device__ static void step_feedback(StepState* st, const int32_t* prompt, int32_t* history, int argmax_tok) {
        const int pos = st->pos;
        const int nxt = (pos + 1 < st->n_prompt) ? prompt[pos + 1] : argmax_tok;
        st->next = argmax_tok;
        history[pos] = nxt;
        st->token = nxt;
        st->pos = pos + 1;
        st->pad[1] += 1;   // monotonic step counter (epochs of the peer-memory all-reduce)
    }
};

// generic kernel over a policy held by value (policies carry per-thread mutable state)
template <int WD, class Policy>
__global__ void __launch_bounds__(kGemvThreads) fused_gemv_kernel(Policy pol) {
    gemv_body<WD>(pol);
    extern __shared__ __align__(16) float smem[];
    pol.finish(smem + pol.cols());
}

}  // namespace sllm
