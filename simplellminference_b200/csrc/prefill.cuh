// csrc/prefill.cuh — batched prompt processing ("prefill"): the same layer loop as the decode step
// (reference source/model/model.cpp:50-128) for T prompt positions at once, so that the dense contractions become
// GEMMs on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
// The reference has no batched prefill (it calls the single-token forward once per prompt token and drops the
// logits, model.cpp:157-166); the result of this path is defined by that loop: same KV-cache contents, same
// next token / last-position logits, within the bf16-operand tolerance stated in tests/test_prefill_gpu.py.
#pragma once
#include "megakernel.cuh"

namespace sllm {

// fused epilogues of the tensor-core GEMM  C[T][N] = A[T][K] . W[N][K]^T  (fp32 accumulators in TMEM)
enum PfEpilogue {
    PF_EPI_STORE = 0,   // C -> out[t][n] fp32 (row-parallel partial sums under tensor parallelism; plain GEMM)
    PF_EPI_RESID = 1,   // out[t][n] += C           (wo / down with the residual add, model.cpp:86,124)
    PF_EPI_QKV = 2,     // unit-ordered rows: RoPE on (j, j+hd/2) pairs -> q (bf16) and the K cache; V rows -> V cache
    PF_EPI_GATEUP = 3,  // unit-ordered rows (up_u, gate_u): sigmoid(gate)*up -> bf16 (swiglu_kernel.cpp:12-13)
    PF_EPI_ACCUM = 4,   // C -> out[t][n] fp32 where out was ZEROED by the caller: like STORE, but K may be split (partial sums meet in L2 with red.add)
};

struct PfGemmArgs {
    const void* A;        // bf16 [T][K] row-major activations
    const void* W;        // bf16 weights: megakernel tiled layout (tiled = 1) or plain row-major [N][K] (tiled = 0)
    int32_t T, N, K;      // N = physical rows (tiled: 2 * units), K = logical row length
    int32_t tiled;
    int32_t epilogue;
    // STORE / RESID
    float* out;           // [T][ld_out]
    int32_t ld_out, n_valid;   // columns >= n_valid are not stored
    // QKV
    uint16_t* q_out;      // bf16 [T][q_loc]
    uint8_t *k_cache, *v_cache;   // this layer's head-major cache [KVH_loc][S][hd]
    int32_t kv_dtype, q_loc, kv_loc, hd, S, pos0;
    const float *sin_t, *cos_t;   // [S][hd/2]
    // GATEUP
    uint16_t* s_out;      // bf16 [T][I_loc]
    int32_t I_loc;
    int32_t bn;           // 0 = choose; 128 / 256 = force the N tile
};

struct PfCache;   // tensor-map cache (one CUtensorMap per weight matrix and tile shape), owned by the engine
PfCache* pf_cache_create();
void pf_cache_destroy(PfCache*);

const char* pf_unsupported_reason(int w_dtype, int hd, int d, int q_loc, int I_loc);   // nullptr = supported
int pf_gemm(PfCache* cache, const PfGemmArgs& a, cudaStream_t st);
// causal attention of T queries at positions pos0.. over cache rows [0, pos0+T): q bf16 [T][q_loc] -> out bf16 [T][q_loc]
int pf_attention(const uint16_t* q, const void* k_cache, const void* v_cache, int kv_dtype, uint16_t* out, int T, int pos0, int S, int hd,
                 int heads, int kv_heads, cudaStream_t st);
// x[t][:] = E[ids[t]][:] from the tiled embedding matrix (emb_kernel.cpp:9-16), fp32
int pf_embed(const int32_t* ids_dev, const void* emb_tiled, int vocab, int d, float* x, int T, cudaStream_t st);
// x[t] += add[t] (optional), y[t] = bf16(rmsnorm(x[t]) * w)   (rms_kernel.cpp:12-22)
int pf_rmsnorm(float* x, const float* add, const float* w, uint16_t* y, int T, int d, float eps, cudaStream_t st);

// ---- tensor-parallel exchange over peer memory (prefill_tp.cu) ----------------------------------------------------
// A rank's exchange block (one cudaMalloc, CUDA-IPC mapped by every peer): flags, the (value, index) pairs of the classifier shards, the
// last residual row, then the [rows][d] fp32 partial sums the row-parallel GEMMs write and the [rows][d] bf16 rows the next GEMM reads.
constexpr size_t kPfxFlagIn = 0, kPfxFlagOut = 1024, kPfxCounter = 2048, kPfxPairs = 2560, kPfxTrace = 3072, kPfxXlast = 4096, kPfxData = 4096 + 32768;
struct PfxParams {
    int32_t tp, rank, T, d;
    int32_t row_lo, row_hi;       // rows handled by this call (their owners do the work): [0, T), or [T-1, T) for the final norm
    uint32_t epoch;               // call counter, the same on every rank
    uint32_t done_target;         // value of the block's CTA counter after this call (calls so far x grid)
    float eps;
    int32_t write_xlast;          // also broadcast the fp32 residual row (emb_output of the reference's buffer table)
    uint8_t* block[kMaxTp];       // every rank's exchange block as mapped HERE (block[rank] = the local one)
    size_t off_part, off_xn;      // off_part: the partial-sum matrix THIS call sums (one of two: wo's and down's)
    float* zero;                  // the OTHER partial-sum matrix of this rank ([T][d] floats): zeroed here for the split-K GEMM that fills it next
    float* x;                     // local residual stream [rows][d]; only this rank's rows are kept current
    const float* norm_w;
};
// one CTA per owned row, at most one per SM (512 threads holding a chunk of every rank's row in registers; two per SM at 2 ranks)
inline int pf_tp_exchange_grid(int T, int tp, int sms) { const int per = (T + tp - 1) / tp, cap = tp == 2 ? 2 * sms : sms; return per < 1 ? 1 : (per < cap ? per : cap); }
size_t pfx_block_bytes(int rows, int d);
size_t pfx_off_part(int which, int rows, int d);
size_t pfx_off_xn(int rows, int d);
int pf_tp_exchange(const PfxParams& p, int sms, cudaStream_t st);
int pf_tp_argmax(const PfxParams& p, const float* logits, const int32_t* idx, int v0, StepState* state, const int32_t* prompt, int32_t* history,
                 cudaStream_t st);

}  // namespace sllm
