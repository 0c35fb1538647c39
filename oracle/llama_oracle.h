/* oracle/llama_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's CPU forward path (Boundwhd/SimpleLLMInference), one function per
 * reference kernel, each citing the file:line it follows. It is the parity CHECKER for the CUDA product path
 * and (kind "port") a CPU baseline. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it. The product path never calls into this file.
 *
 * Pinning: the reference ships no golden vectors, so this restatement is pinned against the
 * reference ITSELF: oracle/_ref (the unmodified reference sources compiled by oracle/Makefile) is run on
 * seeded inputs, op by op and whole-model, and must agree BIT-FOR-BIT (tests/test_oracle_cpu.py), and against
 * the fixtures under tests/golden/ that tests/golden/make_golden.py generated from oracle/_ref.
 *
 * Deliberate differences from the reference (SURVEY.md Appendix D "do not mirror"): K is rotated over
 * kv_hidden elements only (the reference over-runs into the next cache rows for GQA, rope_kernel.cpp:27-38),
 * which leaves results unchanged for pos <= max_len - heads/kv_heads; 64-bit offsets.
 * One extension: orc_set_kv_bf16() rounds every K/V cache row to bfloat16 (after RoPE), which is the
 * definition of the product's bf16-KV-cache variant — the reference has no such mode.
 */
#ifndef ORACLE_LLAMA_ORACLE_H
#define ORACLE_LLAMA_ORACLE_H
#include <stdint.h>
#include "synth_weights.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;

/* ---- per-op restatements --------------------------------------------------------------------------- */
void orc_embedding(int32_t token, const float* table, float* out, int32_t vocab, int32_t d);
void orc_rmsnorm(const float* x, const float* w, float* y, int32_t d, float eps);
void orc_matmul(const float* x, const float* W, float* y, int32_t rows, int32_t cols, float scale);
void orc_rope_cache(int32_t head_dim, int32_t max_len, float theta, float* sin_c, float* cos_c);
void orc_rope(float* q, float* k, int32_t pos, const float* sin_c, const float* cos_c, int32_t q_dim,
              int32_t k_dim, int32_t head_dim);
void orc_softmax(float* x, int32_t n);
void orc_mha(const float* q, float* score, const float* kc, const float* vc, float* out, int32_t layer,
             int32_t pos, int32_t max_len, int32_t head_dim, int32_t heads, int32_t kv_heads);
void orc_add(const float* a, const float* b, float* out, int32_t n);
void orc_swiglu(const float* up, const float* gate, float* out, int32_t n);
int32_t orc_argmax(const float* logits, int32_t n);

/* ---- whole model ------------------------------------------------------------------------------------ */
/* blob: fp32 weights in the reference's order (synth_weights.h); NOT copied, must outlive the model. */
orc_model* orc_create(const syn_shape* shape, const float* blob);
void orc_destroy(orc_model* m);
void orc_set_threads(orc_model* m, int n);      /* row-parallel GEMV; per-row sums unchanged => same bits */
void orc_set_kv_bf16(orc_model* m, int on);
void orc_forward(orc_model* m, int32_t token, int32_t pos, float* logits /* may be NULL */);
int32_t orc_greedy(orc_model* m, const int32_t* prompt, int32_t n_prompt, int32_t n_total, int32_t* out,
                   float* last_logits);
/* buffer ids follow the reference's ModelBufferType (include/model/model.h:14-34) */
void orc_read(orc_model* m, int32_t buffer_id, int64_t offset, int64_t n, float* out);
/* test aid: overwrite floats [offset, offset+n) of the key (2) or value (3) cache, layout [L][S][kv] (model.cpp:264-265) */
void orc_write(orc_model* m, int32_t buffer_id, int64_t offset, int64_t n, const float* src);

#ifdef __cplusplus
}
#endif
#endif
