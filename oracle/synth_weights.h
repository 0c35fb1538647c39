/* oracle/synth_weights.h — TEST INFRASTRUCTURE ONLY (CPU side of the synthetic-weight contract).
 *
 * Counter-based, integer-exact weight generator shared by the CPU oracle and the CUDA product path
 * (simplellminference_b200/csrc/synth.cuh restates the same arithmetic for the device), so both sides
 * materialise BIT-IDENTICAL weights without shipping multi-GiB files (SURVEY.md §8d "Weight values").
 *
 *   stream(t)   = sm64(seed * 0x100000001B3 + t)                  t = tensor id (blob segment, below)
 *   h(t, i)     = sm64(stream(t) + i)                             i = element index inside the segment
 *   s           = sum of the four 16-bit fields of h              Irwin-Hall(4) ~ normal, 0..262140
 *   value       = mean + (float)(s - 131070) * c                  c = (float)(std * sqrt(3)/65535.0)
 *
 * one IEEE fp32 multiply and one fp32 add: identical on x86 (-ffp-contract=off) and sm_100a (__fmul_rn /
 * __fadd_rn). Segments follow the reference blob order (source/model/model.cpp:340-468):
 *   0 E[V][d] (embedding == classifier)   1 norms[(2L+1)][d]   2 wq[L][d][d]   3 wk[L][kv][d]
 *   4 wv[L][kv][d]   5 wo[L][d][d]   6 up[L][I][d]   7 gate[L][I][d]   8 down[L][d][I]
 * std: E 1.0; norms N(1, 0.02); projections 4/sqrt(fan_in) (fan_in = d, or I for down).
 *
 * Weight storage types of the product path are modelled by rounding the generated value and expanding
 * it back to fp32 for the oracle: SYN_F32 (none), SYN_BF16 (RNE to bfloat16), SYN_INT8 (symmetric int8,
 * one fp32 scale per `group` consecutive elements of a row: scale = amax/127, q = rint(w/scale)).
 * Norm vectors (segment 1) always stay fp32.
 */
#ifndef ORACLE_SYNTH_WEIGHTS_H
#define ORACLE_SYNTH_WEIGHTS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { SYN_F32 = 0, SYN_BF16 = 1, SYN_INT8 = 2 };

typedef struct {
    int32_t vocab, head_dim, hidden, kv_hidden, inter, max_len, layers, heads, kv_heads;
    float eps, theta;
} syn_shape;

/* number of floats in the reference blob for this shape */
int64_t syn_blob_floats(const syn_shape* s);
/* offset (floats) and length of segment t in the blob, std/mean/fan-in row length of that segment */
void syn_segment(const syn_shape* s, int t, int64_t* offset, int64_t* count, int64_t* row_len, float* mean,
                 float* c);
/* raw (unrounded) value of element i of segment t */
float syn_value(uint64_t seed, int t, int64_t i, float mean, float c);
/* RNE fp32 -> bf16 -> fp32 */
float syn_round_bf16(float x);
/* fill out[0..count) with elements [first, first+count) of segment t, as the oracle must see them
 * (i.e. after rounding to `wdtype`; first must be a multiple of `group` for SYN_INT8). */
void syn_fill_segment(const syn_shape* s, uint64_t seed, int t, int wdtype, int group, int64_t first,
                      int64_t count, float* out, int n_threads);
/* fill a whole blob (all nine segments) */
void syn_fill_blob(const syn_shape* s, uint64_t seed, int wdtype, int group, float* blob, int n_threads);
/* int8 view of the same data: q[0..count) and scales[0..count/group) for segment t starting at `first` */
void syn_fill_segment_int8(const syn_shape* s, uint64_t seed, int t, int group, int64_t first, int64_t count,
                           int8_t* q, float* scales);

#ifdef __cplusplus
}
#endif
#endif
