/* oracle/llama_oracle.c — TEST INFRASTRUCTURE ONLY. See llama_oracle.h. Build with
 * -O2 -ffp-contract=off (no FMA contraction) so every fp32 operation rounds exactly where the reference's
 * -O2 -ffp-contract=off build rounds. */
#include "llama_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- ops ---- */

/* source/kernel/cpu/emb_kernel.cpp:9-16 — row copy E[token,:] */
void orc_embedding(int32_t token, const float* table, float* out, int32_t vocab, int32_t d) {
    (void)vocab;
    memcpy(out, table + (int64_t)token * d, sizeof(float) * (size_t)d);
}

/* source/kernel/cpu/rms_kernel.cpp:12-22 — serial sum of squares, 1/sqrt(mean+eps), (x*inv)*w */
void orc_rmsnorm(const float* x, const float* w, float* y, int32_t d, float eps) {
    float ss = 0.0f;
    for (int i = 0; i < d; ++i) ss += x[i] * x[i];
    float mean = ss / (float)d;
    float rms = sqrtf(mean + eps);
    float inv = 1.0f / rms;
    for (int i = 0; i < d; ++i) y[i] = (x[i] * inv) * w[i];
}

/* source/kernel/cpu/matmul_kernel.cpp:16-27 — per row: serial ascending dot product, then *scale */
void orc_matmul(const float* x, const float* W, float* y, int32_t rows, int32_t cols, float scale) {
    for (int r = 0; r < rows; ++r) {
        const float* w = W + (int64_t)r * cols;
        float sum = 0.0f;
        for (int j = 0; j < cols; ++j) sum += x[j] * w[j];
        y[r] = sum * scale;
    }
}

/* source/kernel/cpu/rope_kernel.cpp:8-17 — freq = 1/powf(theta, 2k/hd); angle = freq*pos (fp32) */
void orc_rope_cache(int32_t head_dim, int32_t max_len, float theta, float* sin_c, float* cos_c) {
    const int half = head_dim / 2;
    for (int i = 0; i < max_len; ++i) {
        for (int k = 0; k < half; ++k) {
            float freq = 1.0f / powf(theta, (float)(2 * k) / (float)head_dim);
            float val = freq * (float)i;
            cos_c[(int64_t)i * half + k] = cosf(val);
            sin_c[(int64_t)i * half + k] = sinf(val);
        }
    }
}

/* source/kernel/cpu/rope_kernel.cpp:27-38 — rotate-half pairs (j, j+hd/2) of every head, q then k.
 * The reference iterates k over q_dim as well (GQA over-run, Appendix D); here k stops at k_dim. */
void orc_rope(float* q, float* k, int32_t pos, const float* sin_c, const float* cos_c, int32_t q_dim,
              int32_t k_dim, int32_t head_dim) {
    const int half = head_dim / 2;
    const float* s = sin_c + (int64_t)pos * half;
    const float* c = cos_c + (int64_t)pos * half;
    for (int b = 0; b < q_dim; b += head_dim) {
        for (int j = 0; j < half; ++j) {
            float fci = s[j], fcr = c[j];
            float v0 = q[b + j], v1 = q[b + j + half];
            q[b + j] = v0 * fcr - v1 * fci;
            q[b + j + half] = v1 * fcr + v0 * fci;
            if (b < k_dim) {
                v0 = k[b + j];
                v1 = k[b + j + half];
                k[b + j] = v0 * fcr - v1 * fci;
                k[b + j + half] = v1 * fcr + v0 * fci;
            }
        }
    }
}

/* source/kernel/cpu/mha_kernel.cpp:7-20 — max, e=expf(x-max), serial sum, divide */
void orc_softmax(float* x, int32_t n) {
    float mx = x[0];
    for (int i = 1; i < n; ++i)
        if (mx < x[i]) mx = x[i]; /* std::max_element: first maximum */
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) {
        x[i] = expf(x[i] - mx);
        sum += x[i];
    }
    for (int i = 0; i < n; ++i) x[i] /= sum;
}

/* source/kernel/cpu/mha_kernel.cpp:40-76 — per head: scores via 1xhd GEMV with scale=1/sqrt(hd)
 * (matmul_kernel.cpp: (sum)*scale), softmax, zeroed output += s_t * V_t in ascending t. */
void orc_mha(const float* q, float* score, const float* kc, const float* vc, float* out, int32_t layer,
             int32_t pos, int32_t max_len, int32_t head_dim, int32_t heads, int32_t kv_heads) {
    const int64_t kv = (int64_t)kv_heads * head_dim;
    const int64_t layer_off = (int64_t)layer * max_len * kv;
    const int group = heads / kv_heads;
    const float scale = 1.0f / sqrtf((float)head_dim);
    for (int h = 0; h < heads; ++h) {
        float* sc = score + (int64_t)h * max_len;
        const float* qh = q + (int64_t)h * head_dim;
        const int64_t head_off = (int64_t)(h / group) * head_dim;
        for (int t = 0; t <= pos; ++t) {
            const float* kt = kc + layer_off + (int64_t)t * kv + head_off;
            float sum = 0.0f;
            for (int j = 0; j < head_dim; ++j) sum += qh[j] * kt[j];
            sc[t] = sum * scale;
        }
        orc_softmax(sc, pos + 1);
        float* oh = out + (int64_t)h * head_dim;
        for (int j = 0; j < head_dim; ++j) oh[j] = 0.0f;
        for (int t = 0; t <= pos; ++t) {
            const float* vt = vc + layer_off + (int64_t)t * kv + head_off;
            for (int j = 0; j < head_dim; ++j) oh[j] += sc[t] * vt[j];
        }
    }
}

/* source/kernel/cpu/add_kernel.cpp:10-13 — out = in1; out += 1.0f * in2 */
void orc_add(const float* a, const float* b, float* out, int32_t n) {
    for (int i = 0; i < n; ++i) out[i] = a[i] + 1.0f * b[i];
}

/* source/kernel/cpu/swiglu_kernel.cpp:10-14 — sigmoid(gate) * up (NOT silu) */
void orc_swiglu(const float* up, const float* gate, float* out, int32_t n) {
    for (int i = 0; i < n; ++i) {
        float t = 1.0f / (1.0f + expf(-gate[i]));
        out[i] = t * up[i];
    }
}

/* source/op/argmax.cpp:11 — std::max_element: first maximal element */
int32_t orc_argmax(const float* logits, int32_t n) {
    int32_t best = 0;
    for (int32_t i = 1; i < n; ++i)
        if (logits[best] < logits[i]) best = i;
    return best;
}

/* ----------------------------------------------------------------------------------------- model ---- */

struct orc_model {
    syn_shape s;
    const float* blob;
    /* weight segment pointers (source/model/model.cpp:340-468) */
    const float *emb, *norms, *wq, *wk, *wv, *wo, *up, *gate, *down;
    /* activations, named after ModelBufferType (include/model/model.h:14-34) */
    float *key_cache, *value_cache, *emb_output, *rms_output, *query, *score, *mha_output, *att_output,
        *ffn_input, *up_output, *gate_output, *swi_output, *ffn_output, *model_pred, *sin_cache, *cos_cache;
    int threads, kv_bf16;
};

typedef struct {
    const float *x, *W;
    float* y;
    int32_t r0, r1, cols;
} mm_job;

static void* mm_thread(void* p) {
    mm_job* j = (mm_job*)p;
    orc_matmul(j->x, j->W + (int64_t)j->r0 * j->cols, j->y + j->r0, j->r1 - j->r0, j->cols, 1.0f);
    return 0;
}

static void model_matmul(const orc_model* m, const float* x, const float* W, float* y, int32_t rows,
                         int32_t cols) {
    int nt = m->threads;
    if (nt <= 1 || (int64_t)rows * cols < (1 << 18)) {
        orc_matmul(x, W, y, rows, cols, 1.0f);
        return;
    }
    if (nt > 64) nt = 64;
    pthread_t th[64];
    mm_job jobs[64];
    int32_t per = (rows + nt - 1) / nt;
    int n = 0;
    for (int k = 0; k < nt; ++k) {
        int32_t r0 = k * per, r1 = r0 + per > rows ? rows : r0 + per;
        if (r0 >= r1) break;
        jobs[n] = (mm_job){x, W, y, r0, r1, cols};
        pthread_create(&th[n], 0, mm_thread, &jobs[n]);
        n++;
    }
    for (int k = 0; k < n; ++k) pthread_join(th[k], 0);
}

orc_model* orc_create(const syn_shape* shape, const float* blob) {
    orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
    m->s = *shape;
    m->blob = blob;
    m->threads = 1;
    int64_t off[9];
    for (int t = 0; t < 9; ++t) syn_segment(shape, t, &off[t], 0, 0, 0, 0);
    m->emb = blob + off[0]; m->norms = blob + off[1]; m->wq = blob + off[2]; m->wk = blob + off[3];
    m->wv = blob + off[4]; m->wo = blob + off[5]; m->up = blob + off[6]; m->gate = blob + off[7];
    m->down = blob + off[8];
    const int64_t d = shape->hidden, I = shape->inter, S = shape->max_len, kv = shape->kv_hidden,
                  L = shape->layers, V = shape->vocab, hd = shape->head_dim, H = shape->heads;
#define ORC_ALLOC(n) ((float*)calloc((size_t)(n), sizeof(float)))
    m->key_cache = ORC_ALLOC(L * S * kv); m->value_cache = ORC_ALLOC(L * S * kv);
    m->emb_output = ORC_ALLOC(d); m->rms_output = ORC_ALLOC(d); m->query = ORC_ALLOC(d);
    m->score = ORC_ALLOC((H > hd ? H : hd) * S); m->mha_output = ORC_ALLOC(d); m->att_output = ORC_ALLOC(d);
    m->ffn_input = ORC_ALLOC(d); m->up_output = ORC_ALLOC(I); m->gate_output = ORC_ALLOC(I);
    m->swi_output = ORC_ALLOC(I); m->ffn_output = ORC_ALLOC(d); m->model_pred = ORC_ALLOC(V);
    m->sin_cache = ORC_ALLOC(S * (hd / 2)); m->cos_cache = ORC_ALLOC(S * (hd / 2));
#undef ORC_ALLOC
    orc_rope_cache((int32_t)hd, (int32_t)S, shape->theta, m->sin_cache, m->cos_cache); /* model.cpp:312-316 */
    return m;
}

void orc_destroy(orc_model* m) {
    if (!m) return;
    float* bufs[] = {m->key_cache, m->value_cache, m->emb_output, m->rms_output, m->query, m->score,
                     m->mha_output, m->att_output, m->ffn_input, m->up_output, m->gate_output,
                     m->swi_output, m->ffn_output, m->model_pred, m->sin_cache, m->cos_cache};
    for (unsigned i = 0; i < sizeof(bufs) / sizeof(bufs[0]); ++i) free(bufs[i]);
    free(m);
}

void orc_set_threads(orc_model* m, int n) { m->threads = n < 1 ? 1 : n; }
void orc_set_kv_bf16(orc_model* m, int on) { m->kv_bf16 = on; }

/* source/model/model.cpp:40-140 — the one-token layer loop */
void orc_forward(orc_model* m, int32_t token, int32_t pos, float* logits) {
    const syn_shape* s = &m->s;
    const int32_t d = s->hidden, I = s->inter, kv = s->kv_hidden, L = s->layers, V = s->vocab;
    float* x = m->emb_output;
    orc_embedding(token, m->emb, x, V, d);                                             /* :48  */
    for (int l = 0; l < L; ++l) {
        orc_rmsnorm(x, m->norms + (int64_t)(2 * l) * d, m->rms_output, d, s->eps);     /* :52  */
        /* slice_KV_cache (source/memory/tensor.cpp:199-212): row (l*S + pos) of each cache */
        float* krow = m->key_cache + ((int64_t)l * s->max_len + pos) * kv;             /* :54  */
        float* vrow = m->value_cache + ((int64_t)l * s->max_len + pos) * kv;
        model_matmul(m, m->rms_output, m->wq + (int64_t)l * d * d, m->query, d, d);    /* :58  */
        model_matmul(m, m->rms_output, m->wk + (int64_t)l * kv * d, krow, kv, d);      /* :60  */
        model_matmul(m, m->rms_output, m->wv + (int64_t)l * kv * d, vrow, kv, d);      /* :62  */
        orc_rope(m->query, krow, pos, m->sin_cache, m->cos_cache, d, kv, s->head_dim); /* :66  */
        if (m->kv_bf16) {
            for (int j = 0; j < kv; ++j) { krow[j] = syn_round_bf16(krow[j]); vrow[j] = syn_round_bf16(vrow[j]); }
        }
        orc_mha(m->query, m->score, m->key_cache, m->value_cache, m->mha_output, l, pos, s->max_len,
                s->head_dim, s->heads, s->kv_heads);                                   /* :70-78 */
        model_matmul(m, m->mha_output, m->wo + (int64_t)l * d * d, m->att_output, d, d);   /* :80  */
        orc_add(x, m->att_output, m->ffn_input, d);                                    /* :86  */
        orc_rmsnorm(m->ffn_input, m->norms + (int64_t)(2 * l + 1) * d, m->rms_output, d, s->eps); /* :93 */
        model_matmul(m, m->rms_output, m->up + (int64_t)l * I * d, m->up_output, I, d);    /* :99  */
        model_matmul(m, m->rms_output, m->gate + (int64_t)l * I * d, m->gate_output, I, d); /* :105 */
        orc_swiglu(m->up_output, m->gate_output, m->swi_output, I);                    /* :111 */
        model_matmul(m, m->swi_output, m->down + (int64_t)l * d * I, m->ffn_output, d, I); /* :118 */
        orc_add(m->ffn_output, m->ffn_input, x, d);                                    /* :124 */
    }
    orc_rmsnorm(x, m->norms + (int64_t)(2 * L) * d, m->rms_output, d, s->eps);         /* :131 */
    model_matmul(m, m->rms_output, m->emb, m->model_pred, V, d);                       /* :136, tied :350-352 */
    if (logits) memcpy(logits, m->model_pred, sizeof(float) * (size_t)V);
}

/* source/model/model.cpp:148-185 — prompt ids fed one at a time (logits ignored), then argmax feedback */
int32_t orc_greedy(orc_model* m, const int32_t* prompt, int32_t n_prompt, int32_t n_total, int32_t* out,
                   float* last_logits) {
    int32_t pos = 0, n = 0, token = prompt[0];
    while (pos < n_total - 1) {
        orc_forward(m, token, pos, 0);
        if (pos < n_prompt - 1) {
            pos++;
            token = prompt[pos];
        } else {
            pos++;
            token = orc_argmax(m->model_pred, m->s.vocab);
        }
        out[n++] = token;
    }
    if (last_logits) memcpy(last_logits, m->model_pred, sizeof(float) * (size_t)m->s.vocab);
    return n;
}

void orc_read(orc_model* m, int32_t buffer_id, int64_t offset, int64_t n, float* out) {
    const float* src = 0;
    switch (buffer_id) {
        case 2: src = m->key_cache; break;   case 3: src = m->value_cache; break;
        case 4: src = m->emb_output; break;  case 5: src = m->rms_output; break;
        case 6: src = m->query; break;       case 7: src = m->score; break;
        case 8: src = m->mha_output; break;  case 9: src = m->att_output; break;
        case 10: src = m->ffn_input; break;  case 11: src = m->up_output; break;
        case 12: src = m->gate_output; break; case 14: src = m->swi_output; break;
        case 15: src = m->ffn_output; break; case 16: src = m->model_pred; break;
        case 17: src = m->sin_cache; break;  case 18: src = m->cos_cache; break;
        default: break;
    }
    if (src) memcpy(out, src + offset, sizeof(float) * (size_t)n);
}

/* test aid: overwrite part of a buffer (same ids as orc_read) — used to inject a synthetic KV history so that a full-depth
 * model can be checked at a late position without running every earlier position through the CPU path */
void orc_write(orc_model* m, int32_t buffer_id, int64_t offset, int64_t n, const float* src) {
    float* dst = 0;
    switch (buffer_id) {
        case 2: dst = m->key_cache; break;
        case 3: dst = m->value_cache; break;
        default: break;
    }
    if (dst) memcpy(dst + offset, src, sizeof(float) * (size_t)n);
}
