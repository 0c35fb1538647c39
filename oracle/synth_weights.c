/* oracle/synth_weights.c — TEST INFRASTRUCTURE ONLY. See synth_weights.h for the contract. */
#include "synth_weights.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline uint64_t syn_stream(uint64_t seed, int t) { return sm64(seed * 0x100000001B3ULL + (uint64_t)t); }

static inline float syn_from_stream(uint64_t stream, int64_t i, float mean, float c) {
    uint64_t h = sm64(stream + (uint64_t)i);
    int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) +
                (int32_t)(h >> 48);
    float p = (float)(s - 131070) * c; /* one rounding */
    return mean + p;                   /* one rounding */
}

float syn_value(uint64_t seed, int t, int64_t i, float mean, float c) {
    return syn_from_stream(syn_stream(seed, t), i, mean, c);
}

float syn_round_bf16(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) { /* inf / nan: truncate, keep nan quiet */
        u &= 0xFFFF0000u;
    } else {
        u += 0x7FFFu + ((u >> 16) & 1u);
        u &= 0xFFFF0000u;
    }
    memcpy(&x, &u, 4);
    return x;
}

int64_t syn_blob_floats(const syn_shape* s) {
    int64_t off, cnt;
    syn_segment(s, 8, &off, &cnt, 0, 0, 0);
    return off + cnt;
}

void syn_segment(const syn_shape* s, int t, int64_t* offset, int64_t* count, int64_t* row_len, float* mean,
                 float* c) {
    const int64_t V = s->vocab, d = s->hidden, kv = s->kv_hidden, I = s->inter, L = s->layers;
    const int64_t counts[9] = {V * d, (2 * L + 1) * d, L * d * d, L * kv * d, L * kv * d,
                               L * d * d, L * I * d, L * I * d, L * d * I};
    const int64_t rows[9] = {d, d, d, d, d, d, d, d, I};
    int64_t off = 0;
    for (int k = 0; k < t; ++k) off += counts[k];
    if (offset) *offset = off;
    if (count) *count = counts[t];
    if (row_len) *row_len = rows[t];
    double std_ = (t == 0) ? 1.0 : (t == 1) ? 0.02 : 4.0 / sqrt((double)rows[t]);
    if (mean) *mean = (t == 1) ? 1.0f : 0.0f;
    if (c) *c = (float)(std_ * (1.7320508075688772 / 65535.0));
}

void syn_fill_segment_int8(const syn_shape* s, uint64_t seed, int t, int group, int64_t first, int64_t count,
                           int8_t* q, float* scales) {
    float mean, c;
    syn_segment(s, t, 0, 0, 0, &mean, &c);
    const uint64_t st = syn_stream(seed, t);
    for (int64_t g0 = 0; g0 < count; g0 += group) {
        float amax = 0.0f;
        for (int j = 0; j < group; ++j) {
            float a = fabsf(syn_from_stream(st, first + g0 + j, mean, c));
            if (a > amax) amax = a;
        }
        float scale = amax / 127.0f;
        if (!(scale > 0.0f)) scale = 1.0f;
        scales[g0 / group] = scale;
        for (int j = 0; j < group; ++j) {
            float w = syn_from_stream(st, first + g0 + j, mean, c);
            float r = rintf(w / scale); /* RNE, |r| <= 127 by construction */
            if (r > 127.0f) r = 127.0f;
            if (r < -127.0f) r = -127.0f;
            q[g0 + j] = (int8_t)r;
        }
    }
}

typedef struct {
    const syn_shape* s;
    uint64_t seed;
    int t, wdtype, group;
    int64_t first, count;
    float* out;
} fill_job;

static void fill_range(const fill_job* j) {
    float mean, c;
    syn_segment(j->s, j->t, 0, 0, 0, &mean, &c);
    const uint64_t st = syn_stream(j->seed, j->t);
    const int wd = (j->t == 1) ? SYN_F32 : j->wdtype; /* norm vectors stay fp32 */
    if (wd == SYN_INT8) {
        const int G = j->group;
        for (int64_t g0 = 0; g0 < j->count; g0 += G) {
            float amax = 0.0f;
            for (int k = 0; k < G; ++k) {
                float a = fabsf(syn_from_stream(st, j->first + g0 + k, mean, c));
                if (a > amax) amax = a;
            }
            float scale = amax / 127.0f;
            if (!(scale > 0.0f)) scale = 1.0f;
            for (int k = 0; k < G; ++k) {
                float w = syn_from_stream(st, j->first + g0 + k, mean, c);
                float r = rintf(w / scale);
                if (r > 127.0f) r = 127.0f;
                if (r < -127.0f) r = -127.0f;
                j->out[g0 + k] = r * scale; /* dequantised value the GPU uses */
            }
        }
    } else {
        for (int64_t i = 0; i < j->count; ++i) {
            float w = syn_from_stream(st, j->first + i, mean, c);
            j->out[i] = (wd == SYN_BF16) ? syn_round_bf16(w) : w;
        }
    }
}

static void* fill_thread(void* p) {
    fill_range((const fill_job*)p);
    return 0;
}

void syn_fill_segment(const syn_shape* s, uint64_t seed, int t, int wdtype, int group, int64_t first,
                      int64_t count, float* out, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (count < (int64_t)1 << 16) n_threads = 1;
    pthread_t th[64];
    fill_job jobs[64];
    int64_t G = (wdtype == SYN_INT8 && t != 1) ? group : 1;
    int64_t per = ((count / n_threads + G - 1) / G) * G;
    int started = 0;
    for (int k = 0; k < n_threads; ++k) {
        int64_t b = per * k, e = b + per;
        if (k == n_threads - 1 || e > count) e = count;
        if (b >= e) break;
        jobs[k] = (fill_job){s, seed, t, wdtype, group, first + b, e - b, out + b};
        if (n_threads == 1) {
            fill_range(&jobs[k]);
        } else {
            pthread_create(&th[k], 0, fill_thread, &jobs[k]);
            started++;
        }
        if (e == count) break;
    }
    for (int k = 0; k < started; ++k) pthread_join(th[k], 0);
}

void syn_fill_blob(const syn_shape* s, uint64_t seed, int wdtype, int group, float* blob, int n_threads) {
    for (int t = 0; t < 9; ++t) {
        int64_t off, cnt;
        syn_segment(s, t, &off, &cnt, 0, 0, 0);
        syn_fill_segment(s, seed, t, wdtype, group, 0, cnt, blob + off, n_threads);
    }
}
