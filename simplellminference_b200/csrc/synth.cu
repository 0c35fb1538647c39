// csrc/synth.cu — device side of the synthetic-weight contract (oracle/synth_weights.h describes it; the CPU
// oracle and this file must produce bit-identical values: integer hash, one __fmul_rn, one __fadd_rn) and
// the fp32 -> {bf16, int8-group} weight conversion used when a real fp32 blob is loaded.
#include <cmath>

#include "common.cuh"

namespace sllm {

__host__ __device__ __forceinline__ uint64_t sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float synth_value(uint64_t stream, int64_t i, float mean, float c) {
    uint64_t h = sm64(stream + (uint64_t)i);
    int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) + (int32_t)(h >> 48);
    return __fadd_rn(mean, __fmul_rn((float)(s - 131070), c));
}

// symmetric int8 group quantisation of G values held one-per-lane-slot; identical arithmetic to
// oracle/synth_weights.c: scale = amax/127 (IEEE divide), q = rint(w/scale).
__device__ __forceinline__ float quant_scale(float amax) {
    float s = __fdiv_rn(amax, 127.0f);
    return (s > 0.0f) ? s : 1.0f;
}
__device__ __forceinline__ int8_t quant_q(float w, float scale) {
    float r = rintf(__fdiv_rn(w, scale));
    r = fminf(fmaxf(r, -127.0f), 127.0f);
    return (int8_t)r;
}

// One thread per group of `G` consecutive elements of a destination row (G = group for int8, 8 otherwise).
template <int WD, bool FROM_F32>
__global__ void fill_kernel(uint64_t stream, float mean, float c, const float* __restrict__ src, int64_t src_first_row,
                            int64_t n_rows, int64_t src_row_len, int64_t src_col0, int64_t row_len, void* __restrict__ dst,
                            float* __restrict__ scales, int G) {
    const int64_t groups_per_row = row_len / G;
    const int64_t total = n_rows * groups_per_row;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = g / groups_per_row, gc = g - r * groups_per_row;
        const int64_t src_idx = (src_first_row + r) * src_row_len + src_col0 + gc * G;
        const int64_t dst_idx = r * row_len + gc * G;
        auto val = [&](int j) -> float {
            if (FROM_F32) return src[src_idx + j];
            return synth_value(stream, src_idx + j, mean, c);
        };
        if (WD == SLLM_INT8) {
            float amax = 0.0f;
            for (int j = 0; j < G; ++j) amax = fmaxf(amax, fabsf(val(j)));
            const float scale = quant_scale(amax);
            scales[g] = scale;
            int8_t* q = reinterpret_cast<int8_t*>(dst) + dst_idx;
            for (int j = 0; j < G; ++j) q[j] = quant_q(val(j), scale);
        } else if (WD == SLLM_BF16) {
            uint16_t* o = reinterpret_cast<uint16_t*>(dst) + dst_idx;
            for (int j = 0; j < G; ++j) o[j] = f32_to_bf16_bits(val(j));
        } else {
            float* o = reinterpret_cast<float*>(dst) + dst_idx;
            for (int j = 0; j < G; ++j) o[j] = val(j);
        }
    }
}

static void segment_params(const sllm_shape* s, int t, int64_t* row_len, float* mean, float* c) {
    const int64_t rows[9] = {s->hidden, s->hidden, s->hidden, s->hidden, s->hidden, s->hidden, s->hidden, s->hidden, s->inter};
    const double std_ = (t == 0) ? 1.0 : (t == 1) ? 0.02 : 4.0 / std::sqrt((double)rows[t]);
    *row_len = rows[t];
    *mean = (t == 1) ? 1.0f : 0.0f;
    *c = (float)(std_ * (1.7320508075688772 / 65535.0));
}

template <bool FROM_F32>
static int launch_fill(uint64_t stream_key, float mean, float c, const float* src, int64_t src_first_row, int64_t n_rows,
                       int64_t src_row_len, int64_t src_col0, int64_t row_len, void* dst, int w_dtype, float* scales,
                       int group, cudaStream_t st) {
    SLLM_REQUIRE(dst != nullptr && n_rows >= 0 && row_len > 0, SLLM_EINVAL, "synth/convert: bad destination");
    if (n_rows == 0) return SLLM_OK;
    int G = 8;
    if (w_dtype == SLLM_INT8) {
        SLLM_REQUIRE(scales != nullptr && group > 0 && group % 16 == 0 && row_len % group == 0 && src_col0 % group == 0,
                     SLLM_EINVAL, "int8 weights need scales, group %% 16 == 0 and row_len %% group == 0 (row_len=%lld group=%d)",
                     (long long)row_len, group);
        G = group;
    } else {
        SLLM_REQUIRE(row_len % 8 == 0 || w_dtype == SLLM_F32, SLLM_EINVAL, "row_len %% 8 != 0");
        if (row_len % 8 != 0) G = (row_len % 4 == 0) ? 4 : 1;
    }
    const int64_t total = n_rows * (row_len / G);
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, (int64_t)sm_count() * 16);
    switch (w_dtype) {
        case SLLM_F32: fill_kernel<SLLM_F32, FROM_F32><<<blocks, threads, 0, st>>>(stream_key, mean, c, src, src_first_row, n_rows, src_row_len, src_col0, row_len, dst, scales, G); break;
        case SLLM_BF16: fill_kernel<SLLM_BF16, FROM_F32><<<blocks, threads, 0, st>>>(stream_key, mean, c, src, src_first_row, n_rows, src_row_len, src_col0, row_len, dst, scales, G); break;
        case SLLM_INT8: fill_kernel<SLLM_INT8, FROM_F32><<<blocks, threads, 0, st>>>(stream_key, mean, c, src, src_first_row, n_rows, src_row_len, src_col0, row_len, dst, scales, G); break;
        default: SLLM_REQUIRE(false, SLLM_EINVAL, "unknown weight dtype %d", w_dtype);
    }
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

}  // namespace sllm

extern "C" {

int sllm_synth_fill(const sllm_shape* shape, uint64_t seed, int32_t segment, int64_t src_first_row, int64_t n_rows,
                    int64_t src_row_len, int64_t src_col0, int64_t row_len, void* dst, int32_t w_dtype, float* scales,
                    int32_t group, sllm_stream_t stream) {
    SLLM_REQUIRE(shape && segment >= 0 && segment < 9, SLLM_EINVAL, "synth_fill: bad shape/segment");
    int64_t seg_row;
    float mean, c;
    sllm::segment_params(shape, segment, &seg_row, &mean, &c);
    if (segment == 1) w_dtype = SLLM_F32;  // norm vectors always stay fp32
    const uint64_t key = sllm::sm64(seed * 0x100000001B3ULL + (uint64_t)segment);
    return sllm::launch_fill<false>(key, mean, c, nullptr, src_first_row, n_rows, src_row_len, src_col0, row_len, dst,
                                    w_dtype, scales, group, sllm::as_stream(stream));
}

int sllm_convert_weights(const float* src, void* dst, int32_t w_dtype, float* scales, int32_t group, int64_t rows,
                         int64_t cols, sllm_stream_t stream) {
    SLLM_REQUIRE(src != nullptr, SLLM_EINVAL, "convert_weights: null source");
    return sllm::launch_fill<true>(0, 0.f, 0.f, src, 0, rows, cols, 0, cols, dst, w_dtype, scales, group,
                                   sllm::as_stream(stream));
}

}  // extern "C"
