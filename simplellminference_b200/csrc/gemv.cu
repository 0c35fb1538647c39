// csrc/gemv.cu — sllm_gemv: the plain (unfused) decode GEMV behind kernel::matmul_kernel_cuda.
#include "gemv_core.cuh"

namespace sllm {

struct PlainPolicy {
    const float* x;
    const void* W_;
    const float* sc;
    float* y;
    int grp, nrows, ncols;
    float scale;
    __device__ int units() const { return (nrows + 1) >> 1; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const { r0 = 2 * (int64_t)u; r1 = min(2 * u + 1, nrows - 1); }
    __device__ const void* W() const { return W_; }
    __device__ const float* scales() const { return sc; }
    __device__ int group() const { return grp; }
    __device__ int cols() const { return ncols; }
    template <int WD> __device__ void stage_t(float* xs, float*) const { stage_x_plain<WD>(xs, x, ncols); }
    __device__ void emit(int u, float s0, float s1) const {
        y[2 * u] = s0 * scale;  // (sum) * scale, matmul_kernel.cpp:26
        if (2 * u + 1 < nrows) y[2 * u + 1] = s1 * scale;
    }
};

template <int WD>
struct PlainPolicyT : PlainPolicy {
    __device__ void stage(float* xs, float* red) const { this->template stage_t<WD>(xs, red); }
};

template <int WD>
__global__ void __launch_bounds__(kGemvThreads) gemv_plain_kernel(PlainPolicyT<WD> pol) { gemv_body<WD>(pol); }

template <int WD>
static int launch_plain(const PlainPolicy& p, cudaStream_t st) {
    const size_t smem = gemv_smem_bytes(p.ncols);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        SLLM_REQUIRE(smem <= (size_t)smem_optin_bytes(), SLLM_ENOTSUP, "gemv: cols=%d does not fit shared memory", p.ncols);
        SLLM_CUDA(cudaFuncSetAttribute(gemv_plain_kernel<WD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    PlainPolicyT<WD> pt;
    static_cast<PlainPolicy&>(pt) = p;
    const int grid = gemv_grid((p.nrows + 1) / 2, 2);
    gemv_plain_kernel<WD><<<grid, kGemvThreads, smem, st>>>(pt);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

int check_gemv_args(const void* W, int w_dtype, const float* scales, int group, int rows, int cols) {
    SLLM_REQUIRE(W && rows > 0 && cols > 0, SLLM_EINVAL, "gemv: null weight or empty shape");
    SLLM_REQUIRE(((uintptr_t)W & 15) == 0, SLLM_EINVAL, "gemv: weight pointer must be 16-byte aligned");
    const int E = w_dtype == SLLM_F32 ? 4 : w_dtype == SLLM_BF16 ? 8 : w_dtype == SLLM_INT8 ? 16 : 0;
    SLLM_REQUIRE(E != 0, SLLM_EINVAL, "gemv: unknown weight dtype %d", w_dtype);
    SLLM_REQUIRE(cols % E == 0, SLLM_ENOTSUP, "gemv: cols=%d must be a multiple of %d for this weight type", cols, E);
    if (w_dtype == SLLM_INT8)
        SLLM_REQUIRE(scales && group >= 16 && group % 16 == 0 && cols % group == 0, SLLM_EINVAL,
                     "gemv: int8 needs scales, group %% 16 == 0 and cols %% group == 0 (cols=%d group=%d)", cols, group);
    return SLLM_OK;
}

}  // namespace sllm

using namespace sllm;

extern "C" int sllm_gemv(const float* x, const void* W, int32_t w_dtype, const float* scales, int32_t group, float* y,
                         int32_t rows, int32_t cols, float scale, sllm_stream_t stream) {
    SLLM_REQUIRE(x && y, SLLM_EINVAL, "gemv: null vector");
    SLLM_REQUIRE(((uintptr_t)x & 15) == 0, SLLM_EINVAL, "gemv: x must be 16-byte aligned");
    if (int rc = check_gemv_args(W, w_dtype, scales, group, rows, cols)) return rc;
    PlainPolicy p{x, W, scales, y, group, rows, cols, scale};
    switch (w_dtype) {
        case SLLM_F32: return launch_plain<SLLM_F32>(p, as_stream(stream));
        case SLLM_BF16: return launch_plain<SLLM_BF16>(p, as_stream(stream));
        default: return launch_plain<SLLM_INT8>(p, as_stream(stream));
    }
}
