// csrc/runtime.cu — error reporting and device facts behind include/sllm_b200.h.
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace sllm {

static thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;
int g_tune_ctas_per_sm = 0;
int g_tune_pf_bn = 0;   // prefill GEMM: force the N tile (128 / 256), 0 = heuristic
int g_tune_pf_pdl = 1;     // prefill kernels: programmatic dependent launch (set-up overlaps the previous kernel's tail)
int g_tune_pf_ksplit = -1;  // prefill GEMM, residual epilogue: 0 = never split K, n > 0 = up to n parts, -1 = heuristic (up to 4)
int g_tune_batch_graph = 0;   // batched decode: 1 = replay one CUDA graph per live-slot count instead of the launch sequence (experimental)
int g_tune_batch_rows4 = 0;   // batched decode GEMV: 1 = four weight rows per warp at a time for >= 3 vectors (experimental)
int g_tune_batch_ksplit = 0;  // batched decode, with rows4: down projection with K cut in two over grid.y when more vectors are live than fit whole rows (experimental)
int g_tune_mega_debug = 0;   // decode megakernel measurement aid (MegaParams::debug): bit 0 = no grid barriers, bit 1 = no dot products; results are garbage
int g_tune_pf_pair = -1;   // prefill GEMM: 1 / 0 = force / forbid the two-SM (cta_group::2) kernel, -1 = heuristic

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return (int)e;
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int smem_optin_bytes() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (n <= 0) n = 48 * 1024;
    }
    return n;
}

}  // namespace sllm

extern "C" {

const char* sllm_last_error(void) { return sllm::g_err; }
int sllm_abi_version(void) { return 1; }

int sllm_tune(int32_t key, int32_t value) {
    switch (key) {
        case 0: sllm::g_tune_ctas_per_sm = value; return SLLM_OK;
        case 1: sllm::g_tune_pf_bn = value; return SLLM_OK;
        case 2: sllm::g_tune_pf_pair = value; return SLLM_OK;
        case 3: sllm::g_tune_pf_pdl = value; return SLLM_OK;
        case 4: sllm::g_tune_pf_ksplit = value; return SLLM_OK;
        case 5: sllm::g_tune_batch_graph = value; return SLLM_OK;
        case 6: sllm::g_tune_batch_rows4 = value; return SLLM_OK;
        case 7: sllm::g_tune_batch_ksplit = value; return SLLM_OK;
        case 8: sllm::g_tune_mega_debug = value; return SLLM_OK;
        default: sllm::set_error("unknown tunable %d", key); return SLLM_EINVAL;
    }
}

int sllm_device_info(int32_t* sms, int32_t* smem, size_t* total, size_t* free_) {
    int n = 0;
    SLLM_CUDA(cudaGetDeviceCount(&n));
    SLLM_REQUIRE(n > 0, SLLM_ESTATE, "no CUDA device visible: libsllm_b200 has no CPU fallback");
    if (sms) *sms = sllm::sm_count();
    if (smem) *smem = sllm::smem_optin_bytes();
    size_t f = 0, t = 0;
    SLLM_CUDA(cudaMemGetInfo(&f, &t));
    if (total) *total = t;
    if (free_) *free_ = f;
    return SLLM_OK;
}

}  // extern "C"
