// csrc/common.cuh — shared device/host helpers for libsllm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/sllm_b200.h"

namespace sllm {

// ------------------------------------------------------------------------------------- errors --------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SLLM_CUDA(call)                                                      \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) return ::sllm::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define SLLM_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            ::sllm::set_error(__VA_ARGS__); \
            return (code);                 \
        }                                  \
    } while (0)

// Every launcher ends with this: a launch-configuration error surfaces here, not at the next sync.
#define SLLM_LAUNCH_CHECK() SLLM_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(sllm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();             // cached cudaDevAttrMultiProcessorCount of the current device
int smem_optin_bytes();     // cached cudaDevAttrMaxSharedMemoryPerBlockOptin
extern thread_local int64_t g_launches;   // launches issued by this thread (gpu_launches accounting)

// Launch attribute holder: programmatic dependent launch on/off (griddepcontrol in the kernels is a no-op
// when the attribute is absent).
struct LaunchCfg {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl) {
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
    }
};

// ------------------------------------------------------------------------------------- device --------
#ifdef __CUDACC__

constexpr int kWarp = 32;

// PDL: let the next kernel in the stream start its prologue / wait for the previous kernel's results.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// 128-bit streaming load of weights: read-only path, do not pollute L1 (each byte is used exactly once).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum; `red` = shared scratch of >= 33 floats; every thread gets the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? red[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    float r = red[32];
    __syncthreads();
    return r;
}

// bf16 <-> fp32 by bit manipulation (RNE), identical to oracle/synth_weights.c:syn_round_bf16
__device__ __forceinline__ float bf16_lo(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float x) {
    uint32_t u = __float_as_uint(x);
    if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

#endif  // __CUDACC__

}  // namespace sllm
