// tools/microbench/consumer_probe.cu — how fast can an SM CONSUME weight tiles (the decode megakernel's inner loop) when the data is
// always there? 148 CTAs x 16 warps, per-warp rings of 4 KB slots filled by cp.async.bulk exactly as in csrc/megakernel.cu, over an
// L2-resident window (the consumer's own ceiling C) and over an HBM window (what the whole pipe delivers). If C is not well above
// the SM's share of HBM (49 GB/s), nothing that banks data ahead of a dependency stall (deeper rings, L2 prefetch) can pay.
//
//   mode 0  touch one word per lane (pure ingest)
//   mode 1  the megakernel's body: bf16 tile of 4 rows x 64 chunks, x in registers (2 chunks per lane), 64 FMA + 64 unpack per
//           lane, 6-shuffle butterfly, partial sums to shared memory
//   mode 2  the same arithmetic, no butterfly / store (FMA + unpack only)
//   mode 3  body of mode 1 with the unpack replaced by a single PRMT-free reinterpretation (NOT exact: upper bound if unpack were free)
// Output: one JSON line per (mode, slots, source).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_));        \
            std::exit(1);                                                                                     \
        }                                                                                                     \
    } while (0)

constexpr int kThreads = 512, kWarps = 16, kSlotBytes = 4096;
constexpr unsigned kSpinLimit = 1u << 26;

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(bar)), "r"(count)); }
__device__ __forceinline__ void mb_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    unsigned spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s_addr(bar)), "r"(parity) : "memory");
        if (!done && ++spins > kSpinLimit) __trap();
    } while (!done);
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_addr(dst)), "l"(src), "r"(bytes), "r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
template <int MODE>
__device__ __forceinline__ float dot8(const uint4 w, const float* x, float acc) {
    if (MODE == 3) {
        acc = fmaf(__uint_as_float(w.x), x[0], acc); acc = fmaf(__uint_as_float(w.x), x[1], acc);
        acc = fmaf(__uint_as_float(w.y), x[2], acc); acc = fmaf(__uint_as_float(w.y), x[3], acc);
        acc = fmaf(__uint_as_float(w.z), x[4], acc); acc = fmaf(__uint_as_float(w.z), x[5], acc);
        acc = fmaf(__uint_as_float(w.w), x[6], acc); acc = fmaf(__uint_as_float(w.w), x[7], acc);
        return acc;
    }
    acc = fmaf(bf16_lo(w.x), x[0], acc); acc = fmaf(bf16_hi(w.x), x[1], acc);
    acc = fmaf(bf16_lo(w.y), x[2], acc); acc = fmaf(bf16_hi(w.y), x[3], acc);
    acc = fmaf(bf16_lo(w.z), x[4], acc); acc = fmaf(bf16_hi(w.z), x[5], acc);
    acc = fmaf(bf16_lo(w.w), x[6], acc); acc = fmaf(bf16_hi(w.w), x[7], acc);
    return acc;
}

template <int SLOTS, int MODE>
__global__ void __launch_bounds__(kThreads, 1) consumer_kernel(const uint8_t* src, size_t window_bytes, int tiles_per_warp, float* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // [kWarps][SLOTS]
    float* part = reinterpret_cast<float*>(smem + 1024);                // [128][2][16]
    uint8_t* ring = smem + 1024 + 16384;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kWarps * SLOTS) mb_init(bars + tid, 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint64_t* my_bar = bars + warp * SLOTS;
    uint8_t* my_ring = ring + (size_t)warp * SLOTS * kSlotBytes;
    const size_t ntiles_window = window_bytes / kSlotBytes;
    const size_t stream = (size_t)blockIdx.x * kWarps + warp, nstreams = (size_t)gridDim.x * kWarps;
    auto tile_ptr = [&](int t) { return src + ((stream + (size_t)t * nstreams) % ntiles_window) * kSlotBytes; };
    if (lane == 0)
        for (int s = 0; s < SLOTS && s < tiles_per_warp; ++s) {
            mb_expect(my_bar + s, kSlotBytes);
            tma_g2s(my_ring + (size_t)s * kSlotBytes, tile_ptr(s), kSlotBytes, my_bar + s);
        }
    float xr[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) xr[i][e] = 1.0f + 0.001f * (float)(lane + 32 * i + e);
    float total = 0.f;
    const int ks = warp & 7;
    for (int t = 0; t < tiles_per_warp; ++t) {
        const int si = t % SLOTS;
        mb_wait(my_bar + si, (uint32_t)((t / SLOTS) & 1));
        const uint8_t* sp = my_ring + (size_t)si * kSlotBytes + lane * 16;
        if (MODE == 0) {
            const uint4 w = *reinterpret_cast<const uint4*>(sp);
            total += __uint_as_float((w.x ^ w.y ^ w.z ^ w.w) & 0x3fffffffu);
        } else {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint8_t* q = sp + i * 512;
                a0 = dot8<MODE>(*reinterpret_cast<const uint4*>(q), xr[i], a0);
                a1 = dot8<MODE>(*reinterpret_cast<const uint4*>(q + 1024), xr[i], a1);
                a2 = dot8<MODE>(*reinterpret_cast<const uint4*>(q + 2048), xr[i], a2);
                a3 = dot8<MODE>(*reinterpret_cast<const uint4*>(q + 3072), xr[i], a3);
            }
            if (MODE == 2) {
                total += (a0 + a1) + (a2 + a3);
            } else {
                const bool hi = lane & 16;
                float k0 = hi ? a2 : a0, k1 = hi ? a3 : a1;
                k0 += __shfl_xor_sync(0xffffffffu, hi ? a0 : a2, 16);
                k1 += __shfl_xor_sync(0xffffffffu, hi ? a1 : a3, 16);
                const bool hi8 = lane & 8;
                float k = hi8 ? k1 : k0;
                k += __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
                k += __shfl_xor_sync(0xffffffffu, k, 4);
                k += __shfl_xor_sync(0xffffffffu, k, 2);
                k += __shfl_xor_sync(0xffffffffu, k, 1);
                const int r = (lane >> 4) * 2 + ((lane >> 3) & 1);
                if ((lane & 7) == 0) part[(((t & 63) * 2 + (r >> 1)) * 2 + (r & 1)) * kWarps + ks] = k;
            }
        }
        __syncwarp();
        if (lane == 0 && t + SLOTS < tiles_per_warp) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mb_expect(my_bar + si, kSlotBytes);
            tma_g2s(my_ring + (size_t)si * kSlotBytes, tile_ptr(t + SLOTS), kSlotBytes, my_bar + si);
        }
    }
    __syncthreads();
    if (MODE == 1 || MODE == 3) total += part[tid];
    if (total == 12345.678f) sink[0] = total;   // keeps the arithmetic alive
}

template <int SLOTS, int MODE>
static void run(const uint8_t* src, size_t window, const char* source, int sms, float* sink, cudaStream_t st) {
    const size_t smem = 1024 + 16384 + (size_t)kWarps * SLOTS * kSlotBytes;
    CK(cudaFuncSetAttribute(consumer_kernel<SLOTS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = 1024;
    for (int rep = 0; rep < 2; ++rep) {   // first pass warms (fills L2 for the small window)
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, st));
        consumer_kernel<SLOTS, MODE><<<sms, kThreads, smem, st>>>(src, window, tiles, sink);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 1) {
            const double bytes = (double)sms * kWarps * tiles * kSlotBytes;
            std::printf("{\"probe\": \"consumer\", \"mode\": %d, \"slots_per_warp\": %d, \"source\": \"%s\", \"gbs_total\": %.0f, \"gbs_per_sm\": %.1f}\n", MODE, SLOTS,
                        source, bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e9 / sms);
        }
        CK(cudaEventDestroy(e0));
        CK(cudaEventDestroy(e1));
    }
}

int main() {
    std::setvbuf(stdout, nullptr, _IOLBF, 0);
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t big = (size_t)4 << 30, small = (size_t)32 << 20;
    uint8_t* src = nullptr;
    float* sink = nullptr;
    CK(cudaMalloc(&src, big));
    CK(cudaMalloc(&sink, 256));
    CK(cudaMemset(src, 0x3c, big));
    for (int which = 0; which < 2; ++which) {
        const size_t w = which == 0 ? small : big;
        const char* s = which == 0 ? "L2-resident" : "HBM";
        run<2, 0>(src, w, s, sms, sink, st);
        run<2, 1>(src, w, s, sms, sink, st);
        run<2, 2>(src, w, s, sms, sink, st);
        run<2, 3>(src, w, s, sms, sink, st);
        run<3, 1>(src, w, s, sms, sink, st);
    }
    return 0;
}
