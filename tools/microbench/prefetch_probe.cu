// tools/microbench/prefetch_probe.cu — does data asked for with cp.async.bulk.prefetch.L2 during a stall come back at the L2 rate?
// Every round each warp (148 CTAs x 16 warps, the decode megakernel's geometry) (1) optionally asks the L2 for its next N tiles of
// 4 KB from a 6 GB window that is never re-read, (2) idles for `gap` microseconds (a dependency stall), (3) streams those N tiles
// through its two-slot TMA ring. Only step (3) is timed (%globaltimer, max end - min start over the CTAs would need a grid sync, so
// every CTA reports its own time and the host takes mean and max). Cold HBM = 49 GB/s per SM; L2-resident = 140 GB/s per SM.
//   variant 0: no prefetch   1: cp.async.bulk.prefetch.L2   2: the same with an L2::evict_last cache hint
//   variant 3: prefetch.global.L2 of every 128-byte line by all lanes (the non-bulk instruction)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); std::exit(1); } \
    } while (0)

constexpr int kThreads = 512, kWarps = 16, kSlotBytes = 4096, kSlots = 2;
constexpr unsigned kSpinLimit = 1u << 26;
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(bar)), "r"(count)); }
__device__ __forceinline__ void mb_expect(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done; unsigned spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s_addr(bar)), "r"(parity) : "memory");
        if (!done && ++spins > kSpinLimit) __trap();
    } while (!done);
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_addr(dst)), "l"(src), "r"(bytes), "r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(kThreads, 1) prefetch_probe_kernel(const uint8_t* src, size_t window_bytes, int rounds, int ntiles, int gap_ns, int variant,
                                                                     unsigned long long* out /* [grid][2]: sum of stream ns, rounds */, float* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* ring = smem + 1024;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kWarps * kSlots) mb_init(bars + tid, 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint64_t* my_bar = bars + warp * kSlots;
    uint8_t* my_ring = ring + (size_t)warp * kSlots * kSlotBytes;
    const size_t nwin = window_bytes / kSlotBytes;
    const size_t stream = (size_t)blockIdx.x * kWarps + warp, nstreams = (size_t)gridDim.x * kWarps;
    uint64_t policy = 0;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    unsigned count = 0;
    unsigned long long total = 0;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        auto tile_ptr = [&](int t) { return src + ((stream + ((size_t)r * ntiles + t) * nstreams) % nwin) * kSlotBytes; };
        if (variant == 1 || variant == 2) {
            if (lane == 0)
                for (int t = 0; t < ntiles; ++t) {
                    if (variant == 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tile_ptr(t)), "r"(kSlotBytes) : "memory");
                    else asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(tile_ptr(t)), "r"(kSlotBytes), "l"(policy) : "memory");
                }
        } else if (variant == 3) {
            for (int t = 0; t < ntiles; ++t) asm volatile("prefetch.global.L2 [%0];" ::"l"(tile_ptr(t) + lane * 128) : "memory");
        }
        const unsigned long long g0 = gtime();
        while (gtime() - g0 < (unsigned long long)gap_ns) {}
        __syncthreads();
        const unsigned long long t0 = gtime();
        if (lane == 0)
            for (int s = 0; s < kSlots && s < ntiles; ++s) { mb_expect(my_bar + s, kSlotBytes); tma_g2s(my_ring + (size_t)s * kSlotBytes, tile_ptr(s), kSlotBytes, my_bar + s); }
        for (int t = 0; t < ntiles; ++t) {
            const int si = count & 1;
            mb_wait(my_bar + si, (count >> 1) & 1);
            const uint4 w = *reinterpret_cast<const uint4*>(my_ring + (size_t)si * kSlotBytes + lane * 16);
            acc += __uint_as_float((w.x ^ w.y ^ w.z ^ w.w) & 0x3fffffffu);
            count++;
            __syncwarp();
            if (lane == 0 && t + kSlots < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mb_expect(my_bar + si, kSlotBytes);
                tma_g2s(my_ring + (size_t)si * kSlotBytes, tile_ptr(t + kSlots), kSlotBytes, my_bar + si);
            }
        }
        __syncthreads();
        total += gtime() - t0;
    }
    if (tid == 0) { out[2 * blockIdx.x] = total; out[2 * blockIdx.x + 1] = (unsigned long long)rounds; }
    if (acc == 12345.678f) sink[0] = acc;
}

int main() {
    std::setvbuf(stdout, nullptr, _IOLBF, 0);
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t window = (size_t)6 << 30;
    uint8_t* src = nullptr; float* sink = nullptr; unsigned long long* out = nullptr;
    CK(cudaMalloc(&src, window)); CK(cudaMalloc(&sink, 256)); CK(cudaMalloc(&out, sizeof(unsigned long long) * 2 * sms));
    CK(cudaMemset(src, 0x3c, window));
    const size_t smem = 1024 + (size_t)kWarps * kSlots * kSlotBytes;
    CK(cudaFuncSetAttribute(prefetch_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rounds = 32;
    for (int ntiles : {4, 8}) {
        for (int gap_us : {0, 10, 30}) {
            for (int variant = 0; variant < 4; ++variant) {
                if (variant == 0 && gap_us > 0) continue;
                prefetch_probe_kernel<<<sms, kThreads, smem>>>(src, window, rounds, ntiles, gap_us * 1000, variant, out, sink);
                CK(cudaDeviceSynchronize());
                std::vector<unsigned long long> h(2 * sms);
                CK(cudaMemcpy(h.data(), out, sizeof(unsigned long long) * 2 * sms, cudaMemcpyDeviceToHost));
                double mean = 0, mx = 0;
                for (int c = 0; c < sms; ++c) { const double us = (double)h[2 * c] / rounds / 1e3; mean += us; mx = std::max(mx, us); }
                mean /= sms;
                const double mb = (double)sms * kWarps * ntiles * kSlotBytes / 1e6;
                std::printf("{\"probe\": \"prefetch\", \"variant\": %d, \"tiles_per_warp\": %d, \"mb_per_round\": %.1f, \"gap_us\": %d, \"stream_us_mean\": %.2f, \"stream_us_max\": %.2f, "
                            "\"gbs_per_sm_mean\": %.1f}\n", variant, ntiles, mb, gap_us, mean, mx, (double)kWarps * ntiles * kSlotBytes / (mean * 1e-6) / 1e9);
            }
        }
    }
    return 0;
}
