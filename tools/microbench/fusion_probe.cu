// tools/microbench/fusion_probe.cu — three measurements that decide the next step of the decode megakernel
// before any of it is written. Stand-alone (no library): nvcc -gencode arch=compute_100a,code=sm_100a, run on one B200.
//
//   1. What does it cost to combine K-split partial vectors with red.global.add.v4.f32 instead of a grid barrier?
//      148 CTAs x 512 threads each add their 8 floats into the SAME d = 4096 floats (what a K-split down / wo projection fused
//      into its producer phase would do), followed by the megakernel's grid barrier; against the barrier alone and against
//      reductions into per-CTA private vectors (no contention). Per-iteration time from CUDA events over many iterations.
//   2. Do clusters of 2 / 4 / 8 CTAs co-reside with a cooperative, one-CTA-per-SM launch at the megakernel's shared-memory size
//      (the attention -> wo fusion needs the KV splits of a head in one cluster)? cudaOccupancyMaxActiveClusters.
//   3. How fast can an SM ingest TMA bulk copies from L2-resident data compared with HBM-resident data (is prefetching the next
//      phase's weights into L2 during a dependency gap worth anything)? 148 CTAs x 16 warps x 2 slots of 4 KB, the megakernel's
//      ring, over a 32 MB window (L2) and over a 4 GB window (HBM).
//
// Output: one JSON line per measurement on stdout.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <functional>
#include <initializer_list>
#include <vector>

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_));        \
            std::exit(1);                                                                                     \
        }                                                                                                     \
    } while (0)

constexpr int kThreads = 512;
constexpr unsigned kSpinLimit = 1u << 26;

// the megakernel's barrier (csrc/mega_common.cuh grid_barrier, poll-the-counter form)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target, unsigned limit = kSpinLimit) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned v;
        asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(v) : "l"(counter) : "memory");
        v += 1u;
        unsigned spins = 0;
        while ((int)(v - target) < 0) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (++spins > limit) __trap();
        }
    }
    __syncthreads();
}

// mode 0: barrier only; 1: every CTA reduces into the same d floats, then barrier; 2: every CTA reduces into its own d floats, then
// barrier; 3: like 1 but scalar red.global.add.f32 (8 per thread); 4: shared reductions without any barrier (issue rate only)
__global__ void __launch_bounds__(kThreads, 1) red_probe_kernel(float* x, float* priv, unsigned* counter, int d, int iters, int mode, unsigned base) {
    const int tid = threadIdx.x;
    const int per = d / kThreads;   // floats per thread (8 at d = 4096)
    float* dst = (mode == 2) ? priv + (size_t)blockIdx.x * d : x;
    unsigned idx = 0;
    for (int it = 0; it < iters; ++it) {
        const float v = 1.0f + (float)(it & 3);
        if (mode == 1 || mode == 2 || mode == 4) {
            for (int k = 0; k < per; k += 4) {
                float* p = dst + tid * per + k;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
            }
        } else if (mode == 3) {
            for (int k = 0; k < per; ++k) atomicAdd(dst + tid * per + k, v);
        }
        if (mode != 4) grid_barrier(counter, base + (++idx) * gridDim.x);
    }
}

// ---- clusters + cooperative launch: does a grid of clusters really co-reside (every CTA passes one grid barrier) and can a CTA
// add into its neighbour's shared memory (DSMEM) — the two things the attention -> wo fusion would lean on
__global__ void __launch_bounds__(kThreads, 1) cluster_probe_kernel(unsigned* counter, unsigned base, unsigned* ok_count) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    float* mine = reinterpret_cast<float*>(dyn_smem);
    unsigned rank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    if (threadIdx.x == 0) mine[0] = 0.f;
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {   // every CTA adds 1 into slot 0 of cluster rank 0's shared memory
        uint32_t local = (uint32_t)__cvta_generic_to_shared(mine), remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0u));
        asm volatile("red.shared::cluster.add.f32 [%0], %1;" ::"r"(remote), "f"(1.0f) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    grid_barrier(counter, base + gridDim.x, 1u << 21);   // co-residency of the whole grid (a grid that is not co-resident traps within seconds)
    if (threadIdx.x == 0 && rank == 0 && mine[0] == (float)csize) atomicAdd(ok_count, 1u);
}

// ---- TMA ingest probe ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    unsigned spins = 0;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done) : "r"(s_addr(bar)), "r"(parity) : "memory");
        if (!done && ++spins > kSpinLimit) __trap();
    } while (!done);
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s_addr(dst)), "l"(src), "r"(bytes), "r"(s_addr(bar)) : "memory");
}

constexpr int kWarps = kThreads / 32;
constexpr int kSlotBytes = 4096;

// Every warp streams `tiles_per_warp` tiles of 4 KB through SLOTS private slots; the tiles of the whole grid walk a window of
// `window_bytes` (wrapping), so a small window is L2-resident after the first pass and a large one never is. Each lane reads one
// 16-byte word of the landed tile (so that the data is really consumed) and the sum goes to `sink`.
template <int SLOTS>
__global__ void __launch_bounds__(kThreads, 1) ingest_probe_kernel(const uint8_t* src, size_t window_bytes, int tiles_per_warp, unsigned* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // [kWarps][SLOTS]
    uint8_t* ring = smem + 1024;
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
    unsigned acc = 0;
    for (int t = 0; t < tiles_per_warp; ++t) {
        const int si = t % SLOTS;
        mb_wait(my_bar + si, (uint32_t)((t / SLOTS) & 1));
        const uint4 w = *reinterpret_cast<const uint4*>(my_ring + (size_t)si * kSlotBytes + lane * 16);
        acc += w.x ^ w.y ^ w.z ^ w.w;
        __syncwarp();
        if (lane == 0 && t + SLOTS < tiles_per_warp) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mb_expect(my_bar + si, kSlotBytes);
            tma_g2s(my_ring + (size_t)si * kSlotBytes, tile_ptr(t + SLOTS), kSlotBytes, my_bar + si);
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;   // keeps the reads alive
}

static float time_ms(cudaStream_t st, const std::function<void()>& fn) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, st));
    fn();
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
    return ms;
}

static void coop_launch(const void* kernel, int grid, size_t smem, cudaStream_t st, void** args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelExC(&cfg, kernel, args));
}

int main(int argc, char** argv) {
    std::setvbuf(stdout, nullptr, _IOLBF, 0);   // a line is out as soon as it is printed, whatever happens later
    const int iters = argc > 1 ? std::atoi(argv[1]) : 2000;
    int dev = 0, sms = 0, smem_optin = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    std::printf("{\"probe\": \"device\", \"sms\": %d, \"smem_optin\": %d}\n", sms, smem_optin);

    // ---- 1. reductions against the barrier ------------------------------------------------------------------------------
    {
        int d = 4096;
        float *x = nullptr, *priv = nullptr;
        unsigned* counter = nullptr;
        CK(cudaMalloc(&x, sizeof(float) * d));
        CK(cudaMalloc(&priv, sizeof(float) * (size_t)d * sms));
        CK(cudaMalloc(&counter, 256));
        CK(cudaMemset(x, 0, sizeof(float) * d));
        CK(cudaMemset(priv, 0, sizeof(float) * (size_t)d * sms));
        CK(cudaMemset(counter, 0, 256));
        unsigned base = 0;
        const char* names[5] = {"barrier only", "red.v4 into one shared vector + barrier", "red.v4 into per-CTA vectors + barrier",
                                "scalar atomicAdd into one shared vector + barrier", "red.v4 into one shared vector, no barrier"};
        double t_us[5] = {0, 0, 0, 0, 0};
        for (int mode = 0; mode < 5; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
                int it = rep == 0 ? 50 : iters;
                void* args[] = {&x, &priv, &counter, &d, &it, &mode, &base};
                const float ms = time_ms(st, [&] { coop_launch((const void*)red_probe_kernel, sms, 0, st, args); });
                if (mode != 4) base += (unsigned)it * (unsigned)sms;
                if (rep == 1) t_us[mode] = 1e3 * ms / iters;
            }
            std::printf("{\"probe\": \"reduce\", \"mode\": \"%s\", \"ctas\": %d, \"floats\": %d, \"us_per_iteration\": %.3f}\n", names[mode], sms, d,
                        t_us[mode]);
        }
        std::printf("{\"probe\": \"reduce_summary\", \"extra_us_of_shared_red_v4_over_barrier\": %.3f, \"extra_us_private\": %.3f, "
                    "\"extra_us_scalar\": %.3f, \"what\": \"cost of combining %d partial vectors of %d floats by reductions, exposed at the barrier\"}\n",
                    t_us[1] - t_us[0], t_us[2] - t_us[0], t_us[3] - t_us[0], sms, d);
        // check the arithmetic of the vector reductions once (warm-up + timed iterations of modes 1, 3 and 4 hit x)
        std::vector<float> h(d);
        CK(cudaMemcpy(h.data(), x, sizeof(float) * d, cudaMemcpyDeviceToHost));
        bool same = true;
        for (int i = 1; i < d; ++i) same = same && (h[i] == h[0]);
        std::printf("{\"probe\": \"reduce_check\", \"all_elements_equal\": %s, \"value\": %.1f}\n", same ? "true" : "false", h[0]);
        cudaFree(x); cudaFree(priv); cudaFree(counter);
    }

    // ---- 2. clusters under a one-CTA-per-SM launch ------------------------------------------------------------------------
    for (int cs : {2, 4, 8}) {
        for (size_t smem : {(size_t)0, (size_t)200 * 1024}) {
            if (smem) CK(cudaFuncSetAttribute(ingest_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((sms / cs) * cs);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cs;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int nclusters = -1;
            const cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, ingest_probe_kernel<2>, &cfg);
            std::printf("{\"probe\": \"clusters\", \"cluster_size\": %d, \"dynamic_smem\": %zu, \"max_active_clusters\": %d, \"ctas\": %d, \"of_sms\": %d, "
                        "\"status\": \"%s\"}\n", cs, smem, nclusters, nclusters > 0 ? nclusters * cs : 0, sms, cudaGetErrorString(e));
            cudaGetLastError();
        }
    }

    // ---- 3. TMA ingest: L2-resident against HBM-resident source -----------------------------------------------------------
    {
        const size_t big = (size_t)4 << 30, small = (size_t)32 << 20;
        uint8_t* src = nullptr;
        unsigned* sink = nullptr;
        CK(cudaMalloc(&src, big));
        CK(cudaMalloc(&sink, 256));
        CK(cudaMemset(src, 1, big));
        const size_t smem2 = 1024 + (size_t)kWarps * 2 * kSlotBytes, smem3 = 1024 + (size_t)kWarps * 3 * kSlotBytes;
        CK(cudaFuncSetAttribute(ingest_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        CK(cudaFuncSetAttribute(ingest_probe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        for (int slots : {2, 3}) {
            for (int which = 0; which < 2; ++which) {
                const size_t window = which == 0 ? small : big;
                // bytes per launch = grid * 16 warps * tiles * 4 KB: 2048 tiles per warp = 19.9 GB over 148 SMs
                int tiles = 2048;
                const uint8_t* s = src;
                size_t w = window;
                void* args[] = {&s, &w, &tiles, &sink};
                const void* k = slots == 2 ? (const void*)ingest_probe_kernel<2> : (const void*)ingest_probe_kernel<3>;
                const size_t sm = slots == 2 ? smem2 : smem3;
                coop_launch(k, sms, sm, st, args);   // warm-up (fills L2 for the small window)
                CK(cudaStreamSynchronize(st));
                const float ms = time_ms(st, [&] { coop_launch(k, sms, sm, st, args); });
                const double bytes = (double)sms * kWarps * tiles * kSlotBytes;
                std::printf("{\"probe\": \"tma_ingest\", \"slots_per_warp\": %d, \"source\": \"%s\", \"window_mb\": %zu, \"gbs_total\": %.0f, "
                            "\"gbs_per_sm\": %.1f}\n", slots, which == 0 ? "L2-resident" : "HBM", window >> 20, bytes / (ms * 1e-3) / 1e9,
                            bytes / (ms * 1e-3) / 1e9 / sms);
            }
        }
        cudaFree(src); cudaFree(sink);
    }
    // ---- 4. clusters for real (last: a launch that cannot co-reside would end this process): a cooperative launch of clusters at the megakernel's shared-memory size; every CTA passes a grid
    // barrier and the cluster's CTAs add into rank 0's shared memory. ok = clusters whose sum came out right.
    for (int cs : {2, 4}) {
        unsigned *counter = nullptr, *okc = nullptr;
        CK(cudaMalloc(&counter, 256));
        CK(cudaMalloc(&okc, 256));
        CK(cudaMemset(counter, 0, 256));
        CK(cudaMemset(okc, 0, 256));
        const size_t smem = (size_t)200 * 1024;
        cudaError_t e = cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int nclusters = 0;
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3((sms / cs) * cs);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&nclusters, cluster_probe_kernel, &cfg);
        int grid = (e == cudaSuccess && nclusters > 0) ? std::min(nclusters, sms / cs) * cs : 0;
        unsigned okh = 0;
        const char* how = "not launched";
        if (grid > 0) {
            cfg.gridDim = dim3(grid);
            cfg.numAttrs = 2;
            unsigned base = 0;
            void* args[] = {&counter, &base, &okc};
            e = cudaLaunchKernelExC(&cfg, (const void*)cluster_probe_kernel, args);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            how = cudaGetErrorString(e);
            if (e == cudaSuccess) CK(cudaMemcpy(&okh, okc, 4, cudaMemcpyDeviceToHost));
        } else {
            how = cudaGetErrorString(e);
        }
        std::printf("{\"probe\": \"cluster_cooperative_launch\", \"cluster_size\": %d, \"grid\": %d, \"of_sms\": %d, \"status\": \"%s\", "
                    "\"clusters_with_correct_dsmem_sum\": %u, \"clusters\": %d}\n", cs, grid, sms, how, okh, grid / cs);
        cudaGetLastError();
        if (e != cudaSuccess) {   // a failed cooperative launch can leave a sticky error: stop here rather than report nonsense below
            std::printf("{\"probe\": \"aborted\", \"after\": \"cluster_cooperative_launch\"}\n");
            return 0;
        }
        cudaFree(counter); cudaFree(okc);
    }

    return 0;
}
