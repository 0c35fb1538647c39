// csrc/mega_common.cuh — device helpers shared by the two persistent decode-step kernels
// (megakernel.cu: grid-barrier version; megakernel_ll.cu: barrier-free {value, epoch}-word version).
#pragma once
#include "megakernel.cuh"

namespace sllm {

constexpr int kMegaThreads = 512;
constexpr int kMegaWarps = kMegaThreads / 32;
constexpr int kSlotBytes = 4096;
constexpr int kSlots = 2;
constexpr int kCplMax = 4;        // max 16-byte chunks per lane per row slice (slice <= 2 KB)
constexpr int kRoundUnits = 128;  // units per partial-table round
constexpr int kAttTile = 64;
constexpr int kAttP = 3 * kAttTile;   // scores per query head the single-pass attention of megakernel2.cu keeps (two K/V stages + a tail of <= 64 rows)
constexpr unsigned kSpinLimit = 1u << 26;
#ifndef SLLM_BARRIER_POLL_COUNTER
#define SLLM_BARRIER_POLL_COUNTER 1   // measured on B200 (llama2-7b bf16): 350.4 tok/s vs 338.1 with the separate release flag
#endif
constexpr int kAttRecPad = 4;     // partial record = hd floats of O, then m, l (+2 pad: keeps float4 alignment)


struct MegaSmem {
    size_t bars, red, part, ring, att, total;
    size_t att_q, att_p, att_misc, att_k, att_v;
    int kv_stride;
    int att_tile;   // cache positions per K/V stage: 64, or 32 when a row is wider than 256 bytes (fp32 cache with 128-wide heads), so that
                    // two K and two V stages always fit in 64 KB next to the weight rings
};
__host__ __device__ inline int mega_att_tile(int hd, int kv_esz) { return hd * kv_esz > 256 ? kAttTile / 2 : kAttTile; }
__host__ __device__ inline int mega_kv_stride(int row_bytes) {
    int s = (row_bytes / 128) * 128 + 32;
    if (s < row_bytes) s += 128;
    return s;
}
__host__ __device__ inline MegaSmem mega_smem_layout(int hd, int g, int kv_esz) {
    MegaSmem L;
    size_t off = 0;
    L.bars = off; off += 512;
    L.red = off; off += 256;
    L.part = off; off += (size_t)kRoundUnits * 2 * kMegaWarps * 4;
    off = (off + 127) & ~(size_t)127;
    L.ring = off; off += (size_t)kMegaWarps * kSlots * kSlotBytes;
    L.att = off;
    L.att_q = off; off += (size_t)g * hd * 4;
    L.att_p = off; off += (size_t)g * kAttP * 4;
    L.att_misc = off; off += 256;
    off = (off + 127) & ~(size_t)127;
    L.kv_stride = hd * kv_esz;   // dense rows: a tile is one contiguous bulk copy
    L.att_tile = mega_att_tile(hd, kv_esz);
    L.att_k = off; off += (size_t)2 * L.att_tile * L.kv_stride;
    L.att_v = off; off += (size_t)2 * L.att_tile * L.kv_stride;
    L.total = off;
    return L;
}

// ------------------------------------------------------------------------------------- primitives ----
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// grid-wide barrier (cooperative launch guarantees co-residency). counter[0] = arrivals (monotonic across launches; wrap-safe
// compares). Default: thread 0 of every CTA arrives with ONE atom.add.release and then polls the counter itself with ld.acquire
// until the target — the last arriver needs no extra hop to publish anything (measured: 2.85 ms/token instead of 2.96 with the
// older scheme, kept under SLLM_BARRIER_POLL_COUNTER=0: fence + atomicAdd, the last arriver stores the epoch to counter[32] on
// its own 128-byte line, the others poll that line; relaxed polls + one acquire fence were slower (346 tok/s), a 40 ns back-off
// between polls made no difference). No trailing fence: every cross-CTA read in this kernel is an L2 access
// (ld.global.cg or TMA), never an L1-cached load.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
#if SLLM_BARRIER_POLL_COUNTER
        unsigned v;
        asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(v) : "l"(counter) : "memory");
        v += 1u;
        unsigned spins = 0;
        while ((int)(v - target) < 0) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (++spins > kSpinLimit) __trap();
        }
#else
        __threadfence();
        const unsigned prev = atomicAdd(counter, 1u);
        if (prev + 1u == target) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(counter + 32), "r"(target) : "memory");
        } else {
            unsigned spins = 0;
            while (true) {
                unsigned v;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter + 32) : "memory");
                if ((int)(v - target) >= 0) break;
                if (++spins > kSpinLimit) __trap();
            }
        }
#endif
    }
    __syncthreads();
}

__device__ __forceinline__ void unit_rows(const MegaParams& p, const PhaseDesc& ph, int u, int& r0, int& r1) {
    if (ph.kind == PH_QKV) {
        const int half = p.hd >> 1, rope_units = (p.q_loc + p.kv_loc) >> 1;
        if (u < rope_units) {
            const int head = u / half, j = u - head * half;
            r0 = head * p.hd + j;
            r1 = r0 + half;
        } else {
            r0 = p.q_loc + p.kv_loc + 2 * (u - rope_units);
            r1 = r0 + 1;
        }
    } else if (ph.kind == PH_GATEUP) {
        r0 = u;
        r1 = p.I_loc + u;
    } else {
        r0 = 2 * u;
        r1 = min(2 * u + 1, ph.nrows - 1);
    }
}

// balanced split of a phase's tile rows over the CTAs
__device__ __forceinline__ int cta_cut(const uint32_t* cum, int n, int cta, int ncta) {   // first of CTA `cta`'s share of n items
    return cum ? (int)(((uint64_t)n * cum[cta]) >> 24) : (int)(((int64_t)n * cta) / ncta);
}
__device__ __forceinline__ void cta_tiles(const PhaseDesc& ph, int cta, int ncta, int& g0, int& g1) {
    g0 = cta_cut(ph.cum, ph.ntr, cta, ncta);
    g1 = cta_cut(ph.cum, ph.ntr, cta + 1, ncta);
}

// the split a phase's tile rows get in the kernel that fuses the down projection into the gate_up phase: gate_up follows the split
// of the (transposed) down matrix, whose tile rows are kFuseJT units, so that a CTA multiplies exactly the values it produced
template <bool FUSE>
__device__ __forceinline__ void phase_tiles(const PhaseDesc& ph, int cta, int ncta, int& g0, int& g1) {
    if (FUSE && ph.kind == PH_GATEUP) {
        const int per = kFuseJT / (ph.R >> 1);        // gate_up tile rows (R / 2 units each) per down tile row
        const int ntr_e = ph.nunits / kFuseJT;
        g0 = cta_cut(ph.cum, ntr_e, cta, ncta) * per;
        g1 = cta_cut(ph.cum, ntr_e, cta + 1, ncta) * per;
    } else {
        cta_tiles(ph, cta, ncta, g0, g1);
    }
}

// tile rows [a, b) of the transposed down matrix that row group rg of RG takes out of a CTA's [g0, g1): contiguous shares
__device__ __forceinline__ void down_t_rows(int g0, int g1, int rg, int RG, int& a, int& b) {
    const int cnt = g1 - g0;
    a = g0 + (cnt * rg) / RG;
    b = g0 + (cnt * (rg + 1)) / RG;
}

// ---- L2 prefetch of the weight stream: built twice in round 2, measured, removed ------------------------------------------------------
// The rings hold 128 KB per SM; at a dependency point (grid barrier, attention, per-head counters) the consumers stall and HBM runs dry
// (with the dot products removed the step takes exactly as long, with the barriers removed it is 20 % shorter: r02_mega_debug_v1.jsonl).
// Asking the L2 for the tiles beyond the ring with cp.async.bulk.prefetch.L2 works IN ISOLATION (r02_prefetch_probe.jsonl: 39 MB asked
// for, 10 us of idling, then streamed through the rings in 2.1 us instead of 4.2 us = 120-130 GB/s per SM against 49-63 from HBM), and
// loses in the kernel in both forms tried: (1) a continuous look-ahead of 2-16 tiles per warp during streaming: -5 % to -30 %
// (r02_l2_lookahead_sweep.jsonl); (2) 2-12 tiles per warp asked for right before a grid barrier and / or before the attention chain:
// -4 % to -13 % (r02_l2_prefetch_at_stalls_sweep.jsonl). A CTA's "stall" is the time the slowest CTAs still stream (the bulk prefetches
// of the early ones take HBM bandwidth from exactly the copies the barrier is waiting for), and the attention / counter chain is a
// sequence of small latency-critical L2 round trips that queue behind 40 MB of prefetch traffic. HBM is the shared resource: anything
// that is not on the critical path and uses it lengthens the critical path.

// acc[e] += w_e * s for the E weights of one 16-byte chunk (fp32 / bf16 storage)
template <int WD>
__device__ __forceinline__ void axpy_chunk(const uint4 w, float s, float* acc) {
    if (WD == SLLM_F32) {
        acc[0] = fmaf(__uint_as_float(w.x), s, acc[0]); acc[1] = fmaf(__uint_as_float(w.y), s, acc[1]);
        acc[2] = fmaf(__uint_as_float(w.z), s, acc[2]); acc[3] = fmaf(__uint_as_float(w.w), s, acc[3]);
    } else {
        acc[0] = fmaf(bf16_lo(w.x), s, acc[0]); acc[1] = fmaf(bf16_hi(w.x), s, acc[1]);
        acc[2] = fmaf(bf16_lo(w.y), s, acc[2]); acc[3] = fmaf(bf16_hi(w.y), s, acc[3]);
        acc[4] = fmaf(bf16_lo(w.z), s, acc[4]); acc[5] = fmaf(bf16_hi(w.z), s, acc[5]);
        acc[6] = fmaf(bf16_lo(w.w), s, acc[6]); acc[7] = fmaf(bf16_hi(w.w), s, acc[7]);
    }
}
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int KVD> struct MKv;
template <> struct MKv<SLLM_F32> { static constexpr int ESZ = 4, VEC = 4; };
template <> struct MKv<SLLM_BF16> { static constexpr int ESZ = 2, VEC = 8; };

template <int KVD>
__device__ __forceinline__ void kv_unpack(const uint4 v, float* f) {
    if (KVD == SLLM_F32) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    } else {
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    }
}

// 16 bytes of weights (E elements) times E activations held in registers
template <int WD>
__device__ __forceinline__ float reg_dot(const uint4 w, const float* x, float acc) {
    if (WD == SLLM_F32) {
        acc = fmaf(__uint_as_float(w.x), x[0], acc); acc = fmaf(__uint_as_float(w.y), x[1], acc);
        acc = fmaf(__uint_as_float(w.z), x[2], acc); acc = fmaf(__uint_as_float(w.w), x[3], acc);
    } else if (WD == SLLM_INT8) {   // unscaled: the caller multiplies by the chunk's group scale
        acc = word_dot_s8(w.x, make_float4(x[0], x[1], x[2], x[3]), acc);
        acc = word_dot_s8(w.y, make_float4(x[4], x[5], x[6], x[7]), acc);
        acc = word_dot_s8(w.z, make_float4(x[8], x[9], x[10], x[11]), acc);
        acc = word_dot_s8(w.w, make_float4(x[12], x[13], x[14], x[15]), acc);
    } else {
        acc = fmaf(bf16_lo(w.x), x[0], acc); acc = fmaf(bf16_hi(w.x), x[1], acc);
        acc = fmaf(bf16_lo(w.y), x[2], acc); acc = fmaf(bf16_hi(w.y), x[3], acc);
        acc = fmaf(bf16_lo(w.z), x[4], acc); acc = fmaf(bf16_hi(w.z), x[5], acc);
        acc = fmaf(bf16_lo(w.w), x[6], acc); acc = fmaf(bf16_hi(w.w), x[7], acc);
    }
    return acc;
}


static __device__ __noinline__ void mega_timeout(const char* what) {
    printf("sllm mega: %s timed out (cta %d warp %d)\n", what, (int)blockIdx.x, (int)(threadIdx.x >> 5));
    __trap();
}
__device__ __forceinline__ void mb_wait_fast(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    unsigned spins = 0;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done) : "r"(s_addr(bar)), "r"(parity) : "memory");
        if (!done && ++spins > kSpinLimit) mega_timeout("mbarrier");
    } while (!done);
}


}  // namespace sllm
