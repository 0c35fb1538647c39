// csrc/prefill_gemm.cu — the prefill GEMM on the 5th-generation tensor cores (sm_100a only).
//
//   C[T][N] = A[T][K] . W[N][K]^T      A = bf16 activations, W = bf16 weights, C = fp32 accumulators in TMEM
//
// One persistent CTA per SM, 192 threads, three roles (no role ever blocks another with __syncthreads):
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor of a 128 x 64 A box and this CTA's W box per k-block into a ring of
//                      128B-swizzled stages, completion counted on the stage's "full" mbarrier;
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma.kind::f16 (K = 16 each) per stage, both operands through shared-memory
//                      descriptors; tcgen05.commit frees the stage / publishes the accumulator;
//   warps 2-5          epilogue: tcgen05.ld 32 columns at a time (one accumulator row per thread), fused epilogue
//                      (RoPE + KV-cache write, sigmoid(gate)*up, residual add), while the issuer already fills the
//                      second accumulator (TMEM holds two of up to 256 columns).
// Two cluster shapes: one SM per 128 x BN tile (cta_group::1, prompts of <= 128 rows and the T = 1 classifier), or the two SMs of
// a TPC on one 256 x BN tile (cta_group::2: each CTA stages its own 128 rows of A and BN/2 rows of W, the leader issues the MMAs,
// its commits are multicast to both CTAs). BN is a run-time multiple of 32; the residual-epilogue GEMMs may split K (pf_plan).
// W is read straight from the megakernel's TILED decode layout (megakernel.cuh) through a 4-D tensor map
// {k inside a K slice, row inside a tile, K slice, tile row}: prefill and decode share one copy of the weights.
// Rows of the tiled layout are in unit order, so the two values an epilogue needs together (RoPE partners,
// (up, gate)) sit in ADJACENT accumulator columns of the same thread.
// What bounds it (profiles/r01_prefill_ncu_raw.md): the bytes an SM can ingest (~65 GB/s per SM through TMA),
// not the tensor pipe — hence the two-SM tile; L2 prefetch, TMA multicast across pairs and stream-K were measured and dropped.
#include <cuda.h>

#include <map>
#include <tuple>

#include "mega_common.cuh"
#include "prefill.cuh"

namespace sllm {

extern int g_tune_pf_bn, g_tune_pf_pair, g_tune_pf_pdl, g_tune_pf_ksplit;
constexpr int kPfBM = 128, kPfBK = 64, kPfThreads = 192;
constexpr uint32_t kPfABytes = kPfBM * kPfBK * 2;
// PAIR = true: two CTAs of a cluster (the two SMs of a TPC) work on ONE 256 x BN tile with tcgen05.mma.cta_group::2 — each
// CTA stages its own 128 rows of A and its own BN/2 rows of W, so shared-memory traffic (TMA writes + tensor-core reads) and
// L2->SM traffic per flop are 2/3 of the single-CTA 128 x BN tile's (measured bound of the single-CTA kernel: 55 % tensor
// pipe inside a tile, profiles/r01_prefill_ncu_raw.md).
// The N extent of a tile (BN) is a RUN-TIME multiple of 32 up to 256 (one tcgen05.mma takes any N that is a multiple of 16):
// the host picks the BN that fills the last wave best, e.g. 192 instead of 256 for a 512 x 12288 output on 74 SM pairs.
constexpr int kPfMaxStages = 12;
constexpr uint32_t kPfRingBytes = 192 * 1024;          // operand ring; stages = ring / (A box + this CTA's W box)
constexpr uint32_t kPfTmemCols = 512;                  // two accumulators of up to 256 fp32 columns
constexpr size_t kPfSmem = 1024 /*alignment slack*/ + kPfRingBytes + 512 /*barriers + tmem slot*/;

struct PfDev {           // kernel-side view of PfGemmArgs
    int32_t T, N, m_tiles, n_tiles, BN;
    int32_t ksplit;       // > 1 (residual / accumulate epilogues only): a tile's k-blocks are cut into ksplit work units whose partial sums are ADDED to
                          // the residual stream with red.global.add — fills the machine when a GEMM has fewer tiles than SM pairs
    int32_t nkb, kb_per_seg, seg_elems, R, tiled;
    int32_t epilogue;
    int32_t a_rows;       // rows of A one CTA stages per k-block: 128, or (one-SM kernel, T < 128) T rounded up to the 8-row swizzle period —
                          // the rest of the 128-row operand tile is never loaded (its accumulator rows are never stored)
    float* out; int32_t ld_out, n_valid;
    uint16_t* q_out; uint8_t *kc, *vc; int32_t kv_dtype, q_loc, kv_loc, hd, S, pos0;
    const float *sin_t, *cos_t;
    uint16_t* s_out; int32_t I_loc;
};

// ------------------------------------------------------------------------------------------ PTX -------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s_addr(dst)), "l"(tm), "r"(s_addr(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(s_addr(dst)), "l"(tm), "r"(s_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// cta_group::2 forms: the destination is this CTA's shared memory, the mbarrier may live in the peer CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s_addr(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(s_addr(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(const void* local, uint32_t cta) {   // shared::cluster address of `local` in CTA `cta`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(s_addr(local)), "r"(cta));
    return r;
}
__device__ __forceinline__ void mb_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_alloc_pair(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s_addr(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) { asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_addr(bar)) : "memory"); }

__device__ __forceinline__ void tc_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// all tcgen05.mma issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor: K-major operand tile, rows of 64 bf16 = 128 bytes, 128B swizzle (what TMA wrote),
// 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor: start>>4 | LBO<<16 | SBO<<32 | version 1<<46 | layout 2<<61)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* tile) {
    uint64_t d = (uint64_t)((s_addr(tile) & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor, kind::f16: D = fp32 (1<<4), A = B = bf16 (1<<7, 1<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { return (uint32_t)f32_to_bf16_bits(lo) | ((uint32_t)f32_to_bf16_bits(hi) << 16); }

// ------------------------------------------------------------------------------- fused epilogues ------
// v[0..31] = accumulator columns n0 .. n0+31 of token row t (n0 is a multiple of 32)
__device__ __forceinline__ void epi_store(const PfDev& p, int t, int n0, const uint32_t* v, bool add) {
    float* dst = p.out + (size_t)t * p.ld_out + n0;
    if (n0 + 32 <= p.n_valid && (p.ld_out & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float4 o = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            if (add) {
                const float4 r = *reinterpret_cast<const float4*>(dst + i);
                o = make_float4(r.x + o.x, r.y + o.y, r.z + o.z, r.w + o.w);   // residual + projection (add_kernel.cpp:10-13)
            }
            *reinterpret_cast<float4*>(dst + i) = o;
        }
    } else {
        for (int i = 0; i < 32; ++i)
            if (n0 + i < p.n_valid) dst[i] = (add ? dst[i] : 0.f) + __uint_as_float(v[i]);
    }
}

// split-K residual epilogue: out[t][n0..n0+31] += v, one vector reduction per 16 bytes (no return value, resolved in L2)
__device__ __forceinline__ void epi_red_add(const PfDev& p, int t, int n0, const uint32_t* v) {
    float* dst = p.out + (size_t)t * p.ld_out + n0;
    if (n0 + 32 <= p.n_valid && (p.ld_out & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                         "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3])) : "memory");
    } else {
        for (int i = 0; i < 32; ++i)
            if (n0 + i < p.n_valid) atomicAdd(dst + i, __uint_as_float(v[i]));
    }
}

__device__ __forceinline__ void epi_gateup(const PfDev& p, int t, int n0, const uint32_t* v) {
    const int u0 = n0 >> 1;
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float up = __uint_as_float(v[2 * i]), gate = __uint_as_float(v[2 * i + 1]);
        s[i] = (1.0f / (1.0f + expf(-gate))) * up;   // swiglu_kernel.cpp:12-13
    }
    uint16_t* dst = p.s_out + (size_t)t * p.I_loc + u0;
    if (u0 + 16 <= p.I_loc && (p.I_loc & 7) == 0) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]), pack_bf16x2(s[6], s[7]));
        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pack_bf16x2(s[8], s[9]), pack_bf16x2(s[10], s[11]), pack_bf16x2(s[12], s[13]), pack_bf16x2(s[14], s[15]));
    } else {
        for (int i = 0; i < 16; ++i)
            if (u0 + i < p.I_loc) dst[i] = f32_to_bf16_bits(s[i]);
    }
}

__device__ __forceinline__ void kv_store16(uint8_t* cache, int kv_dtype, size_t elem, const float* x) {   // 16 consecutive elements
    if (kv_dtype == SLLM_BF16) {
        uint4* d = reinterpret_cast<uint4*>(cache + elem * 2);
        d[0] = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
        d[1] = make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]), pack_bf16x2(x[14], x[15]));
    } else {
        float4* d = reinterpret_cast<float4*>(cache + elem * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    }
}
__device__ __forceinline__ void kv_store1(uint8_t* cache, int kv_dtype, size_t elem, float x) {
    if (kv_dtype == SLLM_BF16) reinterpret_cast<uint16_t*>(cache)[elem] = f32_to_bf16_bits(x);
    else reinterpret_cast<float*>(cache)[elem] = x;
}

__device__ __forceinline__ void epi_qkv(const PfDev& p, int t, int n0, const uint32_t* v) {
    const int half = p.hd >> 1, rope_units = (p.q_loc + p.kv_loc) >> 1, nunits = (p.q_loc + 2 * p.kv_loc) >> 1;
    const int u0 = n0 >> 1;
    const int pos = p.pos0 + t;
    if (u0 >= nunits) return;
    if ((half & 15) == 0 && u0 + 16 <= nunits) {
        // 16 units = 16 consecutive j of ONE head (half is a multiple of 16) or 32 consecutive V elements
        if (u0 < rope_units) {
            const int head = u0 / half, j0 = u0 - head * half;
            const float4* sp = reinterpret_cast<const float4*>(p.sin_t + (size_t)pos * half + j0);
            const float4* cp = reinterpret_cast<const float4*>(p.cos_t + (size_t)pos * half + j0);
            float lo[16], hi[16];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const float4 s4 = sp[i4], c4 = cp[i4];
                const float ss[4] = {s4.x, s4.y, s4.z, s4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = 4 * i4 + k;
                    const float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
                    lo[i] = a * cc[k] - b * ss[k];   // rope_kernel.cpp:36-37
                    hi[i] = b * cc[k] + a * ss[k];
                }
            }
            const int r0 = head * p.hd + j0;
            if (r0 < p.q_loc) {
                uint16_t* q = p.q_out + (size_t)t * p.q_loc + r0;
                uint4* a = reinterpret_cast<uint4*>(q);
                uint4* b = reinterpret_cast<uint4*>(q + half);
                a[0] = make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]), pack_bf16x2(lo[6], lo[7]));
                a[1] = make_uint4(pack_bf16x2(lo[8], lo[9]), pack_bf16x2(lo[10], lo[11]), pack_bf16x2(lo[12], lo[13]), pack_bf16x2(lo[14], lo[15]));
                b[0] = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]));
                b[1] = make_uint4(pack_bf16x2(hi[8], hi[9]), pack_bf16x2(hi[10], hi[11]), pack_bf16x2(hi[12], hi[13]), pack_bf16x2(hi[14], hi[15]));
            } else {
                const int kvh = (r0 - p.q_loc) / p.hd;
                const size_t e0 = ((size_t)kvh * p.S + pos) * p.hd + j0;
                kv_store16(p.kc, p.kv_dtype, e0, lo);
                kv_store16(p.kc, p.kv_dtype, e0 + half, hi);
            }
        } else {
            const int b2 = 2 * (u0 - rope_units);
            const int kvh = b2 / p.hd, j = b2 - kvh * p.hd;
            const size_t e0 = ((size_t)kvh * p.S + pos) * p.hd + j;
            float x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]);
            kv_store16(p.vc, p.kv_dtype, e0, x);
            kv_store16(p.vc, p.kv_dtype, e0 + 16, x + 16);
        }
        return;
    }
    for (int i = 0; i < 16; ++i) {   // generic shapes: unit by unit
        const int u = u0 + i;
        if (u >= nunits) break;
        const float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
        if (u < rope_units) {
            const int head = u / half, j = u - head * half;
            const float s = p.sin_t[(size_t)pos * half + j], c = p.cos_t[(size_t)pos * half + j];
            const float o0 = a * c - b * s, o1 = b * c + a * s;
            const int r0 = head * p.hd + j;
            if (r0 < p.q_loc) {
                p.q_out[(size_t)t * p.q_loc + r0] = f32_to_bf16_bits(o0);
                p.q_out[(size_t)t * p.q_loc + r0 + half] = f32_to_bf16_bits(o1);
            } else {
                const int kvh = (r0 - p.q_loc) / p.hd;
                const size_t e0 = ((size_t)kvh * p.S + pos) * p.hd + j;
                kv_store1(p.kc, p.kv_dtype, e0, o0);
                kv_store1(p.kc, p.kv_dtype, e0 + half, o1);
            }
        } else {
            const int b2 = 2 * (u - rope_units);
            const int kvh = b2 / p.hd, j = b2 - kvh * p.hd;
            const size_t e0 = ((size_t)kvh * p.S + pos) * p.hd + j;
            kv_store1(p.vc, p.kv_dtype, e0, a);
            kv_store1(p.vc, p.kv_dtype, e0 + 1, b);
        }
    }
}

// ------------------------------------------------------------------------------------------ kernel ----
template <bool PAIR>
__global__ void __launch_bounds__(kPfThreads, 1) pf_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                const PfDev p) {
    constexpr int BM = PAIR ? 2 * kPfBM : kPfBM;          // rows of the (pair's) tile
    const int BN = p.BN;
    const int b_rows = PAIR ? BN / 2 : BN;                 // W rows staged by one CTA
    const uint32_t b_bytes = (uint32_t)b_rows * kPfBK * 2, stage_bytes = kPfABytes + b_bytes;
    const uint32_t tx_bytes = (uint32_t)p.a_rows * kPfBK * 2 + b_bytes;   // bytes the TMA really delivers per stage (a_rows < 128: skinny A)
    const int ST = min((int)(kPfRingBytes / stage_bytes), kPfMaxStages);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader: issues the MMAs of the pair
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ uint8_t pf_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pf_smem_raw) + 1023) & ~(uintptr_t)1023);   // SW128 tiles: 1024-byte aligned
    uint8_t* sA = smem;                                    // [ST] 128 x 64 bf16, 16 KB each
    uint8_t* sB = smem + (size_t)ST * kPfABytes;           // [ST] b_rows x 64 bf16 (multiples of 1024 bytes: b_rows % 8 == 0)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kPfRingBytes);
    uint64_t* empty = full + kPfMaxStages;
    uint64_t* tfull = empty + kPfMaxStages;     // [2] accumulator ready for the epilogue
    uint64_t* tempty = tfull + 2;     // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < ST; ++s) { mb_init(full + s, 1); mb_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mb_init(tfull + a, 1); mb_init(tempty + a, PAIR ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { if (PAIR) tc_alloc_pair(tmem_slot, kPfTmemCols); else tc_alloc(tmem_slot, kPfTmemCols); }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();   // pair: the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nunits = p.m_tiles * p.n_tiles * p.ksplit;
    // programmatic dependent launch: everything above overlapped the previous kernel's tail; its results (A, the residual
    // stream) are visible after the wait. The next kernel may start ITS set-up as soon as every CTA here is past this point.
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = worker; unit < nunits; unit += nworkers) {
                const int tile = unit / p.ksplit, part = unit - tile * p.ksplit;
                const int kb0 = (p.nkb * part) / p.ksplit, kb1 = (p.nkb * (part + 1)) / p.ksplit;
                const int m_tile = tile % p.m_tiles, n_tile = tile / p.m_tiles;   // consecutive workers share the weight tile (L2)
                const int row0 = m_tile * BM + (int)rank * kPfBM;                  // this CTA's 128 rows of A
                const int n0 = n_tile * BN + (int)rank * b_rows;                   // this CTA's rows of W
                for (int kb = kb0; kb < kb1; ++kb) {
                    mb_wait(empty + stage, phase ^ 1);
                    const int ks = kb / p.kb_per_seg, kk = kb - ks * p.kb_per_seg;
                    uint8_t* a_dst = sA + (size_t)stage * kPfABytes;
                    uint8_t* b_dst = sB + (size_t)stage * b_bytes;
                    if (PAIR) {
                        // both CTAs' bytes are counted on the LEADER's barrier (it is the leader that issues the MMAs)
                        if (rank == 0) mb_expect(full + stage, 2 * tx_bytes);
                        const uint32_t bar = mapa_u32(full + stage, 0);
                        tma_load_2d_pair(a_dst, &tmA, ks * p.seg_elems + kk * kPfBK, row0, bar);
                        if (p.tiled) tma_load_4d_pair(b_dst, &tmB, kk * kPfBK, 0, ks, n0 / p.R, bar);
                        else tma_load_4d_pair(b_dst, &tmB, kk * kPfBK, n0, 0, 0, bar);
                    } else {
                        mb_expect(full + stage, tx_bytes);
                        tma_load_2d(a_dst, &tmA, ks * p.seg_elems + kk * kPfBK, row0, full + stage);
                        if (p.tiled) tma_load_4d(b_dst, &tmB, kk * kPfBK, 0, ks, n0 / p.R, full + stage);
                        else tma_load_4d(b_dst, &tmB, kk * kPfBK, n0, 0, 0, full + stage);
                    }
                    if (++stage == ST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {   // ===== MMA issuer (pair: the leader CTA only) =====
            const uint32_t idesc = umma_idesc_bf16(BM, BN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int unit = worker; unit < nunits; unit += nworkers) {
                const int part = unit % p.ksplit;
                const int kb0 = (p.nkb * part) / p.ksplit, kb1 = (p.nkb * (part + 1)) / p.ksplit;
                mb_wait(tempty + acc, acc_phase ^ 1);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mb_wait(full + stage, phase);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_sw128(sA + (size_t)stage * kPfABytes);
                    const uint64_t bd = umma_desc_sw128(sB + (size_t)stage * b_bytes);
#pragma unroll
                    for (int k = 0; k < kPfBK / 16; ++k) {   // +32 bytes along K inside the 128-byte swizzle atom = +2 in the address field
                        if (PAIR) tc_mma_bf16_pair(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0);
                        else tc_mma_bf16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0);
                    }
                    if (PAIR) tc_commit_pair(empty + stage); else tc_commit(empty + stage);   // stage reusable once these MMAs have read it
                    if (++stage == ST) { stage = 0; phase ^= 1; }
                }
                if (PAIR) tc_commit_pair(tfull + acc); else tc_commit(tfull + acc);           // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {               // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t tempty_leader[2] = {PAIR ? mapa_u32(tempty, 0) : 0u, PAIR ? mapa_u32(tempty + 1, 0) : 0u};
        for (int unit = worker; unit < nunits; unit += nworkers) {
            const int tile = unit / p.ksplit;
            const int m_tile = tile % p.m_tiles, n_tile = tile / p.m_tiles;
            const int t = m_tile * BM + (int)rank * kPfBM + row;
            mb_wait(tfull + acc, acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256 + c0), v);
                const int n0 = n_tile * BN + c0;
                if (t < p.T && n0 < p.N) {
                    if (p.epilogue == PF_EPI_QKV) epi_qkv(p, t, n0, v);
                    else if (p.epilogue == PF_EPI_GATEUP) epi_gateup(p, t, n0, v);
                    else if (p.ksplit > 1) epi_red_add(p, t, n0, v);
                    else epi_store(p, t, n0, v, p.epilogue == PF_EPI_RESID);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mb_arrive_cluster(tempty_leader[acc]); else mb_arrive(tempty + acc); }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();   // pair: nobody leaves while the peer may still signal its barriers / read its operands
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tc_dealloc_pair(tmem_base, kPfTmemCols); else tc_dealloc(tmem_base, kPfTmemCols);
    }
}

// -------------------------------------------------------------------------------------------- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct PfCache {
    std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> weights;   // (W, N, K, tiled, rows per box)
};
PfCache* pf_cache_create() { return new PfCache(); }
void pf_cache_destroy(PfCache* c) { delete c; }

static int encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    SLLM_REQUIRE(fn, SLLM_ESTATE, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, ones,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SLLM_REQUIRE(r == CUDA_SUCCESS, SLLM_EINVAL, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu box %u %u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return SLLM_OK;
}

const char* pf_unsupported_reason(int w_dtype, int hd, int d, int q_loc, int I_loc) {
    if (w_dtype != SLLM_BF16) return "batched prefill needs bf16 weights (tensor-core operands)";
    if (hd != 64 && hd != 128) return "batched prefill attention is instantiated for head_dim 64 and 128";
    if (d % 8 || q_loc % 8 || I_loc % 8) return "row lengths must be multiples of 8 elements (16-byte TMA strides)";
    return nullptr;
}

template <bool PAIR>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const PfDev& p, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        SLLM_CUDA(cudaFuncSetAttribute(pf_gemm_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPfSmem));
        configured = true;
    }
    const int workers = std::min(PAIR ? sm_count() / 2 : sm_count(), p.m_tiles * p.n_tiles * p.ksplit);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(PAIR ? 2 * workers : workers);
    cfg.blockDim = dim3(kPfThreads);
    cfg.dynamicSmemBytes = kPfSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_tune_pf_pdl ? 2 : 1;
    SLLM_CUDA(cudaLaunchKernelEx(&cfg, pf_gemm_kernel<PAIR>, tmA, tmB, p));
    g_launches++;
    return SLLM_OK;
}

// Host-side plan of one GEMM: cluster shape, N extent of a tile, K split, k-blocks. Pure arithmetic (no device access beyond the
// cached SM count), exposed as sllm_prefill_gemm_plan for tests.
struct PfPlan { int pair, bn, ksplit, m_tiles, n_tiles, nkb, kb_per_seg, seg_elems, R; TileGeom g; };
static PfPlan pf_plan(int T, int N, int K, int tiled, int epilogue, int bn_forced) {
    PfPlan pl{};
    // more than one 128-row block of tokens: two SMs share a 256-row tile (tcgen05 cta_group::2)
    pl.pair = g_tune_pf_pair >= 0 ? (g_tune_pf_pair != 0) : (T > kPfBM);
    const int BM = pl.pair ? 2 * kPfBM : kPfBM;
    pl.m_tiles = (T + BM - 1) / BM;
    if (tiled) {
        pl.g = mega_tile_geom(N, K, SLLM_BF16);
        pl.seg_elems = pl.g.SC * 8; pl.R = pl.g.R;
        pl.kb_per_seg = (pl.seg_elems + kPfBK - 1) / kPfBK;
        pl.nkb = pl.g.KS * pl.kb_per_seg;
    } else {
        pl.seg_elems = K; pl.R = 1;
        pl.kb_per_seg = (K + kPfBK - 1) / kPfBK;
        pl.nkb = pl.kb_per_seg;
    }
    // N extent of a tile (a multiple of 32 up to 256) and, for the residual epilogue, a split of K: a work unit costs about
    // k-blocks x operand bytes per k-block (the kernel is bound by what an SM can ingest) plus the epilogue of its tile, and the
    // GEMM takes ceil(units / workers) waves of it. Pick the cheapest combination.
    int bn = bn_forced ? bn_forced : g_tune_pf_bn;
    int ksplit = 1;
    if (bn < 32 || bn > 256 || bn % 32) {
        const int workers = pl.pair ? sm_count() / 2 : sm_count();
        const bool adds = epilogue == PF_EPI_RESID || epilogue == PF_EPI_ACCUM;   // the epilogues whose partial sums may meet in L2
        const int max_split = (adds && g_tune_pf_ksplit != 0) ? (g_tune_pf_ksplit > 0 ? g_tune_pf_ksplit : 4) : 1;
        double best = 1e30;
        for (int b = 256; b >= 64; b -= 32) {
            for (int sp = 1; sp <= max_split; ++sp) {
                if (sp > 1 && pl.nkb / sp < 8) break;   // keep at least 8 k-blocks per work unit
                const long units = (long)pl.m_tiles * ((N + b - 1) / b) * sp;
                const int a_rows = (!pl.pair && T < kPfBM) ? (T + 7) / 8 * 8 : kPfBM;   // operand rows a CTA really ingests per k-block
                const double unit_cost = (double)((pl.nkb + sp - 1) / sp) * (a_rows + (pl.pair ? b / 2 : b)) + 10.0 * b;
                const double c = (double)((units + workers - 1) / workers) * unit_cost;
                if (c < best - 1e-9) { best = c; bn = b; ksplit = sp; }
            }
        }
    } else if ((epilogue == PF_EPI_RESID || epilogue == PF_EPI_ACCUM) && g_tune_pf_ksplit > 0 && pl.nkb / g_tune_pf_ksplit >= 1) {
        ksplit = g_tune_pf_ksplit;
    }
    pl.bn = bn; pl.ksplit = ksplit;
    pl.n_tiles = (N + bn - 1) / bn;
    return pl;
}

int pf_gemm(PfCache* cache, const PfGemmArgs& a, cudaStream_t st) {
    SLLM_REQUIRE(cache && a.A && a.W && a.T > 0 && a.N > 0 && a.K > 0, SLLM_EINVAL, "pf_gemm: bad argument");
    SLLM_REQUIRE(a.K % 8 == 0, SLLM_ENOTSUP, "pf_gemm: K=%d must be a multiple of 8", a.K);
    const PfPlan pl = pf_plan(a.T, a.N, a.K, a.tiled, a.epilogue, a.bn);
    const bool pair = pl.pair != 0;
    const int bn = pl.bn, ksplit = pl.ksplit;
    const TileGeom g = pl.g;
    PfDev p{};
    p.T = a.T; p.N = a.N; p.tiled = a.tiled; p.epilogue = a.epilogue;
    p.m_tiles = pl.m_tiles; p.seg_elems = pl.seg_elems; p.R = pl.R; p.kb_per_seg = pl.kb_per_seg; p.nkb = pl.nkb;
    p.ksplit = ksplit;
    p.BN = bn;
    p.n_tiles = pl.n_tiles;
    p.out = a.out; p.ld_out = a.ld_out; p.n_valid = a.n_valid;
    p.q_out = a.q_out; p.kc = a.k_cache; p.vc = a.v_cache; p.kv_dtype = a.kv_dtype; p.q_loc = a.q_loc; p.kv_loc = a.kv_loc; p.hd = a.hd; p.S = a.S;
    p.pos0 = a.pos0; p.sin_t = a.sin_t; p.cos_t = a.cos_t; p.s_out = a.s_out; p.I_loc = a.I_loc;

    CUtensorMap tmA;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.T};
        const cuuint64_t strides[1] = {(cuuint64_t)a.K * 2};
        p.a_rows = (!pair && a.T < kPfBM) ? (a.T + 7) / 8 * 8 : kPfBM;
        const cuuint32_t box[2] = {kPfBK, (cuuint32_t)p.a_rows};
        if (int rc = encode(&tmA, a.A, 2, dims, strides, box)) return rc;
    }
    const int box_rows = pair ? bn / 2 : bn;   // W rows one CTA stages per k-block
    const auto key = std::make_tuple(a.W, a.N, a.K, a.tiled, box_rows);
    auto it = cache->weights.find(key);
    if (it == cache->weights.end()) {
        CUtensorMap tmB;
        if (a.tiled) {
            const cuuint64_t dims[4] = {(cuuint64_t)p.seg_elems, (cuuint64_t)g.R, (cuuint64_t)g.KS, (cuuint64_t)g.ntr};
            const cuuint64_t strides[3] = {(cuuint64_t)g.SC * 16, (cuuint64_t)g.tile_bytes, (cuuint64_t)g.KS * g.tile_bytes};
            const cuuint32_t box[4] = {kPfBK, (cuuint32_t)g.R, 1, (cuuint32_t)(box_rows / g.R)};
            if (int rc = encode(&tmB, a.W, 4, dims, strides, box)) return rc;
        } else {
            const cuuint64_t dims[4] = {(cuuint64_t)a.K, (cuuint64_t)a.N, 1, 1};
            const cuuint64_t strides[3] = {(cuuint64_t)a.K * 2, (cuuint64_t)a.K * 2 * a.N, (cuuint64_t)a.K * 2 * a.N};
            const cuuint32_t box[4] = {kPfBK, (cuuint32_t)box_rows, 1, 1};
            if (int rc = encode(&tmB, a.W, 4, dims, strides, box)) return rc;
        }
        it = cache->weights.emplace(key, tmB).first;
    }
    return pair ? launch<true>(tmA, it->second, p, st) : launch<false>(tmA, it->second, p, st);
}

}  // namespace sllm

extern "C" int sllm_prefill_gemm_bf16(const void* A, const void* W, float* C, int32_t T, int32_t N, int32_t K, int32_t bn, sllm_stream_t stream) {
    using namespace sllm;
    SLLM_REQUIRE(A && W && C, SLLM_EINVAL, "null argument");
    static PfCache* cache = pf_cache_create();
    cache->weights.clear();   // callers may reuse addresses for different matrices
    PfGemmArgs a{};
    a.A = A; a.W = W; a.T = T; a.N = N; a.K = K; a.tiled = 0; a.epilogue = PF_EPI_STORE; a.out = C; a.ld_out = N; a.n_valid = N; a.bn = bn;
    return pf_gemm(cache, a, as_stream(stream));
}

extern "C" int sllm_prefill_gemm_plan(int32_t T, int32_t N, int32_t K, int32_t residual_epilogue, int32_t* two_sm, int32_t* bn, int32_t* ksplit,
                                      int32_t* work_units) {
    using namespace sllm;
    SLLM_REQUIRE(T > 0 && N > 0 && K > 0 && K % 8 == 0, SLLM_EINVAL, "bad GEMM shape");
    const PfPlan pl = pf_plan(T, N, K, /*tiled=*/0, residual_epilogue ? PF_EPI_RESID : PF_EPI_STORE, 0);
    if (two_sm) *two_sm = pl.pair;
    if (bn) *bn = pl.bn;
    if (ksplit) *ksplit = pl.ksplit;
    if (work_units) *work_units = pl.m_tiles * pl.n_tiles * pl.ksplit;
    return SLLM_OK;
}
