// csrc/megakernel_ll.cu — the decode step as ONE persistent kernel with NO grid barriers and the tensor-parallel
// all-reduce INSIDE it (sm_100a, NVLink peer memory).
//
// Same streaming machinery as megakernel.cu (per-warp TMA rings over tiled weights, K split across the 16 warps,
// x in registers), but nothing is exchanged through "store, grid barrier, load" any more. Every activation that
// crosses CTAs — q, the newest K/V row, attention partials, the partial sums of wo and down, sigmoid(gate)*up,
// the ranks' arg-max candidates — travels as an 8-byte word {fp32 bits, epoch}: one aligned 8-byte store is one
// transaction, so value and flag arrive together and a consumer simply spins on the words it needs (NCCL-LL
// style). Epoch = step * (L+2) + layer + 1 from the device-resident step counter; each buffer is written once per
// layer, so an exact epoch match identifies the data. Consequences:
//   * a phase boundary costs one L2 write + one L2 read instead of fence + atomic + poll + load (trace of the
//     barrier version: 14 us of 93 us per layer were grid barriers);
//   * the residual stream is REPLICATED in every CTA's shared memory (x and h never go to global memory): the
//     prologue of qkv / gate_up / classifier adds the incoming partial sums to its own copy, in rank order, so every
//     CTA of every rank holds bit-identical activations;
//   * tensor parallelism is the same code: the row-parallel phases (wo, down) store their partial sums into the
//     areas of ALL ranks (CUDA IPC mappings, NVLink stores), consumers add tp vectors instead of one — the
//     all-reduce costs no launch, no fence and no extra pass.
// Why single-buffered areas are safe: a buffer is rewritten one layer later, and in between lies at least one phase
// whose output every consumer waits for from EVERY CTA (the grid is sized so that each CTA owns >= 1 tile row of
// every weight phase), so no CTA — on any rank — can still be reading the old contents.
#include <algorithm>
#include <cmath>

#include "mega_common.cuh"

#ifndef SLLM_LL_SPLIT_ROWS
#define SLLM_LL_SPLIT_ROWS 128   // cache rows per attention split. A/B at 2 ranks on one box: 128 -> 505.8, 160 -> 498.1, 200 -> 500.4 tok/s
#endif
#ifndef SLLM_LL_KNS
#define SLLM_LL_KNS 2   // attention splits in flight per polling round trip of the wo prologue's merge. A/B on one 2-GPU box (make EXTRA=-DSLLM_LL_KNS=3):
                        // 3 in flight = one round trip less per layer but 60 more bytes of spills in the streaming loop: 481.8 vs 506.1 tok/s
#endif

namespace sllm {

struct MegaLLSmem {
    size_t bars, red, part, ring, resid, att_q, att_p, att_misc, att_k, att_v, total;
    int kv_stride;
    int att_tile;   // cache positions per K/V stage (mega_att_tile: 32 when a row is wider than 256 bytes, else 64)
};
__host__ __device__ inline MegaLLSmem mega_ll_smem_layout(int d, int hd, int g, int kv_esz) {
    MegaLLSmem L;
    size_t off = 0;
    L.bars = off; off += 512;
    L.red = off; off += 256;
    L.part = off; off += (size_t)kRoundUnits * 2 * kMegaWarps * 4;
    // the attention phase's q / probabilities / bookkeeping live IN the partial-sum table (16 KB), which only the weight phases
    // use: with the replicated residual stream (d floats) there is no room for both next to the rings and the K/V stages
    L.att_q = L.part;
    L.att_p = L.att_q + (size_t)g * hd * 4;
    L.att_misc = L.att_p + (size_t)g * kAttTile * 4;
    off = (off + 127) & ~(size_t)127;
    L.ring = off; off += (size_t)kMegaWarps * kSlots * kSlotBytes;
    L.resid = off; off += (size_t)d * 4;
    off = (off + 127) & ~(size_t)127;
    L.kv_stride = hd * kv_esz;
    L.att_tile = mega_att_tile(hd, kv_esz);
    L.att_k = off; off += (size_t)2 * L.att_tile * L.kv_stride;
    L.att_v = off; off += (size_t)2 * L.att_tile * L.kv_stride;
    L.total = off;
    return L;
}

// ---- {value, epoch} words ----------------------------------------------------------------------------------
__device__ __forceinline__ void ll_send(uint2* p, float v, unsigned epoch) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __noinline__ void ll_timeout() {
    printf("sllm mega-ll: a {value,epoch} word never arrived (cta %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
    __trap();
}
__device__ __forceinline__ float2 ll_recv2(const uint2* p, unsigned epoch) {   // two adjacent words (16 bytes)
    uint4 w;
    unsigned spins = 0;
    while (true) {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
        if (w.y == epoch && w.w == epoch) break;
        if (++spins > kSpinLimit) ll_timeout();
    }
    return make_float2(__uint_as_float(w.x), __uint_as_float(w.z));
}
__device__ __forceinline__ uint4 ll_load16(const uint2* p) {
    uint4 w;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p) : "memory");
    return w;
}
// four adjacent words: both 16-byte loads are in flight together
__device__ __forceinline__ float4 ll_recv4(const uint2* p, unsigned epoch) {
    unsigned spins = 0;
    while (true) {
        const uint4 a = ll_load16(p), b = ll_load16(p + 2);
        if (a.y == epoch && a.w == epoch && b.y == epoch && b.w == epoch)
            return make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
        if (++spins > kSpinLimit) ll_timeout();
    }
}
// N groups of four words at p + i*stride: all 2N loads are issued before the first flag is looked at
template <int N>
__device__ __forceinline__ void ll_recv4xN(const uint2* p, int64_t stride, int n, unsigned epoch, float4* out) {
    unsigned spins = 0;
    while (true) {
        uint4 a[N], b[N];
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < n) { a[i] = ll_load16(p + i * stride); b[i] = ll_load16(p + i * stride + 2); }
        bool ok = true;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < n) ok = ok && a[i].y == epoch && a[i].w == epoch && b[i].y == epoch && b[i].w == epoch;
        if (ok) {
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (i < n) out[i] = make_float4(__uint_as_float(a[i].x), __uint_as_float(a[i].z), __uint_as_float(b[i].x), __uint_as_float(b[i].z));
            return;
        }
        if (++spins > kSpinLimit) ll_timeout();
    }
}

// optional per-CTA timeline, same slots as megakernel.cu (tools/mega_trace.py --ll): event = phase in step order (5 per layer:
// qkv, attention, wo, gate_up, down; then the classifier), slots 0 start, 1 activation vector ready, 3 streaming done,
// 4 results sent
constexpr int kLLTraceEvents = 512;
__device__ __forceinline__ unsigned long long ll_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define LL_STAMP(ev, slot)                                                                                  \
    do {                                                                                                    \
        if (p.trace && threadIdx.x == 0 && (ev) < kLLTraceEvents) p.trace[((size_t)blockIdx.x * kLLTraceEvents + (ev)) * 8 + (slot)] = ll_gtime(); \
    } while (0)

template <int WD, int KVD, int G>
__global__ void __launch_bounds__(kMegaThreads, 1) mega_ll_kernel(const MegaLLParams p) {
    constexpr int E = WInfo<WD>::E;
    constexpr int CPL = (WD == SLLM_INT8) ? 2 : kCplMax;   // max chunks per lane per row slice (x of a lane: CPL * E = 32 registers)
    constexpr int KESZ = MKv<KVD>::ESZ, KVEC = MKv<KVD>::VEC;
    extern __shared__ __align__(128) uint8_t mega_ll_smem[];
    uint8_t* const smem = mega_ll_smem;
    const MegaLLSmem SL = mega_ll_smem_layout(p.d, p.hd, G, KESZ);
    const int AT = SL.att_tile;                      // cache positions per K/V stage
    uint64_t* ring_bar = reinterpret_cast<uint64_t*>(smem + SL.bars);
    uint64_t* att_bar = ring_bar + kMegaWarps * kSlots;
    float* red = reinterpret_cast<float*>(smem + SL.red);
    float* part = reinterpret_cast<float*>(smem + SL.part);
    uint8_t* ring = smem + SL.ring;
    float* resid_s = reinterpret_cast<float*>(smem + SL.resid);   // this CTA's copy of the residual stream
    float* xs = reinterpret_cast<float*>(smem + SL.att_k);        // activation staging for wo / down (aliases the idle K/V stages)
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int nwp = 4 * p.L + 1;

    if (tid < kMegaWarps * kSlots + 2) mb_init(ring_bar + tid, 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int pos = p.st->pos;
    const int token = min(max(p.st->token, 0), p.V - 1);
    // attention splits actually used at this position: ~128 cache rows per split (every split costs the wo prologue two
    // polling round trips in the merge, so 32 splits of 16 rows — what a tensor-parallel rank with 4 KV heads would get —
    // are far slower than 5 splits of 128). Same value in every CTA of every rank.
    const int nsplit = min(p.nsplit, max(1, (pos + 1 + SLLM_LL_SPLIT_ROWS - 1) / SLLM_LL_SPLIT_ROWS));
    const unsigned ebase = (unsigned)p.st->pad[1] * (unsigned)(p.L + 2) + 1u;   // epoch of layer l = ebase + l
    uint2* const my_area = p.area[p.rank];

    uint8_t* my_ring = ring + (size_t)warp * kSlots * kSlotBytes;
    uint64_t* my_bar = ring_bar + warp * kSlots;

    // ---------------- producer (lane 0 of every warp): one tile == one bulk copy -------------------------------
    int pr_wp = -1, pr_left = 0;
    unsigned pr_count = 0;
    const uint8_t* pr_ptr = nullptr;
    uint32_t pr_step = 0, pr_bytes = 0;
    auto produce_one = [&]() {
        while (pr_left <= 0) {
            if (++pr_wp >= nwp) { pr_wp = nwp; return; }
            const PhaseDesc ph = p.phases[pr_wp];
            const int ks = warp & (ph.KS - 1), rg = warp / ph.KS, RG = kMegaWarps / ph.KS;
            int g0, g1;
            cta_tiles(ph, cta, ncta, g0, g1);
            pr_left = (g1 - g0 - rg + RG - 1) / RG;
            pr_ptr = ph.W + ((size_t)(g0 + rg) * ph.KS + ks) * ph.tile_bytes;
            pr_step = (uint32_t)RG * ph.KS * ph.tile_bytes;
            pr_bytes = (uint32_t)ph.tile_bytes;
        }
        const int si = pr_count & (kSlots - 1);
        mb_expect(my_bar + si, pr_bytes);
        tma_g2s(my_ring + (size_t)si * kSlotBytes, pr_ptr, pr_bytes, my_bar + si);
        pr_ptr += pr_step;
        pr_left--;
        pr_count++;
    };
    if (lane == 0) {
#pragma unroll 1
        for (int s = 0; s < kSlots; ++s) produce_one();
    }
    unsigned cons_count = 0;
    unsigned kv_use0 = 0, kv_use1 = 0;

#pragma unroll 1
    for (int wp = 0; wp < nwp; ++wp) {
        const PhaseDesc ph = p.phases[wp];
        const int l = ph.layer;                                  // CLS: l == L
        const unsigned e = ebase + (unsigned)min(l, p.L - 1);    // epoch of this layer's buffers (CLS consumes layer L-1's)
        const int ks = warp & (ph.KS - 1), rg = warp / ph.KS, RG = kMegaWarps / ph.KS;
        const int c0 = ks * ph.SC;
        const int nsc = max(0, min(ph.SC, ph.nchunks - c0));
        const int cols = ph.nchunks * E;
        const bool normed = (ph.kind == PH_QKV || ph.kind == PH_GATEUP || ph.kind == PH_CLS);
        const int ev = wp + l + (ph.kind != PH_QKV && ph.kind != PH_CLS ? 1 : 0);
        LL_STAMP(ev, 0);

        // ---- 1. build the activation vector (spinning on the {value, epoch} words it is made of) -------------
        float ss = 0.f;
        const float* xsrc = resid_s;
        if (ph.kind == PH_QKV && l == 0) {                                   // embedding gather (model.cpp:48)
            const PhaseDesc em = p.phases[nwp - 1];
            const uint8_t* trow = p.emb + (size_t)(token / em.R) * em.KS * em.tile_bytes + (size_t)(token % em.R) * em.SC * 16;
            for (int c = tid; c < ph.nchunks; c += kMegaThreads) {
                const int eks = c / em.SC, ecc = c - eks * em.SC;
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(trow + (size_t)eks * em.tile_bytes + (size_t)ecc * 16));
                float f[E];
                if (WD == SLLM_F32) {
                    f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y); f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
                } else if (WD == SLLM_BF16) {
                    kv_unpack<SLLM_BF16>(raw, f);
                } else {   // int8: dequantised value = q * group scale (the scales sit behind the tile's weights)
                    const uint8_t* tile = p.emb + ((size_t)(token / em.R) * em.KS + eks) * em.tile_bytes;
                    const float sc = __ldg(reinterpret_cast<const float*>(tile + (size_t)em.R * em.SC * 16 + (size_t)(token % em.R) * em.srow) + (ecc >> 2));
                    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t wf = w4[k] ^ 0x80808080u;
                        f[4 * k] = s8f<0>(wf) * sc; f[4 * k + 1] = s8f<1>(wf) * sc; f[4 * k + 2] = s8f<2>(wf) * sc; f[4 * k + 3] = s8f<3>(wf) * sc;
                    }
                }
#pragma unroll
                for (int ee = 0; ee < E; ++ee) { resid_s[c * E + ee] = f[ee]; ss = fmaf(f[ee], f[ee], ss); }
            }
        } else if (normed) {
            // residual update: qkv / classifier add the all-reduced DOWN partials of the previous layer, gate_up adds
            // the all-reduced WO partials of this layer — tp vectors, summed in rank order (add_kernel.cpp:10-13)
            const int64_t off = (ph.kind == PH_GATEUP) ? p.off_wop : p.off_dnp;
            const unsigned ein = (ph.kind == PH_QKV) ? e - 1u : e;
            for (int c4 = tid; c4 < p.d / 4; c4 += kMegaThreads) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r0 = 0; r0 < p.tp; r0 += 2) {          // ranks in batches of 2 (register budget: wider batches spill in the hot loop); rank order kept
                    float4 pr[2];
                    ll_recv4xN<2>(my_area + off + (int64_t)r0 * p.d + 4 * c4, p.d, min(2, p.tp - r0), ein, pr);
#pragma unroll
                    for (int r = 0; r < 2; ++r)
                        if (r0 + r < p.tp) a = (r0 + r == 0) ? pr[0] : make_float4(a.x + pr[r].x, a.y + pr[r].y, a.z + pr[r].z, a.w + pr[r].w);
                }
                float4 v = reinterpret_cast<float4*>(resid_s)[c4];
                v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
                reinterpret_cast<float4*>(resid_s)[c4] = v;
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
                if (ph.kind == PH_CLS && cta == 0) reinterpret_cast<float4*>(p.x_out)[c4] = v;   // introspection: final residual
            }
        } else if (ph.kind == PH_WO) {                                       // merge the attention splits
            xsrc = xs;
            const int rec = p.hd + kAttRecPad;
            // step 1: one thread per (head, split) polls that record's {m, l} pair ONCE and turns it into a merge weight in shared
            // memory (the partial-sum table is idle here); the columns of a head then share it — before, every thread polled all
            // splits' {m, l} twice, twelve L2 round trips in a row per layer (trace: 12.5 us of wo prologue, 6.7 us now).
            float* mw = part;                                  // [H_loc][nsplit]: merge weight exp(m - M), L_total
            const int nrec = p.H_loc * nsplit;
            for (int i = tid; i < nrec; i += kMegaThreads) {
                const float2 ml = ll_recv2(my_area + p.off_att + (int64_t)i * rec + p.hd, e);
                mw[2 * i] = ml.x;
                mw[2 * i + 1] = ml.y;
            }
            __syncthreads();
            for (int h = tid; h < p.H_loc; h += kMegaThreads) {
                float M = -INFINITY, Ls = 0.f;
                for (int sp = 0; sp < nsplit; ++sp) M = fmaxf(M, mw[2 * (h * nsplit + sp)]);
                for (int sp = 0; sp < nsplit; ++sp) {
                    const float m = mw[2 * (h * nsplit + sp)];
                    const float w = (m == -INFINITY) ? 0.f : expf(m - M);
                    Ls = fmaf(mw[2 * (h * nsplit + sp) + 1], w, Ls);
                    mw[2 * (h * nsplit + sp)] = w;
                }
                for (int sp = 0; sp < nsplit; ++sp) mw[2 * (h * nsplit + sp) + 1] = Ls;
            }
            __syncthreads();
            // step 2: weighted sum of the splits' outputs, two splits (four 16-byte loads) in flight per polling round trip
            // (four in flight, inline or out of line, pushes the kernel over its 128-register cap into the streaming loop: slower).
            // When the rank has fewer float4 column groups than threads (tensor parallel: q_loc / 4 < 512), P threads share a
            // group and each takes a contiguous block of the splits — 32 splits at 8K context are then 4 round trips, not 16 —
            // and the P partial sums are added in block order from shared memory (fixed order: every CTA and rank agrees bit for bit).
            const int ngr = cols / 4;
            const int P = (2 * nrec > 1024) ? 1 : max(1, min(min(8, kMegaThreads / ngr), nsplit));   // (merge weights must end before the scratch)
            const int blk = (nsplit + P - 1) / P;
            float4* scratch = reinterpret_cast<float4*>(part + 1024);       // behind the merge weights (<= 4 KB), <= 8 KB
            for (int idx = tid; idx < ngr * P; idx += kMegaThreads) {
                const int c4 = idx % ngr, pt = idx / ngr;
                const int col = c4 * 4;
                const int head = col / p.hd, j = col - head * p.hd;
                const uint2* base = my_area + p.off_att + (int64_t)head * nsplit * rec;
                constexpr int kNS = SLLM_LL_KNS;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                const int s_end = min(nsplit, (pt + 1) * blk);
                for (int s0 = pt * blk; s0 < s_end; s0 += kNS) {
                    float4 ov[kNS];
                    ll_recv4xN<kNS>(base + (int64_t)s0 * rec + j, rec, min(kNS, s_end - s0), e, ov);
#pragma unroll
                    for (int i = 0; i < kNS; ++i) {
                        if (s0 + i < s_end) {
                            const float w = mw[2 * (head * nsplit + s0 + i)];
                            o.x = fmaf(ov[i].x, w, o.x); o.y = fmaf(ov[i].y, w, o.y); o.z = fmaf(ov[i].z, w, o.z); o.w = fmaf(ov[i].w, w, o.w);
                        }
                    }
                }
                if (P == 1) {
                    const float Ls = mw[2 * head * nsplit + 1];
                    reinterpret_cast<float4*>(xs)[c4] = make_float4(o.x / Ls, o.y / Ls, o.z / Ls, o.w / Ls);
                } else {
                    scratch[pt * ngr + c4] = o;
                }
            }
            if (P > 1) {
                __syncthreads();
                for (int c4 = tid; c4 < ngr; c4 += kMegaThreads) {
                    float4 o = scratch[c4];
                    for (int pt = 1; pt < P; ++pt) {
                        const float4 q = scratch[pt * ngr + c4];
                        o = make_float4(o.x + q.x, o.y + q.y, o.z + q.z, o.w + q.w);
                    }
                    const float Ls = mw[2 * ((c4 * 4) / p.hd) * nsplit + 1];
                    reinterpret_cast<float4*>(xs)[c4] = make_float4(o.x / Ls, o.y / Ls, o.z / Ls, o.w / Ls);
                }
            }
        } else {                                                              // down: sigmoid(gate)*up of this layer
            xsrc = xs;
            for (int c4 = tid; c4 < cols / 4; c4 += kMegaThreads)
                reinterpret_cast<float4*>(xs)[c4] = ll_recv4(my_area + p.off_swi + 4 * c4, e);
        }
        // norm weights of this lane's columns (requested now so the loads overlap the reduction below)
        float nwr[CPL][E];
        if (normed) {
            const float* nw = p.norms + (size_t)(ph.kind == PH_QKV ? 2 * l : ph.kind == PH_GATEUP ? 2 * l + 1 : 2 * p.L) * p.d;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
#pragma unroll
                for (int e4 = 0; e4 < E; e4 += 4) {
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c < nsc) g = __ldg(reinterpret_cast<const float4*>(nw + (c0 + c) * E + e4));
                    nwr[i][e4] = g.x; nwr[i][e4 + 1] = g.y; nwr[i][e4 + 2] = g.z; nwr[i][e4 + 3] = g.w;
                }
            }
        }

        float inv = 1.f;
        if (normed) {                                                        // RMSNorm, rms_kernel.cpp:12-22
            ss = warp_sum(ss);
            if (lane == 0) red[warp] = ss;
            __syncthreads();
            float tot = 0.f;
#pragma unroll
            for (int k = 0; k < kMegaWarps; ++k) tot += red[k];
            inv = 1.0f / sqrtf(tot / (float)cols + p.eps);
        } else {
            __syncthreads();
        }
        float xr[CPL][E];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
#pragma unroll
            for (int ee = 0; ee < E; ++ee) xr[i][ee] = 0.f;
            if (c < nsc) {
                const int col = (c0 + c) * E;
#pragma unroll
                for (int e4 = 0; e4 < E; e4 += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(xsrc + col + e4);
                    xr[i][e4] = v.x; xr[i][e4 + 1] = v.y; xr[i][e4 + 2] = v.z; xr[i][e4 + 3] = v.w;
                }
                if (normed) {
#pragma unroll
                    for (int ee = 0; ee < E; ++ee) xr[i][ee] = (xr[i][ee] * inv) * nwr[i][ee];
                }
            }
        }

        if (ph.kind == PH_QKV && cta < p.KVH_loc * nsplit) {
            // K/V rows of this CTA's first attention item (all but the newest row were stored by EARLIER launches):
            // start their TMA now so they land while phase A streams its weights. (xs is idle: A reads resid_s.)
            __syncthreads();
            if (warp == 0 && lane == 0) {
                fence_async_smem();
                const int row_bytes = p.hd * KESZ;
                const int npos = pos + 1, per = (npos + nsplit - 1) / nsplit;
                const int kvh = cta / nsplit, split = cta - kvh * nsplit;
                const int t0 = split * per, t1 = min(npos, t0 + per);
                const size_t head_off = ((size_t)l * p.KVH_loc + kvh) * p.S * row_bytes;
                for (int tile = 0; tile < 2; ++tile) {
                    const int ts = t0 + tile * AT;
                    const int rows = min(AT, t1 - ts);
                    if (rows <= 0) break;
                    const int bulk_rows = max(0, min(rows, pos - ts));
                    mb_expect(att_bar + tile, (uint32_t)(2 * bulk_rows * row_bytes));
                    if (bulk_rows > 0) {
                        tma_g2s(smem + SL.att_k + (size_t)tile * AT * row_bytes, p.kc + head_off + (size_t)ts * row_bytes, bulk_rows * row_bytes, att_bar + tile);
                        tma_g2s(smem + SL.att_v + (size_t)tile * AT * row_bytes, p.vc + head_off + (size_t)ts * row_bytes, bulk_rows * row_bytes, att_bar + tile);
                    }
                }
            }
        }

        LL_STAMP(ev, 1);
        // ---- 2. stream this CTA's tile rows through the rings, round by round ------------------------------
        int g0, g1;
        cta_tiles(ph, cta, ncta, g0, g1);
        const int upp = ph.R >> 1;
        const int u0 = g0 * upp;
        const int n = max(0, min(ph.nunits, g1 * upp) - u0);
        const int nslots = g1 - g0;
        const uint32_t sbytes = (uint32_t)ph.SC * 16;
        const int cpl = (nsc + 31) >> 5;
        float best_v = -INFINITY;
        int best_i = 0x7fffffff;

#pragma unroll 1
        for (int rbase = 0; rbase < n || rbase == 0; rbase += kRoundUnits) {
            const int jend = min(nslots, (rbase + kRoundUnits) / upp);
#pragma unroll 1
            for (int j = rbase / upp + rg; j < jend; j += RG) {
                const int si = cons_count & (kSlots - 1);
                mb_wait_fast(my_bar + si, (cons_count / kSlots) & 1);
                const uint8_t* sp = my_ring + (size_t)si * kSlotBytes + lane * 16;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (WD == SLLM_INT8) {
                    // 16 int8 weights per chunk (exact int8 -> fp32 by byte permute); the chunk's group scale (one fp32 per 4 chunks,
                    // stored behind the tile's weights) multiplies the chunk's partial sum — the same arithmetic as megakernel.cu
                    const uint8_t* sc_base = sp - lane * 16 + (size_t)ph.R * sbytes;
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const int c = lane + 32 * i;
                        if (i < cpl && c < nsc) {
                            const uint8_t* q = sp + i * 512;
                            const float* scp = reinterpret_cast<const float*>(sc_base) + (c >> 2);
                            a0 = fmaf(reg_dot<WD>(*reinterpret_cast<const uint4*>(q), xr[i], 0.f), scp[0], a0);
                            a1 = fmaf(reg_dot<WD>(*reinterpret_cast<const uint4*>(q + sbytes), xr[i], 0.f), scp[ph.srow >> 2], a1);
                            if (upp == 2) {
                                a2 = fmaf(reg_dot<WD>(*reinterpret_cast<const uint4*>(q + 2 * sbytes), xr[i], 0.f), scp[2 * (ph.srow >> 2)], a2);
                                a3 = fmaf(reg_dot<WD>(*reinterpret_cast<const uint4*>(q + 3 * sbytes), xr[i], 0.f), scp[3 * (ph.srow >> 2)], a3);
                            }
                        }
                    }
                } else if (upp == 2) {
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        if (i < cpl && lane + 32 * i < nsc) {
                            const uint8_t* q = sp + i * 512;
                            a0 = reg_dot<WD>(*reinterpret_cast<const uint4*>(q), xr[i], a0);
                            a1 = reg_dot<WD>(*reinterpret_cast<const uint4*>(q + sbytes), xr[i], a1);
                            a2 = reg_dot<WD>(*reinterpret_cast<const uint4*>(q + 2 * sbytes), xr[i], a2);
                            a3 = reg_dot<WD>(*reinterpret_cast<const uint4*>(q + 3 * sbytes), xr[i], a3);
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        if (i < cpl && lane + 32 * i < nsc) {
                            const uint8_t* q = sp + i * 512;
                            a0 = reg_dot<WD>(*reinterpret_cast<const uint4*>(q), xr[i], a0);
                            a1 = reg_dot<WD>(*reinterpret_cast<const uint4*>(q + sbytes), xr[i], a1);
                        }
                    }
                }
                {   // multi-value butterfly: 4 row sums over 32 lanes in 6 shuffles (fixed order => deterministic)
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
                    const int ul = j * upp + (r >> 1) - rbase;
                    if ((lane & 7) == 0 && r < 2 * upp && ul + rbase < n) part[(ul * 2 + (r & 1)) * kMegaWarps + ks] = k;
                }
                cons_count++;
                __syncwarp();
                if (lane == 0) {
                    fence_async_smem();
                    produce_one();
                }
            }
            __syncthreads();
            if (rbase + kRoundUnits >= n) LL_STAMP(ev, 3);
            // ---- 3. finish the round's units: sum over K slices, fused epilogue, results leave as words ----------
            const int nround = min(kRoundUnits, n - rbase);
#pragma unroll 1
            for (int t = tid; t < nround; t += kMegaThreads) {
                float s0 = 0.f, s1 = 0.f;
                for (int k = 0; k < ph.KS; ++k) {
                    s0 += part[(t * 2 + 0) * kMegaWarps + k];
                    s1 += part[(t * 2 + 1) * kMegaWarps + k];
                }
                const int u = u0 + rbase + t;
                if (ph.kind == PH_QKV) {
                    const int half = p.hd >> 1, rope_units = (p.q_loc + p.kv_loc) >> 1;
                    auto kv_addr = [&](uint8_t* cache, int idx) -> uint8_t* {
                        const int h = idx / p.hd, j = idx - h * p.hd;
                        return cache + ((((size_t)l * p.KVH_loc + h) * p.S + pos) * p.hd + j) * KESZ;
                    };
                    // cache store (for later steps) + a word with the ROUNDED value (for this step's attention)
                    auto store_kv = [&](uint8_t* cache, int64_t word_off, int idx, float v) {
                        float vr = v;
                        if (KVD == SLLM_BF16) {
                            const uint16_t b = f32_to_bf16_bits(v);
                            *reinterpret_cast<uint16_t*>(kv_addr(cache, idx)) = b;
                            vr = __uint_as_float((uint32_t)b << 16);
                        } else {
                            *reinterpret_cast<float*>(kv_addr(cache, idx)) = v;
                        }
                        ll_send(my_area + word_off + idx, vr, e);
                    };
                    if (u < rope_units) {
                        const int head = u / half, j = u - head * half;
                        const float fci = p.sin_t[(size_t)pos * half + j], fcr = p.cos_t[(size_t)pos * half + j];
                        const float o0 = s0 * fcr - s1 * fci, o1 = s1 * fcr + s0 * fci;   // rope_kernel.cpp:36-37
                        const int r0 = head * p.hd + j;
                        if (r0 < p.q_loc) {
                            ll_send(my_area + p.off_qv + r0, o0, e);
                            ll_send(my_area + p.off_qv + r0 + half, o1, e);
                        } else {
                            store_kv(p.kc, p.off_kvn, r0 - p.q_loc, o0);
                            store_kv(p.kc, p.off_kvn, r0 - p.q_loc + half, o1);
                        }
                    } else {
                        const int b2 = 2 * (u - rope_units);
                        store_kv(p.vc, p.off_kvn + p.kv_loc, b2, s0);
                        store_kv(p.vc, p.off_kvn + p.kv_loc, b2 + 1, s1);
                    }
                } else if (ph.kind == PH_WO || ph.kind == PH_DOWN) {          // partial sums -> every rank (the all-reduce)
                    const int64_t off = (ph.kind == PH_WO ? p.off_wop : p.off_dnp) + (int64_t)p.rank * p.d;
                    const int r = 2 * u;
                    for (int dst = 0; dst < p.tp; ++dst) {
                        ll_send(p.area[dst] + off + r, s0, e);
                        if (r + 1 < ph.nrows) ll_send(p.area[dst] + off + r + 1, s1, e);
                    }
                } else if (ph.kind == PH_GATEUP) {
                    ll_send(my_area + p.off_swi + u, (1.0f / (1.0f + expf(-s1))) * s0, e);   // swiglu_kernel.cpp:12-13
                } else {
                    const int r = 2 * u;
                    p.logits[r] = s0;
                    if (s0 > best_v || (s0 == best_v && p.v0 + r < best_i)) { best_v = s0; best_i = p.v0 + r; }
                    if (r + 1 < ph.nrows) {
                        p.logits[r + 1] = s1;
                        if (s1 > best_v || (s1 == best_v && p.v0 + r + 1 < best_i)) { best_v = s1; best_i = p.v0 + r + 1; }
                    }
                }
            }
            __syncthreads();
        }

        if (ph.kind == PH_CLS) {   // CTA best -> global; last CTA: exchange with the other ranks, arg max, advance the state
            float* sv = part;
            int* si = reinterpret_cast<int*>(part + kMegaThreads);
            sv[tid] = best_v;
            si[tid] = best_i;
            __syncthreads();
            for (int o = kMegaThreads / 2; o > 0; o >>= 1) {
                if (tid < o) {
                    const float ov = sv[tid + o];
                    const int oi = si[tid + o];
                    if (ov > sv[tid] || (ov == sv[tid] && oi < si[tid])) { sv[tid] = ov; si[tid] = oi; }
                }
                __syncthreads();
            }
            if (tid == 0) {
                p.blk_val[cta] = sv[0];
                p.blk_idx[cta] = si[0];
                __threadfence();
                s_last = (atomicAdd(&p.st->ticket, 1) == ncta - 1);
            }
            __syncthreads();
            if (s_last && tid == 0) {
                __threadfence();
                float v = -INFINITY;
                int idx = 0x7fffffff;
                for (int b = 0; b < ncta; ++b) {
                    const float ov = __ldcg(p.blk_val + b);
                    const int oi = __ldcg(p.blk_idx + b);
                    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
                }
                if (p.tp > 1) {   // ranks exchange (value, index) through the same words; first maximum wins
                    const unsigned ea = ebase + (unsigned)p.L;
                    for (int dst = 0; dst < p.tp; ++dst) {
                        ll_send(p.area[dst] + p.off_arg + 2 * p.rank, v, ea);
                        ll_send(p.area[dst] + p.off_arg + 2 * p.rank + 1, __int_as_float(idx), ea);
                    }
                    v = -INFINITY;
                    idx = 0x7fffffff;
                    for (int r = 0; r < p.tp; ++r) {
                        const float2 pr = ll_recv2(my_area + p.off_arg + 2 * r, ea);
                        const int oi = __float_as_int(pr.y);
                        if (pr.x > v || (pr.x == v && oi < idx)) { v = pr.x; idx = oi; }
                    }
                }
                if (idx == 0x7fffffff) idx = 0;
                p.st->ticket = 0;
                p.blk_val[ncta] = v;
                p.blk_idx[ncta] = idx;
                ClsPolicy<SLLM_F32>::step_feedback(p.st, p.prompt, p.history, idx);
            }
            break;
        }
        LL_STAMP(ev, 4);
        if (ph.kind != PH_QKV) continue;
        LL_STAMP(ev + 1, 0);

        // =============================== attention phase of layer l ======================================
        {
            float* q_s = reinterpret_cast<float*>(smem + SL.att_q);
            float* p_s = reinterpret_cast<float*>(smem + SL.att_p);
            float* alpha_s = reinterpret_cast<float*>(smem + SL.att_misc);
            float* ml_s = alpha_s + 16;
            uint8_t* k_s = smem + SL.att_k;
            uint8_t* v_s = smem + SL.att_v;
            const int stride = SL.kv_stride;
            const int row_bytes = p.hd * KESZ;
            const int cpr = row_bytes / 16;
            const int npos = pos + 1;
            const int per = (npos + nsplit - 1) / nsplit;
            const int nitems = p.KVH_loc * nsplit;
            const float scale = 1.0f / sqrtf((float)p.hd);
            constexpr int kStripes = 16;
            const int pv_chunk = tid % cpr, pv_stripe = tid / cpr;
            const bool pv_active = pv_stripe < kStripes;
            const int key = tid >> 3, kpart = tid & 7;

#pragma unroll 1
            for (int item = cta; item < nitems; item += ncta) {
                const int kvh = item / nsplit, split = item - kvh * nsplit;
                const int t0 = split * per, t1 = min(npos, t0 + per);
                const int ntiles = (t1 > t0) ? (t1 - t0 + AT - 1) / AT : 0;
                // q of this KV head's query heads: words written by phase A's epilogues (any CTA)
                for (int i = 2 * tid; i < G * p.hd; i += 2 * kMegaThreads) {
                    const float2 v = ll_recv2(my_area + p.off_qv + (int64_t)(kvh * G) * p.hd + i, e);
                    q_s[i] = v.x; q_s[i + 1] = v.y;
                }
                if (tid < G) { ml_s[2 * tid] = -INFINITY; ml_s[2 * tid + 1] = 0.f; }
                const size_t head_off = ((size_t)l * p.KVH_loc + kvh) * p.S * row_bytes;
                auto issue_tile = [&](int tile) {
                    const int stage = tile & 1;
                    const int ts = t0 + tile * AT;
                    const int rows = min(AT, t1 - ts);
                    const int bulk_rows = max(0, min(rows, pos - ts));
                    if (lane == 0) {
                        mb_expect(att_bar + stage, (uint32_t)(2 * bulk_rows * row_bytes));
                        if (bulk_rows > 0) {
                            tma_g2s(k_s + (size_t)stage * AT * stride, p.kc + head_off + (size_t)ts * row_bytes, bulk_rows * row_bytes, att_bar + stage);
                            tma_g2s(v_s + (size_t)stage * AT * stride, p.vc + head_off + (size_t)ts * row_bytes, bulk_rows * row_bytes, att_bar + stage);
                        }
                    }
                };
                fence_async_smem();
                __syncthreads();
                if (warp == 0 && item != cta) {
                    if (ntiles > 0) issue_tile(0);
                    if (ntiles > 1) issue_tile(1);
                }
                float acc[G][KVEC];
#pragma unroll
                for (int gi = 0; gi < G; ++gi)
#pragma unroll
                    for (int ee = 0; ee < KVEC; ++ee) acc[gi][ee] = 0.f;

#pragma unroll 1
                for (int tile = 0; tile < ntiles; ++tile) {
                    const int stage = tile & 1;
                    const int ts = t0 + tile * AT;
                    const int rows = min(AT, t1 - ts);
                    // the newest row arrives as words from phase A's epilogues (values already rounded to the cache type)
                    if (pos >= ts && pos < ts + rows && warp == 1) {
                        for (int c = lane; c < cpr; c += 32) {
                            float kf[KVEC], vf[KVEC];
                            {   // KVEC consecutive words each of K and V: one polling round trip per matrix
                                const uint2* kp = my_area + p.off_kvn + (int64_t)kvh * p.hd + c * KVEC;
                                const uint2* vp = kp + p.kv_loc;
                                float4 t4[KVEC / 4];
                                ll_recv4xN<KVEC / 4>(kp, 4, KVEC / 4, e, t4);
#pragma unroll
                                for (int q4 = 0; q4 < KVEC / 4; ++q4) { kf[4 * q4] = t4[q4].x; kf[4 * q4 + 1] = t4[q4].y; kf[4 * q4 + 2] = t4[q4].z; kf[4 * q4 + 3] = t4[q4].w; }
                                ll_recv4xN<KVEC / 4>(vp, 4, KVEC / 4, e, t4);
#pragma unroll
                                for (int q4 = 0; q4 < KVEC / 4; ++q4) { vf[4 * q4] = t4[q4].x; vf[4 * q4 + 1] = t4[q4].y; vf[4 * q4 + 2] = t4[q4].z; vf[4 * q4 + 3] = t4[q4].w; }
                            }
                            uint4 kw, vw;
                            if (KVD == SLLM_F32) {
                                kw = make_uint4(__float_as_uint(kf[0]), __float_as_uint(kf[1]), __float_as_uint(kf[2]), __float_as_uint(kf[3]));
                                vw = make_uint4(__float_as_uint(vf[0]), __float_as_uint(vf[1]), __float_as_uint(vf[2]), __float_as_uint(vf[3]));
                            } else {
                                auto pk = [](float lo, float hi) { return (__float_as_uint(lo) >> 16) | (__float_as_uint(hi) & 0xffff0000u); };
                                kw = make_uint4(pk(kf[0], kf[1]), pk(kf[2], kf[3]), pk(kf[4 % KVEC], kf[5 % KVEC]), pk(kf[6 % KVEC], kf[7 % KVEC]));
                                vw = make_uint4(pk(vf[0], vf[1]), pk(vf[2], vf[3]), pk(vf[4 % KVEC], vf[5 % KVEC]), pk(vf[6 % KVEC], vf[7 % KVEC]));
                            }
                            *reinterpret_cast<uint4*>(k_s + ((size_t)stage * AT + (pos - ts)) * stride + c * 16) = kw;
                            *reinterpret_cast<uint4*>(v_s + ((size_t)stage * AT + (pos - ts)) * stride + c * 16) = vw;
                        }
                    }
                    if (stage == 0) { mb_wait_fast(att_bar, kv_use0 & 1); kv_use0++; }
                    else { mb_wait_fast(att_bar + 1, kv_use1 & 1); kv_use1++; }
                    __syncthreads();
                    {
                        float s[G];
#pragma unroll
                        for (int gi = 0; gi < G; ++gi) s[gi] = 0.f;
                        if (key < rows) {
                            const uint8_t* krow = k_s + ((size_t)stage * AT + key) * stride;
                            for (int c = kpart; c < cpr; c += 8) {
                                float kf[KVEC];
                                kv_unpack<KVD>(*reinterpret_cast<const uint4*>(krow + c * 16), kf);
#pragma unroll
                                for (int gi = 0; gi < G; ++gi) {
                                    const float* qv = q_s + gi * p.hd + c * KVEC;
#pragma unroll
                                    for (int ee = 0; ee < KVEC; ++ee) s[gi] = fmaf(qv[ee], kf[ee], s[gi]);
                                }
                            }
                        }
#pragma unroll
                        for (int gi = 0; gi < G; ++gi) {
                            s[gi] += __shfl_xor_sync(0xffffffffu, s[gi], 1);
                            s[gi] += __shfl_xor_sync(0xffffffffu, s[gi], 2);
                            s[gi] += __shfl_xor_sync(0xffffffffu, s[gi], 4);
                            if (kpart == 0) p_s[gi * kAttTile + key] = (key < rows) ? s[gi] * scale : -INFINITY;
                        }
                    }
                    __syncthreads();
                    for (int gi = warp; gi < G; gi += kMegaWarps) {
                        const float s0 = p_s[gi * kAttTile + lane], s1 = p_s[gi * kAttTile + lane + 32];
                        const float m_old = ml_s[2 * gi], l_old = ml_s[2 * gi + 1];
                        const float m_new = fmaxf(m_old, warp_max(fmaxf(s0, s1)));
                        const float e0 = expf(s0 - m_new), e1 = expf(s1 - m_new);
                        const float al = expf(m_old - m_new);
                        const float l_new = l_old * al + warp_sum(e0 + e1);
                        p_s[gi * kAttTile + lane] = e0;
                        p_s[gi * kAttTile + lane + 32] = e1;
                        if (lane == 0) { alpha_s[gi] = al; ml_s[2 * gi] = m_new; ml_s[2 * gi + 1] = l_new; }
                    }
                    __syncthreads();
                    if (pv_active) {
#pragma unroll
                        for (int gi = 0; gi < G; ++gi) {
                            const float al = alpha_s[gi];
#pragma unroll
                            for (int ee = 0; ee < KVEC; ++ee) acc[gi][ee] *= al;
                        }
                        for (int r = pv_stripe; r < rows; r += kStripes) {
                            float vf[KVEC];
                            kv_unpack<KVD>(*reinterpret_cast<const uint4*>(v_s + ((size_t)stage * AT + r) * stride + pv_chunk * 16), vf);
#pragma unroll
                            for (int gi = 0; gi < G; ++gi) {
                                const float pr = p_s[gi * kAttTile + r];
#pragma unroll
                                for (int ee = 0; ee < KVEC; ++ee) acc[gi][ee] = fmaf(pr, vf[ee], acc[gi][ee]);
                            }
                        }
                    }
                    fence_async_smem();
                    __syncthreads();
                    if (warp == 0 && tile + 2 < ntiles) issue_tile(tile + 2);
                }
                float* o_s = reinterpret_cast<float*>(k_s);   // [kStripes][G][hd]
                if (pv_active) {
#pragma unroll
                    for (int gi = 0; gi < G; ++gi)
#pragma unroll
                        for (int ee = 0; ee < KVEC; ++ee) o_s[((size_t)pv_stripe * G + gi) * p.hd + pv_chunk * KVEC + ee] = acc[gi][ee];
                }
                __syncthreads();
                const int rec = p.hd + kAttRecPad;
                for (int i = tid; i < G * p.hd; i += kMegaThreads) {
                    float o = 0.f;
                    for (int s = 0; s < kStripes; ++s) o += o_s[(size_t)s * G * p.hd + i];
                    const int gi = i / p.hd, j = i - gi * p.hd;
                    ll_send(my_area + p.off_att + ((int64_t)(kvh * G + gi) * nsplit + split) * rec + j, o, e);
                }
                if (tid < G) {
                    uint2* r = my_area + p.off_att + ((int64_t)(kvh * G + tid) * nsplit + split) * rec + p.hd;
                    ll_send(r, ml_s[2 * tid], e);
                    ll_send(r + 1, ml_s[2 * tid + 1], e);
                    ll_send(r + 2, 0.f, e);   // pad words: the merge reads the record tail as one group of four
                    ll_send(r + 3, 0.f, e);
                }
            }
            fence_async_smem();
            __syncthreads();   // o_s / xs region is reused by the next prologue
        }
        LL_STAMP(ev + 1, 4);
    }
}

// ------------------------------------------------------------------------------------------- host ----
MegaLLPlan mega_ll_plan(int w_dtype, int kv_dtype, int d, int hd, int q_loc, int kv_loc, int I_loc, int V_loc, int v0, int H_loc, int KVH_loc,
                        int max_len, int tp) {
    return mega_ll_plan_for(sm_count(), smem_optin_bytes(), w_dtype, kv_dtype, d, hd, q_loc, kv_loc, I_loc, V_loc, v0, H_loc, KVH_loc, max_len, tp);
}

// the plan as a pure function of the device facts (testable without a device: sllm_mega_plan)
MegaLLPlan mega_ll_plan_for(int sms, int smem_optin, int w_dtype, int kv_dtype, int d, int hd, int q_loc, int kv_loc, int I_loc, int V_loc, int v0,
                            int H_loc, int KVH_loc, int max_len, int tp) {
    MegaLLPlan pl;
    if (tp > kMaxTp) { pl.why = "more than 8 ranks"; return pl; }
    const int E = w_dtype == SLLM_F32 ? 4 : w_dtype == SLLM_BF16 ? 8 : 16;
    for (int cols : {d, q_loc, I_loc}) {
        // int8: a rank's slice of a row-parallel matrix (q_loc, I_loc columns) must hold whole quantisation groups of 64
        if (cols % E || cols % 4 || (w_dtype == SLLM_INT8 && cols % 64)) { pl.why = "row length not a multiple of 16 bytes / of the int8 group"; return pl; }
        const TileGeom tg = mega_tile_geom(2, cols, w_dtype);
        if (tg.KS * tg.SC < tg.nchunks || (tg.SC * 16 + tg.srow) * 2 > kSlotBytes || tg.SC > 32 * (w_dtype == SLLM_INT8 ? 2 : kCplMax)) {
            pl.why = "rows longer than 32 KB";
            return pl;
        }
    }
    const int g = H_loc / KVH_loc;
    if ((g != 1 && g != 2 && g != 4 && g != 8) || hd % 16 || hd > 256) { pl.why = "head shape"; return pl; }
    const int kesz = kv_dtype == SLLM_F32 ? 4 : 2;
    if ((hd * kesz / 16) > 32) { pl.why = "head_dim chunking"; return pl; }
    const MegaLLSmem SL = mega_ll_smem_layout(d, hd, g, kesz);
    if (SL.total + 1024 > (size_t)smem_optin) { pl.why = "shared memory"; return pl; }
    if ((size_t)g * hd * 4 + (size_t)g * kAttTile * 4 + 256 > (size_t)kRoundUnits * 2 * kMegaWarps * 4) { pl.why = "attention scratch (q/p)"; return pl; }
    if ((size_t)16 * g * hd * 4 > (size_t)4 * SL.att_tile * SL.kv_stride) { pl.why = "attention scratch"; return pl; }   // the cross-stripe reduction buffer spans the (drained, contiguous) K and V stages
    if ((size_t)std::max(q_loc, I_loc) * 4 > (size_t)4 * SL.att_tile * SL.kv_stride) { pl.why = "activation staging"; return pl; }
    if (q_loc % 4 || kv_loc % 2 || I_loc % 4 || d % 4) { pl.why = "dims not multiples of 4"; return pl; }
    // every CTA must own >= 1 tile row of every weight phase (see the safety argument at the top of this file)
    int grid = sms;
    const int rows_kind[4][3] = {{q_loc + 2 * kv_loc, d, PH_QKV}, {d, q_loc, PH_WO}, {2 * I_loc, d, PH_GATEUP}, {d, I_loc, PH_DOWN}};
    for (auto& rk : rows_kind) grid = std::min(grid, mega_tile_geom(rk[2] == PH_GATEUP ? rk[0] : ((rk[0] + 1) / 2) * 2, rk[1], w_dtype).ntr);
    const TileGeom cg = mega_tile_geom(((V_loc + 1) / 2) * 2, d, w_dtype);
    if (v0 % cg.R) { pl.why = "vocab shard not tile aligned"; return pl; }
    if (grid < 1) { pl.why = "empty phase"; return pl; }
    pl.grid = grid;
    pl.smem = SL.total;
    int ns = std::max(1, grid / KVH_loc);
    const int by_len = (max_len + SL.att_tile - 1) / SL.att_tile;
    if (ns > by_len) ns = by_len;
    if (ns > 32) ns = 32;
    pl.nsplit = ns;
    // word offsets inside a rank's area
    int64_t off = 0;
    auto take = [&](int64_t n) { const int64_t o = off; off += (n + 15) / 16 * 16; return o; };
    pl.off_wop = take((int64_t)tp * d);
    pl.off_dnp = take((int64_t)tp * d);
    pl.off_qv = take(q_loc);
    pl.off_kvn = take(2 * (int64_t)kv_loc);
    pl.off_att = take((int64_t)H_loc * pl.nsplit * (hd + kAttRecPad));
    pl.off_swi = take(I_loc);
    pl.off_arg = take(2 * (int64_t)tp);
    pl.area_words = off;
    pl.ok = true;
    return pl;
}

template <int WD, int KVD, int G>
static int mega_ll_launch_t(const MegaLLParams& p, int grid, size_t smem, cudaStream_t st) {
    static size_t configured = 0;
    if (smem > configured) {
        SLLM_CUDA(cudaFuncSetAttribute(mega_ll_kernel<WD, KVD, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kMegaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // co-residency of all CTAs (they wait for each other's words)
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SLLM_CUDA(cudaLaunchKernelEx(&cfg, mega_ll_kernel<WD, KVD, G>, p));
    g_launches++;
    return SLLM_OK;
}

int mega_ll_launch(const MegaLLParams& p, int g, int grid, size_t smem, cudaStream_t st) {
#define MEGA_LL_G(GG)                                                                                       \
    case GG:                                                                                                \
        if (p.w_dtype == SLLM_F32) {                                                                        \
            return p.kv_dtype == SLLM_F32 ? mega_ll_launch_t<SLLM_F32, SLLM_F32, GG>(p, grid, smem, st)     \
                                          : mega_ll_launch_t<SLLM_F32, SLLM_BF16, GG>(p, grid, smem, st);   \
        }                                                                                                   \
        if (p.w_dtype == SLLM_INT8) {                                                                       \
            return p.kv_dtype == SLLM_F32 ? mega_ll_launch_t<SLLM_INT8, SLLM_F32, GG>(p, grid, smem, st)    \
                                          : mega_ll_launch_t<SLLM_INT8, SLLM_BF16, GG>(p, grid, smem, st);  \
        }                                                                                                   \
        return p.kv_dtype == SLLM_F32 ? mega_ll_launch_t<SLLM_BF16, SLLM_F32, GG>(p, grid, smem, st)        \
                                      : mega_ll_launch_t<SLLM_BF16, SLLM_BF16, GG>(p, grid, smem, st);
    switch (g) {
        MEGA_LL_G(1)
        MEGA_LL_G(2)
        MEGA_LL_G(4)
        MEGA_LL_G(8)
        default: set_error("megakernel: %d query heads per KV head not instantiated", g); return SLLM_ENOTSUP;
    }
#undef MEGA_LL_G
}

}  // namespace sllm
