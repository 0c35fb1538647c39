// csrc/megakernel.cu — the whole decode step as ONE persistent cooperative kernel (sm_100a).
//
// Why: with one kernel per fused op (decode_fused.cuh) every launch boundary drains the HBM pipe — measured
// on B200: 2-5 us of ramp + tail per GEMV kernel x 128 kernels, and a 12 us latency chain per attention launch
// (profiles/r01_launches.md). Here the step is one launch of `SM count` CTAs (1 per SM, 16 warps) that walks
// the phases  A qkv | B attention | C wo | D gate_up | E down  of every layer and finally F classifier+argmax,
// separated by grid barriers, and the WEIGHT STREAM NEVER STOPS:
//
//  * every warp owns a private ring of kSlots x kSlotBytes shared-memory slots filled by the TMA bulk-copy
//    engine (cp.async.bulk global->shared, completion on the slot's mbarrier). Lane 0 of the warp is the
//    producer: it re-arms a slot the moment the warp has consumed it and runs AHEAD across phase and layer
//    boundaries — weights do not depend on activations — so while CTAs sit in a grid barrier or in the
//    latency-bound attention phase, 128 KB per SM (19 MB chip-wide) of the next matrices is already landing.
//  * a CTA owns a contiguous block of two-row units of the phase's matrix; its 16 warps split K: warp
//    (ks, rg) streams column slice ks of the rows of row-group rg. A lane therefore multiplies the SAME
//    columns for every row, so its part of the activation vector lives in registers for the whole phase
//    (no shared-memory traffic for x at all). Partial dot products go to a small table; at the end of a
//    round of <= kRoundUnits units the CTA syncs once and thread t finishes unit t (sum over K slices in fixed
//    order — deterministic) and applies the fused epilogue (RoPE + cache write, residual, sigmoid*up, argmax).
//  * attention (phase B) takes (kv head, split) items; K/V tiles are TMA-staged exactly like mha.cu; the
//    merge of the splits is folded into phase C's prologue (every thread merges the partials of the columns
//    it needs), which removes one grid barrier and the atomic ticket.
//  * token and position stay in device memory; the last CTA to finish the classifier picks the arg max
//    (first maximum) and advances the step state, so n tokens = n launches with no host round trip.
//
// Numerics are those of the fused path (fp32 activations/accumulators, IEEE divide/sqrt, accurate expf).
#include <algorithm>
#include <cmath>

#include "mega_common.cuh"

namespace sllm {

// ----------------------------------------------------------------------------------------- kernel ----
// Optional per-CTA timeline (tools/mega_trace.py): stamp s of phase slot `ev` -> trace[(cta*kTraceEvents + ev)*8 + s]
constexpr int kTraceEvents = 512;
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MEGA_STAMP(ev, slot)                                                                              \
    do {                                                                                                  \
        if (p.trace && threadIdx.x == 0 && (ev) < kTraceEvents) p.trace[((size_t)blockIdx.x * kTraceEvents + (ev)) * 8 + (slot)] = gtime(); \
    } while (0)

// producer state of one warp (meaningful in lane 0): constants of the phase it is currently feeding
struct ProdState {
    int wp, j;                 // phase and next slot index (j < 0: phase not entered yet)
    unsigned count;            // slots issued so far
    const uint8_t* W;          // phase matrix + this warp's column offset
    int row_bytes, sbytes, u0, u1, upp, RG, nslots, kind;
};

// FUSE (experimental, SLLM_ENGINE_MEGA_FUSE_DOWN): the down projection runs inside the gate_up phase's CTA as a K-split over the
// values that CTA produced (megakernel.cuh "PH_DOWN_T"): four grid barriers per layer instead of five, no staging of swi.
template <int WD, int KVD, int G, bool FUSE>
__global__ void __launch_bounds__(kMegaThreads, 1) mega_step_kernel(const MegaParams p) {
    constexpr int E = WInfo<WD>::E;                 // weights per 16-byte chunk
    constexpr int CPL = (WD == SLLM_INT8) ? 2 : kCplMax;   // max chunks per lane per row slice (x of a lane: CPL * E = 32 registers)
    constexpr int KESZ = MKv<KVD>::ESZ, KVEC = MKv<KVD>::VEC;
    extern __shared__ __align__(128) uint8_t mega_smem[];
    uint8_t* const smem = mega_smem;
    const MegaSmem SL = mega_smem_layout(p.hd, G, KESZ);
    const int AT = SL.att_tile;                      // cache positions per K/V stage (mega_common.cuh)
    uint64_t* ring_bar = reinterpret_cast<uint64_t*>(smem + SL.bars);            // [16][kSlots]
    uint64_t* att_bar = ring_bar + kMegaWarps * kSlots;                          // [2]
    float* red = reinterpret_cast<float*>(smem + SL.red);
    float* part = reinterpret_cast<float*>(smem + SL.part);                      // [kRoundUnits][2][16]
    uint8_t* ring = smem + SL.ring;
    float* xs = reinterpret_cast<float*>(smem + SL.att_k);    // activation staging: aliases the (idle) K/V stages
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int nwp = 4 * p.L + 1;

    if (tid < kMegaWarps * kSlots + 2) mb_init(ring_bar + tid, 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    if (p.trace && threadIdx.x == 0) {   // which SM this CTA runs on (event 0, slot 7)
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[((size_t)blockIdx.x * kTraceEvents) * 8 + 7] = smid;
    }
    const int pos = p.st->pos;
    const int token = min(max(p.st->token, 0), p.V - 1);
    const unsigned bar_base = (unsigned)p.st->pad[0];
    unsigned bar_idx = 0;

    uint8_t* my_ring = ring + (size_t)warp * kSlots * kSlotBytes;
    uint64_t* my_bar = ring_bar + warp * kSlots;

    // ---------------- producer (lane 0 of every warp): one tile == one bulk copy -------------------------------
    int pr_wp = -1, pr_left = 0;          // phase being fed, tiles of it still to request
    unsigned pr_count = 0;                // slots issued so far
    const uint8_t* pr_ptr = nullptr;      // next tile of this warp's (ks, rg) sub-stream
    uint32_t pr_step = 0, pr_bytes = 0;
    uint32_t pr_tail = 0;                 // FUSE: bytes of the LAST copy of a transposed-down stream (a whole slot otherwise)
    auto produce_one = [&]() {   // lane 0 only: arm the next slot of this warp's stream (if any is left)
        while (pr_left <= 0) {                            // enter the next phase this warp has work in
            if (++pr_wp >= nwp) { pr_wp = nwp; return; }
            const PhaseDesc ph = p.phases[pr_wp];
            const int ks = warp & (ph.KS - 1), rg = warp / ph.KS, RG = kMegaWarps / ph.KS;
            int g0, g1;
            phase_tiles<FUSE>(ph, cta, ncta, g0, g1);
            if constexpr (FUSE) {
                pr_tail = 0;
                if (ph.kind == PH_DOWN_T) {   // stripe-major matrix: this warp's tile rows [a, b) are ONE contiguous byte range, copied a slot at a time
                    int a, b;
                    down_t_rows(g0, g1, rg, RG, a, b);
                    const uint32_t len = (uint32_t)(b - a) * (uint32_t)ph.tile_bytes;
                    pr_left = (int)((len + kSlotBytes - 1) / kSlotBytes);
                    pr_ptr = ph.W + ((size_t)ks * ph.ntr + a) * ph.tile_bytes;
                    pr_step = kSlotBytes;
                    pr_bytes = kSlotBytes;
                    pr_tail = len - (uint32_t)(pr_left - 1) * kSlotBytes;
                    continue;
                }
            }
            pr_left = (g1 - g0 - rg + RG - 1) / RG;       // tile rows g0+rg, g0+rg+RG, ... < g1
            pr_ptr = ph.W + ((size_t)(g0 + rg) * ph.KS + ks) * ph.tile_bytes;
            pr_step = (uint32_t)RG * ph.KS * ph.tile_bytes;
            pr_bytes = (uint32_t)ph.tile_bytes;
        }
        const int si = pr_count & (kSlots - 1);
        if constexpr (FUSE) {
            if (pr_tail && pr_left == 1) pr_bytes = pr_tail;
        }
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
    unsigned cons_count = 0;             // slots consumed so far (warp-uniform)
    unsigned kv_use0 = 0, kv_use1 = 0;   // how often each K/V stage mbarrier was armed (parity)

#pragma unroll 1
    for (int wp = 0; wp < nwp; ++wp) {
        // =============================== weight phase wp ===============================================
        const PhaseDesc ph = p.phases[wp];
        const int l = ph.layer;
        const int ks = warp & (ph.KS - 1), rg = warp / ph.KS, RG = kMegaWarps / ph.KS;
        const int c0 = ks * ph.SC;
        const int nsc = max(0, min(ph.SC, ph.nchunks - c0));   // real chunks of this slice (the tile is zero padded to SC)
        const int cols = ph.nchunks * E;
        const bool normed = (ph.kind == PH_QKV || ph.kind == PH_GATEUP || ph.kind == PH_CLS);
        const int ev = wp + l + (ph.kind != PH_QKV && ph.kind != PH_CLS ? 1 : 0) + (ph.kind == PH_CLS ? 0 : 0);   // event slot: weight phases and attention phases in order
        MEGA_STAMP(ev, 0);
        if constexpr (FUSE) {
            if (ph.kind == PH_DOWN_T) {
                // x += Wdown[:, this CTA's units] . swi[this CTA's units]: the gate_up epilogue left the values in shared memory (xs),
                // the wo epilogue left h in x. A lane owns E outputs; its warp streams the (tile row, stripe) tiles of stripe ks.
                int g0, g1;
                cta_tiles(ph, cta, ncta, g0, g1);
                MEGA_STAMP(ev, 1);   // no prologue: nothing to stage
                float acc[E];
#pragma unroll
                for (int e = 0; e < E; ++e) acc[e] = 0.f;
                bool any = false;
                int ja, jb;
                down_t_rows(g0, g1, rg, RG, ja, jb);      // this warp's tile rows: contiguous in the stripe-major matrix
                constexpr int kRowsPerSlot = kSlotBytes / (kFuseJT * 512);
#pragma unroll 1
                for (int j = ja; j < jb; j += kRowsPerSlot) {
                    const int si = cons_count & (kSlots - 1);
                    mb_wait_fast(my_bar + si, (cons_count / kSlots) & 1);
                    const uint8_t* sp = my_ring + (size_t)si * kSlotBytes + lane * 16;
                    const float* sw = xs + (size_t)(j - g0) * kFuseJT;
                    const int nj = min(kRowsPerSlot, jb - j) * kFuseJT;   // inputs in this slot (the last copy may hold fewer tile rows)
#pragma unroll 1
                    for (int j4 = 0; j4 < nj; j4 += kFuseJT) {
#pragma unroll
                        for (int jj = 0; jj < kFuseJT; ++jj)
                            axpy_chunk<WD>(*reinterpret_cast<const uint4*>(sp + (j4 + jj) * 512), sw[j4 + jj], acc);
                    }
                    any = true;
                    cons_count++;
                    __syncwarp();
                    if (lane == 0) {
                        fence_async_smem();
                        produce_one();
                    }
                }
                MEGA_STAMP(ev, 3);
                if (any) {
                    float* dst = p.x + (size_t)(ks * 32 + lane) * E;
                    red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
                    if (E == 8) red_add_v4(dst + 4, acc[E - 4], acc[E - 3], acc[E - 2], acc[E - 1]);
                }
                MEGA_STAMP(ev, 4);
                grid_barrier(p.bar_counter, bar_base + (++bar_idx) * (unsigned)ncta);
                MEGA_STAMP(ev, 5);
                continue;
            }
        }
        // norm weights of this lane's columns: constant data, requested before anything that has to wait
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
        // ---- 1. stage the activation vector in shared memory (plain loops, all 512 threads) -------------
        float ss = 0.f;
        if (ph.kind == PH_QKV && l == 0) {                                   // embedding gather (model.cpp:48)
            // row `token` of the (tiled) embedding/classifier matrix: tile row token/R, row token%R inside the tile
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
                for (int e = 0; e < E; ++e) {
                    xs[c * E + e] = f[e];
                    ss = fmaf(f[e], f[e], ss);
                    if (cta == 0) p.x[c * E + e] = f[e];
                }
            }
        } else if (ph.kind == PH_WO) {                                       // merge the attention splits
            const int rec = p.hd + kAttRecPad;
            for (int c4 = tid; c4 < cols / 4; c4 += kMegaThreads) {
                const int col = c4 * 4;
                const int head = col / p.hd, j = col - head * p.hd;
                const float* base = p.att_part + (size_t)head * p.nsplit * rec;
                float M = -INFINITY;
                for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, __ldcg(base + (size_t)s * rec + p.hd));
                float Ls = 0.f;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int s = 0; s < p.nsplit; ++s) {
                    const float m = __ldcg(base + (size_t)s * rec + p.hd);
                    const float w = (m == -INFINITY) ? 0.f : expf(m - M);
                    Ls = fmaf(__ldcg(base + (size_t)s * rec + p.hd + 1), w, Ls);
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (size_t)s * rec + j));
                    o.x = fmaf(v.x, w, o.x); o.y = fmaf(v.y, w, o.y); o.z = fmaf(v.z, w, o.z); o.w = fmaf(v.w, w, o.w);
                }
                reinterpret_cast<float4*>(xs)[c4] = make_float4(o.x / Ls, o.y / Ls, o.z / Ls, o.w / Ls);
            }
        } else {
            const float* src = (ph.kind == PH_GATEUP) ? p.h : (ph.kind == PH_DOWN) ? p.swi : p.x;
            for (int c4 = tid; c4 < cols / 4; c4 += kMegaThreads) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + c4);
                reinterpret_cast<float4*>(xs)[c4] = v;
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
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
        // ---- this lane's columns -> registers (chunk c = c0 + lane + 32 i  <->  columns c*E .. c*E+E)
        float xr[CPL][E];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
#pragma unroll
            for (int e = 0; e < E; ++e) xr[i][e] = 0.f;
            if (c < nsc) {
                const int col = (c0 + c) * E;
#pragma unroll
                for (int e4 = 0; e4 < E; e4 += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(xs + col + e4);
                    xr[i][e4] = v.x; xr[i][e4 + 1] = v.y; xr[i][e4 + 2] = v.z; xr[i][e4 + 3] = v.w;
                }
                if (normed) {
#pragma unroll
                    for (int e = 0; e < E; ++e) xr[i][e] = (xr[i][e] * inv) * nwr[i][e];
                }
            }
        }

        if (ph.kind == PH_QKV && cta < p.KVH_loc * p.nsplit) {
            // The K/V rows this CTA's first attention item needs (all but the row written this step) were stored by
            // EARLIER launches: start their TMA now, into the K/V stages that only alias the activation staging
            // buffer (already consumed into registers), so they land while phase A streams its weights.
            __syncthreads();                    // every lane has read its part of xs
            if (warp == 0 && lane == 0) {
                fence_async_smem();
                const int row_bytes = p.hd * KESZ;
                const int npos = pos + 1, per = (npos + p.nsplit - 1) / p.nsplit;
                const int kvh = cta / p.nsplit, split = cta - kvh * p.nsplit;
                const int t0 = split * per, t1 = min(npos, t0 + per);
                const size_t head_off = ((size_t)l * p.KVH_loc + kvh) * p.S * row_bytes;   // [L][KVH][S][hd]
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
        MEGA_STAMP(ev, 1);   // prologue done
        // ---- 2. stream this CTA's tile rows through the rings, round by round ------------------------------
        int g0, g1;
        phase_tiles<FUSE>(ph, cta, ncta, g0, g1);
        const int upp = ph.R >> 1;
        const int u0 = g0 * upp;
        const int n = max(0, min(ph.nunits, g1 * upp) - u0);   // real units of this CTA
        const int nslots = g1 - g0;
        const uint32_t sbytes = (uint32_t)ph.SC * 16;           // row pitch inside a tile
        const int cpl = (nsc + 31) >> 5;
        float best_v = -INFINITY;   // classifier only
        int best_i = 0x7fffffff;

#pragma unroll 1
        for (int rbase = 0; rbase < n || rbase == 0; rbase += kRoundUnits) {
            const int jend = min(nslots, (rbase + kRoundUnits) / upp);
            {
#pragma unroll 1
                for (int j = rbase / upp + rg; j < jend; j += RG) {
                    const int si = cons_count & (kSlots - 1);
                    mb_wait_fast(my_bar + si, (cons_count / kSlots) & 1);
                    const uint8_t* sp = my_ring + (size_t)si * kSlotBytes + lane * 16;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    if (p.debug & 2) {
                        a0 = __uint_as_float(*reinterpret_cast<const uint32_t*>(sp));   // measurement aid: touch the slot, skip the arithmetic
                    } else if (WD == SLLM_INT8) {
                        // 16 int8 weights per chunk, exact int8 -> fp32 by byte permute (gemv_core.cuh s8f); the group scale of
                        // the chunk (one fp32 per 4 chunks, stored behind the tile's weights) multiplies the chunk's partial sum
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
                    // multi-value butterfly: 4 row sums over 32 lanes in 6 shuffles (fixed order => deterministic)
                    {
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
                        // lane 0: row 0, lane 8: row 1, lane 16: row 2, lane 24: row 3 (rows 2,3 only when upp == 2)
                        const int r = (lane >> 4) * 2 + ((lane >> 3) & 1);
                        const int ul = j * upp + (r >> 1) - rbase;
                        if ((lane & 7) == 0 && r < 2 * upp && ul + rbase < n)
                            part[(ul * 2 + (r & 1)) * kMegaWarps + ks] = k;
                    }
                    cons_count++;
                    __syncwarp();
                    if (lane == 0) {
                        fence_async_smem();
                        produce_one();
                    }
                }
            }
            if (rbase + kRoundUnits >= n) MEGA_STAMP(ev, 2);   // this thread-0 warp finished streaming
            __syncthreads();
            if (rbase + kRoundUnits >= n) MEGA_STAMP(ev, 3);   // whole CTA finished streaming
            // ---- 3. finish the round's units: sum over K slices, fused epilogue -------------------------------
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
                    // head-major cache [L][KVH][S][hd]: element j of kv head h at position pos
                    auto kv_addr = [&](uint8_t* cache, int idx) -> uint8_t* {
                        const int h = idx / p.hd, j = idx - h * p.hd;
                        return cache + ((((size_t)l * p.KVH_loc + h) * p.S + pos) * p.hd + j) * KESZ;
                    };
                    auto store_kv = [&](uint8_t* cache, int idx, float v) {
                        if (KVD == SLLM_BF16) *reinterpret_cast<uint16_t*>(kv_addr(cache, idx)) = f32_to_bf16_bits(v);
                        else *reinterpret_cast<float*>(kv_addr(cache, idx)) = v;
                    };
                    if (u < rope_units) {
                        const int head = u / half, j = u - head * half;
                        const float fci = p.sin_t[(size_t)pos * half + j], fcr = p.cos_t[(size_t)pos * half + j];
                        const float o0 = s0 * fcr - s1 * fci, o1 = s1 * fcr + s0 * fci;   // rope_kernel.cpp:36-37
                        const int r0 = head * p.hd + j;
                        if (r0 < p.q_loc) { p.q[r0] = o0; p.q[r0 + half] = o1; }
                        else { store_kv(p.kc, r0 - p.q_loc, o0); store_kv(p.kc, r0 - p.q_loc + half, o1); }
                    } else {
                        const int b2 = 2 * (u - rope_units);
                        store_kv(p.vc, b2, s0);
                        store_kv(p.vc, b2 + 1, s1);
                    }
                } else if (ph.kind == PH_WO) {
                    const int r = 2 * u;
                    const float h0 = __ldcg(p.x + r) + s0;                    // add_kernel.cpp:10-13
                    p.h[r] = h0;
                    if constexpr (FUSE) p.x[r] = h0;                          // the fused down phase adds its partial sums onto h
                    if (r + 1 < ph.nrows) {
                        const float h1 = __ldcg(p.x + r + 1) + s1;
                        p.h[r + 1] = h1;
                        if constexpr (FUSE) p.x[r + 1] = h1;
                    }
                } else if (ph.kind == PH_GATEUP) {
                    const float sv = (1.0f / (1.0f + expf(-s1))) * s0;        // swiglu_kernel.cpp:12-13
                    p.swi[u] = sv;
                    if constexpr (FUSE) xs[rbase + t] = sv;                   // stays in this CTA for the fused down phase
                } else if (ph.kind == PH_DOWN) {
                    const int r = 2 * u;
                    p.x[r] = s0 + __ldcg(p.h + r);
                    if (r + 1 < ph.nrows) p.x[r + 1] = s1 + __ldcg(p.h + r + 1);
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

        if (ph.kind == PH_CLS) {   // CTA best -> global; the last CTA picks the arg max and advances the step state
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
                if (idx == 0x7fffffff) idx = 0;
                p.st->ticket = 0;
                p.st->pad[0] = (int32_t)(bar_base + bar_idx * (unsigned)ncta);   // barrier epoch for the next launch
                p.blk_val[ncta] = v;
                p.blk_idx[ncta] = idx;
                ClsPolicy<SLLM_F32>::step_feedback(p.st, p.prompt, p.history, idx);
            }
            break;
        }
        MEGA_STAMP(ev, 4);   // epilogue done
        if (FUSE && ph.kind == PH_GATEUP) {   // its consumer is this CTA itself (the round's trailing __syncthreads published xs)
            MEGA_STAMP(ev, 5);
            continue;
        }
        if (p.debug & 1) __syncthreads(); else grid_barrier(p.bar_counter, bar_base + (++bar_idx) * (unsigned)ncta);
        MEGA_STAMP(ev, 5);   // barrier passed
        if (ph.kind != PH_QKV) continue;
        MEGA_STAMP(ev + 1, 0);

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
            const int cpr = row_bytes / 16;                    // 16-byte chunks per K/V row
            const int npos = pos + 1;
            const int per = (npos + p.nsplit - 1) / p.nsplit;
            const int nitems = p.KVH_loc * p.nsplit;
            const float scale = 1.0f / sqrtf((float)p.hd);
            constexpr int kStripes = 16;
            const int pv_chunk = tid % cpr, pv_stripe = tid / cpr;
            const bool pv_active = pv_stripe < kStripes;
            const int key = tid >> 3, kpart = tid & 7;         // 8 threads per key

#pragma unroll 1
            for (int item = cta; item < nitems; item += ncta) {
                const int kvh = item / p.nsplit, split = item - kvh * p.nsplit;
                const int t0 = split * per, t1 = min(npos, t0 + per);
                const int ntiles = (t1 > t0) ? (t1 - t0 + AT - 1) / AT : 0;
                for (int i = tid; i < G * p.hd; i += kMegaThreads) q_s[i] = __ldcg(p.q + (size_t)(kvh * G) * p.hd + i);
                if (tid < G) { ml_s[2 * tid] = -INFINITY; ml_s[2 * tid + 1] = 0.f; }
                const size_t head_off = ((size_t)l * p.KVH_loc + kvh) * p.S * row_bytes;   // [L][KVH][S][hd]
                auto issue_tile = [&](int tile) {              // warp 0: TMA for rows written by EARLIER launches;
                    const int stage = tile & 1;                // the row written this step (pos) is copied by hand
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
                fence_async_smem();   // generic accesses to the K/V region (o_s) precede the TMA writes below
                __syncthreads();      // q_s / ml_s visible; previous readers of this smem are done
                if (warp == 0 && item != cta) {   // tiles 0/1 of the FIRST item were prefetched during phase A
                    if (ntiles > 0) issue_tile(0);
                    if (ntiles > 1) issue_tile(1);
                }
                float acc[G][KVEC];
#pragma unroll
                for (int gi = 0; gi < G; ++gi)
#pragma unroll
                    for (int e = 0; e < KVEC; ++e) acc[gi][e] = 0.f;

#pragma unroll 1
                for (int tile = 0; tile < ntiles; ++tile) {
                    const int stage = tile & 1;
                    const int ts = t0 + tile * AT;
                    const int rows = min(AT, t1 - ts);
                    // the newest row (written by phase A of THIS launch with generic stores) bypasses the async proxy
                    if (pos >= ts && pos < ts + rows && warp == 1) {
                        const size_t g_off = head_off + (size_t)pos * row_bytes;
                        for (int c = lane; c < cpr; c += 32) {
                            *reinterpret_cast<uint4*>(k_s + ((size_t)stage * AT + (pos - ts)) * stride + c * 16) =
                                __ldcg(reinterpret_cast<const uint4*>(p.kc + g_off + c * 16));
                            *reinterpret_cast<uint4*>(v_s + ((size_t)stage * AT + (pos - ts)) * stride + c * 16) =
                                __ldcg(reinterpret_cast<const uint4*>(p.vc + g_off + c * 16));
                        }
                    }
                    if (stage == 0) { mb_wait_fast(att_bar, kv_use0 & 1); kv_use0++; }
                    else { mb_wait_fast(att_bar + 1, kv_use1 & 1); kv_use1++; }
                    __syncthreads();
                    {   // scores
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
                                    for (int e = 0; e < KVEC; ++e) s[gi] = fmaf(qv[e], kf[e], s[gi]);
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
                    for (int gi = warp; gi < G; gi += kMegaWarps) {   // online softmax bookkeeping
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
                            for (int e = 0; e < KVEC; ++e) acc[gi][e] *= al;
                        }
                        for (int r = pv_stripe; r < rows; r += kStripes) {
                            float vf[KVEC];
                            kv_unpack<KVD>(*reinterpret_cast<const uint4*>(v_s + ((size_t)stage * AT + r) * stride + pv_chunk * 16), vf);
#pragma unroll
                            for (int gi = 0; gi < G; ++gi) {
                                const float pr = p_s[gi * kAttTile + r];
#pragma unroll
                                for (int e = 0; e < KVEC; ++e) acc[gi][e] = fmaf(pr, vf[e], acc[gi][e]);
                            }
                        }
                    }
                    fence_async_smem();
                    __syncthreads();
                    if (warp == 0 && tile + 2 < ntiles) issue_tile(tile + 2);
                }
                // cross-stripe reduction through the (drained) K stages, then the partial record
                float* o_s = reinterpret_cast<float*>(k_s);   // [kStripes][G][hd]
                if (pv_active) {
#pragma unroll
                    for (int gi = 0; gi < G; ++gi)
#pragma unroll
                        for (int e = 0; e < KVEC; ++e) o_s[((size_t)pv_stripe * G + gi) * p.hd + pv_chunk * KVEC + e] = acc[gi][e];
                }
                __syncthreads();
                const int rec = p.hd + kAttRecPad;
                for (int i = tid; i < G * p.hd; i += kMegaThreads) {
                    float o = 0.f;
                    for (int s = 0; s < kStripes; ++s) o += o_s[(size_t)s * G * p.hd + i];
                    const int gi = i / p.hd, j = i - gi * p.hd;
                    p.att_part[((size_t)(kvh * G + gi) * p.nsplit + split) * rec + j] = o;
                }
                if (tid < G) {
                    float* r = p.att_part + ((size_t)(kvh * G + tid) * p.nsplit + split) * rec + p.hd;
                    r[0] = ml_s[2 * tid];
                    r[1] = ml_s[2 * tid + 1];
                }
            }
            fence_async_smem();
        }
        MEGA_STAMP(ev + 1, 4);
        if (p.debug & 1) __syncthreads(); else grid_barrier(p.bar_counter, bar_base + (++bar_idx) * (unsigned)ncta);
        MEGA_STAMP(ev + 1, 5);
    }
}

// ------------------------------------------------------------------------------------------- host ----
TileGeom mega_tile_geom(int rows_phys, int cols, int w_dtype) {
    TileGeom g;
    const int E = w_dtype == SLLM_F32 ? 4 : w_dtype == SLLM_BF16 ? 8 : 16;
    g.nchunks = cols / E;
    int KS = 1;
    while (KS < 16 && (g.nchunks + KS - 1) / KS * 16 > 1024) KS *= 2;
    g.KS = KS;
    g.SC = (g.nchunks + KS - 1) / KS;
    g.srow = 0;
    if (w_dtype == SLLM_INT8) {   // a quantisation group = 64 weights = 4 chunks must not straddle K slices; SC/4 scales per row
        g.SC = (g.SC + 3) / 4 * 4;
        g.srow = (g.SC + 15) / 16 * 16;
    }
    g.R = ((g.SC * 16 + g.srow) * 4 <= kSlotBytes) ? 4 : 2;
    g.ntr = (rows_phys + g.R - 1) / g.R;
    g.tile_bytes = g.R * (g.SC * 16 + g.srow);
    g.bytes = (size_t)g.ntr * g.KS * g.tile_bytes;
    return g;
}

static int phys_rows(int rows, int kind) { return kind == PH_GATEUP ? rows : ((rows + 1) / 2) * 2; }

size_t mega_matrix_bytes(int rows, int cols, int kind, int w_dtype) { return mega_tile_geom(phys_rows(rows, kind), cols, w_dtype).bytes; }

static void fill_desc(PhaseDesc& ds, const void* W, int rows, int cols, int kind, int layer, int w_dtype) {
    const TileGeom g = mega_tile_geom(phys_rows(rows, kind), cols, w_dtype);
    ds.W = reinterpret_cast<const uint8_t*>(W);
    ds.nchunks = g.nchunks;
    ds.nrows = rows;
    ds.kind = kind;
    ds.layer = layer;
    ds.nunits = (kind == PH_GATEUP) ? rows / 2 : (rows + 1) / 2;
    ds.KS = g.KS;
    ds.SC = g.SC;
    ds.R = g.R;
    ds.ntr = g.ntr;
    ds.tile_bytes = g.tile_bytes;
    ds.srow = g.srow;
    ds.cum = nullptr;
}

// can the megakernel run this shape? (everything else keeps using the per-kernel fused path)
MegaPlan mega_plan(int w_dtype, int group, int kv_dtype, int d, int hd, int q_loc, int kv_loc, int I_loc, int V_loc, int H_loc, int KVH_loc, int max_len) {
    return mega_plan_for(sm_count(), smem_optin_bytes(), w_dtype, group, kv_dtype, d, hd, q_loc, kv_loc, I_loc, V_loc, H_loc, KVH_loc, max_len);
}

// the plan as a pure function of the device facts (SM count, opt-in shared memory per block): testable without a device
MegaPlan mega_plan_for(int sms, int smem_optin, int w_dtype, int group, int kv_dtype, int d, int hd, int q_loc, int kv_loc, int I_loc, int V_loc,
                       int H_loc, int KVH_loc, int max_len) {
    MegaPlan pl;
    if (w_dtype == SLLM_INT8 && group != 64) { pl.why = "int8 group size other than 64"; return pl; }
    const int E = w_dtype == SLLM_F32 ? 4 : w_dtype == SLLM_BF16 ? 8 : 16;
    for (int cols : {d, q_loc, I_loc}) {
        if (cols % E || (w_dtype == SLLM_INT8 && cols % 64)) { pl.why = "row length not a multiple of 16 bytes / of the int8 group"; return pl; }
        const TileGeom tg = mega_tile_geom(2, cols, w_dtype);
        if (tg.KS * tg.SC < tg.nchunks || (tg.SC * 16 + tg.srow) * 2 > kSlotBytes || tg.SC > 32 * (w_dtype == SLLM_INT8 ? 2 : kCplMax)) {
            pl.why = "rows longer than 32 KB";
            return pl;
        }
    }
    const int g = H_loc / KVH_loc;
    if ((g != 1 && g != 2 && g != 4 && g != 8) || hd % 16 || hd > 256) { pl.why = "head shape"; return pl; }   // the instantiated query-head group sizes (mega_launch)
    const int kesz = kv_dtype == SLLM_F32 ? 4 : 2;
    if ((hd * kesz / 16) > 32) { pl.why = "head_dim chunking"; return pl; }
    const MegaSmem SL = mega_smem_layout(hd, g, kesz);
    if (SL.total > (size_t)smem_optin) { pl.why = "shared memory"; return pl; }
    if ((size_t)16 * g * hd * 4 > (size_t)4 * SL.att_tile * SL.kv_stride) { pl.why = "attention scratch"; return pl; }   // the cross-stripe reduction buffer spans the (drained, contiguous) K and V stages
    if ((size_t)std::max(d, I_loc) * 4 > (size_t)4 * SL.att_tile * SL.kv_stride) { pl.why = "activation staging"; return pl; }
    if (q_loc % 2 || kv_loc % 2) { pl.why = "odd dims"; return pl; }
    pl.grid = sms;
    pl.smem = SL.total;
    // splits: fill the grid once, never more splits than 64-position tiles at full context
    int ns = pl.grid / KVH_loc;
    const int by_len = (max_len + kAttTile - 1) / kAttTile;
    if (ns > by_len) ns = by_len;
    if (ns > 32) ns = 32;
    pl.nsplit = ns < 1 ? 1 : ns;
    (void)V_loc;
    pl.att_part_floats = (size_t)H_loc * pl.nsplit * (hd + kAttRecPad);
    pl.ok = true;
    return pl;
}

void mega_fill_phases(PhaseDesc* host, int L, int w_dtype, const void* wqkv, const void* wo, const void* wug, const void* wdown,
                      const void* cls, int d, int q_loc, int kv_loc, int I_loc, int V_loc) {
    auto layer_ptr = [&](const void* base, int rows, int cols, int kind, int l) {
        return reinterpret_cast<const uint8_t*>(base) + (size_t)l * mega_tile_geom(phys_rows(rows, kind), cols, w_dtype).bytes;
    };
    for (int l = 0; l < L; ++l) {
        fill_desc(host[4 * l + 0], layer_ptr(wqkv, q_loc + 2 * kv_loc, d, PH_QKV, l), q_loc + 2 * kv_loc, d, PH_QKV, l, w_dtype);
        fill_desc(host[4 * l + 1], layer_ptr(wo, d, q_loc, PH_WO, l), d, q_loc, PH_WO, l, w_dtype);
        fill_desc(host[4 * l + 2], layer_ptr(wug, 2 * I_loc, d, PH_GATEUP, l), 2 * I_loc, d, PH_GATEUP, l, w_dtype);
        fill_desc(host[4 * l + 3], layer_ptr(wdown, d, I_loc, PH_DOWN, l), d, I_loc, PH_DOWN, l, w_dtype);
    }
    fill_desc(host[4 * L], cls, V_loc, d, PH_CLS, L, w_dtype);
}

// ---- row-major -> tiled, one thread per 16-byte chunk of the destination (padding chunks are zeroed)
__global__ void repack_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int rows, int nchunks, int kind, int KS, int SC, int R,
                              int ntr, int hd, int q_loc, int kv_loc, int I_loc) {
    const int64_t total = (int64_t)ntr * KS * R * SC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % SC);
        const int rr = (int)((i / SC) % R);
        const int ks = (int)((i / ((int64_t)SC * R)) % KS);
        const int g = (int)(i / ((int64_t)SC * R * KS));
        const int pr = g * R + rr;                 // physical row = 2*unit + member
        const int u = pr >> 1, m = pr & 1;
        int r0, r1;
        if (kind == PH_QKV) {
            const int half = hd >> 1, rope_units = (q_loc + kv_loc) >> 1;
            if (u < rope_units) { const int head = u / half; r0 = head * hd + (u - head * half); r1 = r0 + half; }
            else { r0 = q_loc + kv_loc + 2 * (u - rope_units); r1 = r0 + 1; }
        } else if (kind == PH_GATEUP) { r0 = u; r1 = I_loc + u; }
        else { r0 = 2 * u; r1 = r0 + 1; }
        const int row = m ? r1 : r0;
        const int c = ks * SC + cc;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < rows && c < nchunks) v = src[(int64_t)row * nchunks + c];
        dst[i] = v;
    }
}

// int8: a tile = R rows of SC 16-byte chunks, then R rows of srow bytes of group scales (scale j of a row = group ks*SC/4 + j)
__global__ void repack_int8_kernel(const uint4* __restrict__ src, const float* __restrict__ scales, uint4* __restrict__ dst, int rows, int nchunks,
                                   int kind, int KS, int SC, int R, int srow, int ntr, int hd, int q_loc, int kv_loc, int I_loc) {
    const int wu = R * SC, su = R * (srow / 16), tu = wu + su;   // 16-byte units per tile: weights, scales
    const int64_t total = (int64_t)ntr * KS * tu;
    const int ngroups = nchunks / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int unit = (int)(i % tu);
        const int ks = (int)((i / tu) % KS);
        const int g = (int)(i / ((int64_t)tu * KS));
        const bool is_scale = unit >= wu;
        const int rr = is_scale ? (unit - wu) / (srow / 16) : unit / SC;
        const int cc = is_scale ? (unit - wu) % (srow / 16) : unit % SC;
        const int pr = g * R + rr;
        const int u = pr >> 1, m = pr & 1;
        int r0, r1;
        if (kind == PH_QKV) {
            const int half = hd >> 1, rope_units = (q_loc + kv_loc) >> 1;
            if (u < rope_units) { const int head = u / half; r0 = head * hd + (u - head * half); r1 = r0 + half; }
            else { r0 = q_loc + kv_loc + 2 * (u - rope_units); r1 = r0 + 1; }
        } else if (kind == PH_GATEUP) { r0 = u; r1 = I_loc + u; }
        else { r0 = 2 * u; r1 = r0 + 1; }
        const int row = m ? r1 : r0;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (!is_scale) {
            const int c = ks * SC + cc;
            if (row < rows && c < nchunks) v = src[(int64_t)row * nchunks + c];
        } else if (row < rows) {
            float f[4];
            for (int k = 0; k < 4; ++k) {
                const int gi = ks * (SC / 4) + cc * 4 + k;
                f[k] = (cc * 4 + k < SC / 4 && gi < ngroups) ? scales[(int64_t)row * ngroups + gi] : 0.f;
            }
            v = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
        }
        dst[i] = v;
    }
}

int mega_repack(const void* src, const float* scales, void* dst, int rows, int cols, int kind, int w_dtype, int hd, int q_loc, int kv_loc, int I_loc,
                cudaStream_t st) {
    const TileGeom g = mega_tile_geom(phys_rows(rows, kind), cols, w_dtype);
    if (w_dtype == SLLM_INT8) {
        SLLM_REQUIRE(scales, SLLM_EINVAL, "mega_repack: int8 weights without scales");
        const int64_t total = (int64_t)g.ntr * g.KS * (g.tile_bytes / 16);
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
        repack_int8_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), scales, reinterpret_cast<uint4*>(dst), rows, g.nchunks, kind,
                                                   g.KS, g.SC, g.R, g.srow, g.ntr, hd, q_loc, kv_loc, I_loc);
        g_launches++;
        SLLM_LAUNCH_CHECK();
        return SLLM_OK;
    }
    const int64_t total = (int64_t)g.ntr * g.KS * g.R * g.SC;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    repack_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), rows, g.nchunks, kind, g.KS, g.SC,
                                          g.R, g.ntr, hd, q_loc, kv_loc, I_loc);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

// ---- transposed down matrix of the fused kernel (megakernel.cuh "PH_DOWN_T") ---------------------------------------------
bool mega_fuse_down_ok(int w_dtype, int d, int I_loc) {
    if (w_dtype != SLLM_F32 && w_dtype != SLLM_BF16) return false;
    const int row_bytes = d * (w_dtype == SLLM_F32 ? 4 : 2);
    if (row_bytes % 512) return false;
    const int stripes = row_bytes / 512;                       // one stripe = 32 lanes x 16 bytes of outputs
    if (stripes < 1 || stripes > kMegaWarps || (stripes & (stripes - 1))) return false;
    return I_loc % kFuseJT == 0 && I_loc >= kFuseJT;
}
size_t mega_down_t_bytes(int d, int I_loc, int w_dtype) { return (size_t)d * I_loc * (w_dtype == SLLM_F32 ? 4 : 2); }

void mega_fill_down_t(PhaseDesc& ds, const void* Wt, int d, int I_loc, int layer, int w_dtype) {
    const int esz = w_dtype == SLLM_F32 ? 4 : 2;
    ds.W = reinterpret_cast<const uint8_t*>(Wt);
    ds.nchunks = d * esz / 16;
    ds.nunits = I_loc;                 // inputs j
    ds.nrows = d;
    ds.kind = PH_DOWN_T;
    ds.layer = layer;
    ds.KS = d * esz / 512;             // stripes
    ds.SC = 32;
    ds.R = kFuseJT;
    ds.ntr = I_loc / kFuseJT;
    ds.tile_bytes = kFuseJT * 512;
    ds.srow = 0;
    ds.cum = nullptr;
}

template <class T>
__global__ void repack_down_t_kernel(const T* __restrict__ src, T* __restrict__ dst, int d, int I, int E, int KS) {
    const int64_t total = (int64_t)d * I;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ntr = I / kFuseJT;                                   // stripe-major: [ks][g][jj][lane][e]
        const int e = (int)(i % E);
        const int lane = (int)((i / E) % 32);
        const int jj = (int)((i / ((int64_t)E * 32)) % kFuseJT);
        const int64_t g = (i / ((int64_t)E * 32 * kFuseJT)) % ntr;
        const int ks = (int)(i / ((int64_t)E * 32 * kFuseJT * ntr));
        const int r = (ks * 32 + lane) * E + e;
        const int64_t j = g * kFuseJT + jj;
        dst[i] = src[(int64_t)r * I + j];
    }
}

int mega_repack_down_t(const void* src, void* dst, int d, int I_loc, int w_dtype, cudaStream_t st) {
    SLLM_REQUIRE(mega_fuse_down_ok(w_dtype, d, I_loc), SLLM_ENOTSUP, "fused down projection: shape d=%d I=%d weight type %d not supported", d, I_loc, w_dtype);
    const int64_t total = (int64_t)d * I_loc;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    if (w_dtype == SLLM_F32)
        repack_down_t_kernel<uint32_t><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(src), reinterpret_cast<uint32_t*>(dst), d, I_loc, 4, d * 4 / 512);
    else
        repack_down_t_kernel<uint16_t><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(src), reinterpret_cast<uint16_t*>(dst), d, I_loc, 8, d * 2 / 512);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

template <int WD, int KVD, int G, bool FUSE = false>
static int mega_launch_t(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    static size_t configured = 0;
    if (smem > configured) {
        SLLM_CUDA(cudaFuncSetAttribute(mega_step_kernel<WD, KVD, G, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kMegaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SLLM_CUDA(cudaLaunchKernelEx(&cfg, mega_step_kernel<WD, KVD, G, FUSE>, p));
    g_launches++;
    return SLLM_OK;
}

extern int g_tune_mega_debug;
int mega_launch(const MegaParams& p_in, int g, int grid, size_t smem, cudaStream_t st, bool fuse_down) {
    MegaParams p = p_in;
    p.debug = g_tune_mega_debug;
    if (fuse_down) {   // experimental kernel: fp32 / bf16 weights only (mega_fuse_down_ok)
#define MEGA_F(GG)                                                                                                      \
    case GG:                                                                                                            \
        if (p.w_dtype == SLLM_F32) {                                                                                    \
            return p.kv_dtype == SLLM_F32 ? mega_launch_t<SLLM_F32, SLLM_F32, GG, true>(p, grid, smem, st)              \
                                          : mega_launch_t<SLLM_F32, SLLM_BF16, GG, true>(p, grid, smem, st);            \
        }                                                                                                               \
        if (p.w_dtype == SLLM_BF16) {                                                                                   \
            return p.kv_dtype == SLLM_F32 ? mega_launch_t<SLLM_BF16, SLLM_F32, GG, true>(p, grid, smem, st)             \
                                          : mega_launch_t<SLLM_BF16, SLLM_BF16, GG, true>(p, grid, smem, st);           \
        }                                                                                                               \
        break;
        switch (g) {
            MEGA_F(1)
            MEGA_F(2)
            MEGA_F(4)
            MEGA_F(8)
            default: break;
        }
#undef MEGA_F
        set_error("megakernel with the fused down projection: weight type %d / %d query heads per KV head not instantiated", p.w_dtype, g);
        return SLLM_ENOTSUP;
    }
#define MEGA_G(GG)                                                                                        \
    case GG:                                                                                              \
        if (p.w_dtype == SLLM_F32) {                                                                      \
            return p.kv_dtype == SLLM_F32 ? mega_launch_t<SLLM_F32, SLLM_F32, GG>(p, grid, smem, st)      \
                                          : mega_launch_t<SLLM_F32, SLLM_BF16, GG>(p, grid, smem, st);    \
        }                                                                                                 \
        if (p.w_dtype == SLLM_INT8) {                                                                     \
            return p.kv_dtype == SLLM_F32 ? mega_launch_t<SLLM_INT8, SLLM_F32, GG>(p, grid, smem, st)     \
                                          : mega_launch_t<SLLM_INT8, SLLM_BF16, GG>(p, grid, smem, st);   \
        }                                                                                                 \
        return p.kv_dtype == SLLM_F32 ? mega_launch_t<SLLM_BF16, SLLM_F32, GG>(p, grid, smem, st)         \
                                      : mega_launch_t<SLLM_BF16, SLLM_BF16, GG>(p, grid, smem, st);
    switch (g) {
        MEGA_G(1)
        MEGA_G(2)
        MEGA_G(4)
        MEGA_G(8)
        default: set_error("megakernel: %d query heads per KV head not instantiated", g); return SLLM_ENOTSUP;
    }
#undef MEGA_G
}

}  // namespace sllm

// ---- the host-side plan through the C ABI (pure arithmetic, no launch, no device needed when the facts are passed in) ----
extern "C" {

int sllm_mega_plan(const sllm_shape* shape, int32_t w_dtype, int32_t group, int32_t kv_dtype, int32_t tp_size, int32_t word_based,
                   int32_t sm_count_or_0, int32_t smem_optin_or_0, int32_t* ok, int32_t* grid, int64_t* smem_bytes, int32_t* nsplit) {
    using namespace sllm;
    SLLM_REQUIRE(shape && ok, SLLM_EINVAL, "mega_plan: null argument");
    SLLM_REQUIRE(w_dtype == SLLM_F32 || w_dtype == SLLM_BF16 || w_dtype == SLLM_INT8, SLLM_EINVAL, "mega_plan: weight dtype %d", w_dtype);
    SLLM_REQUIRE(kv_dtype == SLLM_F32 || kv_dtype == SLLM_BF16, SLLM_EINVAL, "mega_plan: kv dtype %d", kv_dtype);
    SLLM_REQUIRE(tp_size >= 1 && shape->heads > 0 && shape->kv_heads > 0 && shape->head_dim > 0 && shape->heads % shape->kv_heads == 0 &&
                     shape->heads % tp_size == 0 && shape->kv_heads % tp_size == 0 && shape->inter % tp_size == 0 && shape->vocab % tp_size == 0,
                 SLLM_EINVAL, "mega_plan: heads=%d kv_heads=%d inter=%d vocab=%d must divide by tp_size=%d", shape->heads, shape->kv_heads,
                 shape->inter, shape->vocab, tp_size);
    const int H_loc = shape->heads / tp_size, KVH_loc = shape->kv_heads / tp_size;
    const int sms = sm_count_or_0 > 0 ? sm_count_or_0 : sm_count(), smem = smem_optin_or_0 > 0 ? smem_optin_or_0 : smem_optin_bytes();
    SLLM_REQUIRE(tp_size == 1 || word_based, SLLM_EINVAL, "mega_plan: under tensor parallelism only the word-based kernel runs (it carries the all-reduce)");
    if (word_based) {   // megakernel_ll.cu; rank 0's shard (v0 = 0: every rank's vocabulary offset is a multiple of the same V_loc)
        const MegaLLPlan ll = mega_ll_plan_for(sms, smem, w_dtype, kv_dtype, shape->hidden, shape->head_dim, H_loc * shape->head_dim,
                                               KVH_loc * shape->head_dim, shape->inter / tp_size, shape->vocab / tp_size, 0, H_loc, KVH_loc,
                                               shape->max_len, tp_size);
        if (ll.ok && w_dtype == SLLM_INT8 && group != 64) { *ok = 0; set_error("the word-based megakernel cannot take this shape: int8 group size other than 64"); return SLLM_OK; }
        *ok = ll.ok ? 1 : 0;
        if (grid) *grid = ll.grid;
        if (smem_bytes) *smem_bytes = (int64_t)ll.smem;
        if (nsplit) *nsplit = ll.nsplit;
        if (!ll.ok) set_error("the word-based megakernel cannot take this shape: %s", ll.why);
        return SLLM_OK;
    }
    const MegaPlan pl = mega_plan_for(sms, smem, w_dtype, group, kv_dtype, shape->hidden, shape->head_dim, H_loc * shape->head_dim,
                                      KVH_loc * shape->head_dim, shape->inter / tp_size, shape->vocab / tp_size, H_loc, KVH_loc, shape->max_len);
    *ok = pl.ok ? 1 : 0;
    if (grid) *grid = pl.grid;
    if (smem_bytes) *smem_bytes = (int64_t)pl.smem;
    if (nsplit) *nsplit = pl.nsplit;
    if (!pl.ok) set_error("megakernel cannot take this shape: %s", pl.why);
    return SLLM_OK;
}

/* development aid: the transposed, stripe-major down matrix of the fused kernel from a row-major [d][inter] one (device pointers) */
int sllm_mega_repack_down_t(const void* src_rowmajor, void* dst, int32_t d, int32_t inter, int32_t w_dtype, sllm_stream_t stream) {
    using namespace sllm;
    SLLM_REQUIRE(src_rowmajor && dst, SLLM_EINVAL, "mega_repack_down_t: null pointer");
    return mega_repack_down_t(src_rowmajor, dst, d, inter, w_dtype, as_stream(stream));
}

int sllm_mega_tile_geometry(int32_t rows, int32_t cols, int32_t kind, int32_t w_dtype, int32_t* ks, int32_t* sc, int32_t* r, int32_t* tile_rows,
                            int32_t* tile_bytes, int64_t* matrix_bytes) {
    using namespace sllm;
    SLLM_REQUIRE(rows >= 1 && cols >= 1 && kind >= PH_QKV && kind <= PH_CLS, SLLM_EINVAL, "mega_tile_geometry: rows=%d cols=%d kind=%d", rows, cols, kind);
    SLLM_REQUIRE(w_dtype == SLLM_F32 || w_dtype == SLLM_BF16 || w_dtype == SLLM_INT8, SLLM_EINVAL, "mega_tile_geometry: weight dtype %d", w_dtype);
    const int E = w_dtype == SLLM_F32 ? 4 : w_dtype == SLLM_BF16 ? 8 : 16;
    SLLM_REQUIRE(cols % E == 0, SLLM_EINVAL, "mega_tile_geometry: cols=%d is not a multiple of %d (16 bytes of weights)", cols, E);
    const TileGeom g = mega_tile_geom(phys_rows(rows, kind), cols, w_dtype);
    SLLM_REQUIRE(g.KS * g.SC >= g.nchunks && g.tile_bytes <= kSlotBytes, SLLM_ENOTSUP,
                 "mega_tile_geometry: rows of %d columns are longer than 32 KB (a two-row tile must fit a %d-byte ring slot)", cols, kSlotBytes);
    if (ks) *ks = g.KS;
    if (sc) *sc = g.SC;
    if (r) *r = g.R;
    if (tile_rows) *tile_rows = g.ntr;
    if (tile_bytes) *tile_bytes = g.tile_bytes;
    if (matrix_bytes) *matrix_bytes = (int64_t)g.bytes;
    return SLLM_OK;
}

}  // extern "C"
