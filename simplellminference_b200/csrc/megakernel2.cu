// csrc/megakernel2.cu — the decode step as one persistent cooperative kernel with TWO (or three) grid-wide dependency points per
// layer instead of five (sm_100a). Successor of megakernel.cu (kept: deterministic sums, int8 weights); same rings, same tiles.
//
// What bounds batch-1 decode on a B200 is not bandwidth while a matrix streams (the rings deliver ~7 TB/s) but the chain of
// dependency points between the matrices: at each one every SM drains its pipe, waits for the slowest of 148 CTAs, crosses a grid
// barrier and stages a new activation vector — ~5-6 us during which HBM idles once 128 KB per SM are in flight (profiles/r02_*).
// A layer of megakernel.cu has five (qkv | attention | wo | gate_up | down). Here:
//
//   qkv -> attention   no grid barrier: a (kv head, split) attention item depends on the ~6 CTAs that produced its q / k / v rows,
//                      not on all 148. Every CTA adds the number of units it finished to a per-kv-head counter (red.release);
//                      the item's CTA polls that one counter (ld.acquire).
//   attention -> wo    no grid barrier and no wo phase over all CTAs: the CTAs of a kv head's attention items multiply THAT head
//                      group's columns of Wo (a K split of the projection): they wait for the group's splits through a second
//                      counter, merge the partial (m, l, O) records of the group (a few hundred values instead of the whole
//                      vector), stream their row range of the group's [d][G*hd] block from a column-block copy of Wo ("WoT",
//                      4 KB tiles of whole rows) and add their partial outputs into h with red.global.add.f32. The residual x
//                      is added by exactly one group per tile. h starts from zero: double-buffered, zeroed a layer ahead.
//   gate_up -> down    optionally (FUSE) as in megakernel.cu's experimental variant: the CTA multiplies the columns of Wdown that
//                      belong to the sigmoid(gate)*up values it just produced and adds into x (zeroed a phase ahead; CTA 0 adds h).
//
// So a layer has grid barriers only where the data flow is all-to-all: after wo (h complete -> RMSNorm -> gate_up) and after down
// (x complete -> RMSNorm -> qkv). Summation order of wo (and of the fused down) is not fixed: logits move in the last bits from
// run to run, inside the decode tolerance; tokens are checked against the oracle like every other path.
#include <algorithm>
#include <cmath>

#include "mega_common.cuh"

namespace sllm {

constexpr int kTraceEvents2 = 512;
__device__ __forceinline__ unsigned long long gtime2() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define M2_STAMP(ev, slot)                                                                                \
    do {                                                                                                  \
        if (p.trace && threadIdx.x == 0 && (ev) < kTraceEvents2) p.trace[((size_t)blockIdx.x * kTraceEvents2 + (ev)) * 8 + (slot)] = gtime2(); \
    } while (0)

__device__ __forceinline__ void flag_add_release(unsigned* f, unsigned n) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(f), "r"(n) : "memory");
}
__device__ __forceinline__ void flag_wait(const unsigned* f, unsigned target) {   // ONE thread; callers follow with __syncthreads()
    unsigned v, spins = 0;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (++spins > kSpinLimit) mega_timeout("dependency counter");
    } while ((int)(v - target) < 0);
}
__device__ __forceinline__ void red_add_f32(float* dst, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory"); }

// N values per lane -> N/2: lanes with `bit` set keep (and complete) the upper half, the others the lower half
template <int N>
__device__ __forceinline__ void halve(float (&v)[8], int lane, int bit) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        const float keep = up ? v[k + N / 2] : v[k];
        const float send = up ? v[k] : v[k + N / 2];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
}

// One 4 KB WoT tile = 256 chunks of 16 bytes = NR = 256 / CRP whole rows of CRP chunks. Lane `lane` multiplies chunks lane + 32 i
// (i < 8; a warp pass reads 512 contiguous bytes: conflict-free) with its part of the merged attention output held in registers,
// the row sums are completed with a value-halving butterfly. Returns true in the lanes that end up owning a row sum.
template <int WD, int CRP>
__device__ __forceinline__ bool wot_tile(const uint8_t* slot, int lane, const float (*xw)[WInfo<WD>::E], float& val, int& row) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        constexpr int CPLW = CRP > 32 ? CRP / 32 : 1;
        v[i] = reg_dot<WD>(*reinterpret_cast<const uint4*>(slot + (size_t)(lane + 32 * i) * 16), xw[i % CPLW], 0.f);
    }
    if constexpr (CRP <= 32) {                 // v[i] <-> row lane / CRP + i * (32 / CRP); sum over the CRP lanes of a segment
        halve<8>(v, lane, CRP / 2);
        halve<4>(v, lane, CRP / 4);
        halve<2>(v, lane, CRP / 8);
#pragma unroll
        for (int o = CRP / 16; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        const int i = ((lane & (CRP / 2)) ? 4 : 0) + ((lane & (CRP / 4)) ? 2 : 0) + ((lane & (CRP / 8)) ? 1 : 0);
        row = lane / CRP + i * (32 / CRP);
        val = v[0];
        return (lane & (CRP / 8 - 1)) == 0;
    } else if constexpr (CRP == 64) {          // rows i / 2: two chunks of a row per lane
        v[0] += v[1]; v[1] = v[2] + v[3]; v[2] = v[4] + v[5]; v[3] = v[6] + v[7];
        halve<4>(v, lane, 16);
        halve<2>(v, lane, 8);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        row = ((lane & 16) ? 2 : 0) + ((lane & 8) ? 1 : 0);
        val = v[0];
        return (lane & 7) == 0;
    } else {                                   // CRP == 128: rows i / 4
        v[0] = (v[0] + v[1]) + (v[2] + v[3]);
        v[1] = (v[4] + v[5]) + (v[6] + v[7]);
        halve<2>(v, lane, 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        row = (lane & 16) ? 1 : 0;
        val = v[0];
        return (lane & 15) == 0;
    }
}

constexpr int kMaxGroups = 64;
constexpr int kPhaseCache = 132;   // phase descriptors kept in shared memory (4 * 32 layers + classifier + slack); deeper models read the rest from global memory

template <int WD, int KVD, int G, bool FUSE>
__global__ void __launch_bounds__(kMegaThreads, 1) mega2_step_kernel(const Mega2Params P) {
    const MegaParams& p = P.m;
    constexpr int E = WInfo<WD>::E;                 // weights per 16-byte chunk
    constexpr int CPL = kCplMax;                    // max chunks per lane per row slice
    constexpr int KESZ = MKv<KVD>::ESZ, KVEC = MKv<KVD>::VEC;
    extern __shared__ __align__(128) uint8_t mega_smem[];
    uint8_t* const smem = mega_smem;
    const MegaSmem SL = mega_smem_layout(p.hd, G, KESZ);
    const int AT = SL.att_tile;                      // cache positions per K/V stage (mega_common.cuh)
    uint64_t* ring_bar = reinterpret_cast<uint64_t*>(smem + SL.bars);            // [16][kSlots]
    uint64_t* att_bar = ring_bar + kMegaWarps * kSlots;                          // [3]: K/V stages 0, 1 and the tail stage of the single-pass attention
    float* red = reinterpret_cast<float*>(smem + SL.red);
    float* part = reinterpret_cast<float*>(smem + SL.part);                      // [kRoundUnits][2][16]
    uint8_t* ring = smem + SL.ring;
    float* xs = reinterpret_cast<float*>(smem + SL.att_k);    // activation staging: aliases the (idle) K/V stages
    __shared__ int s_last, s_npub;
    __shared__ int s_pub_g[kMaxGroups], s_pub_n[kMaxGroups];
    // the phase descriptors in shared memory: every phase and every ring producer reads one right after a dependency point, where a
    // global load (an L2 round trip with dependent addresses behind it) sits on the critical path: +1.1 % tokens/s (A/B on one box)
    __shared__ __align__(16) PhaseDesc s_ph[kPhaseCache];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int nwp = 4 * p.L + 1;

    {
        static_assert(sizeof(PhaseDesc) % 16 == 0, "PhaseDesc is copied in 16-byte words");
        const int n16 = min(nwp, kPhaseCache) * (int)(sizeof(PhaseDesc) / 16);
        for (int i = tid; i < n16; i += kMegaThreads) reinterpret_cast<uint4*>(s_ph)[i] = __ldg(reinterpret_cast<const uint4*>(p.phases) + i);
    }
    if (tid < kMegaWarps * kSlots + 3) mb_init(ring_bar + tid, 1);
    if (tid == 0) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_npub = 0;
    }
    __syncthreads();

    if (p.trace && threadIdx.x == 0) {   // which SM this CTA runs on (event 0, slot 7)
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[((size_t)blockIdx.x * kTraceEvents2) * 8 + 7] = smid;
    }
    auto phase = [&](int i) -> PhaseDesc { return i < kPhaseCache ? s_ph[i] : p.phases[i]; };
    const int pos = p.st->pos;
    const int token = min(max(p.st->token, 0), p.V - 1);
    const unsigned bar_base = (unsigned)p.st->pad[0];
    const unsigned epoch = (unsigned)p.st->pad[2];          // launches of THIS kernel so far: every counter below is monotonic
    unsigned bar_idx = 0;
    const int half = p.hd >> 1;
    const int upg = (G + 2) * half;                         // qkv units of one kv head group: G q heads, its k head, its v head
    const int nitems = p.KVH_loc * p.nsplit;                // (kv head, split) items == wo items; nitems <= ncta (plan)
    const unsigned qbase = epoch * (unsigned)(p.L * upg), abase = epoch * (unsigned)(p.L * p.nsplit);
    const unsigned gl0 = epoch * (unsigned)p.L;             // global layer index of this launch's layer 0 (buffer parity)
    unsigned* const qkv_cnt = P.flags;                      // [KVH_loc] counters, one per 128-byte line
    unsigned* const att_cnt = P.flags + (size_t)p.KVH_loc * 32;
    const int my_kvh = cta / p.nsplit, my_part = cta - my_kvh * p.nsplit;   // this CTA's item (meaningful when cta < nitems)
    const bool has_item = cta < nitems;

    // which kv head groups this CTA's qkv units belong to (the same in every layer): group g owns q units [g*G*half, (g+1)*G*half),
    // k units q_units + [g*half, (g+1)*half), v units rope_units + [g*half, (g+1)*half)
    {
        const PhaseDesc ph0 = phase(0);
        int g0, g1;
        cta_tiles(ph0, cta, ncta, g0, g1);
        const int upp = ph0.R >> 1;
        const int a = g0 * upp, b = min(ph0.nunits, g1 * upp);
        if (tid < p.KVH_loc && b > a) {
            const int q_units = p.q_loc >> 1, rope_units = (p.q_loc + p.kv_loc) >> 1;
            auto ov = [&](int lo, int hi) { return max(0, min(b, hi) - max(a, lo)); };
            const int n = ov(tid * G * half, (tid + 1) * G * half) + ov(q_units + tid * half, q_units + (tid + 1) * half) +
                          ov(rope_units + tid * half, rope_units + (tid + 1) * half);
            if (n > 0) {
                const int i = atomicAdd(&s_npub, 1);
                s_pub_g[i] = tid;
                s_pub_n[i] = n;
            }
        }
    }
    __syncthreads();

    uint8_t* my_ring = ring + (size_t)warp * kSlots * kSlotBytes;
    uint64_t* my_bar = ring_bar + warp * kSlots;

    // ---------------- producer (lane 0 of every warp): one tile == one bulk copy -------------------------------
    int pr_wp = -1, pr_left = 0;
    unsigned pr_count = 0;
    const uint8_t* pr_ptr = nullptr;
    uint32_t pr_step = 0, pr_bytes = 0, pr_tail = 0;
    auto produce_one = [&]() {
        while (pr_left <= 0) {
            if (++pr_wp >= nwp) { pr_wp = nwp; return; }
            const PhaseDesc ph = phase(pr_wp);
            if (ph.kind == PH_WO_T) {   // this CTA's row range of its kv head group's block: tiles T0 + warp, + 16, ...
                if (!has_item) continue;
                const int T0 = (int)(((int64_t)ph.ntr * my_part) / p.nsplit), T1 = (int)(((int64_t)ph.ntr * (my_part + 1)) / p.nsplit);
                pr_left = (T1 - T0 - warp + kMegaWarps - 1) / kMegaWarps;
                pr_ptr = ph.W + ((size_t)my_kvh * ph.ntr + T0 + warp) * ph.tile_bytes;
                pr_step = (uint32_t)kMegaWarps * ph.tile_bytes;
                pr_bytes = (uint32_t)ph.tile_bytes;
                pr_tail = 0;
                continue;
            }
            const int ks = warp & (ph.KS - 1), rg = warp / ph.KS, RG = kMegaWarps / ph.KS;
            int g0, g1;
            phase_tiles<FUSE>(ph, cta, ncta, g0, g1);
            pr_tail = 0;
            if constexpr (FUSE) {
                if (ph.kind == PH_DOWN_T) {
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
            pr_left = (g1 - g0 - rg + RG - 1) / RG;
            pr_ptr = ph.W + ((size_t)(g0 + rg) * ph.KS + ks) * ph.tile_bytes;
            pr_step = (uint32_t)RG * ph.KS * ph.tile_bytes;
            pr_bytes = (uint32_t)ph.tile_bytes;
        }
        const int si = pr_count & (kSlots - 1);
        uint32_t nb = pr_bytes;
        if (pr_tail && pr_left == 1) nb = pr_tail;
        mb_expect(my_bar + si, nb);
        tma_g2s(my_ring + (size_t)si * kSlotBytes, pr_ptr, nb, my_bar + si);
        pr_ptr += pr_step;
        pr_left--;
        pr_count++;
    };
    if (lane == 0) {
#pragma unroll 1
        for (int s = 0; s < kSlots; ++s) produce_one();
    }
    unsigned cons_count = 0;
    unsigned kv_use0 = 0, kv_use1 = 0, kv_use2 = 0;
    // value r of the token's embedding row (layer 0's residual): the tiled classifier / embedding matrix
    auto emb_elem = [&](int r) -> float {
        const PhaseDesc em = phase(nwp - 1);
        const int c = r / E, e = r - c * E, eks = c / em.SC, ecc = c - eks * em.SC;
        const uint8_t* a = p.emb + ((size_t)(token / em.R) * em.KS + eks) * em.tile_bytes + (size_t)(token % em.R) * em.SC * 16 + (size_t)ecc * 16;
        if (WD == SLLM_F32) return __ldg(reinterpret_cast<const float*>(a) + e);
        return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const uint16_t*>(a) + e) << 16);
    };

#pragma unroll 1
    for (int wp = 0; wp < nwp; ++wp) {
        const PhaseDesc ph = phase(wp);
        const int l = ph.layer;
        const int par = (int)((gl0 + (unsigned)l) & 1u);
        float* const x_in = P.xbuf[par];
        float* const x_out = P.xbuf[par ^ 1];
        float* const hb = P.hbuf[par];
        const int ev = wp + l + (ph.kind != PH_QKV && ph.kind != PH_CLS ? 1 : 0);
        M2_STAMP(ev, 0);

        // =============================== wo: the kv head group's columns, this CTA's rows =========================
        if (ph.kind == PH_WO_T) {
            if (has_item) {
                const int ghd = G * p.hd;
                if (tid == 0 && !(p.debug & 1)) flag_wait(att_cnt + (size_t)my_kvh * 32, abase + (unsigned)((l + 1) * p.nsplit));
                __syncthreads();
                // merge the group's attention splits: weights of the splits per head first, then the columns
                const int rec = p.hd + kAttRecPad;
                float* wgt = part;                              // [G][nsplit + 1] (the partial table is idle here)
                if (p.nsplit <= 8) {   // every record word a thread needs, requested at once: ONE L2 round trip, the weights recomputed per thread
                    for (int c4 = tid; c4 < ghd / 4; c4 += kMegaThreads) {
                        const int col = c4 * 4;
                        const int gi = col / p.hd, j = col - gi * p.hd;
                        const float* base = p.att_part + (size_t)(my_kvh * G + gi) * p.nsplit * rec;
                        float m[8], ls[8];
                        float4 v[8];
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            m[s] = -INFINITY; ls[s] = 0.f; v[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (s < p.nsplit) {
                                m[s] = __ldcg(base + (size_t)s * rec + p.hd);
                                ls[s] = __ldcg(base + (size_t)s * rec + p.hd + 1);
                                v[s] = __ldcg(reinterpret_cast<const float4*>(base + (size_t)s * rec + j));
                            }
                        }
                        float M = -INFINITY;
#pragma unroll
                        for (int s = 0; s < 8; ++s) M = fmaxf(M, m[s]);
                        float Ls = 0.f;
                        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const float w = (m[s] == -INFINITY) ? 0.f : expf(m[s] - M);
                            Ls = fmaf(ls[s], w, Ls);
                            o.x = fmaf(v[s].x, w, o.x); o.y = fmaf(v[s].y, w, o.y); o.z = fmaf(v[s].z, w, o.z); o.w = fmaf(v[s].w, w, o.w);
                        }
                        reinterpret_cast<float4*>(xs)[c4] = make_float4(o.x / Ls, o.y / Ls, o.z / Ls, o.w / Ls);
                    }
                } else {
                if (tid < G) {
                    const float* base = p.att_part + (size_t)(my_kvh * G + tid) * p.nsplit * rec + p.hd;
                    float M = -INFINITY;
                    for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, __ldcg(base + (size_t)s * rec));
                    float Ls = 0.f;
                    for (int s = 0; s < p.nsplit; ++s) {
                        const float m = __ldcg(base + (size_t)s * rec);
                        const float w = (m == -INFINITY) ? 0.f : expf(m - M);
                        Ls = fmaf(__ldcg(base + (size_t)s * rec + 1), w, Ls);
                        wgt[tid * (p.nsplit + 1) + s] = w;
                    }
                    wgt[tid * (p.nsplit + 1) + p.nsplit] = Ls;
                }
                __syncthreads();
                for (int c4 = tid; c4 < ghd / 4; c4 += kMegaThreads) {
                    const int col = c4 * 4;
                    const int gi = col / p.hd, j = col - gi * p.hd;
                    const float* base = p.att_part + (size_t)(my_kvh * G + gi) * p.nsplit * rec + j;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int s = 0; s < p.nsplit; ++s) {
                        const float w = wgt[gi * (p.nsplit + 1) + s];
                        const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (size_t)s * rec));
                        o.x = fmaf(v.x, w, o.x); o.y = fmaf(v.y, w, o.y); o.z = fmaf(v.z, w, o.z); o.w = fmaf(v.w, w, o.w);
                    }
                    const float Ls = wgt[gi * (p.nsplit + 1) + p.nsplit];
                    reinterpret_cast<float4*>(xs)[c4] = make_float4(o.x / Ls, o.y / Ls, o.z / Ls, o.w / Ls);
                }
                }
                __syncthreads();
                // this lane's chunk(s) of the merged vector -> registers
                const int CRP = ph.SC;
                float xw[4][E];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int cc = (CRP <= 32) ? (lane & (CRP - 1)) : lane + 32 * i;
#pragma unroll
                    for (int e = 0; e < E; ++e) xw[i][e] = 0.f;
                    if (i < (CRP > 32 ? CRP / 32 : 1) && cc < ph.nchunks) {
#pragma unroll
                        for (int e4 = 0; e4 < E; e4 += 4) {
                            const float4 v = *reinterpret_cast<const float4*>(xs + cc * E + e4);
                            xw[i][e4] = v.x; xw[i][e4 + 1] = v.y; xw[i][e4 + 2] = v.z; xw[i][e4 + 3] = v.w;
                        }
                    }
                }
                M2_STAMP(ev, 1);
                const int NR = ph.R;
                const int T0 = (int)(((int64_t)ph.ntr * my_part) / p.nsplit), T1 = (int)(((int64_t)ph.ntr * (my_part + 1)) / p.nsplit);
#pragma unroll 1
                for (int t = T0 + warp; t < T1; t += kMegaWarps) {
                    const int si = cons_count & (kSlots - 1);
                    mb_wait_fast(my_bar + si, (cons_count / kSlots) & 1);
                    const uint8_t* slot = my_ring + (size_t)si * kSlotBytes;
                    float val = 0.f;
                    int row = 0;
                    bool own = false;
                    switch (CRP) {
                        case 8: own = wot_tile<WD, 8>(slot, lane, xw, val, row); break;
                        case 16: own = wot_tile<WD, 16>(slot, lane, xw, val, row); break;
                        case 32: own = wot_tile<WD, 32>(slot, lane, xw, val, row); break;
                        case 64: own = wot_tile<WD, 64>(slot, lane, xw, val, row); break;
                        default: own = wot_tile<WD, 128>(slot, lane, xw, val, row); break;
                    }
                    const int r = t * NR + row;
                    if (own && r < ph.nrows) {
                        if (t % p.KVH_loc == my_kvh) val += (l == 0) ? emb_elem(r) : __ldcg(x_in + r);   // the residual, once per row (add_kernel.cpp:10-13)
                        red_add_f32(hb + r, val);
                    }
                    cons_count++;
                    __syncwarp();
                    if (lane == 0) {
                        fence_async_smem();
                        produce_one();
                    }
                }
                M2_STAMP(ev, 3);
            }
            M2_STAMP(ev, 4);
            if (p.debug & 1) __syncthreads(); else grid_barrier(p.bar_counter, bar_base + (++bar_idx) * (unsigned)ncta);
            M2_STAMP(ev, 5);
            continue;
        }

        const int ks = warp & (ph.KS - 1), rg = warp / ph.KS, RG = kMegaWarps / ph.KS;
        const int c0 = ks * ph.SC;
        const int nsc = max(0, min(ph.SC, ph.nchunks - c0));
        const int cols = ph.nchunks * E;
        const bool normed = (ph.kind == PH_QKV || ph.kind == PH_GATEUP || ph.kind == PH_CLS);

        if constexpr (FUSE) {
            if (ph.kind == PH_DOWN_T) {
                // x_out += Wdown[:, this CTA's units] . swi[this CTA's units] (values left in xs by the gate_up epilogue); x_out was zeroed
                // during the qkv phase, CTA 0 adds the residual h once
                int g0, g1;
                cta_tiles(ph, cta, ncta, g0, g1);
                M2_STAMP(ev, 1);
                float acc[E];
#pragma unroll
                for (int e = 0; e < E; ++e) acc[e] = 0.f;
                if (cta == 0 && rg == 0) {
#pragma unroll
                    for (int e4 = 0; e4 < E; e4 += 4) {
                        const float4 v = __ldcg(reinterpret_cast<const float4*>(hb + (size_t)(ks * 32 + lane) * E + e4));
                        acc[e4] = v.x; acc[e4 + 1] = v.y; acc[e4 + 2] = v.z; acc[e4 + 3] = v.w;
                    }
                }
                bool any = (cta == 0 && rg == 0);
                int ja, jb;
                down_t_rows(g0, g1, rg, RG, ja, jb);
                constexpr int kRowsPerSlot = kSlotBytes / (kFuseJT * 512);
#pragma unroll 1
                for (int j = ja; j < jb; j += kRowsPerSlot) {
                    const int si = cons_count & (kSlots - 1);
                    mb_wait_fast(my_bar + si, (cons_count / kSlots) & 1);
                    const uint8_t* sp = my_ring + (size_t)si * kSlotBytes + lane * 16;
                    const float* sw = xs + (size_t)(j - g0) * kFuseJT;
                    const int nj = min(kRowsPerSlot, jb - j) * kFuseJT;
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
                M2_STAMP(ev, 3);
                if (any) {
                    float* dst = x_out + (size_t)(ks * 32 + lane) * E;
                    red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
                    if (E == 8) red_add_v4(dst + 4, acc[E - 4], acc[E - 3], acc[E - 2], acc[E - 1]);
                }
                M2_STAMP(ev, 4);
                if (p.debug & 1) __syncthreads(); else grid_barrier(p.bar_counter, bar_base + (++bar_idx) * (unsigned)ncta);
                M2_STAMP(ev, 5);
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
        // ---- 1. stage the activation vector in shared memory -----------------------------------------------------
        float ss = 0.f;
        if (ph.kind == PH_QKV) {
            // zero, a layer / a phase ahead, the vectors the reductions of the NEXT dependency points land on: the next layer's h and
            // (fused down) this layer's x_out. Their last readers finished before the grid barrier this CTA has just passed.
            const int n4 = p.d >> 2, z0 = (int)(((int64_t)n4 * cta) / ncta), z1 = (int)(((int64_t)n4 * (cta + 1)) / ncta);
            for (int i = z0 + tid; i < z1; i += kMegaThreads) {
                reinterpret_cast<float4*>(P.hbuf[par ^ 1])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (FUSE) reinterpret_cast<float4*>(x_out)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (ph.kind == PH_QKV && l == 0) {                                   // embedding gather (model.cpp:48)
            const PhaseDesc em = phase(nwp - 1);
            const uint8_t* trow = p.emb + (size_t)(token / em.R) * em.KS * em.tile_bytes + (size_t)(token % em.R) * em.SC * 16;
            for (int c = tid; c < ph.nchunks; c += kMegaThreads) {
                const int eks = c / em.SC, ecc = c - eks * em.SC;
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(trow + (size_t)eks * em.tile_bytes + (size_t)ecc * 16));
                float f[E];
                if (WD == SLLM_F32) {
                    f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y); f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
                } else {
                    kv_unpack<SLLM_BF16>(raw, f);
                }
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    xs[c * E + e] = f[e];
                    ss = fmaf(f[e], f[e], ss);
                }
            }
        } else {
            const float* src = (ph.kind == PH_GATEUP) ? hb : (ph.kind == PH_DOWN) ? p.swi : x_in;
            float* copy = nullptr;                       // introspection copies (emb_output / ffn_input of the reference's buffer table)
            if (cta == 0 && ph.kind == PH_CLS) copy = P.x_copy;
            if (cta == 0 && ph.kind == PH_GATEUP && l == p.L - 1) copy = P.h_copy;
            for (int c4 = tid; c4 < cols / 4; c4 += kMegaThreads) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + c4);
                reinterpret_cast<float4*>(xs)[c4] = v;
                if (copy) reinterpret_cast<float4*>(copy)[c4] = v;
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

        if (ph.kind == PH_QKV && has_item) {
            // K/V rows of this CTA's attention item written by EARLIER launches: start their TMA now (the K/V stages only alias the
            // activation staging buffer, already consumed into registers), so they land while the weights stream
            __syncthreads();
            if (warp == 0 && lane == 0) {
                fence_async_smem();
                const int row_bytes = p.hd * KESZ;
                const int npos = pos + 1, per = (npos + p.nsplit - 1) / p.nsplit;
                const int t0 = my_part * per, t1 = min(npos, t0 + per);
                const size_t head_off = ((size_t)l * p.KVH_loc + my_kvh) * p.S * row_bytes;   // [L][KVH][S][hd]
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
        M2_STAMP(ev, 1);   // prologue done
        // ---- 2. stream this CTA's tile rows through the rings, round by round ------------------------------
        int g0, g1;
        phase_tiles<FUSE>(ph, cta, ncta, g0, g1);
        const int upp = ph.R >> 1;
        const int u0 = g0 * upp;
        const int n = max(0, min(ph.nunits, g1 * upp) - u0);
        const int nslots = g1 - g0;
        const uint32_t sbytes = (uint32_t)ph.SC * 16;
        const int cpl = (nsc + 31) >> 5;
        float best_v = -INFINITY;
        int best_i = 0x7fffffff;
        // what the epilogue of this thread's unit (first round: unit u0 + tid) needs from global memory, requested BEFORE the stream so
        // that its latency is not exposed between the last tile and the dependency point: RoPE table entries / the residual h
        float pre0 = 0.f, pre1 = 0.f;
        if (tid < min(n, kRoundUnits)) {
            const int u = u0 + tid;
            if (ph.kind == PH_QKV) {
                if (u < ((p.q_loc + p.kv_loc) >> 1)) {
                    const int j = u % half;
                    pre0 = __ldg(p.sin_t + (size_t)pos * half + j);
                    pre1 = __ldg(p.cos_t + (size_t)pos * half + j);
                }
            } else if (ph.kind == PH_DOWN) {
                pre0 = __ldcg(hb + 2 * u);
                if (2 * u + 1 < ph.nrows) pre1 = __ldcg(hb + 2 * u + 1);
            }
        }

#pragma unroll 1
        for (int rbase = 0; rbase < n || rbase == 0; rbase += kRoundUnits) {
            const int jend = min(nslots, (rbase + kRoundUnits) / upp);
#pragma unroll 1
            for (int j = rbase / upp + rg; j < jend; j += RG) {
                const int si = cons_count & (kSlots - 1);
                mb_wait_fast(my_bar + si, (cons_count / kSlots) & 1);
                const uint8_t* sp = my_ring + (size_t)si * kSlotBytes + lane * 16;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (p.debug & 2) {
                    a0 = __uint_as_float(*reinterpret_cast<const uint32_t*>(sp));
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
            if (rbase + kRoundUnits >= n) M2_STAMP(ev, 2);
            __syncthreads();
            if (rbase + kRoundUnits >= n) M2_STAMP(ev, 3);
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
                    const int rope_units = (p.q_loc + p.kv_loc) >> 1;
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
                        const float fci = rbase == 0 ? pre0 : p.sin_t[(size_t)pos * half + j], fcr = rbase == 0 ? pre1 : p.cos_t[(size_t)pos * half + j];
                        const float o0 = s0 * fcr - s1 * fci, o1 = s1 * fcr + s0 * fci;   // rope_kernel.cpp:36-37
                        const int r0 = head * p.hd + j;
                        if (r0 < p.q_loc) { p.q[r0] = o0; p.q[r0 + half] = o1; }
                        else { store_kv(p.kc, r0 - p.q_loc, o0); store_kv(p.kc, r0 - p.q_loc + half, o1); }
                    } else {
                        const int b2 = 2 * (u - rope_units);
                        store_kv(p.vc, b2, s0);
                        store_kv(p.vc, b2 + 1, s1);
                    }
                } else if (ph.kind == PH_GATEUP) {
                    const float sv = (1.0f / (1.0f + expf(-s1))) * s0;        // swiglu_kernel.cpp:12-13
                    p.swi[u] = sv;
                    if constexpr (FUSE) xs[rbase + t] = sv;
                } else if (ph.kind == PH_DOWN) {
                    const int r = 2 * u;
                    x_out[r] = s0 + (rbase == 0 ? pre0 : __ldcg(hb + r));     // add_kernel.cpp:10-13
                    if (r + 1 < ph.nrows) x_out[r + 1] = s1 + (rbase == 0 ? pre1 : __ldcg(hb + r + 1));
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

        if (ph.kind == PH_CLS) {
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
                p.st->pad[0] = (int32_t)(bar_base + bar_idx * (unsigned)ncta);
                p.st->pad[2] = (int32_t)(epoch + 1u);
                p.blk_val[ncta] = v;
                p.blk_idx[ncta] = idx;
                ClsPolicy<SLLM_F32>::step_feedback(p.st, p.prompt, p.history, idx);
            }
            break;
        }
        M2_STAMP(ev, 4);   // epilogue done
        if (FUSE && ph.kind == PH_GATEUP) {
            M2_STAMP(ev, 5);
            continue;
        }
        if (ph.kind != PH_QKV) {
            if (p.debug & 1) __syncthreads(); else grid_barrier(p.bar_counter, bar_base + (++bar_idx) * (unsigned)ncta);
            M2_STAMP(ev, 5);
            continue;
        }
        // ---- publish this CTA's q / k / v units to the kv head groups they belong to (all stores above precede the round's
        //      trailing __syncthreads; red.release orders them before the counter)
        if (tid < s_npub) flag_add_release(qkv_cnt + (size_t)s_pub_g[tid] * 32, (unsigned)s_pub_n[tid]);
        // Single-pass attention: when the item's rows all fit on chip — the two K/V stages (128 rows, prefetched while the qkv weights
        // streamed) plus a tail of up to 64 rows in the partial table, which is idle from here to the next weight phase — scores, softmax
        // and P.V run ONCE over all of them instead of tile by tile (4 block-wide syncs per tile and an exposed copy for the third tile).
        // The tail is requested now, so that it lands under the wait for the group's q / k / v.
        const int att_row_bytes = p.hd * KESZ;
        const int tail_cap = min(kAttTile, (kRoundUnits * 2 * kMegaWarps * 4) / (2 * att_row_bytes));
        int att_t0 = 0, att_rows = 0;
        if (has_item) {
            const int npos_ = pos + 1, per_ = (npos_ + p.nsplit - 1) / p.nsplit;
            att_t0 = my_part * per_;
            att_rows = max(0, min(npos_, att_t0 + per_) - att_t0);
        }
        const bool att_single = att_rows <= 2 * AT + tail_cap;   // +0.9 % tokens/s at 512-650 positions (A/B on one box)
        const int att_tail = att_single ? max(0, att_rows - 2 * AT) : 0;
        uint8_t* const tail_k = reinterpret_cast<uint8_t*>(part);
        uint8_t* const tail_v = tail_k + (size_t)tail_cap * att_row_bytes;
        if (att_tail > 0 && warp == 0 && lane == 0) {
            fence_async_smem();   // the epilogue's generic reads of the partial table precede the async writes
            const int ts = att_t0 + 2 * AT;
            const int bulk_rows = max(0, min(att_tail, pos - ts));
            const size_t head_off = ((size_t)l * p.KVH_loc + my_kvh) * p.S * att_row_bytes;
            mb_expect(att_bar + 2, (uint32_t)(2 * bulk_rows * att_row_bytes));
            if (bulk_rows > 0) {
                tma_g2s(tail_k, p.kc + head_off + (size_t)ts * att_row_bytes, bulk_rows * att_row_bytes, att_bar + 2);
                tma_g2s(tail_v, p.vc + head_off + (size_t)ts * att_row_bytes, bulk_rows * att_row_bytes, att_bar + 2);
            }
        }
        M2_STAMP(ev, 5);
        M2_STAMP(ev + 1, 0);
        if (!has_item) { M2_STAMP(ev + 1, 4); M2_STAMP(ev + 1, 5); continue; }

        // =============================== attention item of this CTA ======================================
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
            const int per = (npos + p.nsplit - 1) / p.nsplit;
            const float scale = 1.0f / sqrtf((float)p.hd);
            constexpr int kStripes = 16;
            const int pv_chunk = tid % cpr, pv_stripe = tid / cpr;
            const bool pv_active = pv_stripe < kStripes;
            const int key = tid >> 3, kpart = tid & 7;
            const int kvh = my_kvh, split = my_part;

            if (tid == 0 && !(p.debug & 1)) flag_wait(qkv_cnt + (size_t)kvh * 32, qbase + (unsigned)((l + 1) * upg));
            __syncthreads();
            M2_STAMP(ev + 1, 1);
            const int t0 = split * per, t1 = min(npos, t0 + per);
            const int ntiles = (t1 > t0) ? (t1 - t0 + AT - 1) / AT : 0;
            for (int i = tid; i < G * p.hd; i += kMegaThreads) q_s[i] = __ldcg(p.q + (size_t)(kvh * G) * p.hd + i);
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
            float acc[G][KVEC];
#pragma unroll
            for (int gi = 0; gi < G; ++gi)
#pragma unroll
                for (int e = 0; e < KVEC; ++e) acc[gi][e] = 0.f;

            if (att_single) {
                const int nrows = att_rows;
                auto krow = [&](int r) -> const uint8_t* { return r < 2 * AT ? k_s + (size_t)r * stride : tail_k + (size_t)(r - 2 * AT) * row_bytes; };
                auto vrow = [&](int r) -> const uint8_t* { return r < 2 * AT ? v_s + (size_t)r * stride : tail_v + (size_t)(r - 2 * AT) * row_bytes; };
                if (pos >= t0 && pos < t1 && warp == 1) {   // the newest row (generic stores of this launch) bypasses the async proxy
                    const size_t g_off = head_off + (size_t)pos * row_bytes;
                    for (int c = lane; c < cpr; c += 32) {
                        *reinterpret_cast<uint4*>(const_cast<uint8_t*>(krow(pos - t0)) + c * 16) = __ldcg(reinterpret_cast<const uint4*>(p.kc + g_off + c * 16));
                        *reinterpret_cast<uint4*>(const_cast<uint8_t*>(vrow(pos - t0)) + c * 16) = __ldcg(reinterpret_cast<const uint4*>(p.vc + g_off + c * 16));
                    }
                }
                if (nrows > 0) { mb_wait_fast(att_bar, kv_use0 & 1); kv_use0++; }
                if (nrows > AT) { mb_wait_fast(att_bar + 1, kv_use1 & 1); kv_use1++; }
                if (att_tail > 0) { mb_wait_fast(att_bar + 2, kv_use2 & 1); kv_use2++; }
                __syncthreads();
#pragma unroll 1
                for (int kb = 0; kb < nrows; kb += kAttTile) {        // scores of all rows, 8 threads per key
                    const int r = kb + key;
                    float s[G];
#pragma unroll
                    for (int gi = 0; gi < G; ++gi) s[gi] = 0.f;
                    if (r < nrows) {
                        const uint8_t* kr = krow(r);
                        for (int c = kpart; c < cpr; c += 8) {
                            float kf[KVEC];
                            kv_unpack<KVD>(*reinterpret_cast<const uint4*>(kr + c * 16), kf);
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
                        if (kpart == 0 && r < nrows) p_s[gi * kAttP + r] = s[gi] * scale;
                    }
                }
                __syncthreads();
                for (int gi = warp; gi < G; gi += kMegaWarps) {        // softmax over the split's rows (mha_kernel.cpp:7-20 per split)
                    float mx = -INFINITY;
                    for (int i = lane; i < nrows; i += 32) mx = fmaxf(mx, p_s[gi * kAttP + i]);
                    mx = warp_max(mx);
                    float sum = 0.f;
                    for (int i = lane; i < nrows; i += 32) {
                        const float e = expf(p_s[gi * kAttP + i] - mx);
                        p_s[gi * kAttP + i] = e;
                        sum += e;
                    }
                    sum = warp_sum(sum);
                    if (lane == 0) { ml_s[2 * gi] = mx; ml_s[2 * gi + 1] = sum; }
                }
                __syncthreads();
                if (pv_active) {
                    for (int r = pv_stripe; r < nrows; r += kStripes) {
                        float vf[KVEC];
                        kv_unpack<KVD>(*reinterpret_cast<const uint4*>(vrow(r) + pv_chunk * 16), vf);
#pragma unroll
                        for (int gi = 0; gi < G; ++gi) {
                            const float pr = p_s[gi * kAttP + r];
#pragma unroll
                            for (int e = 0; e < KVEC; ++e) acc[gi][e] = fmaf(pr, vf[e], acc[gi][e]);
                        }
                    }
                }
                fence_async_smem();
                __syncthreads();
            } else {
#pragma unroll 1
            for (int tile = 0; tile < ntiles; ++tile) {
                const int stage = tile & 1;
                const int ts = t0 + tile * AT;
                const int rows = min(AT, t1 - ts);
                if (pos >= ts && pos < ts + rows && warp == 1) {   // the newest row (generic stores of this launch) bypasses the async proxy
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
            }
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
            fence_async_smem();
            __syncthreads();
            if (tid == 0) flag_add_release(att_cnt + (size_t)kvh * 32, 1u);
        }
        M2_STAMP(ev + 1, 4);
        M2_STAMP(ev + 1, 5);
    }
}

// ------------------------------------------------------------------------------------------- host ----
// WoT: for every kv head group g the block Wo[:, g*G*hd : (g+1)*G*hd] as [d rows][CRP chunks] (rows padded with zero chunks to CRP,
// a power of two >= 8, or 64 / 128), cut into 4 KB tiles of 256 / CRP whole rows; group g at g * ntr * 4096.
WotGeom mega2_wot_geom(int d, int ghd, int kvh, int w_dtype) {
    WotGeom g{};
    const int E = w_dtype == SLLM_F32 ? 4 : 8;
    g.nchunks = ghd / E;
    int crp = 8;
    while (crp < g.nchunks) crp *= 2;
    g.crp = crp;
    g.nr = 256 / crp;
    g.ntr = (d + g.nr - 1) / g.nr;
    g.group_bytes = (size_t)g.ntr * kSlotBytes;
    g.bytes = g.group_bytes * kvh;
    return g;
}

bool mega2_ok(int w_dtype, int d, int hd, int H_loc, int KVH_loc, int nsplit, int grid, const char** why) {
    static const char* none = "";
    const char*& w = why ? *why : none;
    if (w_dtype != SLLM_F32 && w_dtype != SLLM_BF16) { w = "int8 weights run in the grid-barrier megakernel"; return false; }
    const int E = w_dtype == SLLM_F32 ? 4 : 8, G = H_loc / KVH_loc;
    if ((G * hd) % E || (G * hd) / E > 128) { w = "kv head group wider than 128 chunks"; return false; }
    if (KVH_loc > kMaxGroups) { w = "more than 64 kv heads per rank"; return false; }
    if (KVH_loc * nsplit > grid) { w = "more attention items than CTAs"; return false; }
    if (d % 4) { w = "hidden size not a multiple of 4"; return false; }
    return true;
}

size_t mega2_wot_bytes(int d, int hd, int H_loc, int KVH_loc, int w_dtype) { return mega2_wot_geom(d, (H_loc / KVH_loc) * hd, KVH_loc, w_dtype).bytes; }

void mega2_fill_wot(PhaseDesc& ds, const void* W, int d, int hd, int H_loc, int KVH_loc, int layer, int w_dtype) {
    const WotGeom g = mega2_wot_geom(d, (H_loc / KVH_loc) * hd, KVH_loc, w_dtype);
    ds.W = reinterpret_cast<const uint8_t*>(W);
    ds.nchunks = g.nchunks;
    ds.nunits = KVH_loc;
    ds.nrows = d;
    ds.kind = PH_WO_T;
    ds.layer = layer;
    ds.KS = 1;
    ds.SC = g.crp;
    ds.R = g.nr;
    ds.ntr = g.ntr;
    ds.tile_bytes = kSlotBytes;
    ds.srow = 0;
    ds.cum = nullptr;
}

// one thread per 16-byte chunk of the destination
__global__ void repack_wot_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int d, int src_chunks, int nchunks, int crp, int nr, int ntr,
                                  int kvh) {
    const int64_t total = (int64_t)kvh * ntr * 256;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % crp);
        const int64_t rowi = i / crp;                      // row index over all groups: g * ntr * nr + r
        const int r = (int)(rowi % ((int64_t)ntr * nr));
        const int g = (int)(rowi / ((int64_t)ntr * nr));
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < d && c < nchunks) v = src[(int64_t)r * src_chunks + (int64_t)g * nchunks + c];
        dst[i] = v;
    }
}

int mega2_repack_wot(const void* src_rowmajor, void* dst, int d, int hd, int H_loc, int KVH_loc, int w_dtype, cudaStream_t st) {
    const int ghd = (H_loc / KVH_loc) * hd, E = w_dtype == SLLM_F32 ? 4 : 8;
    const WotGeom g = mega2_wot_geom(d, ghd, KVH_loc, w_dtype);
    const int64_t total = (int64_t)KVH_loc * g.ntr * 256;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    repack_wot_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src_rowmajor), reinterpret_cast<uint4*>(dst), d, H_loc * hd / E, g.nchunks,
                                             g.crp, g.nr, g.ntr, KVH_loc);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

template <int WD, int KVD, int G, bool FUSE>
static int mega2_launch_t(const Mega2Params& P, int grid, size_t smem, cudaStream_t st) {
    static size_t configured = 0;
    if (smem > configured) {
        SLLM_CUDA(cudaFuncSetAttribute(mega2_step_kernel<WD, KVD, G, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    SLLM_CUDA(cudaLaunchKernelEx(&cfg, mega2_step_kernel<WD, KVD, G, FUSE>, P));
    g_launches++;
    return SLLM_OK;
}

extern int g_tune_mega_debug;
int mega2_launch(const Mega2Params& P_in, int g, int grid, size_t smem, cudaStream_t st, bool fuse_down) {
    Mega2Params P = P_in;
    P.m.debug = g_tune_mega_debug;
#define MEGA2_KV(WDT, GG, FU)                                                                      \
    return P.m.kv_dtype == SLLM_F32 ? mega2_launch_t<WDT, SLLM_F32, GG, FU>(P, grid, smem, st)     \
                                    : mega2_launch_t<WDT, SLLM_BF16, GG, FU>(P, grid, smem, st);
#define MEGA2_G(GG)                                                                                \
    case GG:                                                                                       \
        if (P.m.w_dtype == SLLM_F32) { if (fuse_down) { MEGA2_KV(SLLM_F32, GG, true) } else { MEGA2_KV(SLLM_F32, GG, false) } } \
        if (fuse_down) { MEGA2_KV(SLLM_BF16, GG, true) } else { MEGA2_KV(SLLM_BF16, GG, false) }
    switch (g) {
        MEGA2_G(1)
        MEGA2_G(2)
        MEGA2_G(4)
        MEGA2_G(8)
        default: set_error("megakernel v2: %d query heads per KV head not instantiated", g); return SLLM_ENOTSUP;
    }
#undef MEGA2_G
#undef MEGA2_KV
}

}  // namespace sllm
