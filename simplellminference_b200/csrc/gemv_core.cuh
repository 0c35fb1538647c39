// csrc/gemv_core.cuh — the bandwidth-bound decode GEMV building block (sm_100a).
//
// y[r] = sum_j x[j] * W[r][j] for a row-major weight matrix that is read from HBM exactly once per token.
// The kernel shape shared by every fused variant (decode_fused.cu) and by the plain launcher (gemv.cu):
//
//  * the activation vector x (fp32, `cols` floats) is staged ONCE per CTA in shared memory, split into
//    E/4 "planes" of float4 (E = weights per 16-byte load) so that lane l's float4 reads are 16 bytes apart:
//    conflict-free LDS.128 regardless of the weight type;
//  * a warp owns a *unit* of two weight rows at a time (the pairs the epilogues need: RoPE partners
//    (j, j+hd/2), (up_i, gate_i), or two consecutive rows) and streams them with 128-bit
//    ld.global.nc.L1::no_allocate loads, U loads per row in flight per lane (2*U*512 B per warp);
//  * products accumulate in fp32 registers; one butterfly of warp shuffles per row finishes the dot product;
//  * units are dealt round-robin to the warps of a persistent grid (a multiple of the SM count);
//  * programmatic dependent launch: the first batch of WEIGHT loads is issued before griddepcontrol.wait —
//    weights never depend on the previous kernel — so HBM keeps streaming across kernel boundaries; only the
//    staging of x waits for the producer kernel.
//
// HBM-bound integer/byte work: no tensor cores here on purpose (batch 1: 2 flop per weight byte).
#pragma once
#include "common.cuh"

namespace sllm {

template <int WD> struct WInfo;
template <> struct WInfo<SLLM_F32> { static constexpr int E = 4; static constexpr int BYTES = 4; };
template <> struct WInfo<SLLM_BF16> { static constexpr int E = 8; static constexpr int BYTES = 2; };
template <> struct WInfo<SLLM_INT8> { static constexpr int E = 16; static constexpr int BYTES = 1; };

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;
#ifndef SLLM_GEMV_U
#define SLLM_GEMV_U 8
#endif
constexpr int kGemvU = SLLM_GEMV_U;  // 16-byte loads in flight per row per lane

__host__ __device__ inline size_t gemv_smem_bytes(int cols) { return (size_t)cols * sizeof(float) + 64 * sizeof(float); }

// ---- staging of x into planes -----------------------------------------------------------------------
// xs layout: plane p (0..E/4) is float4[nchunks]; element j of x lives in chunk c=j/E, plane (j%E)/4, slot j%4.
template <int WD>
__device__ __forceinline__ int plane_index(int j4 /* index of a float4 of x */, int nchunks) {
    constexpr int P = WInfo<WD>::E / 4;
    const int c = j4 / P, p = j4 - c * P;
    return p * nchunks + c;
}

// plain copy
template <int WD>
__device__ __forceinline__ void stage_x_plain(float* xs, const float* __restrict__ x, int cols) {
    const int nchunks = cols / WInfo<WD>::E;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* s4 = reinterpret_cast<float4*>(xs);
    for (int j4 = threadIdx.x; j4 < cols / 4; j4 += blockDim.x) s4[plane_index<WD>(j4, nchunks)] = x4[j4];
}

struct P2PComm;
__device__ __forceinline__ unsigned p2p_slots(const P2PComm* c, int op, const uint2** slots, int* tp);
__device__ __forceinline__ float2 p2p_recv2(const uint2* p, unsigned epoch);

// fused RMSNorm (rms_kernel.cpp:12-22): xs = (x * 1/sqrt(mean(x^2)+eps)) * w, recomputed by every CTA from
// the L2-resident residual stream. `red` = 33 floats of shared scratch. Optional fused residual add for the
// tensor-parallel path: the vector normalised is x + add (add = all-reduced partial sums of the previous
// row-parallel GEMV) and CTA 0 writes that sum to sum_out so later kernels see the updated residual stream.
template <int WD>
__device__ __forceinline__ void stage_x_rmsnorm(float* xs, float* red, const float* __restrict__ x,
                                                const float* __restrict__ w, int cols, float eps,
                                                const float* __restrict__ add = nullptr, float* __restrict__ sum_out = nullptr,
                                                const P2PComm* p2p = nullptr, int p2p_op = 0) {
    const int nchunks = cols / WInfo<WD>::E;
    const uint2* slots[8];
    int ntp = 0;
    unsigned epoch = 0;
    if (p2p) epoch = p2p_slots(p2p, p2p_op, slots, &ntp);   // peer-memory all-reduce: partial vectors of op p2p_op
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* a4 = reinterpret_cast<const float4*>(add);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    float4* o4 = reinterpret_cast<float4*>(sum_out);
    float4* s4 = reinterpret_cast<float4*>(xs);
    float ss = 0.0f;
    for (int j4 = threadIdx.x; j4 < cols / 4; j4 += blockDim.x) {
        float4 v = x4[j4];
        if (p2p) {   // spin on the {value, epoch} words themselves; sum in rank order (identical on every rank)
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < ntp; ++r) {
                const float2 lo = p2p_recv2(slots[r] + 4 * j4, epoch), hi = p2p_recv2(slots[r] + 4 * j4 + 2, epoch);
                a = (r == 0) ? make_float4(lo.x, lo.y, hi.x, hi.y) : make_float4(a.x + lo.x, a.y + lo.y, a.z + hi.x, a.w + hi.y);
            }
            v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
            if (sum_out && blockIdx.x == 0) o4[j4] = v;
        } else if (add) {
            const float4 a = a4[j4];
            v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
            if (sum_out && blockIdx.x == 0) o4[j4] = v;
        }
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        s4[plane_index<WD>(j4, nchunks)] = v;
    }
    ss = block_sum(ss, red);  // contains the __syncthreads that orders the stores above
    const float inv = 1.0f / sqrtf(ss / (float)cols + eps);
    for (int j4 = threadIdx.x; j4 < cols / 4; j4 += blockDim.x) {
        const int idx = plane_index<WD>(j4, nchunks);
        const float4 v = s4[idx], g = w4[j4];
        s4[idx] = make_float4((v.x * inv) * g.x, (v.y * inv) * g.y, (v.z * inv) * g.z, (v.w * inv) * g.w);
    }
}

// ---- one 16-byte chunk of weights times the matching x ----------------------------------------------
template <int WD>
__device__ __forceinline__ float chunk_dot(const uint4 w, const float4* __restrict__ xs4, int c, int nchunks, float acc);

template <>
__device__ __forceinline__ float chunk_dot<SLLM_F32>(const uint4 w, const float4* __restrict__ xs4, int c, int, float acc) {
    const float4 x = xs4[c];
    acc = fmaf(__uint_as_float(w.x), x.x, acc);
    acc = fmaf(__uint_as_float(w.y), x.y, acc);
    acc = fmaf(__uint_as_float(w.z), x.z, acc);
    acc = fmaf(__uint_as_float(w.w), x.w, acc);
    return acc;
}

template <>
__device__ __forceinline__ float chunk_dot<SLLM_BF16>(const uint4 w, const float4* __restrict__ xs4, int c, int nchunks, float acc) {
    const float4 a = xs4[c], b = xs4[nchunks + c];
    acc = fmaf(bf16_lo(w.x), a.x, acc);
    acc = fmaf(bf16_hi(w.x), a.y, acc);
    acc = fmaf(bf16_lo(w.y), a.z, acc);
    acc = fmaf(bf16_hi(w.y), a.w, acc);
    acc = fmaf(bf16_lo(w.z), b.x, acc);
    acc = fmaf(bf16_hi(w.z), b.y, acc);
    acc = fmaf(bf16_lo(w.w), b.z, acc);
    acc = fmaf(bf16_hi(w.w), b.w, acc);
    return acc;
}

// int8 -> fp32 without I2F: byte b (two's complement) -> u = b ^ 0x80 in [0,255]; the float with bits
// 0x4B000000|u is 2^23 + u exactly; subtracting 2^23 + 128 gives b exactly.
template <int B>
__device__ __forceinline__ float s8f(uint32_t wf) {
    // selector: result bytes (lsb..msb) = {wf.byteB, 0x00, 0x00, 0x4B}  -> second operand bytes are 4..7
    constexpr uint32_t sel = 0x7440u | (uint32_t)B;  // b0 = wf[B], b1 = y[0]=0x00, b2 = y[0]=0x00, b3 = y[3]=0x4B
    return __uint_as_float(__byte_perm(wf, 0x4B000000u, sel)) - 8388736.0f;
}
__device__ __forceinline__ float word_dot_s8(uint32_t w, const float4 x, float acc) {
    const uint32_t wf = w ^ 0x80808080u;
    acc = fmaf(s8f<0>(wf), x.x, acc);
    acc = fmaf(s8f<1>(wf), x.y, acc);
    acc = fmaf(s8f<2>(wf), x.z, acc);
    acc = fmaf(s8f<3>(wf), x.w, acc);
    return acc;
}
// returns the UNSCALED partial sum of the 16 products (the caller applies the group scale)
template <>
__device__ __forceinline__ float chunk_dot<SLLM_INT8>(const uint4 w, const float4* __restrict__ xs4, int c, int nchunks, float acc) {
    acc = word_dot_s8(w.x, xs4[c], acc);
    acc = word_dot_s8(w.y, xs4[nchunks + c], acc);
    acc = word_dot_s8(w.z, xs4[2 * nchunks + c], acc);
    acc = word_dot_s8(w.w, xs4[3 * nchunks + c], acc);
    return acc;
}

// ---- a batch of weight loads for one unit (two rows) ------------------------------------------------
template <int WD>
struct Batch {
    uint4 w0[kGemvU], w1[kGemvU];
    float s0[kGemvU], s1[kGemvU];  // int8 only: group scale of each chunk
};

template <int WD>
__device__ __forceinline__ void load_batch(Batch<WD>& b, const uint4* __restrict__ r0, const uint4* __restrict__ r1,
                                           const float* __restrict__ sc0, const float* __restrict__ sc1, int chunks_per_group,
                                           int base, int lane, int nchunks) {
#pragma unroll
    for (int u = 0; u < kGemvU; ++u) {
        const int c = base + u * 32 + lane;
        if (c < nchunks) {
            b.w0[u] = ldg_stream(r0 + c);
            b.w1[u] = ldg_stream(r1 + c);
            if (WD == SLLM_INT8) {
                b.s0[u] = __ldg(sc0 + c / chunks_per_group);
                b.s1[u] = __ldg(sc1 + c / chunks_per_group);
            }
        } else {
            b.w0[u] = make_uint4(0, 0, 0, 0);
            b.w1[u] = make_uint4(0, 0, 0, 0);
            if (WD == SLLM_INT8) { b.s0[u] = 0.f; b.s1[u] = 0.f; }
        }
    }
}

template <int WD>
__device__ __forceinline__ void fma_batch(const Batch<WD>& b, const float4* __restrict__ xs4, int base, int lane, int nchunks,
                                          float& a0, float& a1) {
#pragma unroll
    for (int u = 0; u < kGemvU; ++u) {
        const int c = base + u * 32 + lane;
        if (c < nchunks) {
            if (WD == SLLM_INT8) {
                a0 = fmaf(chunk_dot<WD>(b.w0[u], xs4, c, nchunks, 0.f), b.s0[u], a0);
                a1 = fmaf(chunk_dot<WD>(b.w1[u], xs4, c, nchunks, 0.f), b.s1[u], a1);
            } else {
                a0 = chunk_dot<WD>(b.w0[u], xs4, c, nchunks, a0);
                a1 = chunk_dot<WD>(b.w1[u], xs4, c, nchunks, a1);
            }
        }
    }
}

// ---- the shared kernel body ---------------------------------------------------------------------------
// Policy concept (all methods __device__):
//   int  units() const;                                  number of two-row units
//   void rows(int unit, int64_t& r0, int64_t& r1) const; weight row indices (into the policy's W)
//   const void* W() const; const float* scales() const; int group() const; int cols() const;
//   void stage(float* xs, float* red) const;             fills xs (after the dependency wait)
//   void emit(int unit, float s0, float s1) const;       called by lane 0 with the two finished sums
template <int WD, class Policy>
__device__ __forceinline__ void gemv_body(Policy& pol) {
    extern __shared__ __align__(16) float smem[];
    const int cols = pol.cols();
    float* xs = smem;
    float* red = smem + cols;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kGemvWarps + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * kGemvWarps;
    const int nchunks = cols / WInfo<WD>::E;
    const int nunits = pol.units();
    const int cpg = (WD == SLLM_INT8) ? pol.group() / 16 : 1;
    const int groups_per_row = (WD == SLLM_INT8) ? cols / pol.group() : 0;
    const uint4* Wv = reinterpret_cast<const uint4*>(pol.W());
    const float* Sc = pol.scales();

    // first batch of the first unit: weights only, so it may run ahead of the producer kernel (PDL)
    Batch<WD> cur;
    int unit = warp_global;
    int64_t r0 = 0, r1 = 0;
    if (unit < nunits) {
        pol.rows(unit, r0, r1);
        load_batch<WD>(cur, Wv + r0 * nchunks, Wv + r1 * nchunks, Sc + r0 * groups_per_row, Sc + r1 * groups_per_row, cpg, 0, lane, nchunks);
    }
#ifdef SLLM_PDL_EARLY
    pdl_launch_dependents();
#endif
    pdl_wait();
    pol.stage(xs, red);
    __syncthreads();
    const float4* xs4 = reinterpret_cast<const float4*>(xs);

    while (unit < nunits) {
        const uint4* w0 = Wv + r0 * nchunks;
        const uint4* w1 = Wv + r1 * nchunks;
        const float* s0 = Sc + r0 * groups_per_row;
        const float* s1 = Sc + r1 * groups_per_row;
        float a0 = 0.f, a1 = 0.f;
        for (int base = 0; base < nchunks; base += 32 * kGemvU) {
            fma_batch<WD>(cur, xs4, base, lane, nchunks, a0, a1);
            const int nb = base + 32 * kGemvU;
            if (nb < nchunks) load_batch<WD>(cur, w0, w1, s0, s1, cpg, nb, lane, nchunks);
        }
        // next unit's first batch goes out before this unit's reduction/epilogue
        const int this_unit = unit;
        unit += warps_total;
        if (unit < nunits) {
            pol.rows(unit, r0, r1);
            load_batch<WD>(cur, Wv + r0 * nchunks, Wv + r1 * nchunks, Sc + r0 * groups_per_row, Sc + r1 * groups_per_row, cpg, 0, lane, nchunks);
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) pol.emit(this_unit, a0, a1);
    }
#ifndef SLLM_PDL_EARLY
    // late trigger: this warp has issued all of its loads; once every CTA got here (or exited) the next
    // kernel's CTAs may take over the freed SM slots and start prefetching ITS weights
    pdl_launch_dependents();
#endif
}

// grid size for a GEMV-shaped kernel: a multiple of the SM count, no more CTAs than there are warps' worth
// of units.
extern int g_tune_ctas_per_sm;   // 0 = use the caller's value (sllm_tune)
inline int gemv_grid(int units, int ctas_per_sm) {
    const int sms = sm_count();
    if (g_tune_ctas_per_sm > 0) ctas_per_sm = g_tune_ctas_per_sm;
    const int need = (units + kGemvWarps - 1) / kGemvWarps;
    int g = sms * ctas_per_sm;
    if (need < g) g = need;
    return g < 1 ? 1 : g;
}

}  // namespace sllm
