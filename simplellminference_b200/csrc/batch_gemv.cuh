// csrc/batch_gemv.cuh — the decode GEMV for several sequences at once (batched multi-sequence decode, batch.cu).
//
// Y[b][r] = sum_j X[b][j] * W[r][j] for up to NB activation vectors against ONE pass over the weight matrix: the
// weights are the HBM traffic of a decode step, so B sequences stepping together read them once instead of B times
// ("turns GEMV into skinny GEMM"). Same kernel shape as gemv_core.cuh, whose building blocks it
// reuses unchanged (plane-staged activations, 128-bit streaming weight loads, two-row units, warp-shuffle
// reductions): every sequence's dot product is accumulated in exactly the order of the one-sequence kernel, so a
// GEMV result does not depend on which other sequences share the launch. Activations stay fp32 (the decode parity
// contract), so the contraction runs on the FMA pipes: 8 FMAs per weight at NB = 8; larger batches go through the launcher
// in groups of 8 (or fewer when 8 vectors of `cols` floats do not fit shared memory).
#pragma once
#include "decode_fused.cuh"   // gemv_core.cuh + the peer-memory helpers its RMSNorm staging refers to
#include "paged_kv.cuh"

namespace sllm {

constexpr int kBatchMaxNb = 8;   // activation vectors per launch

__host__ __device__ inline size_t bgemv_smem_bytes(int cols, int nb) { return (size_t)nb * cols * sizeof(float) + 64 * sizeof(float); }

// Policy concept: GemvBase plus
//   int  nb() const;                                       vectors in this launch (<= NB)
//   int  units() const; void rows(int unit, int64_t&, int64_t&) const;     as in gemv_core.cuh
//   void stage(int b, float* xs, float* red) const;        fills vector b (all threads of the CTA)
//   void emit(int unit, int b, float s0, float s1) const;  lane 0, the unit's two finished sums of vector b
template <int WD, int NB, class Policy>
__device__ __forceinline__ void bgemv_body(Policy& pol) {
    extern __shared__ __align__(16) float smem[];
    const int cols = pol.cols();
    const int nb = pol.nb();
    float* red = smem + (size_t)nb * cols;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kGemvWarps + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * kGemvWarps;
    const int nchunks = cols / WInfo<WD>::E;
    const int nunits = pol.units();
    const int cpg = (WD == SLLM_INT8) ? pol.group() / 16 : 1;
    const int groups_per_row = (WD == SLLM_INT8) ? cols / pol.group() : 0;
    const uint4* Wv = reinterpret_cast<const uint4*>(pol.W());
    const float* Sc = pol.scales();

    // the first batch of weight loads is in flight while the activations are staged
    Batch<WD> cur;
    int unit = warp_global;
    int64_t r0 = 0, r1 = 0;
    if (unit < nunits) {
        pol.rows(unit, r0, r1);
        load_batch<WD>(cur, Wv + r0 * nchunks, Wv + r1 * nchunks, Sc + r0 * groups_per_row, Sc + r1 * groups_per_row, cpg, 0, lane, nchunks);
    }
    for (int b = 0; b < nb; ++b) pol.stage(b, smem + (size_t)b * cols, red);
    __syncthreads();
    const float4* xs4 = reinterpret_cast<const float4*>(smem);
    const int xstride4 = cols / 4;

    while (unit < nunits) {
        const uint4* w0 = Wv + r0 * nchunks;
        const uint4* w1 = Wv + r1 * nchunks;
        const float* s0 = Sc + r0 * groups_per_row;
        const float* s1 = Sc + r1 * groups_per_row;
        float a0[NB], a1[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) { a0[b] = 0.f; a1[b] = 0.f; }
        for (int base = 0; base < nchunks; base += 32 * kGemvU) {
#pragma unroll
            for (int u = 0; u < kGemvU; ++u) {
                const int c = base + u * 32 + lane;
                if (c < nchunks) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        if (b < nb) {
                            const float4* xb = xs4 + (size_t)b * xstride4;
                            if (WD == SLLM_INT8) {
                                a0[b] = fmaf(chunk_dot<WD>(cur.w0[u], xb, c, nchunks, 0.f), cur.s0[u], a0[b]);
                                a1[b] = fmaf(chunk_dot<WD>(cur.w1[u], xb, c, nchunks, 0.f), cur.s1[u], a1[b]);
                            } else {
                                a0[b] = chunk_dot<WD>(cur.w0[u], xb, c, nchunks, a0[b]);
                                a1[b] = chunk_dot<WD>(cur.w1[u], xb, c, nchunks, a1[b]);
                            }
                        }
                    }
                }
            }
            const int nxt = base + 32 * kGemvU;
            if (nxt < nchunks) load_batch<WD>(cur, w0, w1, s0, s1, cpg, nxt, lane, nchunks);
        }
        const int this_unit = unit;
        unit += warps_total;
        if (unit < nunits) {   // next unit's first batch goes out before this unit's reductions
            pol.rows(unit, r0, r1);
            load_batch<WD>(cur, Wv + r0 * nchunks, Wv + r1 * nchunks, Sc + r0 * groups_per_row, Sc + r1 * groups_per_row, cpg, 0, lane, nchunks);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < nb) {   // uniform over the warp
                const float t0 = warp_sum(a0[b]), t1 = warp_sum(a1[b]);
                if (lane == 0) pol.emit(this_unit, b, t0, t1);
            }
        }
    }
}

template <int WD, int NB, class Policy>
__global__ void __launch_bounds__(kGemvThreads) bgemv_kernel(Policy pol) {
    bgemv_body<WD, NB>(pol);
}

// ---- experimental variant (sllm_tune key 6): TWO units = four weight rows per warp at a time -----------------------------
// Every 16-byte chunk of activations read from shared memory then meets four weight rows instead of two: half the LDS
// traffic per FMA, which is what bounds the two-row body at 8 vectors (2 x NB LDS.128 per chunk column against 16 x NB FMAs;
// an SM moves 128 B of shared memory per clock). Four loads per row per lane in flight, so the bytes in flight per warp stay
// 8 KB. A lane visits its chunks in the same ascending order, so every sum is bit-identical to the two-row body's.
constexpr int kBgemvU4 = 4;

template <int WD>
struct Batch4 {
    uint4 w[4][kBgemvU4];
    float s[4][kBgemvU4];   // int8 only: group scale of each chunk
};

template <int WD>
__device__ __forceinline__ void load_batch4(Batch4<WD>& b, const uint4* const* rp, const float* const* sp, int chunks_per_group, int base,
                                            int lane, int nchunks) {
#pragma unroll
    for (int u = 0; u < kBgemvU4; ++u) {
        const int c = base + u * 32 + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (c < nchunks) {
                b.w[k][u] = ldg_stream(rp[k] + c);
                if (WD == SLLM_INT8) b.s[k][u] = __ldg(sp[k] + c / chunks_per_group);
            } else {
                b.w[k][u] = make_uint4(0, 0, 0, 0);
                if (WD == SLLM_INT8) b.s[k][u] = 0.f;
            }
        }
    }
}

template <int WD, int NB, class Policy>
__device__ __forceinline__ void bgemv_body4(Policy& pol) {
    extern __shared__ __align__(16) float smem[];
    const int cols = pol.cols();
    const int nb = pol.nb();
    float* red = smem + (size_t)nb * cols;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kGemvWarps + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * kGemvWarps;
    const int nchunks = cols / WInfo<WD>::E;
    const int nunits = pol.units();
    const int cpg = (WD == SLLM_INT8) ? pol.group() / 16 : 1;
    const int groups_per_row = (WD == SLLM_INT8) ? cols / pol.group() : 0;
    const uint4* Wv = reinterpret_cast<const uint4*>(pol.W());
    const float* Sc = pol.scales();

    // a policy may contract only a column range [col0, col0 + cols) of rows that are row_cols long (K split over grid.y)
    const int row_chunks = pol.row_cols() / WInfo<WD>::E, chunk0 = pol.col0() / WInfo<WD>::E;
    const int row_groups = (WD == SLLM_INT8) ? pol.row_cols() / pol.group() : 0, group0 = (WD == SLLM_INT8) ? pol.col0() / pol.group() : 0;
    (void)groups_per_row;

    // this warp's units: (unit, unit + warps_total), then both advance by 2 * warps_total
    const uint4* rp[4];
    const float* sp[4];
    auto set_rows = [&](int ua) {
        int64_t r[4];
        pol.rows(ua, r[0], r[1]);
        if (ua + warps_total < nunits) pol.rows(ua + warps_total, r[2], r[3]);
        else { r[2] = r[0]; r[3] = r[1]; }   // no second unit: it aliases the first and is not emitted
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            rp[k] = Wv + r[k] * row_chunks + chunk0;
            sp[k] = Sc + r[k] * row_groups + group0;
        }
    };
    Batch4<WD> cur;
    int unit = warp_global;
    if (unit < nunits) {
        set_rows(unit);
        load_batch4<WD>(cur, rp, sp, cpg, 0, lane, nchunks);
    }
    for (int b = 0; b < nb; ++b) pol.stage(b, smem + (size_t)b * cols, red);
    __syncthreads();
    const float4* xs4 = reinterpret_cast<const float4*>(smem);
    const int xstride4 = cols / 4;

    while (unit < nunits) {
        float a[4][NB];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < NB; ++b) a[k][b] = 0.f;
        for (int base = 0; base < nchunks; base += 32 * kBgemvU4) {
#pragma unroll
            for (int u = 0; u < kBgemvU4; ++u) {
                const int c = base + u * 32 + lane;
                if (c < nchunks) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        if (b < nb) {
                            const float4* xb = xs4 + (size_t)b * xstride4;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (WD == SLLM_INT8) a[k][b] = fmaf(chunk_dot<WD>(cur.w[k][u], xb, c, nchunks, 0.f), cur.s[k][u], a[k][b]);
                                else a[k][b] = chunk_dot<WD>(cur.w[k][u], xb, c, nchunks, a[k][b]);
                            }
                        }
                    }
                }
            }
            const int nxt = base + 32 * kBgemvU4;
            if (nxt < nchunks) load_batch4<WD>(cur, rp, sp, cpg, nxt, lane, nchunks);
        }
        const int ua = unit, ub = unit + warps_total;
        unit += 2 * warps_total;
        if (unit < nunits) {   // the next pair's first loads go out before this pair's reductions
            set_rows(unit);
            load_batch4<WD>(cur, rp, sp, cpg, 0, lane, nchunks);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < nb) {   // uniform over the warp
                const float t0 = warp_sum(a[0][b]), t1 = warp_sum(a[1][b]), t2 = warp_sum(a[2][b]), t3 = warp_sum(a[3][b]);
                if (lane == 0) {
                    pol.emit(ua, b, t0, t1);
                    if (ub < nunits) pol.emit(ub, b, t2, t3);
                }
            }
        }
    }
}

template <int WD, int NB, class Policy>
__global__ void __launch_bounds__(kGemvThreads) bgemv4_kernel(Policy pol) {
    bgemv_body4<WD, NB>(pol);
}

// ---- policies: the decode_fused.cuh epilogues with a sequence-slot index -------------------------------------
// Slot s = b0 + b (b0 = first slot of this launch's group). Per-slot arrays are [slots][...] with the stated stride.
struct BatchBase : GemvBase {
    int nb_, b0;
    __device__ int nb() const { return nb_; }
    __device__ int row_cols() const { return cols_; }   // length of a stored weight row (four-row body only: see BDownSplitPolicy)
    __device__ int col0() const { return 0; }           // first column this launch contracts
};

// A: RMSNorm -> [Wq;Wk;Wv] -> RoPE(q, k) -> q buffer, K/V rows of position pos[s] in the slot's page
template <int WD>
struct BQkvPolicy : BatchBase {
    const float* x;          // residual stream [slots][cols]
    const float* norm_w;     // [cols]
    float eps;
    const float* sin_t;      // [S][hd/2]
    const float* cos_t;
    float* q_out;            // [slots][q_dim]
    void* k_pool;            // paged_kv.cuh layout
    void* v_pool;
    PagedKv pk;              // block table, positions, page geometry (q_stride / heads unused here)
    int layer, kv_dtype, q_dim, kv_dim, hd;
    __device__ int units() const { return (q_dim + 2 * kv_dim) >> 1; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const {
        const int half = hd >> 1;
        const int rope_units = (q_dim + kv_dim) >> 1;
        if (u < rope_units) {  // q or k head: RoPE partners (j, j + hd/2)
            const int head = u / half, j = u - head * half;
            r0 = (int64_t)head * hd + j;
            r1 = r0 + half;
        } else {  // v: two consecutive rows
            r0 = (int64_t)q_dim + kv_dim + 2 * (int64_t)(u - rope_units);
            r1 = r0 + 1;
        }
    }
    __device__ void stage(int b, float* xs, float* red) const {
        stage_x_rmsnorm<WD>(xs, red, x + (size_t)(b0 + b) * cols_, norm_w, cols_, eps);
    }
    __device__ void store_kv(void* pool, size_t idx, float v) const {
        if (kv_dtype == SLLM_BF16) reinterpret_cast<uint16_t*>(pool)[idx] = f32_to_bf16_bits(v);
        else reinterpret_cast<float*>(pool)[idx] = v;
    }
    __device__ void emit(int u, int b, float s0, float s1) const {
        const int slot = b0 + b;
        const int pos = pk.pos[slot];
        if (pos < 0) return;   // slot not in use
        const int half = hd >> 1;
        const int rope_units = (q_dim + kv_dim) >> 1;
        const int kv_heads = kv_dim / hd;
        const int pi = pos / pk.page_len;
        const int page = pk.block_table[(size_t)slot * pk.max_pages + pi];
        const int in_page = pos - pi * pk.page_len;
        if (u < rope_units) {
            const int head = u / half, j = u - head * half;
            const float fci = sin_t[(int64_t)pos * half + j], fcr = cos_t[(int64_t)pos * half + j];
            const float o0 = s0 * fcr - s1 * fci;  // rope_kernel.cpp:36-37
            const float o1 = s1 * fcr + s0 * fci;
            const int r0 = head * hd + j;
            if (r0 < q_dim) {
                q_out[(size_t)slot * q_dim + r0] = o0;
                q_out[(size_t)slot * q_dim + r0 + half] = o1;
            } else {
                const int kvh = head - q_dim / hd;
                const size_t idx = paged_row_index(page, pk.layers, layer, kv_heads, kvh, pk.page_len, in_page, hd) + j;
                store_kv(k_pool, idx, o0);
                store_kv(k_pool, idx + half, o1);
            }
        } else {
            const int c = 2 * (u - rope_units);   // element of the value row (even; its partner c + 1 is in the same head)
            const int kvh = c / hd, j = c - kvh * hd;
            const size_t idx = paged_row_index(page, pk.layers, layer, kv_heads, kvh, pk.page_len, in_page, hd) + j;
            store_kv(v_pool, idx, s0);
            store_kv(v_pool, idx + 1, s1);
        }
    }
};

// C / E: GEMV + residual
template <int WD>
struct BResidualPolicy : BatchBase {
    const float* x;       // GEMV input [slots][cols]
    const float* resid;   // [slots][nrows]
    float* y;             // [slots][nrows]
    int nrows;
    __device__ int units() const { return (nrows + 1) >> 1; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const { r0 = 2 * (int64_t)u; r1 = min(2 * u + 1, nrows - 1); }
    __device__ void stage(int b, float* xs, float*) const { stage_x_plain<WD>(xs, x + (size_t)(b0 + b) * cols_, cols_); }
    __device__ void emit(int u, int b, float s0, float s1) const {
        const size_t o = (size_t)(b0 + b) * nrows + 2 * (size_t)u;
        y[o] = resid[o] + s0;  // add_kernel.cpp:10-13: out = in1 + in2
        if (2 * u + 1 < nrows) y[o + 1] = resid[o + 1] + s1;
    }
};

// E', experimental (sllm_tune key 7, four-row body only): the down projection with K cut in two over grid.y, so that 8 vectors
// of HALF a row fit shared memory and Wdown is read once for up to 8 sequences instead of once per 5. Each half writes its
// partial sums; batch_add_partials_kernel then forms x = h + (p0 + p1). cols_ = the half's length.
template <int WD>
struct BDownSplitPolicy : BatchBase {
    const float* x;       // GEMV input [slots][row_cols_]
    float* part;          // [2][slots_total][nrows] partial sums
    int nrows, row_cols_, slots_total;
    __device__ int row_cols() const { return row_cols_; }
    __device__ int col0() const { return (int)blockIdx.y * cols_; }
    __device__ int units() const { return (nrows + 1) >> 1; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const { r0 = 2 * (int64_t)u; r1 = min(2 * u + 1, nrows - 1); }
    __device__ void stage(int b, float* xs, float*) const { stage_x_plain<WD>(xs, x + (size_t)(b0 + b) * row_cols_ + col0(), cols_); }
    __device__ void emit(int u, int b, float s0, float s1) const {
        const size_t o = ((size_t)blockIdx.y * slots_total + (b0 + b)) * nrows + 2 * (size_t)u;
        part[o] = s0;
        if (2 * u + 1 < nrows) part[o + 1] = s1;
    }
};

// D: RMSNorm -> up/gate -> sigmoid(gate)*up
template <int WD>
struct BGateUpPolicy : BatchBase {
    const float* h;        // [slots][cols]
    const float* norm_w;
    float eps;
    float* s_out;          // [slots][inter]
    int inter;             // W = [up rows (inter)][gate rows (inter)]
    __device__ int units() const { return inter; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const { r0 = u; r1 = (int64_t)inter + u; }
    __device__ void stage(int b, float* xs, float* red) const {
        stage_x_rmsnorm<WD>(xs, red, h + (size_t)(b0 + b) * cols_, norm_w, cols_, eps);
    }
    __device__ void emit(int u, int b, float up, float gate) const {
        const float sg = 1.0f / (1.0f + expf(-gate));  // swiglu_kernel.cpp:12-13
        s_out[(size_t)(b0 + b) * inter + u] = sg * up;
    }
};

// F: RMSNorm -> tied classifier -> logits (arg-max and token feedback: batch_argmax_feedback_kernel)
template <int WD>
struct BClsPolicy : BatchBase {
    const float* x;        // [slots][cols]
    const float* norm_w;
    float eps;
    float* logits;         // [slots][nrows]
    int nrows;
    __device__ int units() const { return (nrows + 1) >> 1; }
    __device__ void rows(int u, int64_t& r0, int64_t& r1) const { r0 = 2 * (int64_t)u; r1 = min(2 * u + 1, nrows - 1); }
    __device__ void stage(int b, float* xs, float* red) const {
        stage_x_rmsnorm<WD>(xs, red, x + (size_t)(b0 + b) * cols_, norm_w, cols_, eps);
    }
    __device__ void emit(int u, int b, float s0, float s1) const {
        const size_t o = (size_t)(b0 + b) * nrows + 2 * (size_t)u;
        logits[o] = s0;
        if (2 * u + 1 < nrows) logits[o + 1] = s1;
    }
};

}  // namespace sllm
