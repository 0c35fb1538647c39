// csrc/prefill_attn.cu — causal attention for a block of T prompt positions (+ the small batched ops of the prefill
// path: embedding gather, RMSNorm with bf16 output).
//
// Attention replaces T calls of mha_kernel_cpu (reference source/kernel/cpu/mha_kernel.cpp:36-77): query t at
// position pos0+t attends to cache rows 0..pos0+t of its KV head. It is ~1 % of the prefill flops (the GEMMs on
// tcgen05 are the other 99 %), so it uses the warp-level mma.sync.m16n8k16 bf16 path with fp32 online softmax:
// CTA = 64 queries of one head (4 warps x 16 rows), K/V tiles of 64 positions double-buffered in shared memory
// with cp.async, XOR-swizzled 16-byte chunks (conflict-free ldmatrix), S = QK^T and O += PV with the P fragments
// taken straight from the S accumulators. Heaviest query tiles are scheduled first (causal imbalance).
#include "mega_common.cuh"
#include "prefill.cuh"

namespace sllm {
extern int g_tune_pf_pdl;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;   // src-size 0: destination is zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s_addr(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s_addr(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) { return (uint32_t)f32_to_bf16_bits(lo) | ((uint32_t)f32_to_bf16_bits(hi) << 16); }

constexpr int kPaBQ = 64, kPaBK = 64, kPaThreads = 128;

template <int HD, int KVD>
__global__ void __launch_bounds__(kPaThreads) pf_attn_kernel(const uint16_t* __restrict__ Q, const uint8_t* __restrict__ Kc,
                                                            const uint8_t* __restrict__ Vc, uint16_t* __restrict__ O, int T, int pos0, int S, int ldq,
                                                            int group, float scale_log2) {
    constexpr int CPR = HD / 8;                   // 16-byte chunks per (bf16) row
    constexpr int ESZ = (KVD == SLLM_F32) ? 4 : 2;
    extern __shared__ __align__(128) uint8_t pa_smem[];
    uint8_t* sQ = pa_smem;
    uint8_t* sK = sQ + kPaBQ * HD * 2;            // [2][BK][HD] bf16
    uint8_t* sV = sK + 2 * kPaBK * HD * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;
    const int h = blockIdx.y, kvh = h / group;
    const int t0 = qt * kPaBQ;
    const int nkeys = pos0 + min(T, t0 + kPaBQ);  // cache rows this tile may look at
    const int nkt = (nkeys + kPaBK - 1) / kPaBK;
    const uint8_t* Kh = Kc + (size_t)kvh * S * HD * ESZ;
    const uint8_t* Vh = Vc + (size_t)kvh * S * HD * ESZ;
    auto sw = [](int row, int chunk) { return (row * CPR + (chunk ^ (row & 7))) * 16; };
    pdl_wait();
    pdl_launch_dependents();

    for (int i = tid; i < kPaBQ * CPR; i += kPaThreads) {
        const int row = i / CPR, c = i - row * CPR;
        const int t = t0 + row;
        cp_async16(sQ + sw(row, c), Q + (size_t)min(t, T - 1) * ldq + h * HD + c * 8, t < T);
    }
    auto load_kv = [&](int kt, int stage) {
        uint8_t* dk = sK + (size_t)stage * kPaBK * HD * 2;
        uint8_t* dv = sV + (size_t)stage * kPaBK * HD * 2;
        for (int i = tid; i < kPaBK * CPR; i += kPaThreads) {
            const int row = i / CPR, c = i - row * CPR;
            const int key = kt * kPaBK + row;
            const bool valid = key < nkeys;
            const size_t off = ((size_t)min(key, nkeys - 1) * HD + c * 8) * ESZ;
            if (KVD == SLLM_BF16) {
                cp_async16(dk + sw(row, c), Kh + off, valid);
                cp_async16(dv + sw(row, c), Vh + off, valid);   // rows past the prompt are ZERO: 0 * garbage could be NaN in PV
            } else {
                uint4 kw = make_uint4(0, 0, 0, 0), vw = kw;
                if (valid) {
                    const float4 a = *reinterpret_cast<const float4*>(Kh + off), b = *reinterpret_cast<const float4*>(Kh + off + 16);
                    const float4 c4 = *reinterpret_cast<const float4*>(Vh + off), d4 = *reinterpret_cast<const float4*>(Vh + off + 16);
                    kw = make_uint4(bf16x2(a.x, a.y), bf16x2(a.z, a.w), bf16x2(b.x, b.y), bf16x2(b.z, b.w));
                    vw = make_uint4(bf16x2(c4.x, c4.y), bf16x2(c4.z, c4.w), bf16x2(d4.x, d4.y), bf16x2(d4.z, d4.w));
                }
                *reinterpret_cast<uint4*>(dk + sw(row, c)) = kw;
                *reinterpret_cast<uint4*>(dv + sw(row, c)) = vw;
            }
        }
    };
    load_kv(0, 0);
    cp_async_commit();

    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    uint32_t qf[HD / 16][4];
    const int qpos0 = pos0 + t0 + warp * 16 + g;   // position of this thread's first row (second: +8)

#pragma unroll 1
    for (int kt = 0; kt < nkt; ++kt) {
        const int stage = kt & 1;
        if (kt + 1 < nkt) {
            load_kv(kt + 1, stage ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) ldsm_x4(qf[ks], sQ + sw(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)));
        }
        const uint8_t* kS = sK + (size_t)stage * kPaBK * HD * 2;
        const uint8_t* vS = sV + (size_t)stage * kPaBK * HD * 2;
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {   // two 8-key tiles per ldmatrix.x4
                uint32_t r[4];
                ldsm_x4(r, kS + sw(np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, ks * 2 + ((lane >> 3) & 1)));
                mma_bf16_16816(s[2 * np], qf[ks], r[0], r[1]);
                mma_bf16_16816(s[2 * np + 1], qf[ks], r[2], r[3]);
            }
        }
        // scale (score = (q.k)/sqrt(hd), mha_kernel.cpp:59), causal mask, online softmax (softmax_kernel_cpu :7-20)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int key = kt * kPaBK + nt * 8 + 2 * tig + (i & 1);
                const int qp = qpos0 + (i >> 1) * 8;
                const float v = (key <= qp) ? s[nt][i] * scale_log2 : -INFINITY;
                s[nt][i] = v;
                mx[i >> 1] = fmaxf(mx[i >> 1], v);
            }
        }
        float alpha[2], mnew[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            mnew[r] = fmaxf(m_run[r], mx[r]);           // finite from the first tile on: key 0 is visible to every query
            alpha[r] = exp2f(m_run[r] - mnew[r]);
            m_run[r] = mnew[r];
            l_run[r] *= alpha[r];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float pv = exp2f(s[nt][i] - mnew[i >> 1]);
                s[nt][i] = pv;
                l_run[i >> 1] += pv;
            }
        }
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) { o[i][0] *= alpha[0]; o[i][1] *= alpha[0]; o[i][2] *= alpha[1]; o[i][3] *= alpha[1]; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {       // 16 keys per step: the S accumulators of key tiles 2kk, 2kk+1 ARE the A fragment
            uint32_t a[4] = {bf16x2(s[2 * kk][0], s[2 * kk][1]), bf16x2(s[2 * kk][2], s[2 * kk][3]), bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]),
                             bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
            for (int dn = 0; dn < HD / 8; dn += 2) {
                uint32_t r[4];
                ldsm_x4_trans(r, vS + sw(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dn + (lane >> 4)));
                mma_bf16_16816(o[dn], a, r[0], r[1]);
                mma_bf16_16816(o[dn + 1], a, r[2], r[3]);
            }
        }
        __syncthreads();   // everyone is done with this stage before the next iteration refills it
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const int ta = t0 + warp * 16 + g, tb = ta + 8;
    const float ia = 1.0f / l_run[0], ib = 1.0f / l_run[1];
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
        const int col = h * HD + dn * 8 + 2 * tig;
        if (ta < T) *reinterpret_cast<uint32_t*>(O + (size_t)ta * ldq + col) = bf16x2(o[dn][0] * ia, o[dn][1] * ia);
        if (tb < T) *reinterpret_cast<uint32_t*>(O + (size_t)tb * ldq + col) = bf16x2(o[dn][2] * ib, o[dn][3] * ib);
    }
}

template <int HD, int KVD>
static int attn_launch(const uint16_t* q, const void* kc, const void* vc, uint16_t* out, int T, int pos0, int S, int heads, int kv_heads, cudaStream_t st) {
    const size_t smem = (size_t)kPaBQ * HD * 2 + (size_t)4 * kPaBK * HD * 2;
    static bool configured = false;
    if (!configured) {
        SLLM_CUDA(cudaFuncSetAttribute(pf_attn_kernel<HD, KVD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const dim3 grid((T + kPaBQ - 1) / kPaBQ, heads);
    const float scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
    LaunchCfg lc(grid, dim3(kPaThreads), smem, st, g_tune_pf_pdl != 0);
    SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, pf_attn_kernel<HD, KVD>, q, reinterpret_cast<const uint8_t*>(kc), reinterpret_cast<const uint8_t*>(vc), out, T, pos0, S,
                                 heads * HD, heads / kv_heads, scale_log2));
    g_launches++;
    return SLLM_OK;
}

int pf_attention(const uint16_t* q, const void* kc, const void* vc, int kv_dtype, uint16_t* out, int T, int pos0, int S, int hd, int heads,
                 int kv_heads, cudaStream_t st) {
    SLLM_REQUIRE(q && kc && vc && out && T > 0 && pos0 >= 0 && pos0 + T <= S, SLLM_EINVAL, "pf_attention: bad argument");
    if (hd == 128) return kv_dtype == SLLM_F32 ? attn_launch<128, SLLM_F32>(q, kc, vc, out, T, pos0, S, heads, kv_heads, st)
                                               : attn_launch<128, SLLM_BF16>(q, kc, vc, out, T, pos0, S, heads, kv_heads, st);
    if (hd == 64) return kv_dtype == SLLM_F32 ? attn_launch<64, SLLM_F32>(q, kc, vc, out, T, pos0, S, heads, kv_heads, st)
                                              : attn_launch<64, SLLM_BF16>(q, kc, vc, out, T, pos0, S, heads, kv_heads, st);
    set_error("pf_attention: head_dim %d not instantiated (64, 128)", hd);
    return SLLM_ENOTSUP;
}

// ------------------------------------------------------------------------------------ small ops -------
__global__ void pf_embed_kernel(const int32_t* __restrict__ ids, const uint8_t* __restrict__ emb, int vocab, int nchunks, int KS, int SC, int R,
                                int tile_bytes, float* __restrict__ x, int d) {
    const int t = blockIdx.x;
    const int tok = min(max(ids[t], 0), vocab - 1);
    const uint8_t* trow = emb + (size_t)(tok / R) * KS * tile_bytes + (size_t)(tok % R) * SC * 16;
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
        const int ks = c / SC, cc = c - ks * SC;
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(trow + (size_t)ks * tile_bytes + (size_t)cc * 16));
        float4* dst = reinterpret_cast<float4*>(x + (size_t)t * d + c * 8);
        dst[0] = make_float4(bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y));
        dst[1] = make_float4(bf16_lo(raw.z), bf16_hi(raw.z), bf16_lo(raw.w), bf16_hi(raw.w));
    }
}

int pf_embed(const int32_t* ids_dev, const void* emb_tiled, int vocab, int d, float* x, int T, cudaStream_t st) {
    const TileGeom g = mega_tile_geom(((vocab + 1) / 2) * 2, d, SLLM_BF16);
    pf_embed_kernel<<<T, 128, 0, st>>>(ids_dev, reinterpret_cast<const uint8_t*>(emb_tiled), vocab, g.nchunks, g.KS, g.SC, g.R, g.tile_bytes, x, d);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

__global__ void __launch_bounds__(256) pf_rmsnorm_kernel(float* __restrict__ x, const float* __restrict__ add, const float* __restrict__ w,
                                                        uint16_t* __restrict__ y, int d, float eps) {
    __shared__ float red[33];
    const int t = blockIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    float4* xr = reinterpret_cast<float4*>(x + (size_t)t * d);
    const float4* ar = add ? reinterpret_cast<const float4*>(add + (size_t)t * d) : nullptr;
    constexpr int kKeep = 8;   // float4 per thread kept in registers between the two passes (rows up to 8192 floats are read once)
    float4 keep[kKeep];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kKeep; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < d / 4) {
            float4 v = xr[i];
            if (ar) {
                const float4 a = ar[i];
                v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
                xr[i] = v;
            }
            keep[k] = v;
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
    }
    for (int i = threadIdx.x + kKeep * 256; i < d / 4; i += 256) {
        float4 v = xr[i];
        if (ar) {
            const float4 a = ar[i];
            v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
            xr[i] = v;
        }
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = block_sum(ss, red);
    const float inv = 1.0f / sqrtf(ss / (float)d + eps);   // rms_kernel.cpp:17-19
    auto emit = [&](int i, const float4 v) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + i);
        *reinterpret_cast<uint2*>(y + (size_t)t * d + 4 * i) =
            make_uint2(bf16x2((v.x * inv) * g.x, (v.y * inv) * g.y), bf16x2((v.z * inv) * g.z, (v.w * inv) * g.w));
    };
#pragma unroll
    for (int k = 0; k < kKeep; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < d / 4) emit(i, keep[k]);
    }
    for (int i = threadIdx.x + kKeep * 256; i < d / 4; i += 256) emit(i, xr[i]);
}

int pf_rmsnorm(float* x, const float* add, const float* w, uint16_t* y, int T, int d, float eps, cudaStream_t st) {
    SLLM_REQUIRE(d % 4 == 0, SLLM_ENOTSUP, "pf_rmsnorm: d must be a multiple of 4");
    LaunchCfg lc(dim3(T), dim3(256), 0, st, g_tune_pf_pdl != 0);
    SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, pf_rmsnorm_kernel, x, add, w, y, d, eps));
    g_launches++;
    return SLLM_OK;
}

}  // namespace sllm

extern "C" int sllm_prefill_attention(const void* q, const void* kc, const void* vc, int32_t kv_dtype, void* out, int32_t T, int32_t pos0, int32_t max_len,
                                      int32_t head_dim, int32_t heads, int32_t kv_heads, sllm_stream_t stream) {
    using namespace sllm;
    SLLM_REQUIRE(heads > 0 && kv_heads > 0 && heads % kv_heads == 0, SLLM_EINVAL, "bad head counts");
    return pf_attention(reinterpret_cast<const uint16_t*>(q), kc, vc, kv_dtype, reinterpret_cast<uint16_t*>(out), T, pos0, max_len, head_dim, heads,
                        kv_heads, as_stream(stream));
}
