// csrc/mha.cu — sllm_mha_decode: split-KV flash-decoding for one query token (replaces
// kernel::mha_kernel_cuda, reference source/kernel/cuda/mha_kernel.cu:133-169; semantics from the CPU kernel
// source/kernel/cpu/mha_kernel.cpp:40-76).
//
// Work decomposition: grid = (kv_heads, nsplit). A CTA owns one KV head and a contiguous range of cache
// positions; it serves all `g = heads/kv_heads` query heads that share that KV head, so each K/V byte is read
// from HBM once per token (GQA sharing). Tiles of T=64 positions are staged in shared memory by the TMA bulk
// copy engine (cp.async.bulk, one 16-byte-aligned row segment per copy, completion on an mbarrier), two
// stages deep, rows padded to a stride = 32 (mod 128) bytes so that both the score phase (2 threads per key,
// 16-byte LDS) and the PV phase (16-byte chunks along the head dimension) are bank-conflict free.
// Softmax is the online (running max / running sum) form; the nsplit partial results (m, l, unnormalised O)
// are merged by whichever CTA of the KV head finishes last (atomic ticket), so there is no second kernel and
// no host-visible scratch besides `workspace`.
//
// The same body serves the batched multi-sequence decode over a PAGED cache (batch.cu): grid.z = sequence slot, every
// slot has its own position, and a cache row is found through the slot's block table (pool layout
// [pages][layers][kv_heads][page_len][hd]: the rows of one (page, layer, head) are contiguous). PAGED is a compile-time
// switch; the dense instantiation is the kernel above, unchanged.
//
// Numerics: fp32 scores and accumulators, accurate expf, score = (q.k) * (1/sqrt(hd)) applied after the sum
// like the reference (matmul_kernel.cpp:26 via mha_kernel.cpp:59). Only the summation order differs from the
// oracle's serial loops.
#include <cmath>

#include "common.cuh"
#include "paged_kv.cuh"

namespace sllm {

constexpr int kMhaThreads = 128;
constexpr int kMhaTile = 64;     // positions per shared-memory tile
constexpr int kMhaStages = 2;
constexpr int kMhaMaxGroup = 8;  // query heads per KV head handled by one CTA
constexpr int kMhaMaxSplit = 64;

__host__ __device__ inline int mha_row_stride(int row_bytes) {  // smallest s >= row_bytes with s % 128 == 32
    int s = (row_bytes / 128) * 128 + 32;
    if (s < row_bytes) s += 128;
    return s;
}

// fixed (graph-capturable) number of KV splits: enough CTAs for ~2 per SM, never more than one tile's worth
// of positions per split at full context.
inline int mha_nsplit(int kv_heads, int max_len) {
    int n = (2 * sm_count() + kv_heads - 1) / kv_heads;
    const int by_len = (max_len + kMhaTile - 1) / kMhaTile;
    if (n > by_len) n = by_len;
    if (n > kMhaMaxSplit) n = kMhaMaxSplit;
    return n < 1 ? 1 : n;
}

struct MhaSmem {
    size_t q_off, p_off, alpha_off, ml_off, o_off, k_off, v_off, total;
    int stride;
};
// padded = rows copied one by one (dense cache: a head's rows are kv_dim apart). Paged pools keep a head's rows of a page contiguous:
// unpadded rows there, so that ONE bulk copy brings a whole run of rows (the 16-byte accesses of QK / PV are 128 contiguous bytes per
// quarter warp either way: no bank conflicts without the padding)
__host__ __device__ inline MhaSmem mha_smem_layout(int hd, int g, int esz, bool padded = true) {
    MhaSmem L;
    L.stride = padded ? mha_row_stride(hd * esz) : hd * esz;
    size_t off = 64;                                   // mbarriers
    L.q_off = off; off += (size_t)g * hd * 4;
    L.p_off = off; off += (size_t)g * kMhaTile * 4;
    L.alpha_off = off; off += (size_t)kMhaMaxGroup * 4;
    L.ml_off = off; off += (size_t)kMhaMaxGroup * 2 * 4;
    off = (off + 15) & ~(size_t)15;
    L.k_off = off; off += (size_t)kMhaStages * kMhaTile * L.stride;
    L.v_off = off; off += (size_t)kMhaStages * kMhaTile * L.stride;
    // the cross-stripe reduction scratch for O ([stripes][g][hd] fp32) reuses the K/V stages once they are drained
    L.o_off = L.k_off;
    const size_t stripes = kMhaThreads / (size_t)(hd * esz / 16);
    const size_t o_end = L.o_off + stripes * g * hd * 4;
    L.total = off > o_end ? off : o_end;
    return L;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int KVD> struct KvInfo;
template <> struct KvInfo<SLLM_F32> { static constexpr int ESZ = 4; static constexpr int VEC = 4; };
template <> struct KvInfo<SLLM_BF16> { static constexpr int ESZ = 2; static constexpr int VEC = 8; };

template <int KVD>
__device__ __forceinline__ void unpack16(const uint4 v, float* f) {
    if (KVD == SLLM_F32) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    } else {
        f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
        f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
    }
}

// G = query heads per KV head (compile-time so the accumulators live in registers)
template <int KVD, int G, bool PAGED>
__device__ __forceinline__ void
mha_decode_body(const float* __restrict__ q, const uint8_t* __restrict__ kc, const uint8_t* __restrict__ vc,
                float* __restrict__ out, float* __restrict__ partials, int* __restrict__ counters, int layer,
                const int32_t* __restrict__ pos_dev, int pos_val, int max_len, int hd, int kv_heads, int nsplit,
                const PagedKv& pk) {
    constexpr int ESZ = KvInfo<KVD>::ESZ, VEC = KvInfo<KVD>::VEC;
    extern __shared__ __align__(128) uint8_t smem[];
    const MhaSmem L = mha_smem_layout(hd, G, ESZ, !PAGED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* q_s = reinterpret_cast<float*>(smem + L.q_off);
    float* p_s = reinterpret_cast<float*>(smem + L.p_off);
    float* alpha_s = reinterpret_cast<float*>(smem + L.alpha_off);
    float* ml_s = reinterpret_cast<float*>(smem + L.ml_off);  // [G][2] running (max, sum)
    uint8_t* k_s = smem + L.k_off;
    uint8_t* v_s = smem + L.v_off;
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kvh = blockIdx.x, split = blockIdx.y;
    const int row_bytes = hd * ESZ;
    const int kv = kv_heads * hd;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();  // q and the newest cache row come from the previous kernel; pos from the previous step
    int pos;
    const int32_t* bt = nullptr;
    if constexpr (PAGED) {
        const int slot = blockIdx.z;
        pos = pk.pos[slot];
        if (pos < 0) return;   // slot not in use (uniform over the CTA; nothing is pending on the mbarriers yet)
        bt = pk.block_table + (size_t)slot * pk.max_pages;
        q += (size_t)slot * pk.q_stride;
        out += (size_t)slot * pk.q_stride;
        partials += (size_t)slot * pk.heads * nsplit * (hd + 2);
        counters += (size_t)slot * kv_heads;
    } else {
        pos = pos_dev ? *pos_dev : pos_val;
    }
    const int n = pos + 1;
    const int per = (n + nsplit - 1) / nsplit;
    const int t0 = split * per;
    const int t1 = min(n, t0 + per);
    const int ntiles = (t1 > t0) ? (t1 - t0 + kMhaTile - 1) / kMhaTile : 0;

    for (int i = tid; i < G * hd; i += kMhaThreads) q_s[i] = q[(size_t)(kvh * G) * hd + i];
    if (tid < G) { ml_s[2 * tid] = -INFINITY; ml_s[2 * tid + 1] = 0.f; }
    __syncthreads();  // also publishes the mbarrier init

    const size_t head_off = ((size_t)layer * max_len) * kv * ESZ + (size_t)kvh * row_bytes;
    auto issue_tile = [&](int tile) {  // warp 0 only
        const int stage = tile & 1;
        const int ts = t0 + tile * kMhaTile;
        const int rows = min(kMhaTile, t1 - ts);
        if (lane == 0) mbar_expect_tx(&bars[stage], (uint32_t)(2 * rows * row_bytes));
        __syncwarp();
        if constexpr (PAGED) {
            // one bulk copy per run of rows inside a page (a 64-position tile over pages of 64 = ONE 16 KB copy for K and one for V, instead
            // of 128 copies of 256 bytes: the copy engine's request rate, not HBM, bounded the row-by-row version at ~2.7 TB/s)
            const int first_page = ts / pk.page_len, last_page = (ts + rows - 1) / pk.page_len;
            for (int pi = first_page + lane; pi <= last_page; pi += 32) {
                const int ta = max(ts, pi * pk.page_len), tb = min(ts + rows, (pi + 1) * pk.page_len);
                const size_t g_off = paged_row_index(bt[pi], pk.layers, layer, kv_heads, kvh, pk.page_len, ta - pi * pk.page_len, hd) * ESZ;
                const uint32_t bytes = (uint32_t)(tb - ta) * (uint32_t)row_bytes;
                bulk_g2s(k_s + ((size_t)stage * kMhaTile + (ta - ts)) * L.stride, kc + g_off, bytes, &bars[stage]);
                bulk_g2s(v_s + ((size_t)stage * kMhaTile + (ta - ts)) * L.stride, vc + g_off, bytes, &bars[stage]);
            }
            return;
        }
        for (int r = lane; r < rows; r += 32) {
            size_t g_off;
            if constexpr (PAGED) {
                const int t = ts + r, pi = t / pk.page_len;
                g_off = paged_row_index(bt[pi], pk.layers, layer, kv_heads, kvh, pk.page_len, t - pi * pk.page_len, hd) * ESZ;
            } else {
                g_off = head_off + (size_t)(ts + r) * kv * ESZ;
            }
            bulk_g2s(k_s + ((size_t)stage * kMhaTile + r) * L.stride, kc + g_off, row_bytes, &bars[stage]);
            bulk_g2s(v_s + ((size_t)stage * kMhaTile + r) * L.stride, vc + g_off, row_bytes, &bars[stage]);
        }
    };
    if (warp == 0) {
        if (ntiles > 0) issue_tile(0);
        if (ntiles > 1) issue_tile(1);
    }

    // PV mapping: thread -> (16-byte chunk of the head dim, stripe of positions)
    const int chunks_per_row = row_bytes / 16;
    const int nstripes = kMhaThreads / chunks_per_row;
    const int pv_chunk = tid % chunks_per_row, pv_stripe = tid / chunks_per_row;
    const bool pv_active = pv_stripe < nstripes;
    float acc[G][VEC];
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[gi][e] = 0.f;

    // score mapping: two adjacent lanes share a key; lane parity picks the even / odd 16-byte chunks
    const int key = tid >> 1, part = tid & 1;
    const float scale = 1.0f / sqrtf((float)hd);

    for (int tile = 0; tile < ntiles; ++tile) {
        const int stage = tile & 1;
        const int ts = t0 + tile * kMhaTile;
        const int rows = min(kMhaTile, t1 - ts);
        mbar_wait(&bars[stage], (uint32_t)((tile >> 1) & 1));

        // ---- scores: s[gi][key] = (q_gi . K[key]) * scale
        {
            float s[G];
#pragma unroll
            for (int gi = 0; gi < G; ++gi) s[gi] = 0.f;
            if (key < rows) {
                const uint8_t* krow = k_s + ((size_t)stage * kMhaTile + key) * L.stride;
                for (int c = part; c < chunks_per_row; c += 2) {
                    float kf[VEC];
                    unpack16<KVD>(*reinterpret_cast<const uint4*>(krow + c * 16), kf);
#pragma unroll
                    for (int gi = 0; gi < G; ++gi) {
                        const float* qv = q_s + gi * hd + c * VEC;
#pragma unroll
                        for (int e = 0; e < VEC; ++e) s[gi] = fmaf(qv[e], kf[e], s[gi]);
                    }
                }
            }
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                s[gi] += __shfl_xor_sync(0xffffffffu, s[gi], 1);
                if (part == 0) p_s[gi * kMhaTile + key] = (key < rows) ? s[gi] * scale : -INFINITY;
            }
        }
        __syncthreads();

        // ---- online softmax bookkeeping: warp w handles heads w, w+4, ...
        for (int gi = warp; gi < G; gi += kMhaThreads / 32) {
            const float s0 = p_s[gi * kMhaTile + lane], s1 = p_s[gi * kMhaTile + lane + 32];
            const float m_old = ml_s[2 * gi], l_old = ml_s[2 * gi + 1];
            const float m_new = fmaxf(m_old, warp_max(fmaxf(s0, s1)));
            const float e0 = expf(s0 - m_new), e1 = expf(s1 - m_new);
            const float a = expf(m_old - m_new);  // 0 on the first tile (m_old = -inf)
            const float l_new = l_old * a + warp_sum(e0 + e1);
            p_s[gi * kMhaTile + lane] = e0;
            p_s[gi * kMhaTile + lane + 32] = e1;
            if (lane == 0) { alpha_s[gi] = a; ml_s[2 * gi] = m_new; ml_s[2 * gi + 1] = l_new; }
        }
        __syncthreads();

        // ---- PV: acc = acc*alpha + sum_t p_t * V[t]
        if (pv_active) {
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                const float a = alpha_s[gi];
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[gi][e] *= a;
            }
            for (int r = pv_stripe; r < rows; r += nstripes) {
                float vf[VEC];
                unpack16<KVD>(*reinterpret_cast<const uint4*>(v_s + ((size_t)stage * kMhaTile + r) * L.stride + pv_chunk * 16), vf);
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    const float p = p_s[gi * kMhaTile + r];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[gi][e] = fmaf(p, vf[e], acc[gi][e]);
                }
            }
        }
        __syncthreads();  // everyone is done with this stage and with p_s
        if (warp == 0 && tile + 2 < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_tile(tile + 2);
        }
    }

    pdl_launch_dependents();  // late trigger: all of this CTA's HBM reads are done
    // ---- reduce the stripes' accumulators through shared memory (reuses the K stages)
    float* o_s = reinterpret_cast<float*>(smem + L.o_off);  // [nstripes][G][hd]
    if (pv_active) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi)
#pragma unroll
            for (int e = 0; e < VEC; ++e) o_s[((size_t)pv_stripe * G + gi) * hd + pv_chunk * VEC + e] = acc[gi][e];
    }
    __syncthreads();

    const int heads = kv_heads * G;
    (void)heads;
    if (nsplit == 1) {
        for (int i = tid; i < G * hd; i += kMhaThreads) {
            float o = 0.f;
            for (int s = 0; s < nstripes; ++s) o += o_s[(size_t)s * G * hd + i];
            out[(size_t)(kvh * G) * hd + i] = o / ml_s[2 * (i / hd) + 1];
        }
        return;
    }

    // partial record per (head, split): [hd] unnormalised O, then m, l
    const int rec = hd + 2;
    for (int i = tid; i < G * hd; i += kMhaThreads) {
        float o = 0.f;
        for (int s = 0; s < nstripes; ++s) o += o_s[(size_t)s * G * hd + i];
        const int gi = i / hd, j = i - gi * hd;
        partials[((size_t)(kvh * G + gi) * nsplit + split) * rec + j] = o;
    }
    if (tid < G) {
        float* r = partials + ((size_t)(kvh * G + tid) * nsplit + split) * rec + hd;
        r[0] = ml_s[2 * tid];
        r[1] = ml_s[2 * tid + 1];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&counters[kvh], 1) == nsplit - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = tid; i < G * hd; i += kMhaThreads) {
        const int gi = i / hd, j = i - gi * hd;
        const float* base = partials + (size_t)(kvh * G + gi) * nsplit * rec;
        float M = -INFINITY;
        for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(base + (size_t)s * rec + hd));
        float Lsum = 0.f, o = 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float m = __ldcg(base + (size_t)s * rec + hd);
            const float w = (m == -INFINITY) ? 0.f : expf(m - M);
            Lsum = fmaf(__ldcg(base + (size_t)s * rec + hd + 1), w, Lsum);
            o = fmaf(__ldcg(base + (size_t)s * rec + j), w, o);
        }
        out[(size_t)(kvh * G) * hd + i] = o / Lsum;
    }
    if (tid == 0) counters[kvh] = 0;  // leave the workspace zeroed for the next launch
}

template <int KVD, int G>
__global__ void __launch_bounds__(kMhaThreads)
mha_decode_kernel(const float* __restrict__ q, const uint8_t* __restrict__ kc, const uint8_t* __restrict__ vc,
                  float* __restrict__ out, float* __restrict__ partials, int* __restrict__ counters, int layer,
                  const int32_t* __restrict__ pos_dev, int pos_val, int max_len, int hd, int kv_heads, int nsplit) {
    mha_decode_body<KVD, G, false>(q, kc, vc, out, partials, counters, layer, pos_dev, pos_val, max_len, hd, kv_heads, nsplit, PagedKv{});
}

template <int KVD, int G>
__global__ void __launch_bounds__(kMhaThreads)
mha_paged_kernel(const float* __restrict__ q, const uint8_t* __restrict__ kc, const uint8_t* __restrict__ vc,
                 float* __restrict__ out, float* __restrict__ partials, int* __restrict__ counters, int layer, int hd,
                 int kv_heads, int nsplit, PagedKv pk) {
    mha_decode_body<KVD, G, true>(q, kc, vc, out, partials, counters, layer, nullptr, 0, 0, hd, kv_heads, nsplit, pk);
}

template <int KVD, int G>
static int launch_mha(const float* q, const void* kc, const void* vc, float* out, void* ws, int layer, const int32_t* pos_dev,
                      int pos, int max_len, int hd, int kv_heads, cudaStream_t st, bool pdl) {
    const int nsplit = mha_nsplit(kv_heads, max_len);
    const MhaSmem L = mha_smem_layout(hd, G, KvInfo<KVD>::ESZ);
    SLLM_REQUIRE(L.total <= (size_t)smem_optin_bytes(), SLLM_ENOTSUP, "mha: tile does not fit shared memory (hd=%d)", hd);
    static size_t configured = 0;
    if (L.total > 48 * 1024 && L.total > configured) {
        SLLM_CUDA(cudaFuncSetAttribute(mha_decode_kernel<KVD, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        configured = L.total;
    }
    int* counters = reinterpret_cast<int*>(ws);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + 256);
    LaunchCfg lc(dim3(kv_heads, nsplit), dim3(kMhaThreads), L.total, st, pdl);
    SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, mha_decode_kernel<KVD, G>, q, reinterpret_cast<const uint8_t*>(kc),
                                 reinterpret_cast<const uint8_t*>(vc), out, partials, counters, layer, pos_dev, pos, max_len, hd,
                                 kv_heads, nsplit));
    g_launches++;
    return SLLM_OK;
}

int mha_decode_dispatch(const float* q, const void* kc, const void* vc, int kv_dtype, float* out, void* ws, int layer,
                        const int32_t* pos_dev, int pos, int max_len, int hd, int heads, int kv_heads, cudaStream_t st, bool pdl) {
    SLLM_REQUIRE(q && kc && vc && out && ws, SLLM_EINVAL, "mha: null pointer");
    SLLM_REQUIRE(heads > 0 && kv_heads > 0 && heads % kv_heads == 0, SLLM_EINVAL, "mha: heads=%d not a multiple of kv_heads=%d", heads, kv_heads);
    SLLM_REQUIRE(kv_heads <= 64, SLLM_ENOTSUP, "mha: kv_heads=%d > 64", kv_heads);
    SLLM_REQUIRE(hd >= 16 && hd <= 256 && hd % 16 == 0, SLLM_ENOTSUP, "mha: head_dim=%d must be a multiple of 16 in [16,256]", hd);
    SLLM_REQUIRE(kv_dtype == SLLM_F32 || kv_dtype == SLLM_BF16, SLLM_EINVAL, "mha: kv dtype %d", kv_dtype);
    SLLM_REQUIRE(pos_dev || (pos >= 0 && pos < max_len), SLLM_EINVAL, "mha: pos=%d outside [0,%d)", pos, max_len);
    const int g = heads / kv_heads;
#define SLLM_MHA_CASE(GG)                                                                                                   \
    case GG:                                                                                                                \
        return kv_dtype == SLLM_F32 ? launch_mha<SLLM_F32, GG>(q, kc, vc, out, ws, layer, pos_dev, pos, max_len, hd, kv_heads, st, pdl) \
                                    : launch_mha<SLLM_BF16, GG>(q, kc, vc, out, ws, layer, pos_dev, pos, max_len, hd, kv_heads, st, pdl);
    switch (g) {
        SLLM_MHA_CASE(1)
        SLLM_MHA_CASE(2)
        SLLM_MHA_CASE(3)
        SLLM_MHA_CASE(4)
        SLLM_MHA_CASE(5)
        SLLM_MHA_CASE(6)
        SLLM_MHA_CASE(7)
        SLLM_MHA_CASE(8)
        default: SLLM_REQUIRE(false, SLLM_ENOTSUP, "mha: heads/kv_heads=%d > 8 query heads per KV head", g);
    }
#undef SLLM_MHA_CASE
}

// ---- paged, batched launch (batch.cu) -------------------------------------------------------------------
// Fixed number of KV splits per slot: ~2 CTAs per SM over all slots, never more than one tile per split at max_ctx.
int mha_paged_nsplit(int kv_heads, int slots, int max_ctx) {
    const int ctas = kv_heads * (slots < 1 ? 1 : slots);
    int n = (2 * sm_count() + ctas - 1) / ctas;
    const int by_len = (max_ctx + kMhaTile - 1) / kMhaTile;
    if (n > by_len) n = by_len;
    if (n > kMhaMaxSplit) n = kMhaMaxSplit;
    return n < 1 ? 1 : n;
}

static size_t paged_counter_bytes(int slots, int kv_heads) { return ((size_t)slots * kv_heads * sizeof(int) + 255) / 256 * 256; }

size_t mha_paged_workspace_bytes(int slots, int heads, int kv_heads, int head_dim) {
    return paged_counter_bytes(slots, kv_heads) + (size_t)slots * heads * kMhaMaxSplit * (head_dim + 2) * sizeof(float);
}

template <int KVD, int G>
static int launch_mha_paged(const float* q, const void* kp, const void* vp, float* out, void* ws, int layer, const PagedKv& pk,
                            int slots, int max_slots, int nsplit, int hd, int kv_heads, cudaStream_t st) {
    const MhaSmem L = mha_smem_layout(hd, G, KvInfo<KVD>::ESZ, /*padded=*/false);
    SLLM_REQUIRE(L.total <= (size_t)smem_optin_bytes(), SLLM_ENOTSUP, "paged mha: tile does not fit shared memory (hd=%d)", hd);
    static size_t configured = 0;
    if (L.total > 48 * 1024 && L.total > configured) {
        SLLM_CUDA(cudaFuncSetAttribute(mha_paged_kernel<KVD, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        configured = L.total;
    }
    int* counters = reinterpret_cast<int*>(ws);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + paged_counter_bytes(max_slots, kv_heads));
    LaunchCfg lc(dim3(kv_heads, nsplit, slots), dim3(kMhaThreads), L.total, st, false);
    SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, mha_paged_kernel<KVD, G>, q, reinterpret_cast<const uint8_t*>(kp),
                                 reinterpret_cast<const uint8_t*>(vp), out, partials, counters, layer, hd, kv_heads, nsplit, pk));
    g_launches++;
    return SLLM_OK;
}

// q / out: [slots][pk.q_stride]; pools [pages][layers][kv_heads][page_len][hd] in kv_dtype; ws: mha_paged_workspace_bytes(max_slots, ..),
// zeroed once (the kernel leaves it zeroed); the first `slots` slots are served (unused ones carry pos < 0).
int mha_paged_dispatch(const float* q, const void* k_pool, const void* v_pool, int kv_dtype, float* out, void* ws, int layer,
                       const PagedKv& pk, int slots, int max_slots, int nsplit, int hd, int kv_heads, cudaStream_t st) {
    SLLM_REQUIRE(q && k_pool && v_pool && out && ws && pk.block_table && pk.pos, SLLM_EINVAL, "paged mha: null pointer");
    SLLM_REQUIRE(slots >= 1 && slots <= max_slots && slots <= 65535, SLLM_EINVAL, "paged mha: %d slots (max %d)", slots, max_slots);
    SLLM_REQUIRE(pk.heads > 0 && kv_heads > 0 && pk.heads % kv_heads == 0, SLLM_EINVAL, "paged mha: heads=%d not a multiple of kv_heads=%d", pk.heads, kv_heads);
    SLLM_REQUIRE(hd >= 16 && hd <= 256 && hd % 16 == 0, SLLM_ENOTSUP, "paged mha: head_dim=%d must be a multiple of 16 in [16,256]", hd);
    SLLM_REQUIRE(kv_dtype == SLLM_F32 || kv_dtype == SLLM_BF16, SLLM_EINVAL, "paged mha: kv dtype %d", kv_dtype);
    SLLM_REQUIRE(nsplit >= 1 && nsplit <= kMhaMaxSplit && pk.page_len >= 1 && pk.max_pages >= 1, SLLM_EINVAL, "paged mha: bad split/page geometry");
    const int g = pk.heads / kv_heads;
#define SLLM_MHA_CASE(GG)                                                                                                             \
    case GG:                                                                                                                          \
        return kv_dtype == SLLM_F32 ? launch_mha_paged<SLLM_F32, GG>(q, k_pool, v_pool, out, ws, layer, pk, slots, max_slots, nsplit, hd, kv_heads, st) \
                                    : launch_mha_paged<SLLM_BF16, GG>(q, k_pool, v_pool, out, ws, layer, pk, slots, max_slots, nsplit, hd, kv_heads, st);
    switch (g) {
        SLLM_MHA_CASE(1)
        SLLM_MHA_CASE(2)
        SLLM_MHA_CASE(3)
        SLLM_MHA_CASE(4)
        SLLM_MHA_CASE(5)
        SLLM_MHA_CASE(6)
        SLLM_MHA_CASE(7)
        SLLM_MHA_CASE(8)
        default: SLLM_REQUIRE(false, SLLM_ENOTSUP, "paged mha: heads/kv_heads=%d > 8 query heads per KV head", g);
    }
#undef SLLM_MHA_CASE
}

size_t mha_workspace_bytes(int heads, int head_dim, int max_len) {
    (void)max_len;
    return 256 + (size_t)heads * kMhaMaxSplit * (head_dim + 2) * sizeof(float);
}

}  // namespace sllm

extern "C" {

size_t sllm_mha_workspace_bytes(int32_t heads, int32_t head_dim, int32_t max_len) {
    return sllm::mha_workspace_bytes(heads, head_dim, max_len);
}

int sllm_mha_decode(const float* q, const void* key_cache, const void* value_cache, int32_t kv_dtype, float* out,
                    void* workspace, int32_t layer, const int32_t* pos_dev, int32_t pos, int32_t max_len, int32_t head_dim,
                    int32_t heads, int32_t kv_heads, sllm_stream_t stream) {
    return sllm::mha_decode_dispatch(q, key_cache, value_cache, kv_dtype, out, workspace, layer, pos_dev, pos, max_len,
                                     head_dim, heads, kv_heads, sllm::as_stream(stream), false);
}

}  // extern "C"
