// csrc/prefill_tp.cu — the tensor-parallel exchange of the batched prefill as ONE kernel over NVLink peer memory.
//
// After a row-parallel GEMM (wo, down: matmul.cpp:14-26 split along the input dimension) every rank holds a [T][d] fp32 matrix of
// PARTIAL sums. The reference semantics that follow are: sum over ranks, add to the residual stream (add_kernel.cpp:10-13), RMSNorm
// (rms_kernel.cpp:12-22), and the normalised rows feed the next column-parallel GEMM on every rank. Round 1 did that as
// ncclAllReduce(8 MB fp32) + an RMSNorm kernel: 64 un-overlapped collectives per prompt, which made prefill SLOWER on 8 GPUs than on 1.
// Here the rows are sharded instead (rank r owns rows [r*per, (r+1)*per)): for each of its rows a rank
//   * reads the N partial rows straight out of the peers' memory (coalesced 16-byte loads over NVLink), adds them in rank order
//     and then the residual row it owns (the fp32 residual stream is row-sharded from here on: nobody else needs it);
//   * RMSNorm in registers;
//   * stores the bf16 row into EVERY rank's GEMM operand buffer (16-byte stores over NVLink).
// Per rank and call: (N-1)/N * T*d*4 bytes in, (N-1)/N * T*d*2 bytes out — 3/8 of the all-reduce's traffic, no RMSNorm launch, and
// two flag rounds (entry: every rank's partial matrix is complete; exit: every rank's rows have landed here) instead of a collective.
// The flags are monotonic epochs in each rank's exchange block (release/acquire at system scope); every wait is bounded.
#include "mega_common.cuh"
#include "prefill.cuh"

namespace sllm {

extern int g_tune_pf_pdl;

namespace {

constexpr unsigned kPfxSpinLimit = 1u << 24;   // x ~1 us per poll (a system-scope load + 40 ns of sleep): ~20 s, then the kernel traps instead of hanging the GPU

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __noinline__ void pfx_timeout(int what, int peer) {
    printf("sllm prefill exchange: rank flag %d of peer %d never arrived (cta %d)\n", what, peer, (int)blockIdx.x);
    __trap();
}
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned epoch, int what, int peer) {
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(p) - epoch) < 0) {
        __nanosleep(40);
        if (++spins > kPfxSpinLimit) pfx_timeout(what, peer);
    }
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { return (uint32_t)f32_to_bf16_bits(lo) | ((uint32_t)f32_to_bf16_bits(hi) << 16); }

constexpr int kPfxThreads = 512;
// TP = number of ranks (compile time: the loads of a chunk from ALL ranks are issued before the first one is used — a remote load is a
// ~2.5 us round trip over NVLink, a serial chain of them would be the whole cost of the kernel)
template <int TP>
__global__ void __launch_bounds__(kPfxThreads, TP == 2 ? 2 : 1) pf_tp_exchange_kernel(const PfxParams p) {
    __shared__ float red[33];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    uint8_t* const mine = p.block[p.rank];
    unsigned long long* const tr = reinterpret_cast<unsigned long long*>(mine + kPfxTrace) + (p.write_xlast ? 8 : 0);   // timeline of the most recent full / final call (CTA 0 + the last CTA)
    const bool stamp = blockIdx.x == 0 && tid == 0;
    if (stamp) tr[0] = gtime();
    pdl_wait();   // programmatic dependent launch: resident early, but the GEMM before this kernel must have completed before anything below
    if (stamp) tr[1] = gtime();
    // ---- entry: my partial matrix is complete (the GEMM before this kernel on my stream) -> tell everyone; wait for everyone's
    if (blockIdx.x == 0 && tid < TP) st_release_sys(reinterpret_cast<unsigned*>(p.block[tid] + kPfxFlagIn) + 16 * p.rank, p.epoch);
    // while the flags travel: zero this rank's OTHER partial-sum matrix (its readers finished before the previous call's exit flags; the next
    // row-parallel GEMM accumulates K-split partial sums into it with red.add)
    if (p.zero) {
        const size_t n4 = (size_t)p.T * p.d >> 2;
        for (size_t i = (size_t)blockIdx.x * kPfxThreads + tid; i < n4; i += (size_t)gridDim.x * kPfxThreads)
            reinterpret_cast<float4*>(p.zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < TP) wait_flag(reinterpret_cast<const unsigned*>(mine + kPfxFlagIn) + 16 * tid, p.epoch, 0, tid);
    __syncthreads();
    // the next kernel may become resident now (its own griddepcontrol.wait holds it until this one has completed). Only here: the trigger takes
    // effect once EVERY CTA of this grid has issued it, i.e. when all of them are running — a dependent grid that took SMs away from CTAs of
    // this one that were still waiting to be scheduled would deadlock on the exit flags.
    pdl_launch_dependents();
    if (stamp) tr[2] = gtime();

    const int per = (p.T + TP - 1) / TP;
    const int lo = max(p.row_lo, p.rank * per), hi = min(min(p.row_hi, p.T), (p.rank + 1) * per);
    const int nch = p.d >> 3;                       // 8 columns (two float4 in, one 16-byte bf16 chunk out) per thread and pass
    constexpr int kKeep = 2;                        // chunks per thread kept in registers (rows up to 8192 floats)
    for (int row = lo + (int)blockIdx.x; row < hi; row += (int)gridDim.x) {
        float v[kKeep][8];
        float ss = 0.f;
        float* xr = p.x + (size_t)row * p.d;
#pragma unroll
        for (int k = 0; k < kKeep; ++k) {
            const int c = tid + k * kPfxThreads;
            if (c < nch) {
                float4 b0[TP], b1[TP];
#pragma unroll
                for (int q = 0; q < TP; ++q) {
                    const float4* src = reinterpret_cast<const float4*>(p.block[q] + p.off_part) + ((size_t)row * p.d >> 2) + 2 * c;
                    b0[q] = __ldcg(src);
                    b1[q] = __ldcg(src + 1);
                }
                float4 x0 = reinterpret_cast<float4*>(xr)[2 * c], x1 = reinterpret_cast<float4*>(xr)[2 * c + 1];
                float4 a0 = b0[0], a1 = b1[0];
#pragma unroll
                for (int q = 1; q < TP; ++q) {      // rank order: every run adds the same way (the owner alone computes a row)
                    a0 = make_float4(a0.x + b0[q].x, a0.y + b0[q].y, a0.z + b0[q].z, a0.w + b0[q].w);
                    a1 = make_float4(a1.x + b1[q].x, a1.y + b1[q].y, a1.z + b1[q].z, a1.w + b1[q].w);
                }
                x0 = make_float4(x0.x + a0.x, x0.y + a0.y, x0.z + a0.z, x0.w + a0.w);       // add_kernel.cpp:10-13
                x1 = make_float4(x1.x + a1.x, x1.y + a1.y, x1.z + a1.z, x1.w + a1.w);
                reinterpret_cast<float4*>(xr)[2 * c] = x0;
                reinterpret_cast<float4*>(xr)[2 * c + 1] = x1;
                v[k][0] = x0.x; v[k][1] = x0.y; v[k][2] = x0.z; v[k][3] = x0.w; v[k][4] = x1.x; v[k][5] = x1.y; v[k][6] = x1.z; v[k][7] = x1.w;
#pragma unroll
                for (int e = 0; e < 8; ++e) ss = fmaf(v[k][e], v[k][e], ss);
            }
        }
        ss = block_sum(ss, red);
        const float inv = 1.0f / sqrtf(ss / (float)p.d + p.eps);   // rms_kernel.cpp:17-19
#pragma unroll
        for (int k = 0; k < kKeep; ++k) {
            const int c = tid + k * kPfxThreads;
            if (c < nch) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.norm_w) + 2 * c), g1 = __ldg(reinterpret_cast<const float4*>(p.norm_w) + 2 * c + 1);
                uint4 o;
                o.x = pack_bf16x2((v[k][0] * inv) * g0.x, (v[k][1] * inv) * g0.y);
                o.y = pack_bf16x2((v[k][2] * inv) * g0.z, (v[k][3] * inv) * g0.w);
                o.z = pack_bf16x2((v[k][4] * inv) * g1.x, (v[k][5] * inv) * g1.y);
                o.w = pack_bf16x2((v[k][6] * inv) * g1.z, (v[k][7] * inv) * g1.w);
#pragma unroll
                for (int q = 0; q < TP; ++q) {
                    const int dst = (p.rank + q) % TP;              // start with myself, then round the ring: the ranks' stores spread over the links
                    reinterpret_cast<uint4*>(p.block[dst] + p.off_xn)[((size_t)row * p.d >> 3) + c] = o;
                    if (p.write_xlast) {
                        float4* xl = reinterpret_cast<float4*>(p.block[dst] + kPfxXlast) + 2 * c;
                        xl[0] = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
                        xl[1] = make_float4(v[k][4], v[k][5], v[k][6], v[k][7]);
                    }
                }
            }
        }
    }
    // ---- exit: my rows are stored everywhere -> tell everyone; the kernel ends when everyone's rows are here
    // Ordering: every thread's stores -> CTA barrier -> thread 0: gpu-scope release (fence + counter) -> the last CTA acquires the counter at gpu
    // scope -> its st.release.sys of the flags is cumulative over everything that happened before it. Only that ONE release is system scope:
    // a system-scope fence costs ~7 us here whether or not the CTA stored anything (measured), a gpu-scope one a fraction of that.
    if (stamp) tr[3] = gtime();
    __syncthreads();
    if (tid == 0) {
        unsigned* cnt = reinterpret_cast<unsigned*>(mine + kPfxCounter);
        __threadfence();
        unsigned old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(cnt) : "memory");
        s_last = (old + 1u == p.done_target);
    }
    if (stamp) tr[4] = gtime();
    __syncthreads();
    if (!s_last) return;
    if (tid < TP) {
        st_release_sys(reinterpret_cast<unsigned*>(p.block[tid] + kPfxFlagOut) + 16 * p.rank, p.epoch);
        wait_flag(reinterpret_cast<const unsigned*>(mine + kPfxFlagOut) + 16 * tid, p.epoch, 1, tid);
    }
    __syncwarp();
    if (tid == 0) { tr[5] = gtime(); tr[6] = gtime() - tr[0]; }
}

// the (value, index) pairs of the ranks' vocabulary shards: exchange through the blocks, first maximum wins (argmax.cpp:12-15), then the
// step state advances exactly as after a decode step (the same merge as tp_merge_kernel, without a collective)
__global__ void pf_tp_argmax_kernel(const PfxParams p, const float* logits, const int32_t* idx, int v0, StepState* st, const int32_t* prompt,
                                    int32_t* history) {
    const int tid = threadIdx.x;
    if (tid < p.tp) {
        uint8_t* dst = p.block[tid] + kPfxPairs + 16 * p.rank;
        const int i = *idx;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(__float_as_uint(logits[i])), "r"((unsigned)(v0 + i)), "r"(p.epoch), "r"(p.epoch) : "memory");
    }
    __shared__ float sv[kMaxTp];
    __shared__ int si[kMaxTp];
    if (tid < p.tp) {
        const uint8_t* src = p.block[p.rank] + kPfxPairs + 16 * tid;
        uint4 w;
        unsigned spins = 0;
        while (true) {
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(src) : "memory");
            if (w.z == p.epoch && w.w == p.epoch) break;
            __nanosleep(40);
            if (++spins > kPfxSpinLimit) pfx_timeout(2, tid);
        }
        sv[tid] = __uint_as_float(w.x);
        si[tid] = (int)w.y;
    }
    __syncthreads();
    if (tid == 0) {
        float bv = sv[0];
        int bi = si[0];
        for (int r = 1; r < p.tp; ++r)
            if (sv[r] > bv || (sv[r] == bv && si[r] < bi)) { bv = sv[r]; bi = si[r]; }
        ClsPolicy<SLLM_F32>::step_feedback(st, prompt, history, bi);
    }
}

}  // namespace

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// block = header | partial sums A [rows][d] fp32 | partial sums B | bf16 operand rows [rows][d]
size_t pfx_block_bytes(int rows, int d) { return kPfxData + 2 * align_up((size_t)rows * d * 4, 1024) + align_up((size_t)rows * d * 2, 1024); }
size_t pfx_off_part(int which, int rows, int d) { return kPfxData + (size_t)which * align_up((size_t)rows * d * 4, 1024); }
size_t pfx_off_xn(int rows, int d) { return kPfxData + 2 * align_up((size_t)rows * d * 4, 1024); }

int pf_tp_exchange(const PfxParams& p, int sms, cudaStream_t st) {
    SLLM_REQUIRE(p.d % 8 == 0 && p.d <= 8 * kPfxThreads * 2, SLLM_ENOTSUP, "prefill exchange: hidden size must be a multiple of 8, at most 8192");
    SLLM_REQUIRE((size_t)p.d * 4 <= kPfxData - kPfxXlast, SLLM_ENOTSUP, "prefill exchange: residual row does not fit its slot");
    LaunchCfg lc(dim3(pf_tp_exchange_grid(p.T, p.tp, sms)), dim3(kPfxThreads), 0, st, g_tune_pf_pdl != 0);
    switch (p.tp) {
        case 2: SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, pf_tp_exchange_kernel<2>, p)); break;
        case 4: SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, pf_tp_exchange_kernel<4>, p)); break;
        case 8: SLLM_CUDA(cudaLaunchKernelEx(&lc.cfg, pf_tp_exchange_kernel<8>, p)); break;
        default: set_error("prefill exchange: %d ranks (2, 4 or 8)", p.tp); return SLLM_ENOTSUP;
    }
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

int pf_tp_argmax(const PfxParams& p, const float* logits, const int32_t* idx, int v0, StepState* state, const int32_t* prompt, int32_t* history,
                 cudaStream_t st) {
    pf_tp_argmax_kernel<<<1, 32, 0, st>>>(p, logits, idx, v0, state, prompt, history);
    g_launches++;
    SLLM_LAUNCH_CHECK();
    return SLLM_OK;
}

}  // namespace sllm
