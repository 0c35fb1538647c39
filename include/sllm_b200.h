/* include/sllm_b200.h — the drop-in boundary: C ABI of libsllm_b200.so.
 *
 * B200-native (sm_100a) replacement for the transformer forward hot path of Boundwhd/SimpleLLMInference.
 * Plain pointers and sizes only: no C++ types, no torch types, no CUDA headers needed to include this file.
 * The reference has no FFI of its own (it is one C++ program); its hot-path boundary is the set of free
 * functions kernel::<op>_kernel_cuda (include/kernel/cuda/<op>_kernel.cuh) that the op layers call (source/op/<op>.cpp) plus
 * model::LlamaModel::forward (source/model/model.cpp:40-140). Each entry point below names the reference
 * interface it replaces. The C++ host mirror in simplellminference_b200/host/ (mem::Tensor, op::*Layer,
 * kernel::*_cuda, model::LlamaModel) is a thin layer over exactly these functions; INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer named *_dev / every tensor argument is DEVICE memory of the current CUDA device unless the
 *    name ends in _host; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - launchers never allocate, never synchronise, never throw; they enqueue on `stream` and return;
 *  - return value: 0 = SLLM_OK; >0 = a cudaError_t; <0 = an SLLM_E* code; sllm_last_error() gives the text
 *    (thread-local). The C++ shims turn non-zero into the reference's convention: LOG(...) prints
 *    "file: .. line: .. - msg" and std::exit(EXIT_FAILURE) (include/base/base.h:6-10);
 *  - `token` and `pos` may be read from device memory (token_dev / pos_dev non-NULL) so that the launch
 *    sequence is CUDA-graph capturable — a deliberate deviation from the reference, which reads them on the
 *    host (emb_kernel.cu:15, rope_kernel.cu:49). With the *_dev pointer NULL the by-value argument is used;
 *  - activations, RoPE tables and accumulators are fp32. Matrix weights are stored as fp32, bf16 or
 *    int8-with-group-scales (SLLM_F32 / SLLM_BF16 / SLLM_INT8); the KV cache as fp32 or bf16.
 *  - there is NO CPU fallback anywhere behind this header;
 *  - threading: the model of the reference (source/model/model.cpp drives one model from one host thread) and of the
 *    tensor-parallel runtime here — ONE process per GPU, one host thread driving an engine / batch at a time. The library keeps
 *    per-process caches of device facts and of per-kernel shared-memory attributes for the device that was current at first use, and
 *    the development knobs of sllm_tune are process-wide: do not drive two devices, or one engine from two threads, in one process.
 *    Pure-arithmetic entry points (sllm_*_plan, sllm_mega_tile_geometry, sllm_batch_arena_bytes, sllm_kvpages_*) touch no device.
 */
#ifndef SLLM_B200_H
#define SLLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLLM_OK 0
#define SLLM_EINVAL (-1)   /* bad argument (null pointer, size, alignment, divisibility) */
#define SLLM_ENOTSUP (-2)  /* shape or type combination not supported by the kernels */
#define SLLM_ENOMEM (-3)   /* device allocation failed */
#define SLLM_ESTATE (-4)   /* call sequence error (e.g. decode before weights are loaded) */
#define SLLM_ECOMM (-5)    /* NCCL / peer-memory error */

enum { SLLM_F32 = 0, SLLM_BF16 = 1, SLLM_INT8 = 2 };

typedef void* sllm_stream_t;

const char* sllm_last_error(void);
int sllm_abi_version(void);
/* development knobs (not part of the drop-in surface): key 0 = GEMV CTAs per SM (0 = built-in choice); prefill GEMM: key 1 = force
 * the N extent of a tile (multiple of 32, 0 = cost model), key 2 = two-SM tiles (1 / 0 force / forbid, -1 = by T), key 3 = programmatic
 * dependent launch (1 default), key 4 = K split of the residual-epilogue GEMMs (0 never, n up to n ranges, -1 cost model); batched decode:
 * key 5 = replay one CUDA graph per live-slot count instead of the launch sequence, key 6 = GEMV body with four weight rows per warp at
 * a time when 3 or more sequences share a launch, key 7 (with key 6) = down projection with K cut in two over grid.y when more sequences
 * are live than whole rows fit shared memory, so that Wdown is read once (all 0 by default: experimental until measured); key 8 = decode
 * megakernel MEASUREMENT aid, results are garbage: bit 0 = skip the grid barriers, bit 1 = skip the dot products (tools/mega_debug.py),
 * bit 4 = tensor-parallel prefill without its peer-memory exchanges (the GEMM / attention time alone) */
int sllm_tune(int32_t key, int32_t value);
/* device facts the host side sizes things by: sm count, max opt-in shared memory per block, total/free HBM */
int sllm_device_info(int32_t* sm_count, int32_t* smem_optin_bytes, size_t* hbm_total, size_t* hbm_free);

/* ------------------------------------------------------------------------------------------------------
 * Op launchers — one per reference CUDA kernel entry point.
 * ---------------------------------------------------------------------------------------------------- */

/* replaces kernel::add_kernel_cuda (include/kernel/cuda/add_kernel.cuh:6): out[i] = a[i] + b[i] */
int sllm_add_f32(const float* a, const float* b, float* out, int32_t n, sllm_stream_t stream);

/* replaces kernel::emb_kernel_cuda (emb_kernel.cuh:6-7): out[0..d) = table[token][:], dequantised to fp32.
 * scales: [vocab][d/group] fp32, only for SLLM_INT8. Fails with SLLM_EINVAL if the by-value token is outside
 * [0, vocab) (the reference's guard is `token > vocab`, off by one: emb_kernel.cu:16); a device-side token
 * outside the range is clamped into it. */
int sllm_embedding(const int32_t* token_dev, int32_t token, const void* table, int32_t w_dtype,
                   const float* scales, int32_t group, float* out, int32_t vocab, int32_t d,
                   sllm_stream_t stream);

/* replaces kernel::rmsnorm_kernel_cuda (rms_kernel.cuh:6-7): y = (x * 1/sqrt(mean(x^2)+eps)) * w */
int sllm_rmsnorm_f32(const float* x, const float* w, float* y, int32_t d, float eps, sllm_stream_t stream);

/* replaces kernel::matmul_kernel_cuda (matmul_kernel.cuh:6-7): y[r] = (sum_j x[j] * W[r][j]) * scale,
 * W row-major [rows][cols]. Unlike the reference's CUDA kernel, `scale` is honoured (matmul_kernel.cu:41-55
 * drops it; the CPU kernel applies it, matmul_kernel.cpp:26). cols must be a multiple of 16 bytes' worth of
 * elements (4 / 8 / 16 for f32 / bf16 / int8), W 16-byte aligned, and for SLLM_INT8 cols % group == 0,
 * group % 16 == 0, scales = [rows][cols/group] fp32. */
int sllm_gemv(const float* x, const void* W, int32_t w_dtype, const float* scales, int32_t group, float* y,
              int32_t rows, int32_t cols, float scale, sllm_stream_t stream);

/* replaces kernel::rope_cache_cal_cuda (rope_kernel.cuh:5): sin/cos tables [max_len][head_dim/2]. Computed
 * on the HOST with the same libm calls the reference CPU path makes (rope_kernel.cpp:8-17) and uploaded, so
 * the tables are bit-identical to the oracle's (GPU powf/sinf/cosf differ by ulps). Synchronous. */
int sllm_rope_tables(int32_t head_dim, int32_t max_len, float theta, float* sin_dev, float* cos_dev,
                     sllm_stream_t stream);

/* replaces kernel::rope_kernel_cuda (rope_kernel.cuh:7-8): rotate-half RoPE in place on q[0..q_dim) and
 * k[0..k_dim). The reference rotates k over q_dim too (GQA over-run); here k_dim is
 * explicit. k may be fp32 only (it is a row of an fp32 cache or a scratch vector). */
int sllm_rope_f32(float* q, float* k, const int32_t* pos_dev, int32_t pos, const float* sin_tab,
                  const float* cos_tab, int32_t q_dim, int32_t k_dim, int32_t head_dim, sllm_stream_t stream);

/* replaces kernel::mha_kernel_cuda (mha_kernel.cuh:6-21): single-query attention over cache rows 0..pos of
 * `layer`, caches laid out [layers][max_len][kv_heads*head_dim] (fp32 or bf16). Split-KV flash decoding with
 * an in-kernel combine; the reference's `score` scratch is not needed. workspace: device scratch of at least
 * sllm_mha_workspace_bytes(heads, head_dim, max_len) bytes, 16-byte aligned, zero-initialised once (the
 * kernel leaves it zeroed). head_dim must be a multiple of 8 and <= 256. */
size_t sllm_mha_workspace_bytes(int32_t heads, int32_t head_dim, int32_t max_len);
int sllm_mha_decode(const float* q, const void* key_cache, const void* value_cache, int32_t kv_dtype,
                    float* out, void* workspace, int32_t layer, const int32_t* pos_dev, int32_t pos,
                    int32_t max_len, int32_t head_dim, int32_t heads, int32_t kv_heads, sllm_stream_t stream);

/* replaces kernel::swiglu_kernel_cuda (swiglu_kernel.cuh:5): out = sigmoid(gate) * up  (sic: not SiLU,
 * swiglu_kernel.cpp:12-13) */
int sllm_swiglu_f32(const float* up, const float* gate, float* out, int32_t n, sllm_stream_t stream);

/* device version of op::argmaxLayer::forward (source/op/argmax.cpp:7-17; CPU-only in the reference): index of
 * the FIRST maximum of logits[0..n). idx_dev receives the int32 index. */
int sllm_argmax_f32(const float* logits, int32_t n, int32_t* idx_dev, sllm_stream_t stream);

/* KV-row store used by the op-by-op path when the cache is bf16: cache_row[i] = bf16(src[i]) */
int sllm_store_kv_row(const float* src, void* cache_row, int32_t kv_dtype, int32_t n, sllm_stream_t stream);

/* Sampling beyond arg-max (additive: the reference only has argmaxLayer; LayerType::kLayerSoftmax is declared but unused,
 * include/op/layer.h:17). One draw from softmax(logits / temperature) restricted to the top_k largest logits (0 = all; ties at
 * the k-th value are kept) and then to the smallest set of largest logits whose mass reaches top_p (0 or 1 = all). The draw is a
 * pure function of (logits, parameters, seed, step): u = hash(seed, step), first index whose running kept mass exceeds u * mass.
 * temperature <= 0 = arg-max (first maximum). idx_dev: device int32. */
int sllm_sample_f32(const float* logits, int32_t n, float temperature, int32_t top_k, float top_p, uint64_t seed,
                    uint64_t step, int32_t* idx_dev, sllm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Synthetic weights (device side of oracle/synth_weights.h: same integer hash, bit-identical values).
 * Fills `count` elements [first, first+count) of blob segment `segment` (0 E, 1 norms, 2 wq, 3 wk, 4 wv,
 * 5 wo, 6 up, 7 gate, 8 down) of the given shape into dst in w_dtype; int8 also writes count/group scales.
 * row_len/row_stride_* let a tensor-parallel shard be cut out of the full matrix: the destination is
 * [n_rows][row_len] taken from source rows of length src_row_len starting at column src_col0.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t vocab, head_dim, hidden, kv_hidden, inter, max_len, layers, heads, kv_heads;
    float eps, theta;
} sllm_shape;

int sllm_synth_fill(const sllm_shape* shape, uint64_t seed, int32_t segment, int64_t src_first_row,
                    int64_t n_rows, int64_t src_row_len, int64_t src_col0, int64_t row_len, void* dst,
                    int32_t w_dtype, float* scales, int32_t group, sllm_stream_t stream);

/* fp32 -> storage type conversion of an uploaded fp32 matrix [rows][cols] (the path real checkpoints take) */
int sllm_convert_weights(const float* src_f32_dev, void* dst, int32_t w_dtype, float* scales, int32_t group,
                         int64_t rows, int64_t cols, sllm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Engine — replaces model::LlamaModel::{init_mem, create_param_layers, forward} and the greedy loop of
 * predict (source/model/model.cpp:40-187, 247-469) with a static device arena, a KV cache and a
 * CUDA-graph-captured, fused decode step. One engine per GPU; tensor parallelism = one process per GPU.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    sllm_shape shape;
    int32_t w_dtype;   /* SLLM_F32 / SLLM_BF16 / SLLM_INT8 */
    int32_t kv_dtype;  /* SLLM_F32 / SLLM_BF16 */
    int32_t group;     /* int8 group size along the input dimension */
    int32_t tp_rank, tp_size;
    int32_t flags;     /* SLLM_ENGINE_* */
} sllm_engine_config;

#define SLLM_ENGINE_UNFUSED 1u   /* run the 13-ops-per-layer sequence of the reference instead of fused kernels */
#define SLLM_ENGINE_NO_GRAPH 2u  /* launch kernels directly instead of replaying a CUDA graph */
#define SLLM_ENGINE_PDL 4u       /* programmatic dependent launch between the kernels of a step (measured slower
                                    inside a CUDA graph on B200 than plain graph edges: off by default) */
#define SLLM_ENGINE_MEGAKERNEL 16u /* whole decode step as ONE persistent cooperative kernel whose TMA weight rings
                                    keep streaming across phases (single GPU; shapes it cannot take fall back to the
                                    per-kernel fused path — sllm_engine_mode() tells which one runs) */
#define SLLM_ENGINE_MEGA_LL 32u   /* with MEGAKERNEL on one GPU: the barrier-free {value, epoch}-word version
                                    (megakernel_ll.cu) instead of the grid-barrier one (megakernel.cu, faster on one GPU).
                                    Under tensor parallelism (MEGAKERNEL | P2P_ALLREDUCE) the word version is the only
                                    one: it carries the all-reduce inside the kernel */
#define SLLM_ENGINE_MEGA_FUSE_DOWN 64u /* experimental, with MEGAKERNEL on one GPU (fp32 / bf16 weights, hidden size = 2^k x 512 bytes per
                                    row <= 8 KB): the down projection runs inside the gate_up phase as a K-split over the values each
                                    CTA produced, partial outputs added to the residual stream with red.global.add — four grid barriers
                                    per layer instead of five, but a summation order that is not fixed (logits vary in the last bits
                                    from run to run). Costs a second, transposed copy of the down matrices. Off until measured;
                                    sllm_engine_mode() says "megakernel(fused-down)" when it is in effect */
#define SLLM_ENGINE_MEGA_V2 128u /* with MEGAKERNEL on one GPU (fp32 / bf16 weights): the megakernel with two grid-wide dependency points
                                   * per layer instead of five (csrc/megakernel2.cu): qkv -> attention -> wo chained through per-kv-head
                                   * counters, wo as a K split by kv head group adding into h with red.global.add.f32. With
                                   * MEGA_FUSE_DOWN the down projection is fused into the gate_up phase as well. Summation order of wo
                                   * (and of a fused down) is not fixed: logits differ in the last bits from run to run (inside the
                                   * decode tolerance). Shapes it cannot take fall back to the grid-barrier kernel (sllm_engine_mode). */
#define SLLM_ENGINE_P2P_ALLREDUCE 8u /* TP: one-shot all-reduce over NVLink peer memory instead of NCCL */

typedef struct sllm_engine sllm_engine;

/* stream: the stream every engine operation is enqueued on; NULL = the engine creates and owns a
 * non-blocking stream (the legacy default stream cannot be captured into a CUDA graph). */
int sllm_engine_create(const sllm_engine_config* cfg, sllm_stream_t stream, sllm_engine** out);
void sllm_engine_destroy(sllm_engine* e);

/* weights: generated on the device (synthetic), or taken from a HOST fp32 blob in the reference's order
 * (source/model/model.cpp:340-468), converted to w_dtype and cut to this rank's tensor-parallel shard. */
int sllm_engine_load_synthetic(sllm_engine* e, uint64_t seed);
int sllm_engine_load_blob_f32(sllm_engine* e, const float* blob_host, int64_t n_floats);

/* tensor-parallel communicator: id_bytes = 128-byte ncclUniqueId produced by sllm_comm_unique_id on rank 0
 * and distributed by the caller (torch.distributed / MPI / a file — plumbing is the host's business). */
int sllm_comm_unique_id(void* id_bytes_128);
int sllm_engine_init_comm(sllm_engine* e, const void* id_bytes_128);
/* peer-memory exchange for SLLM_ENGINE_P2P_ALLREDUCE: each rank exports a 64-byte IPC handle, the caller
 * all-gathers them, every rank imports all of them. */
int sllm_engine_p2p_export(sllm_engine* e, void* handle_bytes_64);
int sllm_engine_p2p_import(sllm_engine* e, const void* all_handles /* tp_size * 64 bytes */);
/* Batched prefill under tensor parallelism without a collective library: every rank exports the CUDA-IPC handle of its prefill exchange
   block and imports everybody's (same messenger protocol as above). From then on sllm_engine_prefill sums the row-parallel partial
   matrices, adds the residual, normalises and distributes the rows in one kernel over NVLink peer memory (csrc/prefill_tp.cu).
   SLLM_ENOTSUP: the engine has no such block (one rank, no SLLM_ENGINE_P2P_ALLREDUCE, shape not taken by the batched prefill) — prefill
   then needs the NCCL communicator. Every rank must issue the same sequence of sllm_engine_prefill calls (the exchange kernels wait for each
   other's epoch flags; a wait that is never answered traps after ~20 s instead of hanging the GPU): a prefill that failed on one rank only leaves
   the ranks out of step — destroy the engines. Replaces nothing in the reference (it has neither prefill nor a working multi-GPU build). */
int sllm_engine_prefill_p2p_export(sllm_engine* e, void* handle_bytes_64);
int sllm_engine_prefill_p2p_import(sllm_engine* e, const void* all_handles /* tp_size * 64 bytes */);

/* One token, one position: the semantics of LlamaModel::forward(). token/pos by value (host), synchronous.
 * logits_host (vocab floats, may be NULL) receives model_pred; next_token_host (may be NULL) the greedy
 * argmax (first maximum). */
int sllm_engine_forward(sllm_engine* e, int32_t token, int32_t pos, float* logits_host,
                        int32_t* next_token_host);

/* Greedy loop of LlamaModel::predict without tokenizer/printing: feeds prompt[0..n_prompt) one token at a
 * time, then feeds back the argmax; writes the n_total-1 tokens that follow prompt[0] to tokens_out_host.
 * Runs entirely on the device (token and position live in device memory, one graph replay per token);
 * asynchronous until the final copy of the token list. */
int sllm_engine_greedy(sllm_engine* e, const int32_t* prompt_host, int32_t n_prompt, int32_t n_total,
                       int32_t* tokens_out_host);

/* Device-resident stepping for benchmarks: set state, enqueue n steps (no host sync), read results. */
int sllm_engine_set_state(sllm_engine* e, int32_t token, int32_t pos);
int sllm_engine_enqueue_steps(sllm_engine* e, int32_t n_steps);
int sllm_engine_read_tokens(sllm_engine* e, int32_t* tokens_out_host, int32_t n);

/* Batched prefill of prompt[0..n) at positions start_pos..: the layer loop of LlamaModel::forward
 * (source/model/model.cpp:50-128) for all prompt rows at once — dense contractions as tcgen05/TMEM tensor-core
 * GEMMs fed by TMA (bf16 operands, fp32 accumulators), causal attention over the block — instead of the
 * reference's one forward() per prompt token (model.cpp:157-166). Fills the KV cache for start_pos..start_pos+n-1,
 * leaves the LAST prompt token's logits in model_pred, its arg-max as the current token and the position at
 * start_pos+n, exactly as the token-by-token loop would (the last row's classifier is a T = 1 GEMM + arg-max).
 * Needs SLLM_ENGINE_MEGAKERNEL, bf16 weights and head_dim 64/128: otherwise SLLM_ENOTSUP (feed the prompt through
 * sllm_engine_greedy / sllm_engine_forward, which is what the reference does); sllm_engine_prefill_supported
 * returns 1/0 (0: reason in sllm_last_error). Tensor parallel: also needs sllm_engine_init_comm (NCCL all-reduce of
 * the row-parallel partial sums, one per wo/down GEMM). */
int sllm_engine_prefill(sllm_engine* e, const int32_t* prompt_host, int32_t n, int32_t start_pos);
int sllm_engine_prefill_supported(const sllm_engine* e);

/* How the persistent decode megakernel (SLLM_ENGINE_MEGAKERNEL; replaces the layer loop of LlamaModel::forward,
 * source/model/model.cpp:48-139, by one launch per token) would run a shape — host arithmetic only, no launch, and no device
 * either when the two device facts are passed in (sm_count_or_0 / smem_optin_or_0 > 0; 0 = ask the current device):
 * *ok = 1 and grid (= CTAs = SMs), dynamic shared memory per CTA, KV splits per head of the attention phase; or *ok = 0 with the
 * reason in sllm_last_error (the engine then runs the per-kernel fused path: sllm_engine_mode()). word_based = 0: the grid-barrier kernel
 * (one GPU); 1: the barrier-free {value, epoch}-word kernel (SLLM_ENGINE_MEGA_LL; the one that runs under tensor parallelism,
 * tp_size > 1 plans one rank's shard; its grid may be smaller than the SM count: every CTA must own a tile row of every phase). */
int sllm_mega_plan(const sllm_shape* shape, int32_t w_dtype, int32_t group, int32_t kv_dtype, int32_t tp_size, int32_t word_based,
                   int32_t sm_count_or_0, int32_t smem_optin_or_0, int32_t* ok, int32_t* grid, int64_t* smem_bytes, int32_t* nsplit);
/* The tiled weight layout of that kernel for one [rows][cols] matrix (kind: 0 qkv, 1 wo, 2 [up; gate], 3 down, 4 classifier):
 * tiles of r physical rows x sc 16-byte chunks, ks K slices per row, tile_rows tile rows, one tile = tile_bytes contiguous bytes
 * (= one TMA bulk copy = one ring slot), the matrix = tile_rows * ks * tile_bytes = matrix_bytes (zero padding included). */
int sllm_mega_tile_geometry(int32_t rows, int32_t cols, int32_t kind, int32_t w_dtype, int32_t* ks, int32_t* sc, int32_t* r,
                            int32_t* tile_rows, int32_t* tile_bytes, int64_t* matrix_bytes);

/* Development aid for the experimental SLLM_ENGINE_MEGA_FUSE_DOWN kernel: its transposed, stripe-major copy of a row-major
 * [d][inter] down matrix (device pointers; fp32 / bf16; d * element size = 2^k * 512 bytes <= 8 KB, inter % 4 == 0). Element e of lane
 * `lane`, input jj of tile row g, output stripe ks — at index (((ks * (inter/4) + g) * 4 + jj) * 32 + lane) * E + e, E = 16 bytes of
 * elements — is W[(ks * 32 + lane) * E + e][g * 4 + jj]. */
int sllm_mega_repack_down_t(const void* src_rowmajor, void* dst, int32_t d, int32_t inter, int32_t w_dtype, sllm_stream_t stream);

/* The two prefill kernels on their own (device pointers), for op-level parity tests:
 * C[T][N] (fp32) = A[T][K] (bf16) . W[N][K]^T (bf16, row-major) on tcgen05; bn = 0 (auto), 128 or 256 = N tile. */
int sllm_prefill_gemm_bf16(const void* A, const void* W, float* C, int32_t T, int32_t N, int32_t K, int32_t bn,
                           sllm_stream_t stream);
/* How sllm_prefill_gemm_bf16 / the engine would run a [T][K] x [N][K]^T GEMM (host arithmetic only, no launch): two_sm = 1 when SM
 * pairs share 256-row tiles (tcgen05 cta_group::2), bn = N extent of a tile, ksplit = K ranges per tile (> 1 only with the
 * residual epilogue, whose partial sums are added with red.global.add), work_units = tiles x ksplit. */
int sllm_prefill_gemm_plan(int32_t T, int32_t N, int32_t K, int32_t residual_epilogue, int32_t* two_sm, int32_t* bn,
                           int32_t* ksplit, int32_t* work_units);
/* causal attention of T queries (bf16 [T][heads*head_dim]) at positions pos0.. over a HEAD-MAJOR cache
 * [kv_heads][max_len][head_dim] (kv_dtype f32 or bf16), rows 0..pos0+T-1 valid; out bf16 like q. */
int sllm_prefill_attention(const void* q, const void* key_cache, const void* value_cache, int32_t kv_dtype, void* out,
                           int32_t T, int32_t pos0, int32_t max_len, int32_t head_dim, int32_t heads, int32_t kv_heads,
                           sllm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Batched multi-sequence decode over a paged KV cache (additive: the reference decodes one sequence at a time,
 * include/model/model.h:15-18 input_token{1} / position{1}, source/model/model.cpp:148-185). Up to 64 sequences,
 * each at its own position, advance one token per step and share ONE pass over the weights (the HBM traffic of a
 * decode step); every sequence follows the semantics of sllm_engine_greedy (prompt tokens fed verbatim, then the
 * first-max arg-max fed back, no EOS stop). Activations stay fp32: the decode parity contract holds per sequence.
 * ---------------------------------------------------------------------------------------------------- */

/* Page bookkeeping of the paged cache — host arithmetic only (no device, no CUDA call): a stack of free pages and, per
 * sequence, the list of pages that hold its positions [i*page_len, (i+1)*page_len). sllm_batch uses one internally; it
 * is exported so that a scheduler can plan admissions with the same rules. */
typedef struct sllm_kvpages sllm_kvpages;
sllm_kvpages* sllm_kvpages_create(int32_t n_pages, int32_t page_len, int32_t max_seqs, int32_t max_pages_per_seq);
void sllm_kvpages_destroy(sllm_kvpages* kp);
/* grow `seq` so that positions [0, n_positions) are covered; all or nothing. Returns the number of pages newly taken
 * (>= 0), SLLM_ENOMEM when the free pages do not suffice (nothing changes), SLLM_EINVAL beyond max_pages_per_seq. */
int32_t sllm_kvpages_reserve(sllm_kvpages* kp, int32_t seq, int32_t n_positions);
int sllm_kvpages_release(sllm_kvpages* kp, int32_t seq);   /* all pages of seq return to the free stack */
int32_t sllm_kvpages_free_count(const sllm_kvpages* kp);
int32_t sllm_kvpages_held(const sllm_kvpages* kp, int32_t seq);
const int32_t* sllm_kvpages_table(const sllm_kvpages* kp); /* [max_seqs][max_pages_per_seq] page ids, -1 = none */

typedef struct sllm_batch sllm_batch;
/* A batch borrows the weights, RoPE tables and stream of `e`, which must outlive it: one GPU, weights loaded, created
 * WITHOUT SLLM_ENGINE_MEGAKERNEL (the batched kernels read row-major matrices). It owns the paged cache: two pools of
 * n_pages pages of page_len positions, [pages][layers][kv_heads][page_len][head_dim] in kv_dtype, plus per-slot
 * activations. A sequence may grow to the engine's max_len (the RoPE tables' extent). */
int sllm_batch_create(sllm_engine* e, int32_t max_seqs, int32_t page_len, int32_t n_pages, int32_t kv_dtype, sllm_batch** out);
void sllm_batch_destroy(sllm_batch* b);
/* admit a sequence into the lowest free slot at position 0 (pages are taken when it steps); slot_out = its handle */
int sllm_batch_add(sllm_batch* b, const int32_t* prompt_host, int32_t n_prompt, int32_t* slot_out);
/* retire a sequence: its slot and pages are free for the next sllm_batch_add at once (stream order protects them) */
int sllm_batch_remove(sllm_batch* b, int32_t slot);
/* Sampling instead of arg-max for one sequence (additive, like sllm_sample_f32; not part of the parity contract): from now on the token
 * that follows position p of this slot is sllm_sample_f32(logits_p, temperature, top_k, top_p, seed, step = p) — a pure function of the
 * sequence, whatever its neighbours. temperature <= 0 switches back to arg-max. Prompt tokens are still fed verbatim. Retiring the
 * sequence clears the setting. */
int sllm_batch_set_sampling(sllm_batch* b, int32_t slot, float temperature, int32_t top_k, float top_p, uint64_t seed);
/* n_steps tokens for EVERY live sequence, asynchronous. Takes the pages the steps will write first, for all sequences
 * or none: SLLM_ENOMEM (nothing enqueued) when the pool is short, SLLM_EINVAL when a sequence would pass max_len. */
int sllm_batch_step(sllm_batch* b, int32_t n_steps);
/* Opt-in (on != 0): the steps of this batch run their five projection groups as tcgen05 GEMMs over the live sequences' rows (the prefill GEMM
 * kernel of csrc/prefill_gemm.cu) instead of batched GEMVs, so that every weight byte is read once per step however many sequences are live —
 * the batched GEMVs stop scaling at ~4 sequences. GEMM operands are bf16 (activations rounded after RMSNorm / attention / SwiGLU, fp32
 * accumulation): a sequence's results then agree with the reference within the bf16-operand tolerance of the prefill, not bit for bit.
 * bf16 weights only (SLLM_ENOTSUP otherwise). The reference has no batched decode (model.h:15-18: one token, one position). */
int sllm_batch_set_tensor_cores(sllm_batch* b, int32_t on);
/* the tokens that followed positions 0.. of the slot's sequence (prompt tokens included, as sllm_engine_greedy
 * reports them), at most max_tokens; *n_out = how many. Synchronises the stream. */
int sllm_batch_read(sllm_batch* b, int32_t slot, int32_t* tokens_out_host, int32_t max_tokens, int32_t* n_out);
/* logits (vocab floats) of the slot's latest step. Synchronises the stream. */
int sllm_batch_logits(sllm_batch* b, int32_t slot, float* logits_host);
/* Introspection for parity tests: per-slot buffers under the reference's ModelBufferType numbers (include/model/model.h:14-34),
 * each [max_seqs][...] fp32: 4 emb_output (the residual stream) [d], 6 query [q], 8 mha_output [q], 10 ffn_input [d], 14 swi_output
 * [I], 16 model_pred [vocab]; 2 / 3 = the key / value page pools in kv_dtype. Synchronises the stream. */
int sllm_batch_buffer(sllm_batch* b, int32_t buffer_id, void** dev_ptr, int64_t* n_elems, int32_t* dtype);
int32_t sllm_batch_free_pages(const sllm_batch* b);
int32_t sllm_batch_position(const sllm_batch* b, int32_t slot);   /* position of the slot's next step; -1 = free slot */
/* algorithmic HBM bytes of the NEXT step: every weight once, plus per live sequence its embedding row, the K/V rows
 * 0..pos it reads and the row it writes (SURVEY.md 8d B(p) with the weight term shared) */
int64_t sllm_batch_step_bytes(const sllm_batch* b);
int64_t sllm_batch_total_launches(const sllm_batch* b);
/* Device bytes sllm_batch_create(max_seqs, page_len, n_pages, kv_dtype) takes on an engine of this shape (the two page pools
 * [n_pages][layers][kv_heads][page_len][head_dim], per-slot activations and logits, attention workspace; rounded up to 1 MiB) —
 * host arithmetic only, so that a host sizes n_pages against sllm_device_info's free HBM first; -1 = bad argument. */
int64_t sllm_batch_arena_bytes(const sllm_shape* shape, int32_t max_seqs, int32_t page_len, int32_t n_pages, int32_t kv_dtype);

/* Calibrated partition of the persistent decode kernels (SLLM_ENGINE_MEGAKERNEL, one GPU, not MEGA_LL). The SMs of a B200 do not all
 * stream from HBM at the same rate (a stable pattern of about +-5 % per GPU) and every phase of the kernel ends with its slowest CTA.
 * This call runs `rounds` x 9 decode steps from (token 1, position 0) with the kernel's timeline on, measures every CTA's time per
 * byte over the big weight phases, and from then on gives each CTA a share of every phase's tile rows in inverse proportion. It
 * changes WHICH CTA computes a row, never how a row is computed: results are unchanged (bit for bit in the deterministic kernel).
 * Call it after loading weights and before use: it overwrites the first positions of the KV cache and resets the step state to
 * (token 0, position 0). sllm_engine_calibration(cta) = the measured relative time per byte (1.0 = mean, 0 = never calibrated). */
int sllm_engine_calibrate(sllm_engine* e, int32_t rounds);
float sllm_engine_calibration(const sllm_engine* e, int32_t cta);
/* Introspection for parity tests and the roofline: named buffers follow the reference's ModelBufferType
 * numbering (include/model/model.h:14-34); returns a device pointer and its element count/dtype. */
int sllm_engine_buffer(sllm_engine* e, int32_t buffer_id, void** dev_ptr, int64_t* n_elems, int32_t* dtype);
/* algorithmic HBM bytes one decode step at position pos must move on this rank */
int64_t sllm_engine_step_bytes(const sllm_engine* e, int32_t pos);
/* Enqueue ONE fused kernel of the decode step for `layer` at the current device-side position, for per-kernel
 * timing and profiling. kind: 0 embedding, 1 qkv (RMSNorm+QKV GEMV+RoPE+cache write), 2 mha (flash decoding),
 * 3 wo (+residual), 4 gate_up (RMSNorm+GEMV+sigmoid*up), 5 down (+residual), 6 classifier (+argmax; the step
 * state is left untouched). sllm_engine_kernel_bytes = the algorithmic HBM bytes of that launch (its share of
 * B(p): weights it streams, norm vector, KV rows read/written). */
int sllm_engine_enqueue_kernel(sllm_engine* e, int32_t kind, int32_t layer);
int64_t sllm_engine_kernel_bytes(const sllm_engine* e, int32_t kind, int32_t pos);
/* layout of the key_cache / value_cache buffers: 0 = [layers][max_len][kv_heads*head_dim] (the reference's,
 * model.cpp:264-265), 1 = head-major [layers][kv_heads][max_len][head_dim] (megakernel mode: a tile of positions
 * of one head is one contiguous TMA copy). Weight matrices of a megakernel engine are stored tiled (buffer ids
 * 100-105 are then opaque). */
int32_t sllm_engine_kv_layout(const sllm_engine* e);
/* which decode path this engine runs: "megakernel", "fused+graph", "fused", "unfused", ... (static string) */
const char* sllm_engine_mode(const sllm_engine* e);
/* number of kernel launches (graph kernel nodes) one decode step issues on this rank */
int32_t sllm_engine_step_launches(const sllm_engine* e);
/* total launches issued by this engine since creation */
int64_t sllm_engine_total_launches(const sllm_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* SLLM_B200_H */
