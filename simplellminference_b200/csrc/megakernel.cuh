// csrc/megakernel.cuh — host-side interface of the persistent decode-step kernel (megakernel.cu).
#pragma once
#include "decode_fused.cuh"

namespace sllm {

// Megakernel weight layout ("tiled"): the rows of a phase's matrix are first put in UNIT ORDER — unit u owns
// physical rows 2u, 2u+1 = the two logical rows its epilogue needs together (RoPE partners j / j+hd/2 of a q or k
// head, two consecutive v rows, (up_u, gate_u), or two consecutive rows) — and then cut into tiles of R physical
// rows x SC 16-byte chunks; tile (g, ks) (row group g, K slice ks) is CONTIGUOUS at (g*KS + ks) * tile_bytes,
// rows inside a tile SC*16 bytes apart. One ring slot == one tile == ONE cp.async.bulk (measured on B200: a bulk
// copy costs ~46 cycles of TMA issue per SM whatever its size, so 1 KB copies cap the stream at the HBM rate).
// K is zero-padded to KS*SC chunks, rows to a multiple of R.
struct PhaseDesc {       // one weight phase, built on the host
    const uint8_t* W;    // tiled matrix
    int32_t nchunks;     // real 16-byte chunks per logical row
    int32_t nunits;      // two-row units
    int32_t nrows;       // logical rows
    int32_t kind, layer;
    int32_t KS;          // K slices (power of two <= 16); RG = 16 / KS row groups stream concurrently
    int32_t SC;          // chunks per slice
    int32_t R;           // physical rows per tile (4 or 2)
    int32_t ntr;         // tile rows = ceil(2*nunits / R)
    int32_t tile_bytes;  // R * (SC * 16 + srow)
    int32_t srow;        // int8 only: bytes of one row's group scales inside a tile (SC/4 fp32 scales padded to 16 bytes), else 0
    const uint32_t* cum; // optional [grid + 1] cumulative shares of the CTAs in 2^-24 units (cum[0] = 0, cum[grid] = 2^24): CTA c takes tile
                         // rows [ntr * cum[c], ntr * cum[c + 1]) instead of an equal cut. SMs do not all stream from HBM at the same rate
                         // (a stable +-5 % pattern per GPU, profiles/r02_mega_trace_v1.txt) and every phase ends with the slowest one:
                         // sllm_engine_calibrate measures the pattern and sizes the shares by it. nullptr = equal shares.
};

// geometry of a [rows][cols] matrix in the tiled layout
struct TileGeom { int nchunks, KS, SC, R, ntr, tile_bytes, srow; size_t bytes; };
TileGeom mega_tile_geom(int rows_phys, int cols, int w_dtype);
// bytes of one [rows][cols] matrix of phase kind `kind` in the tiled layout
size_t mega_matrix_bytes(int rows, int cols, int kind, int w_dtype);

struct MegaParams {
    const PhaseDesc* phases;   // [4L+1]
    int32_t d, hd, L, S, V, V_loc, v0, q_loc, kv_loc, I_loc, H_loc, KVH_loc, nsplit;
    int32_t w_dtype, kv_dtype;
    float eps;
    const uint8_t* emb;        // [V][d] storage dtype
    const float* norms;        // [(2L+1)][d]
    uint8_t *kc, *vc;          // HEAD-MAJOR cache [L][KVH_loc][S][hd] kv dtype: a tile of positions of one head is contiguous
    const float *sin_t, *cos_t;
    float *x, *h, *q, *swi, *logits;
    float* att_part;           // [H_loc][nsplit][hd+4]
    float* blk_val;
    int32_t* blk_idx;
    StepState* st;
    const int32_t* prompt;
    int32_t* history;
    unsigned* bar_counter;
    unsigned long long* trace;  // optional [grid][512][8] %globaltimer stamps (nullptr = off)
    int32_t debug;              // measurement aid (sllm_tune key 8; results are garbage): bit 0 = skip the grid barriers, bit 1 = skip the dot products
};

// ---- barrier-free {value, epoch}-word version (megakernel_ll.cu), single GPU and tensor parallel -------------
struct MegaLLParams {
    const PhaseDesc* phases;   // [4L+1]
    int32_t d, hd, L, S, V, V_loc, v0, q_loc, kv_loc, I_loc, H_loc, KVH_loc, nsplit;
    int32_t w_dtype, kv_dtype;
    float eps;
    const uint8_t* emb;        // full tiled [V][d] (gather); the classifier phase points at this rank's vocab rows
    const float* norms;
    uint8_t *kc, *vc;          // head-major cache [L][KVH_loc][S][hd]
    const float *sin_t, *cos_t;
    float* logits;
    float* x_out;              // introspection only: CTA 0 stores the final residual stream here (emb_output)
    float* blk_val;
    int32_t* blk_idx;
    StepState* st;
    const int32_t* prompt;
    int32_t* history;
    int32_t tp, rank;
    uint2* area[8];            // area[r] = rank r's word area (area[rank] is local memory, the others CUDA-IPC mappings)
    int64_t off_wop, off_dnp, off_qv, off_kvn, off_att, off_swi, off_arg;   // word offsets inside an area
    unsigned long long* trace; // optional [grid][512][8] %globaltimer stamps (nullptr = off)
};
struct MegaLLPlan {
    bool ok = false;
    int grid = 0;
    size_t smem = 0;
    int nsplit = 1;
    int64_t off_wop = 0, off_dnp = 0, off_qv = 0, off_kvn = 0, off_att = 0, off_swi = 0, off_arg = 0, area_words = 0;
    const char* why = "";
};
MegaLLPlan mega_ll_plan(int w_dtype, int kv_dtype, int d, int hd, int q_loc, int kv_loc, int I_loc, int V_loc, int v0, int H_loc, int KVH_loc,
                        int max_len, int tp);
MegaLLPlan mega_ll_plan_for(int sms, int smem_optin, int w_dtype, int kv_dtype, int d, int hd, int q_loc, int kv_loc, int I_loc, int V_loc, int v0,
                            int H_loc, int KVH_loc, int max_len, int tp);
int mega_ll_launch(const MegaLLParams& p, int g, int grid, size_t smem, cudaStream_t st);

struct MegaPlan {
    bool ok = false;
    int grid = 0;
    size_t smem = 0;
    int nsplit = 1;
    size_t att_part_floats = 0;
    const char* why = "";
};

MegaPlan mega_plan(int w_dtype, int group, int kv_dtype, int d, int hd, int q_loc, int kv_loc, int I_loc, int V_loc, int H_loc, int KVH_loc, int max_len);
MegaPlan mega_plan_for(int sms, int smem_optin, int w_dtype, int group, int kv_dtype, int d, int hd, int q_loc, int kv_loc, int I_loc, int V_loc,
                       int H_loc, int KVH_loc, int max_len);
void mega_fill_phases(PhaseDesc* host, int L, int w_dtype, const void* wqkv, const void* wo, const void* wug, const void* wdown,
                      const void* cls, int d, int q_loc, int kv_loc, int I_loc, int V_loc);
// row-major [rows][cols] (storage dtype) -> tiled layout in unit order. kind: PH_* (row pairing rule).
// scales_rowmajor: int8 only, fp32 [rows][cols/64] (the tiled layout carries each tile's scales behind its weights)
int mega_repack(const void* src_rowmajor, const float* scales_rowmajor, void* dst_tiled, int rows, int cols, int kind, int w_dtype, int hd,
                int q_loc, int kv_loc, int I_loc, cudaStream_t st);
enum { PH_QKV = 0, PH_WO = 1, PH_GATEUP = 2, PH_DOWN = 3, PH_CLS = 4, PH_DOWN_T = 5, PH_WO_T = 6 };
int mega_launch(const MegaParams& p, int g, int grid, size_t smem, cudaStream_t st, bool fuse_down = false);

// ---- experimental: down projection fused into the gate_up phase (SLLM_ENGINE_MEGA_FUSE_DOWN) ---------------------------------
// K-split of Wdown over the CTAs: a CTA multiplies the columns of Wdown that belong to the sigma(gate)*up values IT produced and
// adds its partial output vector into the residual stream with red.global.add.v4.f32, so gate_up -> down needs no grid barrier, no
// activation staging and no cross-lane reduction. Layout of the transposed matrix ("PH_DOWN_T"): tile row g = kFuseJT consecutive
// inputs j; K slice ks = a stripe of 32 lanes x 16 bytes of outputs r (d = KS * 32 * E outputs, KS a power of two <= 16); tile
// (g, ks) = kFuseJT x 512 bytes; the matrix is STRIPE-MAJOR, tile (g, ks) at (ks * ntr + g) * tile_bytes, so that the tile rows a
// warp takes (a contiguous share of its CTA's range) are one contiguous byte range that it copies a whole ring slot (two tile rows)
// at a time. Element e of lane `lane` in row jj of the tile is Wdown[(ks * 32 + lane) * E + e][g * kFuseJT + jj]. A lane owns E
// outputs for the whole phase.
constexpr int kFuseJT = 4;
bool mega_fuse_down_ok(int w_dtype, int d, int I_loc);
size_t mega_down_t_bytes(int d, int I_loc, int w_dtype);
void mega_fill_down_t(PhaseDesc& ds, const void* Wt, int d, int I_loc, int layer, int w_dtype);
int mega_repack_down_t(const void* src_rowmajor /* [d][I_loc] */, void* dst, int d, int I_loc, int w_dtype, cudaStream_t st);


// ---- megakernel2.cu: two (or three) grid-wide dependency points per layer instead of five (SLLM_ENGINE_MEGA_V2) -------------------
// qkv -> attention and attention -> wo go through per-kv-head counters instead of grid barriers; wo is a K split by kv head group
// over the CTAs that hold the group's attention items, reading a column-block copy of Wo ("WoT", PH_WO_T: for group g the block
// Wo[:, g*G*hd:(g+1)*G*hd] as [d rows][CRP chunks], 4 KB tiles of 256 / CRP whole rows) and adding into h with red.global.add.f32.
struct Mega2Params {
    MegaParams m;          // phases[4l + 1] is the PH_WO_T descriptor; m.x / m.h are not used
    float* xbuf[2];        // residual stream entering layer gl (global layer index = launches * L + l): xbuf[gl & 1]
    float* hbuf[2];        // h of layer gl: hbuf[gl & 1], zeroed during layer gl - 1
    float* x_copy;         // introspection: final residual stream (emb_output)
    float* h_copy;         // introspection: h of the last layer (ffn_input)
    unsigned* flags;       // [2][KVH_loc] monotonic counters, one per 128-byte line: units of q/k/v finished, attention splits finished
};
struct WotGeom { int nchunks, crp, nr, ntr; size_t group_bytes, bytes; };
WotGeom mega2_wot_geom(int d, int ghd, int kvh, int w_dtype);
bool mega2_ok(int w_dtype, int d, int hd, int H_loc, int KVH_loc, int nsplit, int grid, const char** why);
size_t mega2_wot_bytes(int d, int hd, int H_loc, int KVH_loc, int w_dtype);
void mega2_fill_wot(PhaseDesc& ds, const void* W, int d, int hd, int H_loc, int KVH_loc, int layer, int w_dtype);
int mega2_repack_wot(const void* src_rowmajor /* [d][H_loc*hd] */, void* dst, int d, int hd, int H_loc, int KVH_loc, int w_dtype, cudaStream_t st);
int mega2_launch(const Mega2Params& P, int g, int grid, size_t smem, cudaStream_t st, bool fuse_down);

}  // namespace sllm
