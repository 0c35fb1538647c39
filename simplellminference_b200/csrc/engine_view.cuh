// csrc/engine_view.cuh — what other translation units may see of an engine (engine.cu owns the struct): the weights in
// their plain row-major layout, the RoPE tables and the shape, for subsystems that share an engine's weights
// (batch.cu: batched multi-sequence decode).
#pragma once
#include "common.cuh"

namespace sllm {

struct EngineView {
    sllm_shape shape;
    int w_dtype, group;
    int tp, mega;            // tensor-parallel size; 1 = weights are stored tiled for the megakernel (not usable here)
    int weights_loaded;
    cudaStream_t stream;
    const void* emb;         // [vocab][d]                      embedding = classifier
    const float* emb_sc;     // int8 group scales (or nullptr), same for the matrices below
    const float* norms;      // [(2L+1)][d] fp32
    const void* wqkv;        // [L][q + 2kv][d]
    const float* wqkv_sc;
    const void* wo;          // [L][d][q]
    const float* wo_sc;
    const void* wug;         // [L][2I][d]   up rows, then gate rows
    const float* wug_sc;
    const void* wdown;       // [L][d][I]
    const float* wdown_sc;
    const float *sin_t, *cos_t;   // [max_len][hd/2]
};

int engine_view(const sllm_engine* e, EngineView* out);

}  // namespace sllm
