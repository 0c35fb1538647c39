// csrc/paged_kv.cuh — addressing of the paged KV cache of the batched multi-sequence decode (batch.cu, mha.cu).
//
// Pool layout (keys and values each): [pages][layers][kv_heads][page_len][head_dim] in the KV dtype — a page holds
// page_len consecutive positions of ONE sequence for every layer and head, and the rows of one (page, layer, head) are
// contiguous (page_len * head_dim elements: one bulk copy per tile when page_len is the attention tile).
// block_table[slot][i] = page that holds positions i*page_len .. (i+1)*page_len-1 of the sequence in `slot`.
#pragma once
#include "common.cuh"

namespace sllm {

struct PagedKv {
    const int32_t* block_table;   // [slots][max_pages]; -1 = no page
    const int32_t* pos;           // [slots]: position of the slot's current token; < 0 = slot not in use
    int max_pages, page_len, layers;
    int q_stride;                 // floats between consecutive slots in q / out
    int heads;                    // query heads (a slot's split-KV partials: heads * nsplit records)
};

// element index of (page, layer, kv head, slot-in-page, 0) in a pool
__host__ __device__ inline size_t paged_row_index(int page, int layers, int layer, int kv_heads, int kvh, int page_len, int in_page, int hd) {
    return ((((size_t)page * layers + layer) * kv_heads + kvh) * page_len + in_page) * hd;
}

int mha_paged_nsplit(int kv_heads, int slots, int max_ctx);
size_t mha_paged_workspace_bytes(int slots, int heads, int kv_heads, int head_dim);
int mha_paged_dispatch(const float* q, const void* k_pool, const void* v_pool, int kv_dtype, float* out, void* ws, int layer,
                       const PagedKv& pk, int slots, int max_slots, int nsplit, int hd, int kv_heads, cudaStream_t st);

}  // namespace sllm
