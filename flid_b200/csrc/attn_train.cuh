// Training-mode attention stream: forward + backward kernels (see attn_train.cu).
#pragma once
#include "common.cuh"

namespace flid {

struct AttnTrainArgs {
    const float* u;          // [n, H, kd] folded queries (natural-log score domain, scaling included)
    const float* table;      // [R, dn] rows the neighbour slots read
    const int64_t* hrow;     // [n, k] row of `table` per slot; null: row = hrow_offset + i * k + j
    const int64_t* nbr;      // [n, k] neighbour ids, 0 = padded slot (models/modules.py:206-215)
    const int64_t* eid;      // [n, k]
    const float* dt;         // [n, k]
    const float* edge_feat;  // [E+1, de]
    const float* time_w;     // [T]
    const float* time_b;     // [T]
    int64_t n;
    int k, dn, de, T;
    float p_drop;            // dropout on the attention probabilities (modules.py:224)
    uint64_t seed;
    float* z;                // [n, H, kd]            (forward only)
    float* probs;            // [n, H, k] softmax probabilities before dropout (forward only; saved for backward)
    int64_t hrow_offset = 0;
};

struct AttnTrainGrads {
    const float* probs;      // [n, H, k]
    const float* dz;         // [n, H, kd]
    float* du;               // [n, H, kd]
    float* dtable;           // nullable; [R, dn], ACCUMULATED into with atomics
    float* dtime_partial;    // nullable; [attn_train_blocks(n), 2, T] per-block sums of (dL/dw, dL/db)
};

int64_t attn_train_blocks(int64_t n);
int launch_attn_train_fwd(const AttnTrainArgs& a, int H, cudaStream_t st);
int launch_attn_train_bwd(const AttnTrainArgs& a, const AttnTrainGrads& g, int H, cudaStream_t st);
int launch_keep_mask(uint64_t seed, int64_t n, int H, int k, float p, uint8_t* keep, cudaStream_t st);

}  // namespace flid
