// Attention stream: gather + cos time encoding + masked softmax + weighted sum (see attn_packed.cu).
#pragma once
#include "common.cuh"

namespace flid {

struct AttnArgs {
    const float* u_base;      // query folds (scores in log2 domain), row stride H * kd
    const int32_t* u_index;   // nullable: u row of target i is u_base[u_index[i]] (per-node table) else row i
    const float* hrow_base;   // neighbour layer-(l-1) rows, row stride dn
    int hrow_by_id;           // 1: row = neighbour id (feature table); 0: row = hrow_offset + i*k + j
    int64_t hrow_offset;
    const int32_t* hrow_idx = nullptr;  // non-null: row = hrow_idx[i*k + j] (layer memo table, by CSR position)
    const float* edge_feat;   // [E+1, de]
    const int32_t* nbr;       // [n, k]
    const int32_t* eid;
    const float* dt;
    const float* time_w;
    const float* time_b;
    const float* time_bound;  // [0] = max |w|, [1] = max |b| over the T channels
    float* z;                 // [n, H * kd]
    int64_t n;
    int k, dn, de, T;
};

int launch_attn(const AttnArgs& a, int H, cudaStream_t st);  // attn_packed.cu

}  // namespace flid
