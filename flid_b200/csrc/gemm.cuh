// fp32 SIMT "NT" GEMM used by the fp32-parity mode of the projection chain:
//   C[m, n] (+)= sum_k A[row(m), k] * W[n, k]  (+ bias[n]) (ReLU)
// W is an [out, in] row-major weight exactly as the reference's nn.Linear stores it,
// A rows may be gathered through an int32 index (feature-table lookups fused into the
// GEMM's A-tile load).  All dims that are used as K must be multiples of 4 (float4 loads).
#pragma once
#include "common.cuh"

namespace flid {

struct GemmArgs {
    const float* A;
    int64_t lda;
    const int32_t* a_idx;  // nullable: row m reads A + a_idx[m] * lda
    const float* W;
    int64_t ldw;
    float* C;
    int64_t ldc;
    const float* bias;  // nullable [N]
    int64_t M;
    int N, K;
    int accumulate;  // C += ...
    int relu;
    const int32_t* c_idx = nullptr;  // nullable: output row m is written to C + c_idx[m] * ldc (scatter)
};

int launch_gemm(const GemmArgs& g, cudaStream_t st);

}  // namespace flid
