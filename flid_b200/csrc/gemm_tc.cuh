// tcgen05 (5th-gen tensor core) "NT" GEMM with fp32-grade accuracy via the 3xTF32 split:
//   C[m, n] = sum_k A[m, k] * W[n, k] (+ bias[n]) (ReLU),  A = [A0 | A1] (two column segments,
//   each optionally row-gathered), accumulators in TMEM, operands staged in shared memory in
//   the UMMA canonical K-major no-swizzle layout.
// x = hi + lo with hi = x & 0xffffe000 (exact tf32), lo = x - hi (exact fp32, truncated to tf32
// by the tensor core): A.W ~= Ahi.Whi + Ahi.Wlo + Alo.Whi, error ~2^-21 relative per product.
#pragma once
#include "common.cuh"

namespace flid {

constexpr int TC_KC = 16;  // K floats per pipeline stage (2 UMMA k-steps of 8)
constexpr int TC_SMALL_M = 2048;  // launches up to this many rows use the 32-column weight image

// weight pre-split into (hi, lo) and pre-tiled so that one (n-block, k-chunk) stage is a
// single contiguous bulk copy:  [n_block][k_chunk][half][c4 = 4][n_tile][4 floats]
struct TcWeight {
    float* buf = nullptr;
    int N = 0, K = 0, n_tile = 0, n_blocks = 0, k_chunks = 0;
    // 0: fp32-grade product (3xTF32: hi/lo split of both operands, three MMAs per product)
    // 1: "bf16 projections" numeric mode: both operands rounded to bfloat16 (round-to-nearest-even; a bf16 value is
    //    exactly representable in tf32), one MMA per product, fp32 accumulation -- the arithmetic of a bf16-in /
    //    fp32-acc GEMM on the same tensor pipe, with a third of the MMAs and half of the operand traffic
    int single = 0;
    // second image with 32-column tiles for launches with few rows (per-batch calls, M <= TC_SMALL_M): a 128-row tile
    // then spreads over N / 32 CTAs instead of one, which is what a B = 200 drop-in call (400 rows = 4 tiles) needs
    float* small_buf = nullptr;
    int small_tile = 0, small_blocks = 0;
    size_t bytes() const { return (size_t)n_blocks * k_chunks * 2 * (TC_KC / 4) * n_tile * 16; }
};

struct TcGemmArgs {
    const float* A0 = nullptr;
    int64_t lda0 = 0;
    const int32_t* idx0 = nullptr;
    int w0 = 0;  // columns [0, w0) come from segment 0
    const float* A1 = nullptr;
    int64_t lda1 = 0;
    const int32_t* idx1 = nullptr;
    int w1 = 0;  // columns [w0, w0 + w1) from segment 1 (0 = unused)
    float* C = nullptr;
    int64_t ldc = 0;
    const int32_t* cidx = nullptr;  // nullable: output row m is written to C + cidx[m] * ldc (scatter)
    const float* bias = nullptr;
    int64_t M = 0;
    int relu = 0;
    // dense layers of the other sampler consumers (dense.cu): out = act(acc + bias + resid[m]); resid has C's row order
    int gelu = 0;                        // exact (erf) GELU instead of ReLU
    const float* resid = nullptr;
    int64_t ldr = 0;
    // LayerNorm folded into this product (fc1 of the MergeLayer after the attention block, models/modules.py:235-238 then
    // :66): A0 holds residual_fc's outputs; the producers add the residual row [ln_self | ln_tail] on the fly and
    // accumulate each row's sum and sum of squares (float64); the weight image is W diag(gamma); the epilogue applies
    //     out = act( rstd * (acc - mean * ln_c1[n]) + ln_add[row][n] )
    // with ln_c1 = row sums of W diag(gamma) and ln_add = everything that does not depend on the normalised row
    // (W beta + bias + the raw-feature part of fc1, one row per node).  A-from-TMEM kernel, single n block only.
    const float* ln_self = nullptr;      // [*, ln_self_w], row = ln_self_idx ? ln_self_idx[m] : m
    const int32_t* ln_self_idx = nullptr;
    int ln_self_w = 0;
    const float* ln_tail = nullptr;      // [w0 - ln_self_w] constant tail of the residual
    const float* ln_c1 = nullptr;        // [N]
    const float* ln_add = nullptr;       // [*, ln_add_ld], row = ln_add_idx[m]
    const int32_t* ln_add_idx = nullptr;
    int64_t ln_add_ld = 0;
    float ln_eps = 1e-5f;
};

// (re)build the tiled hi/lo image of W[N, K] (row stride ldw); allocates w->buf on first use
int tc_prepare_weight(const float* W, int64_t ldw, int N, int K, TcWeight* w, cudaStream_t st, int single = 0);
// same for a weight stored transposed, Wt[K, N] (row stride ldw): the data-gradient GEMMs of train_layer.cu
int tc_prepare_weight_t(const float* Wt, int64_t ldw, int N, int K, TcWeight* w, cudaStream_t st);
void tc_free_weight(TcWeight* w);
int tc_gemm(const TcGemmArgs& g, const TcWeight& w, cudaStream_t st);
// A-operand-in-TMEM variant, gemm_tc_ts.cu
int tc_gemm_ts(const TcGemmArgs& g, const TcWeight& w, int sm_count, int smem_max, cudaStream_t st);

}  // namespace flid
