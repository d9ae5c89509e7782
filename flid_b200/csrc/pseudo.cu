// Pseudo-label scoring epilogue:
//   MLPClassifier.forward (models/modules.py:86-97, eval) -> softmax -> argmax
//   (PTCL/E_step.py:334-335), entropy_filter / prob_filter (PTCL/utils.py:38-67).
// Reads 4*in bytes per event and writes 4*C + 8: HBM-bound and tiny next to the embedding.
#include <math.h>

#include "common.cuh"

namespace flid {

constexpr int MAXC = 16;

// one warp per row; fc1 rows are reduced across the warp, fc2/fc3 are small enough to
// finish redundantly in every lane.
__global__ void __launch_bounds__(256) pseudo_label_kernel(flid_mlp_weights w, const float* __restrict__ emb,
                                                           int64_t n, float* __restrict__ probs,
                                                           int64_t* __restrict__ labels, float* __restrict__ logits) {
    constexpr int MAXI = 16, MAXH1 = 8;
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    float x[MAXI];
#pragma unroll
    for (int r = 0; r < MAXI; ++r) {
        const int c = lane + 32 * r;
        x[r] = c < w.input_dim ? __ldg(emb + i * w.input_dim + c) : 0.f;
    }
    float h1[MAXH1];
#pragma unroll
    for (int q = 0; q < MAXH1; ++q) h1[q] = 0.f;
    for (int o = 0; o < w.hidden1; ++o) {
        const float* row = w.fc1_w + (int64_t)o * w.input_dim;
        float p = 0.f;
#pragma unroll
        for (int r = 0; r < MAXI; ++r) {
            const int c = lane + 32 * r;
            if (c < w.input_dim) p = fmaf(x[r], __ldg(row + c), p);
        }
        p = fmaxf(warp_sum(p) + __ldg(w.fc1_b + o), 0.f);
        if ((o & 31) == lane) {
#pragma unroll
            for (int q = 0; q < MAXH1; ++q)
                if (q == (o >> 5)) h1[q] = p;
        }
    }
    float lg[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) lg[c] = c < w.num_classes ? __ldg(w.fc3_b + c) : -INFINITY;
    for (int o = 0; o < w.hidden2; ++o) {
        const float* row = w.fc2_w + (int64_t)o * w.hidden1;
        float p = 0.f;
#pragma unroll
        for (int q = 0; q < MAXH1; ++q) {
            const int c = lane + 32 * q;
            if (c < w.hidden1) p = fmaf(h1[q], __ldg(row + c), p);
        }
        const float h2 = fmaxf(warp_sum(p) + __ldg(w.fc2_b + o), 0.f);
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < w.num_classes) lg[c] = fmaf(h2, __ldg(w.fc3_w + c * w.hidden2 + o), lg[c]);
    }
    float mx = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < w.num_classes && lg[c] > mx) mx = lg[c], arg = c;
    float den = 0.f, e[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        e[c] = c < w.num_classes ? expf(lg[c] - mx) : 0.f;
        den += e[c];
    }
    // the reference takes argmax over the probabilities; ties after rounding resolve to the first index
    float pm = -1.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if (c < w.num_classes) {
            const float pr = e[c] / den;
            if (pr > pm) pm = pr, arg = c;
            if (lane == (c & 31)) probs[i * w.num_classes + c] = pr;
            if (logits && lane == (c & 31)) logits[i * w.num_classes + c] = lg[c];
        }
    }
    if (lane == 0) labels[i] = arg;
}

__global__ void entropy_filter_kernel(const float* const* __restrict__ store, int iters, int64_t n, int C, float thr,
                                      float* __restrict__ labels) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
    for (int it = 0; it < iters; ++it) {
        const float* p = store[it] + i * C;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) acc[c] += __ldg(p + c);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < C) mx = fmaxf(mx, acc[c]);
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        acc[c] = c < C ? expf(acc[c] - mx) : 0.f;
        den += acc[c];
    }
    float ent = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
            const float p = acc[c] / den;
            ent += p * log2f(p + 1e-10f);
        }
    }
    if (-ent > thr) labels[i] = -1.f;
}

__global__ void prob_filter_kernel(const float* __restrict__ probs, int64_t n, int C, float thr,
                                   float* __restrict__ labels) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(probs + i * C + c));
    if (mx < thr) labels[i] = -1.f;
}

static DevBuf g_ptrs;

}  // namespace flid

extern "C" {

int flid_pseudo_label(const flid_mlp_weights* w, const float* emb, int64_t n, float* probs, int64_t* labels,
                      float* logits_or_null, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(w && emb && probs && labels, "flid_pseudo_label: null argument");
    FLID_REQUIRE(w->input_dim > 0 && w->input_dim <= 512 && w->hidden1 > 0 && w->hidden1 <= 256 && w->hidden2 > 0 &&
                     w->num_classes > 0 && w->num_classes <= MAXC,
                 "flid_pseudo_label: unsupported decoder shape (input<=512, hidden1<=256, classes<=16)");
    if (n <= 0) return FLID_OK;
    pseudo_label_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(*w, emb, n, probs, labels,
                                                                                         logits_or_null);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

int flid_entropy_filter(const float* const* probs_store_host, int num_iters, int64_t n, int num_classes,
                        float threshold, float* labels, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(probs_store_host && labels && num_iters > 0, "flid_entropy_filter: bad argument");
    FLID_REQUIRE(num_classes > 0 && num_classes <= MAXC, "flid_entropy_filter: classes must be in 1..16");
    if (n <= 0) return FLID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    FLID_TRY(g_ptrs.reserve(sizeof(void*) * num_iters));
    FLID_CUDA(cudaMemcpyAsync(g_ptrs.p, probs_store_host, sizeof(void*) * num_iters, cudaMemcpyHostToDevice, st));
    entropy_filter_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(g_ptrs.as<const float*>(), num_iters, n,
                                                                     num_classes, threshold, labels);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

int flid_prob_filter(const float* probs_last, int64_t n, int num_classes, float threshold, float* labels,
                     flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(probs_last && labels && num_classes > 0, "flid_prob_filter: bad argument");
    if (n <= 0) return FLID_OK;
    prob_filter_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(probs_last, n, num_classes,
                                                                                    threshold, labels);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

}  // extern "C"
