// Pseudo-label scoring epilogue:
//   MLPClassifier.forward (models/modules.py:86-97, eval) -> softmax -> argmax
//   (PTCL/E_step.py:334-335), entropy_filter / prob_filter (PTCL/utils.py:38-67).
// Reads 4*in bytes per event and writes 4*C + 8: HBM-bound and tiny next to the embedding.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "gemm_tc.cuh"

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

// scratch is keyed by device ordinal (a process may drive an E-model and an M-model on two GPUs); calls on one
// device are expected on one stream at a time
constexpr int MAX_DEV = 32;
static DevBuf g_ptrs_dev[MAX_DEV];
static int cur_device(int* dev) {
    FLID_CUDA(cudaGetDevice(dev));
    FLID_REQUIRE(*dev >= 0 && *dev < MAX_DEV, "device ordinal %d not supported", *dev);
    return FLID_OK;
}

// Bulk path: fc1 (in -> hidden1, ReLU) runs on the tcgen05 GEMM, this kernel finishes the row:
// fc2 (ReLU) -> fc3 -> softmax / argmax.  128 rows per block, the h1 tile and the two small
// weight matrices staged in shared memory, one thread per row.
constexpr int TAIL_ROWS = 128, MAXH2 = 32;
__global__ void __launch_bounds__(TAIL_ROWS) pseudo_tail_kernel(flid_mlp_weights w, const float* __restrict__ h1,
                                                                int64_t n, float* __restrict__ probs,
                                                                int64_t* __restrict__ labels,
                                                                float* __restrict__ logits) {
    extern __shared__ float sm[];
    const int H1 = w.hidden1, H2 = w.hidden2, C = w.num_classes, ld = H1 + 1;
    float* tile = sm;                       // [TAIL_ROWS][H1 + 1]
    float* w2 = tile + TAIL_ROWS * ld;      // [H2][H1]
    float* b2 = w2 + H2 * H1;               // [H2]
    float* w3 = b2 + H2;                    // [C][H2]
    float* b3 = w3 + C * H2;                // [C]
    const int64_t r0 = (int64_t)blockIdx.x * TAIL_ROWS;
    const int rows = (int)min((int64_t)TAIL_ROWS, n - r0);
    for (int i = threadIdx.x; i < rows * H1; i += TAIL_ROWS) tile[(i / H1) * ld + (i % H1)] = __ldg(h1 + r0 * H1 + i);
    for (int i = threadIdx.x; i < H2 * H1; i += TAIL_ROWS) w2[i] = __ldg(w.fc2_w + i);
    for (int i = threadIdx.x; i < H2; i += TAIL_ROWS) b2[i] = __ldg(w.fc2_b + i);
    for (int i = threadIdx.x; i < C * H2; i += TAIL_ROWS) w3[i] = __ldg(w.fc3_w + i);
    for (int i = threadIdx.x; i < C; i += TAIL_ROWS) b3[i] = __ldg(w.fc3_b + i);
    __syncthreads();
    if ((int)threadIdx.x >= rows) return;
    const float* x = tile + threadIdx.x * ld;
    float h2[MAXH2];
#pragma unroll
    for (int o = 0; o < MAXH2; ++o) {
        h2[o] = 0.f;
        if (o < H2) {
            float p0 = 0.f, p1 = 0.f;
            int c = 0;
            for (; c + 1 < H1; c += 2) p0 = fmaf(x[c], w2[o * H1 + c], p0), p1 = fmaf(x[c + 1], w2[o * H1 + c + 1], p1);
            if (c < H1) p0 = fmaf(x[c], w2[o * H1 + c], p0);
            h2[o] = fmaxf(p0 + p1 + b2[o], 0.f);
        }
    }
    float lg[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        lg[c] = -INFINITY;
        if (c < C) {
            float p = b3[c];
#pragma unroll
            for (int o = 0; o < MAXH2; ++o)
                if (o < H2) p = fmaf(h2[o], w3[c * H2 + o], p);
            lg[c] = p;
        }
    }
    const int64_t i = r0 + threadIdx.x;
    float mx = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (c < C && lg[c] > mx) mx = lg[c], arg = c;
    float den = 0.f, e[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        e[c] = c < C ? expf(lg[c] - mx) : 0.f;
        den += e[c];
    }
    float pm = -1.f;  // argmax over the probabilities, first index on ties (as in pseudo_label_kernel)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
            const float pr = e[c] / den;
            if (pr > pm) pm = pr, arg = c;
            probs[i * C + c] = pr;
            if (logits) logits[i * C + c] = lg[c];
        }
    }
    labels[i] = arg;
}

static TcWeight g_fc1_dev[MAX_DEV];   // tiled fc1 image, one per device ordinal
static DevBuf g_h1_dev[MAX_DEV];

}  // namespace flid

extern "C" {

int flid_pseudo_label(const flid_mlp_weights* w, const float* emb, int64_t n, float* probs, int64_t* labels,
                      float* logits_or_null, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(w != nullptr, "flid_pseudo_label: null argument");
    if (n <= 0) return FLID_OK;   // an empty batch has no buffers (a rank that owns no source endpoint)
    FLID_REQUIRE(emb && probs && labels, "flid_pseudo_label: null argument");
    FLID_REQUIRE(w->input_dim > 0 && w->input_dim <= 512 && w->hidden1 > 0 && w->hidden1 <= 256 && w->hidden2 > 0 &&
                     w->num_classes > 0 && w->num_classes <= MAXC,
                 "flid_pseudo_label: unsupported decoder shape (input<=512, hidden1<=256, classes<=16)");
    if (n <= 0) return FLID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tail_smem = sizeof(float) * ((size_t)TAIL_ROWS * (w->hidden1 + 1) + (size_t)w->hidden2 * (w->hidden1 + 1) +
                                              (size_t)w->num_classes * (w->hidden2 + 1));
    const bool bulk = n >= 4096 && (w->input_dim % 4) == 0 && (w->hidden1 % 4) == 0 && w->hidden2 <= MAXH2 &&
                      tail_smem <= 48 * 1024;
    if (!bulk) {  // per-batch calls (B = 200): one warp per row, no workspace
        pseudo_label_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(*w, emb, n, probs, labels, logits_or_null);
        FLID_LAUNCH_CHECK();
        return FLID_OK;
    }
    // weights may have changed since the last call: re-tile fc1 (a few KB) every time
    int dev = 0;
    FLID_TRY(cur_device(&dev));
    TcWeight& g_fc1 = g_fc1_dev[dev];
    DevBuf& g_h1 = g_h1_dev[dev];
    FLID_TRY(tc_prepare_weight(w->fc1_w, w->input_dim, w->hidden1, w->input_dim, &g_fc1, st));
    const int64_t chunk = 262144;
    FLID_TRY(g_h1.reserve(sizeof(float) * (size_t)std::min<int64_t>(chunk, n) * w->hidden1));
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t nc = std::min<int64_t>(chunk, n - r0);
        TcGemmArgs a;
        a.A0 = emb + r0 * w->input_dim, a.lda0 = w->input_dim, a.w0 = w->input_dim;
        a.C = g_h1.as<float>(), a.ldc = w->hidden1, a.bias = w->fc1_b, a.M = nc, a.relu = 1;
        FLID_TRY(tc_gemm(a, g_fc1, st));
        pseudo_tail_kernel<<<(unsigned)ceil_div(nc, TAIL_ROWS), TAIL_ROWS, tail_smem, st>>>(
            *w, g_h1.as<float>(), nc, probs + r0 * w->num_classes, labels + r0,
            logits_or_null ? logits_or_null + r0 * w->num_classes : nullptr);
        FLID_LAUNCH_CHECK();
    }
    return FLID_OK;
}

int flid_entropy_filter(const float* const* probs_store_host, int num_iters, int64_t n, int num_classes,
                        float threshold, float* labels, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(probs_store_host && labels && num_iters > 0, "flid_entropy_filter: bad argument");
    FLID_REQUIRE(num_classes > 0 && num_classes <= MAXC, "flid_entropy_filter: classes must be in 1..16");
    if (n <= 0) return FLID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    FLID_TRY(cur_device(&dev));
    DevBuf& g_ptrs = g_ptrs_dev[dev];
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
