// Dense halves of the other two consumers of the device CSR, evaluation path:
//   GraphMixer link encoder   models/GraphMixer.py:91-117, :172-246   (time tokens -> Linear(T, 100) -> MLP-Mixer blocks -> mean)
//   TCL encoder               models/TCL.py:108-157, models/modules.py:248-312 (projections + depth rows, post-norm
//                             transformer blocks over nn.MultiheadAttention on (k + 1)-token sequences)
// Every Linear runs on the tcgen05 3xTF32 GEMM (gemm_tc*.cu) through flid_dense: up to two row-gathered input
// segments, bias, residual add and ReLU / exact GELU in the epilogue.  What is not a GEMM is one small kernel each:
// time-token rows, token mixing (LayerNorm over the tokens + the k -> k/2 -> k feed-forward, one thread per
// (query, channel), tokens in registers), row LayerNorm, the mean over tokens, the periodic depth-row add and the
// (k + 1) x (k + 1) sequence attention with key padding mask (one CTA per sequence, Q/K/V rows in shared memory).
// Training keeps the torch modules (autograd); these entry points are forward only.
#include "gemm_tc.cuh"
#include "flid_b200.h"

struct flid_dense_weight {
    flid::TcWeight w;
};

namespace flid {
namespace {

// ---------------------------------------------------------------- time-token rows
// out[r, :] = cos(dt[r] * w + b), or zeros where ids[r] == 0 (GraphMixer.py:103-108); ids may be null (TCL.py:118)
__global__ void __launch_bounds__(256) time_rows_kernel(const float* __restrict__ dt, const int64_t* __restrict__ ids,
                                                        const float* __restrict__ w, const float* __restrict__ b, int T,
                                                        float* __restrict__ out, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * T) return;
    const int64_t r = i / T;
    const int c = (int)(i - r * T);
    float v = 0.f;
    if (ids == nullptr || __ldg(ids + r) != 0) v = cos_accurate(fmaf(__ldg(dt + r), __ldg(w + c), __ldg(b + c)));
    out[i] = v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// ---------------------------------------------------------------- token mixing (GraphMixer.py:226-236)
// x [m, k, C]: per (query, channel) the k token values -> LayerNorm over tokens -> Linear(k, hid) -> GELU ->
// Linear(hid, k) -> + input.  Weights in shared memory; the token vector and the output vector live in registers.
template <int KMAX, bool EXACT>   // EXACT: k == KMAX, the token loops carry no predicates
__global__ void __launch_bounds__(128) token_mix_kernel(const float* __restrict__ x, int k_rt, int C, const float* __restrict__ ln_w,
                                                        const float* __restrict__ ln_b, float eps, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, int hid, float* __restrict__ out,
                                                        int64_t m) {
    extern __shared__ __align__(16) float sm[];
    const int k = EXACT ? KMAX : k_rt;
    float* s_w1 = sm;                 // [hid][k]
    float* s_w2t = s_w1 + hid * k;    // [hid][k]  (W2 transposed)
    float* s_b1 = s_w2t + hid * k;    // [hid]
    float* s_b2 = s_b1 + hid;         // [k]
    float* s_g = s_b2 + k;            // [k]
    float* s_be = s_g + k;            // [k]
    for (int i = threadIdx.x; i < hid * k; i += blockDim.x) {
        s_w1[i] = __ldg(w1 + i);
        const int t = i / hid, j = i - t * hid;   // w2 is [k][hid]
        s_w2t[j * k + t] = __ldg(w2 + i);
    }
    for (int i = threadIdx.x; i < hid; i += blockDim.x) s_b1[i] = __ldg(b1 + i);
    for (int i = threadIdx.x; i < k; i += blockDim.x) s_b2[i] = __ldg(b2 + i), s_g[i] = __ldg(ln_w + i), s_be[i] = __ldg(ln_b + i);
    __syncthreads();
    const float inv_k = 1.0f / (float)k;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < m * C; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = idx / C;
        const int c = (int)(idx - q * C);
        const float* xp = x + q * (int64_t)k * C + c;
        float v[KMAX], y[KMAX], o[KMAX];
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
            v[t] = (EXACT || t < k) ? __ldg(xp + (int64_t)t * C) : 0.f;
            sum += v[t];
        }
        const float mean = sum * inv_k;
        float var = 0.f;
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
            const float d = (EXACT || t < k) ? v[t] - mean : 0.f;
            var = fmaf(d, d, var);
        }
        const float rstd = rsqrtf(var * inv_k + eps);
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
            y[t] = (EXACT || t < k) ? fmaf((v[t] - mean) * rstd, s_g[t], s_be[t]) : 0.f;
            o[t] = (EXACT || t < k) ? s_b2[t] : 0.f;
        }
        for (int j = 0; j < hid; ++j) {
            float h = s_b1[j];
            const float* wr = s_w1 + j * k;
            const float* wt = s_w2t + j * k;
            if (EXACT && (KMAX & 3) == 0) {   // rows of k floats stay 16-byte aligned: 128-bit shared loads
#pragma unroll
                for (int t = 0; t < KMAX; t += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wr + t);
                    h = fmaf(w4.x, y[t], h), h = fmaf(w4.y, y[t + 1], h), h = fmaf(w4.z, y[t + 2], h), h = fmaf(w4.w, y[t + 3], h);
                }
                h = gelu_erf(h);
#pragma unroll
                for (int t = 0; t < KMAX; t += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wt + t);
                    o[t] = fmaf(w4.x, h, o[t]), o[t + 1] = fmaf(w4.y, h, o[t + 1]), o[t + 2] = fmaf(w4.z, h, o[t + 2]),
                    o[t + 3] = fmaf(w4.w, h, o[t + 3]);
                }
            } else {
#pragma unroll
                for (int t = 0; t < KMAX; ++t)
                    if (EXACT || t < k) h = fmaf(wr[t], y[t], h);
                h = gelu_erf(h);
#pragma unroll
                for (int t = 0; t < KMAX; ++t)
                    if (EXACT || t < k) o[t] = fmaf(wt[t], h, o[t]);
            }
        }
        float* op = out + q * (int64_t)k * C + c;
#pragma unroll
        for (int t = 0; t < KMAX; ++t)
            if (EXACT || t < k) op[(int64_t)t * C] = v[t] + o[t];
    }
}

// ---------------------------------------------------------------- mean over tokens (GraphMixer.py:117)
__global__ void __launch_bounds__(256) token_mean_kernel(const float* __restrict__ x, int k, int C, float* __restrict__ out,
                                                         int64_t ldo, int64_t m) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= m * C) return;
    const int64_t q = idx / C;
    const int c = (int)(idx - q * C);
    const float* xp = x + q * (int64_t)k * C + c;
    float s = 0.f;
    for (int t = 0; t < k; ++t) s += __ldg(xp + (int64_t)t * C);
    out[q * ldo + c] = s / (float)k;
}

// ---------------------------------------------------------------- row LayerNorm
// one warp per row, D <= 32 * 4 * LN_C floats, the row in registers (two-pass variance, as torch's kernel)
constexpr int LN_C = 8;
__global__ void __launch_bounds__(256) row_layernorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* __restrict__ y,
                                                            int64_t ldy, int64_t M, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= M) return;
    const int c4 = D >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + r * ldx);
    float4 v[LN_C];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_C; ++i) {
        const int f = lane + 32 * i;
        v[i] = f < c4 ? __ldg(xr + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    const float mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_C; ++i) {
        const int f = lane + 32 * i;
        if (f < c4) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(FULL, q, o);
    const float rstd = rsqrtf(q / (float)D + eps);
    float4* yr = reinterpret_cast<float4*>(y + r * ldy);
#pragma unroll
    for (int i = 0; i < LN_C; ++i) {
        const int f = lane + 32 * i;
        if (f < c4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + f), b = __ldg(reinterpret_cast<const float4*>(beta) + f);
            yr[f] = make_float4(fmaf((v[i].x - mean) * rstd, g.x, b.x), fmaf((v[i].y - mean) * rstd, g.y, b.y),
                                fmaf((v[i].z - mean) * rstd, g.z, b.z), fmaf((v[i].w - mean) * rstd, g.w, b.w));
        }
    }
}

// ---------------------------------------------------------------- x[r, :] += table[r % S, :]  (depth embedding, TCL.py:127-131)
__global__ void __launch_bounds__(256) add_periodic_kernel(float* __restrict__ x, const float* __restrict__ table, int S, int D,
                                                           int64_t M) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c4 = D >> 2;
    if (i >= M * c4) return;
    const int64_t r = i / c4;
    const int f = (int)(i - r * c4);
    float4 a = reinterpret_cast<float4*>(x + r * D)[f];
    const float4 t = __ldg(reinterpret_cast<const float4*>(table + (r % S) * D) + f);
    a.x += t.x, a.y += t.y, a.z += t.z, a.w += t.w;
    reinterpret_cast<float4*>(x + r * D)[f] = a;
}

// ---------------------------------------------------------------- sequence attention (nn.MultiheadAttention core)
// One CTA per sequence: Q (scaled by head_dim^-1/2, as torch scales the projected query), K, V rows of all heads in
// shared memory (row stride d + 1: the per-key dot products of a warp hit distinct banks), S x S scores per head,
// key padding mask = -inf where key_ids == 0, softmax, P V.  modules.py:287-300 calls it with S = k + 1 <= 33.
__global__ void __launch_bounds__(128) seq_attention_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                            int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                            const int64_t* __restrict__ key_ids, int S, int H, int hd,
                                                            float* __restrict__ out, int64_t ldo, int q_rows) {
    extern __shared__ float sm[];
    const int d = H * hd, ld = d + 1, ps = S + 1;
    float* s_q = sm;
    float* s_k = s_q + S * ld;
    float* s_v = s_k + S * ld;
    float* s_p = s_v + S * ld;          // [H][S][S + 1]
    int* s_mask = reinterpret_cast<int*>(s_p + H * S * ps);
    const int64_t e = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int c4 = d >> 2;
    const float scale = rsqrtf((float)hd);
    for (int i = tid; i < S * c4; i += nt) {
        const int r = i / c4, f = i - r * c4;
        const float4 b = __ldg(reinterpret_cast<const float4*>(k + (e * S + r) * ldk) + f);
        const float4 c = __ldg(reinterpret_cast<const float4*>(v + (e * S + r) * ldv) + f);
        float* kp = s_k + r * ld + 4 * f;
        float* vp = s_v + r * ld + 4 * f;
        kp[0] = b.x, kp[1] = b.y, kp[2] = b.z, kp[3] = b.w;
        vp[0] = c.x, vp[1] = c.y, vp[2] = c.z, vp[3] = c.w;
        if (r < q_rows) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(q + (e * S + r) * ldq) + f);
            float* qp = s_q + r * ld + 4 * f;
            qp[0] = a.x * scale, qp[1] = a.y * scale, qp[2] = a.z * scale, qp[3] = a.w * scale;
        }
    }
    for (int i = tid; i < S; i += nt) s_mask[i] = __ldg(key_ids + e * S + i) == 0;
    __syncthreads();
    // scores in 3 x 3 register tiles: 6 shared loads per 9 products (S = 21: 2 x 7 x 7 tiles for the 128 threads)
    const int ti_n = (q_rows + 2) / 3, tj_n = (S + 2) / 3;
    for (int idx = tid; idx < H * ti_n * tj_n; idx += nt) {
        const int tj = idx % tj_n, ti = (idx / tj_n) % ti_n, h = idx / (tj_n * ti_n);
        const int i0 = ti * 3, j0 = tj * 3;
        // rows past the end are clamped (their products are computed and dropped)
        const float* q0 = s_q + min(i0, q_rows - 1) * ld + h * hd;
        const float* q1 = s_q + min(i0 + 1, q_rows - 1) * ld + h * hd;
        const float* q2 = s_q + min(i0 + 2, q_rows - 1) * ld + h * hd;
        const float* k0 = s_k + min(j0, S - 1) * ld + h * hd;
        const float* k1 = s_k + min(j0 + 1, S - 1) * ld + h * hd;
        const float* k2 = s_k + min(j0 + 2, S - 1) * ld + h * hd;
        float a[3][3] = {};
        for (int c = 0; c < hd; ++c) {
            const float x0 = q0[c], x1 = q1[c], x2 = q2[c], y0 = k0[c], y1 = k1[c], y2 = k2[c];
            a[0][0] = fmaf(x0, y0, a[0][0]), a[0][1] = fmaf(x0, y1, a[0][1]), a[0][2] = fmaf(x0, y2, a[0][2]);
            a[1][0] = fmaf(x1, y0, a[1][0]), a[1][1] = fmaf(x1, y1, a[1][1]), a[1][2] = fmaf(x1, y2, a[1][2]);
            a[2][0] = fmaf(x2, y0, a[2][0]), a[2][1] = fmaf(x2, y1, a[2][1]), a[2][2] = fmaf(x2, y2, a[2][2]);
        }
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int dj = 0; dj < 3; ++dj)
                if (i0 + di < q_rows && j0 + dj < S) s_p[(h * S + i0 + di) * ps + j0 + dj] = s_mask[j0 + dj] ? -INFINITY : a[di][dj];
    }
    __syncthreads();
    for (int row = tid; row < H * q_rows; row += nt) {
        const int h = row / q_rows, i = row - h * q_rows;
        float* p = s_p + (h * S + i) * ps;
        float mx = -INFINITY;
        for (int j = 0; j < S; ++j) mx = fmaxf(mx, p[j]);
        float sum = 0.f;
        for (int j = 0; j < S; ++j) {
            const float w = expf(p[j] - mx);    // all keys masked: -inf - -inf = NaN, as torch's softmax over a fully masked row
            p[j] = w;
            sum += w;
        }
        const float inv = 1.0f / sum;
        for (int j = 0; j < S; ++j) p[j] *= inv;
    }
    __syncthreads();
    // P V: one thread per (3 query rows, column): 4 shared loads per 3 products
    for (int idx = tid; idx < ti_n * d; idx += nt) {
        const int ti = idx / d, c = idx - ti * d, h = c / hd;
        const int i0 = ti * 3;
        const float* p0 = s_p + (h * S + min(i0, q_rows - 1)) * ps;
        const float* p1 = s_p + (h * S + min(i0 + 1, q_rows - 1)) * ps;
        const float* p2 = s_p + (h * S + min(i0 + 2, q_rows - 1)) * ps;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int j = 0; j < S; ++j) {
            const float vv = s_v[j * ld + c];
            a0 = fmaf(p0[j], vv, a0), a1 = fmaf(p1[j], vv, a1), a2 = fmaf(p2[j], vv, a2);
        }
        out[(e * S + i0) * ldo + c] = a0;
        if (i0 + 1 < q_rows) out[(e * S + i0 + 1) * ldo + c] = a1;
        if (i0 + 2 < q_rows) out[(e * S + i0 + 2) * ldo + c] = a2;
    }
}

}  // namespace
}  // namespace flid

using namespace flid;

// ---------------------------------------------------------------- C ABI
extern "C" flid_dense_weight* flid_dense_weight_create(const float* weight, int64_t ldw, int n_out, int n_in, flid_stream stream) {
    if (!weight || n_out <= 0 || n_in <= 0 || (n_in & 3) != 0) {
        set_error("flid_dense_weight_create: weight must be [n_out, n_in] with n_in a multiple of 4");
        return nullptr;
    }
    flid_dense_weight* h = new flid_dense_weight();
    if (tc_prepare_weight(weight, ldw, n_out, n_in, &h->w, (cudaStream_t)stream, 0) != FLID_OK) {
        tc_free_weight(&h->w);
        delete h;
        return nullptr;
    }
    return h;
}

extern "C" int flid_dense_weight_update(flid_dense_weight* h, const float* weight, int64_t ldw, flid_stream stream) {
    FLID_REQUIRE(h && weight, "flid_dense_weight_update: null argument");
    return tc_prepare_weight(weight, ldw, h->w.N, h->w.K, &h->w, (cudaStream_t)stream, 0);
}

extern "C" void flid_dense_weight_free(flid_dense_weight* h) {
    if (!h) return;
    tc_free_weight(&h->w);
    delete h;
}

extern "C" int flid_dense(const flid_dense_weight* h, const float* a0, const int32_t* idx0, int64_t lda0, int w0, const float* a1,
                          const int32_t* idx1, int64_t lda1, int w1, const float* bias, const float* resid, int64_t ldr, int act,
                          float* c, int64_t ldc, int64_t m, flid_stream stream) {
    if (m <= 0) return FLID_OK;
    FLID_REQUIRE(h && a0 && c, "flid_dense: null argument");
    FLID_REQUIRE(act >= 0 && act <= 2, "flid_dense: act must be 0 (none), 1 (ReLU) or 2 (GELU)");
    FLID_REQUIRE(w1 == 0 || a1 != nullptr, "flid_dense: second segment without data");
    TcGemmArgs g;
    g.A0 = a0, g.idx0 = idx0, g.lda0 = lda0, g.w0 = w0;
    g.A1 = a1, g.idx1 = idx1, g.lda1 = lda1, g.w1 = w1;
    g.C = c, g.ldc = ldc, g.bias = bias, g.M = m;
    g.relu = act == 1, g.gelu = act == 2;
    g.resid = resid, g.ldr = ldr;
    return tc_gemm(g, h->w, (cudaStream_t)stream);
}

extern "C" int flid_time_rows(const float* dt, const int64_t* ids, const float* w, const float* b, int time_dim, float* out, int64_t n,
                              flid_stream stream) {
    if (n <= 0) return FLID_OK;
    FLID_REQUIRE(dt && w && b && out && time_dim > 0, "flid_time_rows: bad argument");
    time_rows_kernel<<<(unsigned)ceil_div(n * time_dim, 256), 256, 0, (cudaStream_t)stream>>>(dt, ids, w, b, time_dim, out, n);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

extern "C" int flid_token_mix(const float* x, int num_tokens, int channels, const float* ln_w, const float* ln_b, float eps,
                              const float* w1, const float* b1, const float* w2, const float* b2, int hidden, float* out, int64_t m,
                              flid_stream stream) {
    if (m <= 0) return FLID_OK;
    FLID_REQUIRE(x && ln_w && ln_b && w1 && b1 && w2 && b2 && out, "flid_token_mix: null argument");
    FLID_REQUIRE(num_tokens > 0 && num_tokens <= 64 && hidden > 0 && hidden <= 256 && channels > 0,
                 "flid_token_mix: supports 1..64 tokens and 1..256 hidden units (got %d, %d)", num_tokens, hidden);
    const size_t smem = sizeof(float) * (2 * (size_t)hidden * num_tokens + hidden + 3 * num_tokens);
    int dev = 0, sms = 0;
    FLID_CUDA(cudaGetDevice(&dev));
    FLID_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = ceil_div(m * channels, 128);
    const unsigned blocks = (unsigned)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);   // weights staged once per CTA
    cudaStream_t st = (cudaStream_t)stream;
#define FLID_TOKEN_MIX(KM, EX)                                                                                                     \
    do {                                                                                                                           \
        if (smem > 48 * 1024)                                                                                                      \
            FLID_CUDA(cudaFuncSetAttribute(token_mix_kernel<KM, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        token_mix_kernel<KM, EX><<<blocks, 128, smem, st>>>(x, num_tokens, channels, ln_w, ln_b, eps, w1, b1, w2, b2, hidden, out, m); \
    } while (0)
    if (num_tokens == 20) FLID_TOKEN_MIX(20, true);        // the reference's default num_neighbors
    else if (num_tokens == 10) FLID_TOKEN_MIX(10, true);
    else if (num_tokens == 32) FLID_TOKEN_MIX(32, true);
    else if (num_tokens <= 32) FLID_TOKEN_MIX(32, false);
    else FLID_TOKEN_MIX(64, false);
#undef FLID_TOKEN_MIX
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

extern "C" int flid_token_mean(const float* x, int num_tokens, int channels, float* out, int64_t ldo, int64_t m, flid_stream stream) {
    if (m <= 0) return FLID_OK;
    FLID_REQUIRE(x && out && num_tokens > 0 && channels > 0 && ldo >= channels, "flid_token_mean: bad argument");
    token_mean_kernel<<<(unsigned)ceil_div(m * channels, 256), 256, 0, (cudaStream_t)stream>>>(x, num_tokens, channels, out, ldo, m);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

extern "C" int flid_row_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, float* y, int64_t ldy,
                                  int64_t m, int dim, flid_stream stream) {
    if (m <= 0) return FLID_OK;
    FLID_REQUIRE(x && gamma && beta && y, "flid_row_layernorm: null argument");
    FLID_REQUIRE(dim > 0 && (dim & 3) == 0 && dim <= 128 * LN_C && (ldx & 3) == 0 && (ldy & 3) == 0,
                 "flid_row_layernorm: dim must be a multiple of 4, <= %d, rows 16-byte aligned", 128 * LN_C);
    row_layernorm_kernel<<<(unsigned)ceil_div(m * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, eps, y, ldy, m, dim);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

extern "C" int flid_add_periodic_rows(float* x, const float* table, int period, int dim, int64_t m, flid_stream stream) {
    if (m <= 0) return FLID_OK;
    FLID_REQUIRE(x && table && period > 0 && dim > 0 && (dim & 3) == 0, "flid_add_periodic_rows: bad argument");
    add_periodic_kernel<<<(unsigned)ceil_div(m * (dim >> 2), 256), 256, 0, (cudaStream_t)stream>>>(x, table, period, dim, m);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

extern "C" int flid_seq_attention(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                  const int64_t* key_ids, int seq_len, int num_heads, int head_dim, float* out, int64_t ldo,
                                  int q_rows, int64_t num_seqs, flid_stream stream) {
    if (num_seqs <= 0) return FLID_OK;
    FLID_REQUIRE(q && k && v && key_ids && out, "flid_seq_attention: null argument");
    const int d = num_heads * head_dim;
    FLID_REQUIRE(seq_len > 0 && num_heads > 0 && head_dim > 0 && (d & 3) == 0 && (ldq & 3) == 0 && (ldk & 3) == 0 && (ldv & 3) == 0,
                 "flid_seq_attention: model dim and row strides must be multiples of 4 floats");
    FLID_REQUIRE(q_rows > 0 && q_rows <= seq_len, "flid_seq_attention: q_rows must be in 1..seq_len");
    const size_t smem = sizeof(float) * (3 * (size_t)seq_len * (d + 1) + (size_t)num_heads * seq_len * (seq_len + 1)) + sizeof(int) * seq_len;
    int dev = 0, smem_max = 0;
    FLID_CUDA(cudaGetDevice(&dev));
    FLID_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    FLID_REQUIRE(smem <= (size_t)smem_max, "flid_seq_attention: sequence of %d tokens x %d does not fit shared memory", seq_len, d);
    if (smem > 48 * 1024)
        FLID_CUDA(cudaFuncSetAttribute(seq_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLID_REQUIRE(num_seqs <= 0x7fffffff, "flid_seq_attention: too many sequences in one call");
    seq_attention_kernel<<<(unsigned)num_seqs, 128, smem, (cudaStream_t)stream>>>(q, ldq, k, ldk, v, ldv, key_ids, seq_len, num_heads,
                                                                                  head_dim, out, ldo, q_rows);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}
