// One MultiHeadAttention + MergeLayer evaluation in training mode, forward and backward, as two
// C-ABI calls (models/modules.py:167-245 and :58-69 under autograd, as the M-step batches of
// PTCL/M_step.py:196-325 run them).  Same re-association as the inference path:
//
//   u   = q . Mq^T                     Mq = scale * Wk_h^T Wq_h stacked over heads   [H*kd, qd]
//   z   = attention stream (attn_train.cu: gather, time encoding, masked softmax, score dropout, weighted sum)
//   pre = z . Fo^T + b_res             Fo = residual_fc.weight . blockdiag(Wv_h)     [qd, H*kd]
//   y   = dropout(pre) + q ;  ln = LayerNorm(y)
//   hid = relu([ln | merge_self] . W1^T + b1) ;  out = hid . W2^T + b2
//
// The caller (flid_b200/train.py) builds Mq and Fo from the module parameters with differentiable torch
// ops, so this file returns the gradients of the FOLDED matrices and autograd carries them to
// query / key / value_projection and residual_fc.  Forward and data-gradient products run on the
// tcgen05 3xTF32 GEMM (gemm_tc.cu; the transposed weights are re-tiled per call); the
// weight-gradient products (reduction over the n rows) on a split-row fp32 kernel below.
#include <stdlib.h>

#include "attn_train.cuh"
#include "gemm_tc.cuh"
#include "graph.cuh"

namespace flid {
namespace {

constexpr int MAXQ4 = 4;  // float4 per lane of one [qd] row: qd <= 512

__device__ __forceinline__ uint4 philox_out(uint4 c, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
        key.x += 0x9E3779B9u, key.y += 0xBB67AE85u;
    }
    return c;
}
// multipliers (0 or 1 / (1 - p)) of the residual_fc-output dropout (modules.py:235) for columns 4f .. 4f+3 of row i
__device__ __forceinline__ float4 out_keep(uint64_t seed, int64_t i, int f, unsigned thr, float scale) {
    if (thr == 0u) return make_float4(1.f, 1.f, 1.f, 1.f);
    const uint4 r = philox_out(make_uint4((unsigned)i, (unsigned)((uint64_t)i >> 32), (unsigned)f, 0x4F757444u),
                               make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    return make_float4(r.x >= thr ? scale : 0.f, r.y >= thr ? scale : 0.f, r.z >= thr ? scale : 0.f,
                       r.w >= thr ? scale : 0.f);
}
inline unsigned threshold_of(float p) {
    if (!(p > 0.f)) return 0u;
    const double t = (double)p * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (unsigned)t;
}

// y = dropout(pre) + q ;  ln = LayerNorm(y) * gamma + beta.  One warp per row.
__global__ void __launch_bounds__(256) ln_train_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ q,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* __restrict__ y, float* __restrict__ ln, int64_t n, int qd,
                                                           uint64_t seed, unsigned thr, float scale) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int q4 = qd >> 2;
    float4 x[MAXQ4];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < MAXQ4; ++v) {
        const int f = lane + 32 * v;
        x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < q4) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(pre + i * qd) + f);
            const float4 r = __ldg(reinterpret_cast<const float4*>(q + i * qd) + f);
            const float4 m = out_keep(seed, i, f, thr, scale);
            x[v] = make_float4(fmaf(p.x, m.x, r.x), fmaf(p.y, m.y, r.y), fmaf(p.z, m.z, r.z), fmaf(p.w, m.w, r.w));
            reinterpret_cast<float4*>(y + i * qd)[f] = x[v];
            s += (x[v].x + x[v].y) + (x[v].z + x[v].w);
        }
    }
    const float mean = warp_sum(s) / (float)qd;
    float ss = 0.f;
#pragma unroll
    for (int v = 0; v < MAXQ4; ++v)
        if (lane + 32 * v < q4) {
            const float a = x[v].x - mean, b = x[v].y - mean, c = x[v].z - mean, d = x[v].w - mean;
            ss = fmaf(a, a, ss), ss = fmaf(b, b, ss), ss = fmaf(c, c, ss), ss = fmaf(d, d, ss);
        }
    const float rstd = 1.0f / sqrtf(warp_sum(ss) / (float)qd + 1e-5f);
#pragma unroll
    for (int v = 0; v < MAXQ4; ++v) {
        const int f = lane + 32 * v;
        if (f < q4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + f), b = __ldg(reinterpret_cast<const float4*>(beta) + f);
            reinterpret_cast<float4*>(ln + i * qd)[f] =
                make_float4(fmaf((x[v].x - mean) * rstd, g.x, b.x), fmaf((x[v].y - mean) * rstd, g.y, b.y),
                            fmaf((x[v].z - mean) * rstd, g.z, b.z), fmaf((x[v].w - mean) * rstd, g.w, b.w));
        }
    }
}

// LayerNorm backward from the saved pre-norm rows y: d_y (the residual / query gradient), d_pre = d_y * dropout
// multiplier, and block-accumulated d_gamma / d_beta (atomic adds into zero-initialised vectors).
__global__ void __launch_bounds__(256) ln_train_bwd_kernel(const float* __restrict__ y, const float* __restrict__ d_ln,
                                                           int64_t ld_dln, const float* __restrict__ gamma,
                                                           float* __restrict__ d_y, float* __restrict__ d_pre,
                                                           float* __restrict__ d_gamma, float* __restrict__ d_beta, int64_t n,
                                                           int qd, uint64_t seed, unsigned thr, float scale) {
    __shared__ float red[8][2 * 128 * MAXQ4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q4 = qd >> 2;
    float4 gsum[MAXQ4], bsum[MAXQ4], gm[MAXQ4];
#pragma unroll
    for (int v = 0; v < MAXQ4; ++v) {
        gsum[v] = bsum[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int f = lane + 32 * v;
        gm[v] = f < q4 ? __ldg(reinterpret_cast<const float4*>(gamma) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t warps = (int64_t)gridDim.x * 8;
    for (int64_t i = (int64_t)blockIdx.x * 8 + warp; i < n; i += warps) {
        float4 x[MAXQ4], g[MAXQ4];
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < MAXQ4; ++v) {
            const int f = lane + 32 * v;
            x[v] = g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < q4) {
                x[v] = __ldg(reinterpret_cast<const float4*>(y + i * qd) + f);
                g[v] = __ldg(reinterpret_cast<const float4*>(d_ln + i * ld_dln) + f);
                s += (x[v].x + x[v].y) + (x[v].z + x[v].w);
            }
        }
        const float mean = warp_sum(s) / (float)qd;
        float ss = 0.f;
#pragma unroll
        for (int v = 0; v < MAXQ4; ++v)
            if (lane + 32 * v < q4) {
                x[v].x -= mean, x[v].y -= mean, x[v].z -= mean, x[v].w -= mean;
                ss = fmaf(x[v].x, x[v].x, ss), ss = fmaf(x[v].y, x[v].y, ss), ss = fmaf(x[v].z, x[v].z, ss),
                ss = fmaf(x[v].w, x[v].w, ss);
            }
        const float rstd = 1.0f / sqrtf(warp_sum(ss) / (float)qd + 1e-5f);
        float s1 = 0.f, s2 = 0.f;  // sum(g gamma), sum(g gamma xhat)
#pragma unroll
        for (int v = 0; v < MAXQ4; ++v)
            if (lane + 32 * v < q4) {
                x[v].x *= rstd, x[v].y *= rstd, x[v].z *= rstd, x[v].w *= rstd;  // xhat
                gsum[v].x = fmaf(g[v].x, x[v].x, gsum[v].x), gsum[v].y = fmaf(g[v].y, x[v].y, gsum[v].y);
                gsum[v].z = fmaf(g[v].z, x[v].z, gsum[v].z), gsum[v].w = fmaf(g[v].w, x[v].w, gsum[v].w);
                bsum[v].x += g[v].x, bsum[v].y += g[v].y, bsum[v].z += g[v].z, bsum[v].w += g[v].w;
                g[v].x *= gm[v].x, g[v].y *= gm[v].y, g[v].z *= gm[v].z, g[v].w *= gm[v].w;
                s1 += (g[v].x + g[v].y) + (g[v].z + g[v].w);
                s2 = fmaf(g[v].x, x[v].x, s2), s2 = fmaf(g[v].y, x[v].y, s2), s2 = fmaf(g[v].z, x[v].z, s2),
                s2 = fmaf(g[v].w, x[v].w, s2);
            }
        s1 = warp_sum(s1) / (float)qd, s2 = warp_sum(s2) / (float)qd;
#pragma unroll
        for (int v = 0; v < MAXQ4; ++v) {
            const int f = lane + 32 * v;
            if (f < q4) {
                float4 d;
                d.x = rstd * (g[v].x - s1 - x[v].x * s2), d.y = rstd * (g[v].y - s1 - x[v].y * s2);
                d.z = rstd * (g[v].z - s1 - x[v].z * s2), d.w = rstd * (g[v].w - s1 - x[v].w * s2);
                reinterpret_cast<float4*>(d_y + i * qd)[f] = d;
                const float4 m = out_keep(seed, i, f, thr, scale);
                reinterpret_cast<float4*>(d_pre + i * qd)[f] = make_float4(d.x * m.x, d.y * m.y, d.z * m.z, d.w * m.w);
            }
        }
    }
    // block reduction of the column sums, then one atomic per column and block
    float* mine = red[warp];
#pragma unroll
    for (int v = 0; v < MAXQ4; ++v) {
        const int f = lane + 32 * v;
        reinterpret_cast<float4*>(mine)[f] = gsum[v];
        reinterpret_cast<float4*>(mine + 128 * MAXQ4)[f] = bsum[v];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * qd; c += blockDim.x) {
        const int col = c < qd ? c : c - qd, off = c < qd ? col : 128 * MAXQ4 + col;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][off];
        atomicAdd((c < qd ? d_gamma : d_beta) + col, t);
    }
}

__global__ void out_keep_kernel(uint64_t seed, int64_t n, int qd, unsigned thr, uint8_t* __restrict__ keep) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int q4 = qd >> 2;
    if (t >= n * q4) return;
    const int64_t i = t / q4;
    const int f = (int)(t % q4);
    const float4 m = out_keep(seed, i, f, thr, 1.f);
    uint8_t* o = keep + i * qd + 4 * f;
    o[0] = m.x != 0.f, o[1] = m.y != 0.f, o[2] = m.z != 0.f, o[3] = m.w != 0.f;
}

// One level of the top-down sampling of models/TGAT.py:68-144, level-batched: the n targets of this level are
// sampled (utils/utils.py:149-214, 'recent'), their time differences formed with the reference's dtype rules
// (TGAT.py:120-125: float64 root time minus the float32 neighbour time, rounded once; float32 minus float32 in the
// recursion), and the next level's targets [these n targets ; their n*k neighbours] written out.
__global__ void __launch_bounds__(256) train_level_kernel(const int64_t* __restrict__ indptr, const int2* __restrict__ adj,
                                                          const double* __restrict__ ts, const int64_t* __restrict__ ids,
                                                          const double* __restrict__ t64, int64_t n, int64_t n_f64, int k,
                                                          int64_t* __restrict__ nbr, int64_t* __restrict__ eid,
                                                          float* __restrict__ dt, int64_t* __restrict__ next_ids,
                                                          double* __restrict__ next_t64) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= n) return;
    const int64_t v = __ldg(ids + q);
    const double t = __ldg(t64 + q);
    const int64_t start = __ldg(indptr + v);
    const int64_t cut = warp_lower_bound(ts, start, __ldg(indptr + v + 1), t, lane);
    const int64_t have = cut - start;
    const int cnt = have < (int64_t)k ? (int)have : k;
    if (next_ids && lane == 0) next_ids[q] = v, next_t64[q] = t;
    for (int j = lane; j < k; j += 32) {
        int64_t a = 0, e = 0;
        float tn = 0.f;
        if (j >= k - cnt) {
            const int64_t p = cut - k + j;
            const int2 ne = __ldg(adj + p);
            a = ne.x, e = ne.y, tn = (float)__ldg(ts + p);
        }
        nbr[q * k + j] = a, eid[q * k + j] = e;
        dt[q * k + j] = q < n_f64 ? (float)(t - (double)tn) : (float)t - tn;
        if (next_ids) next_ids[n + q * k + j] = a, next_t64[n + q * k + j] = (double)tn;
    }
}

// q[i] = [h row | te0] and merge_self[i] = node_feat[ids[i]]: the query / residual and MergeLayer's second input
// (models/TGAT.py:88-92, :137-141).  h row = h_src[h_by_id ? ids[i] : i].  One warp per target.
__global__ void __launch_bounds__(256) build_inputs_kernel(const float* __restrict__ h_src, int h_by_id,
                                                           const float* __restrict__ node_feat, const int64_t* __restrict__ ids,
                                                           const float* __restrict__ te0, float* __restrict__ q,
                                                           float* __restrict__ merge_self, int64_t n, int dn, int T) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int64_t id = __ldg(ids + i);
    const float4* h = reinterpret_cast<const float4*>(h_src + (h_by_id ? id : i) * dn);
    const float4* raw = reinterpret_cast<const float4*>(node_feat + id * dn);
    const int qd = dn + T;
    for (int f = lane; f < dn / 4; f += 32) {
        reinterpret_cast<float4*>(q + i * qd)[f] = __ldg(h + f);
        reinterpret_cast<float4*>(merge_self + i * dn)[f] = __ldg(raw + f);
    }
    for (int f = lane; f < T / 4; f += 32)
        reinterpret_cast<float4*>(q + i * qd + dn)[f] = __ldg(reinterpret_cast<const float4*>(te0) + f);
}

// dst[ids ? ids[i] : i, :dn] += src[i, :dn]  (atomic when rows may repeat)
__global__ void __launch_bounds__(256) add_rows_kernel(const float* __restrict__ src, int64_t ld_src,
                                                       const int64_t* __restrict__ ids, float* __restrict__ dst, int64_t n,
                                                       int dn) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const float4* s = reinterpret_cast<const float4*>(src + i * ld_src);
    float4* d = reinterpret_cast<float4*>(dst + (ids ? __ldg(ids + i) : i) * dn);
    for (int f = lane; f < dn / 4; f += 32) {
        const float4 v = __ldg(s + f);
        if (ids) {
            atomicAdd(d + f, v);
        } else {
            float4 o = d[f];
            o.x += v.x, o.y += v.y, o.z += v.z, o.w += v.w;
            d[f] = o;
        }
    }
}

// d_hid *= (hid > 0), in place (ReLU backward)
__global__ void relu_mask_kernel(float4* __restrict__ d, const float4* __restrict__ h, int64_t n4) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 g = d[i];
    const float4 a = __ldg(h + i);
    g.x = a.x > 0.f ? g.x : 0.f, g.y = a.y > 0.f ? g.y : 0.f, g.z = a.z > 0.f ? g.z : 0.f, g.w = a.w > 0.f ? g.w : 0.f;
    d[i] = g;
}
__global__ void add_kernel(float4* __restrict__ d, const float4* __restrict__ s, int64_t n4) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = d[i];
    const float4 b = __ldg(s + i);
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    d[i] = a;
}

// out[c] += sum_m X[m, c]   (bias gradients, per-block partial sums)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t ld, int64_t n, int cols,
                                                     int64_t rows_per_slab, float* __restrict__ out) {
    __shared__ float red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int64_t lo = blockIdx.y * rows_per_slab, hi = min(n, lo + rows_per_slab);
    float s = 0.f;
    if (c < cols)
        for (int64_t m = lo + ty; m < hi; m += 8) s += __ldg(X + m * ld + c);
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][tx];
        atomicAdd(out + c, t);
    }
}

// dW[i, j] += sum_m Y[m, i] X[m, j].  A block owns a (16 RI) x (16 RJ) tile of dW and one slab of rows
// (blockIdx.z); each of its 256 threads accumulates RI x RJ outputs in registers from two shared-memory
// row tiles, then adds them to dW atomically.  RI x RJ = 8 x 4 / 4 x 8 keeps the FFMA pipe, not the
// shared-memory port, the limiter (32 FMAs per three 16-byte loads).
constexpr int WG_R = 32;
template <int RI, int RJ>
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ Y, int64_t ldy, int NI,
                                                    const float* __restrict__ X, int64_t ldx, int NJ,
                                                    float* __restrict__ dW, int64_t ldw, int64_t n, int64_t rows_per_slab) {
    constexpr int TI = 16 * RI, TJ = 16 * RJ;
    __shared__ __align__(16) float Ys[WG_R][TI + 4];
    __shared__ __align__(16) float Xs[WG_R][TJ + 4];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int i0 = blockIdx.x * TI, j0 = blockIdx.y * TJ;
    const int64_t lo = blockIdx.z * rows_per_slab, hi = min(n, lo + rows_per_slab);
    float acc[RI][RJ];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < RJ; ++j) acc[i][j] = 0.f;
    // register prefetch: the next row tile is in flight while this one is multiplied out of shared memory
    constexpr int LY = WG_R * TI / 4 / 256, LX = WG_R * TJ / 4 / 256;
    float4 py[LY], px[LX];
    auto fetch = [&](int64_t m0) {
#pragma unroll
        for (int u = 0; u < LY; ++u) {
            const int v = t + 256 * u, r = v / (TI / 4), c4 = (v % (TI / 4)) * 4;
            py[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < hi && i0 + c4 < NI) py[u] = __ldg(reinterpret_cast<const float4*>(Y + (m0 + r) * ldy + i0 + c4));
        }
#pragma unroll
        for (int u = 0; u < LX; ++u) {
            const int v = t + 256 * u, r = v / (TJ / 4), c4 = (v % (TJ / 4)) * 4;
            px[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < hi && j0 + c4 < NJ) px[u] = __ldg(reinterpret_cast<const float4*>(X + (m0 + r) * ldx + j0 + c4));
        }
    };
    if (lo < hi) fetch(lo);
    for (int64_t m0 = lo; m0 < hi; m0 += WG_R) {
#pragma unroll
        for (int u = 0; u < LY; ++u) {
            const int v = t + 256 * u;
            *reinterpret_cast<float4*>(&Ys[v / (TI / 4)][(v % (TI / 4)) * 4]) = py[u];
        }
#pragma unroll
        for (int u = 0; u < LX; ++u) {
            const int v = t + 256 * u;
            *reinterpret_cast<float4*>(&Xs[v / (TJ / 4)][(v % (TJ / 4)) * 4]) = px[u];
        }
        __syncthreads();
        if (m0 + WG_R < hi) fetch(m0 + WG_R);
#pragma unroll 4
        for (int r = 0; r < WG_R; ++r) {
            float av[RI], bv[RJ];
#pragma unroll
            for (int i = 0; i < RI; i += 4) {
                const float4 a = *reinterpret_cast<const float4*>(&Ys[r][ty * RI + i]);
                av[i] = a.x, av[i + 1] = a.y, av[i + 2] = a.z, av[i + 3] = a.w;
            }
#pragma unroll
            for (int j = 0; j < RJ; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(&Xs[r][(j / 4) * 64 + tx * 4]);  // 16-byte lane stride
                bv[j] = b.x, bv[j + 1] = b.y, bv[j + 2] = b.z, bv[j + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < RI; ++i)
#pragma unroll
                for (int j = 0; j < RJ; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < RJ; ++j) {
            const int gi = i0 + ty * RI + i, gj = j0 + (j / 4) * 64 + tx * 4 + (j % 4);
            if (gi < NI && gj < NJ) atomicAdd(dW + (int64_t)gi * ldw + gj, acc[i][j]);
        }
}

template <int RI, int RJ>
int wgrad_launch(const float* Y, int64_t ldy, int NI, const float* X, int64_t ldx, int NJ, float* dW, int64_t ldw, int64_t n,
                 cudaStream_t st) {
    const int ti = (int)ceil_div(NI, 16 * RI), tj = (int)ceil_div(NJ, 16 * RJ);
    // target blocks per SM: measured 4 best at n ~ 8 800 rows (one B = 200 batch), 8 at n ~ 88 000; FLID_WGRAD_BLOCKS overrides
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("FLID_WGRAD_BLOCKS");
        forced = (e && atoi(e) > 0) ? atoi(e) : 0;
    }
    const int per_sm = forced ? forced : (n >= 32768 ? 8 : 4);
    int64_t slabs = ceil_div((int64_t)per_sm * 148, (int64_t)ti * tj);
    const int64_t max_slabs = ceil_div(n, 2 * WG_R);
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    const int64_t rows = ceil_div(ceil_div(n, slabs), WG_R) * WG_R;
    wgrad_kernel<RI, RJ><<<dim3(ti, tj, (unsigned)ceil_div(n, rows)), 256, 0, st>>>(Y, ldy, NI, X, ldx, NJ, dW, ldw, n, rows);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

int wgrad(const float* Y, int64_t ldy, int NI, const float* X, int64_t ldx, int NJ, float* dW, int64_t ldw, int64_t n,
          cudaStream_t st) {
    // the long side of the tile goes to the larger dimension; small square outputs keep 64 x 64
    if (NI >= 256 && NI >= NJ) return wgrad_launch<8, 4>(Y, ldy, NI, X, ldx, NJ, dW, ldw, n, st);
    if (NJ >= 256) return wgrad_launch<4, 8>(Y, ldy, NI, X, ldx, NJ, dW, ldw, n, st);
    return wgrad_launch<4, 4>(Y, ldy, NI, X, ldx, NJ, dW, ldw, n, st);
}

int colsum(const float* X, int64_t ld, int64_t n, int cols, float* out, cudaStream_t st) {
    const int cb = (int)ceil_div(cols, 32);
    int64_t slabs = ceil_div(2 * 148, cb);
    const int64_t max_slabs = ceil_div(n, 64);
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    const int64_t rows = ceil_div(n, slabs);
    colsum_kernel<<<dim3(cb, (unsigned)ceil_div(n, rows)), 256, 0, st>>>(X, ld, n, cols, rows, out);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

int gemm(const float* A0, int64_t lda0, int w0, const float* A1, int64_t lda1, int w1, const TcWeight& w, const float* bias,
         int relu, float* C, int64_t ldc, int64_t n, cudaStream_t st) {
    TcGemmArgs g;
    g.A0 = A0, g.lda0 = lda0, g.w0 = w0, g.A1 = A1, g.lda1 = lda1, g.w1 = w1;
    g.C = C, g.ldc = ldc, g.bias = bias, g.M = n, g.relu = relu;
    return tc_gemm(g, w, st);
}

// tiled weight images, rebuilt at the start of every call (the parameters change between optimizer steps)
struct Images {
    TcWeight q, o, f1, f2;
};
// one set per device: the images live in that device's memory
constexpr int MAX_DEVICES = 32;
Images g_fwd_dev[MAX_DEVICES], g_bwd_dev[MAX_DEVICES];
int current_device(int* dev) {
    FLID_CUDA(cudaGetDevice(dev));
    FLID_REQUIRE(*dev >= 0 && *dev < MAX_DEVICES, "train layer: device ordinal %d not supported", *dev);
    return FLID_OK;
}

struct Dims {
    int64_t n;
    int k, H, dn, de, T, kd, qd, zw;
};
int make_dims(int64_t n, int k, int H, int dn, int de, int T, Dims* d) {
    FLID_REQUIRE(n >= 0 && dn > 0 && de > 0 && T > 0, "train layer: bad shape");
    FLID_REQUIRE(dn % 4 == 0 && de % 4 == 0 && T % 4 == 0, "train layer: feature widths must be multiples of 4");
    FLID_REQUIRE(dn + T <= 128 * MAXQ4, "train layer: node_dim + time_dim must be <= %d", 128 * MAXQ4);
    *d = Dims{n, k, H, dn, de, T, dn + de + T, dn + T, H * (dn + de + T)};
    return FLID_OK;
}

}  // namespace
}  // namespace flid

using namespace flid;

extern "C" int flid_train_sample_levels(const flid_graph* g, const int64_t* roots, const double* times,
                                        int times_are_f32, int64_t n, int k, int num_levels, int64_t* const* ids_host,
                                        double* const* t64_host, int64_t* const* nbr_host, int64_t* const* eid_host,
                                        float* const* dt_host, flid_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FLID_REQUIRE(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    FLID_REQUIRE(num_levels >= 1 && n >= 0, "flid_train_sample_levels: bad shape");
    if (n == 0) return FLID_OK;
    FLID_REQUIRE(g && roots && times && ids_host && t64_host && nbr_host && eid_host && dt_host, "flid_train_sample_levels: null argument");
    const int top = num_levels - 1;
    FLID_CUDA(cudaMemcpyAsync(ids_host[top], roots, n * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    FLID_CUDA(cudaMemcpyAsync(t64_host[top], times, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    int64_t nl = n;
    for (int l = top; l >= 0; --l) {
        train_level_kernel<<<(unsigned)ceil_div(nl * 32, 256), 256, 0, st>>>(
            g->indptr, g->adj, g->ts, ids_host[l], t64_host[l], nl, times_are_f32 ? 0 : n, k, nbr_host[l], eid_host[l],
            dt_host[l], l > 0 ? ids_host[l - 1] : nullptr, l > 0 ? t64_host[l - 1] : nullptr);
        FLID_LAUNCH_CHECK();
        nl *= (1 + k);
    }
    return FLID_OK;
}

extern "C" int flid_train_layer_out_keep_mask(uint64_t seed, int64_t n, int qd, float p_drop, uint8_t* keep,
                                              flid_stream stream) {
    FLID_REQUIRE(n >= 0 && qd > 0 && qd % 4 == 0 && keep, "out_keep_mask: bad argument");
    if (n == 0) return FLID_OK;
    out_keep_kernel<<<(unsigned)ceil_div(n * (qd / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        seed ^ 0x9E3779B97F4A7C15ull, n, qd, threshold_of(p_drop), keep);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

static int layer_fwd(const flid_train_weights* w, const float* q, const float* merge_self, const float* table,
                     const int64_t* hrow, int64_t hrow_offset, const int64_t* nbr, const int64_t* eid, const float* dt,
                     const float* edge_feat, const Dims& d, float p_drop, uint64_t seed, const flid_train_saved* s,
                     float* pre_scratch, float* out, cudaStream_t st) {
    const int64_t n = d.n;
    const int k = d.k, num_heads = d.H;
    int dev = 0;
    FLID_TRY(current_device(&dev));
    Images& g_fwd = g_fwd_dev[dev];
    FLID_TRY(tc_prepare_weight(w->fold_q, d.qd, d.zw, d.qd, &g_fwd.q, st));
    FLID_TRY(tc_prepare_weight(w->fold_o, d.zw, d.qd, d.zw, &g_fwd.o, st));
    FLID_TRY(tc_prepare_weight(w->fc1_w, d.qd + d.dn, d.dn, d.qd + d.dn, &g_fwd.f1, st));
    FLID_TRY(tc_prepare_weight(w->fc2_w, d.dn, d.dn, d.dn, &g_fwd.f2, st));
    // u = q . Mq^T
    FLID_TRY(gemm(q, d.qd, d.qd, nullptr, 0, 0, g_fwd.q, nullptr, 0, s->u, d.zw, n, st));
    AttnTrainArgs a{s->u, table, hrow, nbr, eid, dt, edge_feat, w->time_w, w->time_b, n, k, d.dn, d.de, d.T, p_drop, seed, s->z, s->probs, hrow_offset};
    FLID_TRY(launch_attn_train_fwd(a, num_heads, st));
    // pre = z . Fo^T + b_res ; y = dropout(pre) + q ; ln = LayerNorm(y)
    FLID_TRY(gemm(s->z, d.zw, d.zw, nullptr, 0, 0, g_fwd.o, w->res_b, 0, pre_scratch, d.qd, n, st));
    ln_train_fwd_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(pre_scratch, q, w->ln_w, w->ln_b, s->y, s->ln, n, d.qd,
                                                                         seed ^ 0x9E3779B97F4A7C15ull, threshold_of(p_drop),
                                                                         1.0f / (1.0f - p_drop));
    FLID_LAUNCH_CHECK();
    // MergeLayer
    FLID_TRY(gemm(s->ln, d.qd, d.qd, merge_self, d.dn, d.dn, g_fwd.f1, w->fc1_b, 1, s->hid, d.dn, n, st));
    FLID_TRY(gemm(s->hid, d.dn, d.dn, nullptr, 0, 0, g_fwd.f2, w->fc2_b, 0, out, d.dn, n, st));
    return FLID_OK;
}

static int layer_bwd(const flid_train_weights* w, const float* q, const float* merge_self, const float* table,
                     const int64_t* hrow, int64_t hrow_offset, const int64_t* nbr, const int64_t* eid, const float* dt,
                     const float* edge_feat, const Dims& d, float p_drop, uint64_t seed, const flid_train_saved* s,
                     const float* d_out, float* d_q, float* d_cat, float* d_table, const flid_train_grads* g,
                     float* scratch, cudaStream_t st) {
    const int64_t n = d.n;
    const int k = d.k, num_heads = d.H;
    int dev = 0;
    FLID_TRY(current_device(&dev));
    Images& g_bwd = g_bwd_dev[dev];
    float* d_hid = scratch;
    float* d_pre = d_hid + n * d.dn;
    float* dz = d_pre + n * d.qd;
    float* du = dz + n * d.zw;
    float* tpart = du + n * d.zw;
    const int cw = d.qd + d.dn;
    // transposed images for the data-gradient products
    FLID_TRY(tc_prepare_weight_t(w->fc2_w, d.dn, d.dn, d.dn, &g_bwd.f2, st));   // [in = dn, out = dn]
    FLID_TRY(tc_prepare_weight_t(w->fc1_w, cw, cw, d.dn, &g_bwd.f1, st));       // [in = qd + dn, out = dn]
    FLID_TRY(tc_prepare_weight_t(w->fold_o, d.zw, d.zw, d.qd, &g_bwd.o, st));   // [zw, qd]
    FLID_TRY(tc_prepare_weight_t(w->fold_q, d.qd, d.qd, d.zw, &g_bwd.q, st));   // [qd, zw]
    // ---- MergeLayer
    FLID_TRY(wgrad(d_out, d.dn, d.dn, s->hid, d.dn, d.dn, g->fc2_w, d.dn, n, st));
    FLID_TRY(colsum(d_out, d.dn, n, d.dn, g->fc2_b, st));
    FLID_TRY(gemm(d_out, d.dn, d.dn, nullptr, 0, 0, g_bwd.f2, nullptr, 0, d_hid, d.dn, n, st));
    relu_mask_kernel<<<(unsigned)ceil_div(n * d.dn / 4, 256), 256, 0, st>>>(reinterpret_cast<float4*>(d_hid),
                                                                           reinterpret_cast<const float4*>(s->hid), n * d.dn / 4);
    FLID_LAUNCH_CHECK();
    FLID_TRY(wgrad(d_hid, d.dn, d.dn, s->ln, d.qd, d.qd, g->fc1_w, cw, n, st));
    FLID_TRY(wgrad(d_hid, d.dn, d.dn, merge_self, d.dn, d.dn, g->fc1_w + d.qd, cw, n, st));
    FLID_TRY(colsum(d_hid, d.dn, n, d.dn, g->fc1_b, st));
    FLID_TRY(gemm(d_hid, d.dn, d.dn, nullptr, 0, 0, g_bwd.f1, nullptr, 0, d_cat, cw, n, st));   // [d_ln | d_merge_self]
    // ---- LayerNorm, residual, output dropout: d_q <- d_y for now
    {
        int64_t blocks = ceil_div(n, 8);
        if (blocks > 2 * 148) blocks = 2 * 148;
        ln_train_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(s->y, d_cat, cw, w->ln_w, d_q, d_pre, g->ln_w,
                                                              g->ln_b, n, d.qd,
                                                              seed ^ 0x9E3779B97F4A7C15ull, threshold_of(p_drop),
                                                              1.0f / (1.0f - p_drop));
        FLID_LAUNCH_CHECK();
    }
    // ---- folded out-projection
    FLID_TRY(wgrad(d_pre, d.qd, d.qd, s->z, d.zw, d.zw, g->fold_o, d.zw, n, st));
    FLID_TRY(colsum(d_pre, d.qd, n, d.qd, g->res_b, st));
    FLID_TRY(gemm(d_pre, d.qd, d.qd, nullptr, 0, 0, g_bwd.o, nullptr, 0, dz, d.zw, n, st));
    // ---- attention stream
    {
        AttnTrainArgs a{s->u, table, hrow, nbr, eid, dt, edge_feat, w->time_w, w->time_b, n, k, d.dn, d.de, d.T, p_drop, seed, nullptr, nullptr, hrow_offset};
        AttnTrainGrads ag{s->probs, dz, du, d_table, tpart};
        FLID_TRY(launch_attn_train_bwd(a, ag, num_heads, st));
        const int64_t blocks = attn_train_blocks(n);
        FLID_TRY(colsum(tpart, 2 * d.T, blocks, d.T, g->time_w, st));
        FLID_TRY(colsum(tpart + d.T, 2 * d.T, blocks, d.T, g->time_b, st));
    }
    // ---- query fold: d_Mq = du^T q ;  d_q = d_y + du . Mq
    FLID_TRY(wgrad(du, d.zw, d.zw, q, d.qd, d.qd, g->fold_q, d.qd, n, st));
    FLID_TRY(gemm(du, d.zw, d.zw, nullptr, 0, 0, g_bwd.q, nullptr, 0, d_pre, d.qd, n, st));   // d_pre is free again
    add_kernel<<<(unsigned)ceil_div(n * d.qd / 4, 256), 256, 0, st>>>(reinterpret_cast<float4*>(d_q),
                                                                     reinterpret_cast<const float4*>(d_pre), n * d.qd / 4);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

// floats of per-layer backward scratch for n targets: d_hid | d_pre | dz | du | time partials | d_q | d_cat
static int64_t layer_scratch(const Dims& d, int64_t n) {
    return n * (d.dn + d.qd + 2 * (int64_t)d.zw) + attn_train_blocks(n) * 2 * d.T + n * (2 * d.qd + d.dn) + 64;
}

extern "C" int64_t flid_train_model_scratch_floats(int64_t n_roots, int k, int num_layers, int num_heads, int node_dim,
                                                   int edge_dim, int time_dim) {
    Dims d;
    if (make_dims(n_roots, k, num_heads, node_dim, edge_dim, time_dim, &d) != FLID_OK || num_layers < 1) return -1;
    int64_t n1 = n_roots, dh = 0;
    for (int l = num_layers; l > 1; --l) n1 *= (1 + k), dh += n1 * node_dim;   // gradient buffers of layers 1 .. L-1
    return layer_scratch(d, n1) + dh;
}

extern "C" int flid_train_model_fwd(const flid_train_weights* w, const flid_train_level* lv, const flid_train_saved* sv,
                                    const float* node_feat, const float* edge_feat, const float* te0, int num_layers, int k,
                                    int num_heads, int node_dim, int edge_dim, int time_dim, float p_drop,
                                    const uint64_t* seeds_host, float* pre_scratch, flid_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FLID_REQUIRE(w && lv && sv && node_feat && edge_feat && te0 && seeds_host && pre_scratch && num_layers >= 1,
                 "flid_train_model_fwd: bad argument");
    FLID_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "train layer: dropout probability must be in [0, 1)");
    FLID_REQUIRE(k > 0 && k <= 32, "training path: num_neighbors must be in 1..32 (got %d)", k);
    for (int l = 1; l <= num_layers; ++l) {
        const flid_train_level& L = lv[l - 1];
        const flid_train_saved& S = sv[l - 1];
        Dims d;
        FLID_TRY(make_dims(L.n, k, num_heads, node_dim, edge_dim, time_dim, &d));
        if (L.n == 0) continue;
        const float* prev = l == 1 ? node_feat : sv[l - 2].out;
        build_inputs_kernel<<<(unsigned)ceil_div(L.n * 32, 256), 256, 0, st>>>(prev, l == 1, node_feat, L.ids, te0, S.q,
                                                                               S.merge_self, L.n, d.dn, d.T);
        FLID_LAUNCH_CHECK();
        FLID_TRY(layer_fwd(&w[l - 1], S.q, S.merge_self, prev, l == 1 ? L.nbr : nullptr, l == 1 ? 0 : L.n, L.nbr, L.eid, L.dt,
                           edge_feat, d, p_drop, seeds_host[l - 1], &S, pre_scratch, S.out, st));
    }
    return FLID_OK;
}

extern "C" int flid_train_model_bwd(const flid_train_weights* w, const flid_train_level* lv, const flid_train_saved* sv,
                                    const float* node_feat, const float* edge_feat, int num_layers, int k, int num_heads,
                                    int node_dim, int edge_dim, int time_dim, float p_drop, const uint64_t* seeds_host,
                                    const float* d_out, const flid_train_grads* g, float* d_te0, float* d_node_feat,
                                    float* scratch, flid_stream stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FLID_REQUIRE(w && lv && sv && node_feat && edge_feat && seeds_host && d_out && g && d_te0 && scratch && num_layers >= 1,
                 "flid_train_model_bwd: bad argument");
    Dims d1;
    FLID_TRY(make_dims(lv[0].n, k, num_heads, node_dim, edge_dim, time_dim, &d1));
    // gradient buffers of the intermediate layer outputs, zeroed: the attention backward accumulates into them
    float* dh[16] = {nullptr};
    FLID_REQUIRE(num_layers <= 16, "flid_train_model_bwd: too many layers");
    float* p = scratch + layer_scratch(d1, lv[0].n);
    for (int l = 1; l < num_layers; ++l) {
        dh[l] = p;
        p += lv[l - 1].n * node_dim;
        FLID_CUDA(cudaMemsetAsync(dh[l], 0, lv[l - 1].n * node_dim * sizeof(float), st));
    }
    const float* d_cur = d_out;
    for (int l = num_layers; l >= 1; --l) {
        const flid_train_level& L = lv[l - 1];
        const flid_train_saved& S = sv[l - 1];
        Dims d;
        FLID_TRY(make_dims(L.n, k, num_heads, node_dim, edge_dim, time_dim, &d));
        if (L.n == 0) continue;
        const int cw = d.qd + d.dn;
        float* d_q = scratch + layer_scratch(d, L.n) - 64 - L.n * (2 * d.qd + d.dn);
        float* d_cat = d_q + L.n * d.qd;
        const float* table = l == 1 ? node_feat : sv[l - 2].out;
        float* d_table = l == 1 ? d_node_feat : dh[l - 1];
        FLID_TRY(layer_bwd(&w[l - 1], S.q, S.merge_self, table, l == 1 ? L.nbr : nullptr, l == 1 ? 0 : L.n, L.nbr, L.eid,
                           L.dt, edge_feat, d, p_drop, seeds_host[l - 1], &S, d_cur, d_q, d_cat, d_table, &g[l - 1], scratch,
                           st));
        FLID_TRY(colsum(d_q + d.dn, d.qd, L.n, d.T, d_te0, st));                     // te0 = cos(time_b) part of the query
        const unsigned blocks = (unsigned)ceil_div(L.n * 32, 256);
        if (d_node_feat) {                                                             // MergeLayer's raw-feature input
            add_rows_kernel<<<blocks, 256, 0, st>>>(d_cat + d.qd, cw, L.ids, d_node_feat, L.n, d.dn);
            FLID_LAUNCH_CHECK();
        }
        if (l == 1) {
            if (d_node_feat) {
                add_rows_kernel<<<blocks, 256, 0, st>>>(d_q, d.qd, L.ids, d_node_feat, L.n, d.dn);
                FLID_LAUNCH_CHECK();
            }
        } else {
            add_rows_kernel<<<blocks, 256, 0, st>>>(d_q, d.qd, nullptr, dh[l - 1], L.n, d.dn);   // targets = rows [0, n)
            FLID_LAUNCH_CHECK();
            d_cur = dh[l - 1];
        }
    }
    return FLID_OK;
}
