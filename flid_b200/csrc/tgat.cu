// TGAT temporal-embedding path (models/TGAT.py:50-144, models/modules.py:7-69,126-245),
// eval mode, re-associated so that the per-neighbour work is a pure streaming pass:
//
//   scores_h[j] = scale * (Wq_h q)^T (Wk_h x_j)  =  x_j . u_h ,  u_h = (scale Wk_h^T Wq_h) q
//   out         = Wr [ Wv_h sum_j a_hj x_j ]_h + br = sum_h (Wr_h Wv_h) z_h + br , z_h = sum_j a_hj x_j
//
// with q = [h_self | te(0)], x_j = [h_nbr_j | e_j | te(dt_j)].  Per attention evaluation
// the kernels are: level_sample (index work) -> [u GEMM] -> attn (gather + cos time
// encoding + masked online softmax + weighted sum; HBM-bound) -> out GEMM -> LayerNorm
// -> MergeLayer GEMMs.  This file holds the fp32 reference-parity mode.
#include "tgat.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "attn.cuh"
#include "gemm.cuh"

namespace flid {

// ------------------------------------------------------------------ weight folding
__global__ void fold_qk_kernel(const float* __restrict__ wq, const float* __restrict__ wk, int qd, int kd, int H,
                               double scale, float* __restrict__ mfoldT) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)H * kd * qd) return;
    const int row = (int)(idx / qd), b = (int)(idx % qd), h = row / kd, a = row % kd, hd = qd / H;
    // four independent partial sums: the loop is bound by the latency of its loads, not by the float64 FMAs
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const float* pk = wk + (int64_t)(h * hd) * kd + a;
    const float* pq = wq + (int64_t)(h * hd) * qd + b;
    int r = 0;
#pragma unroll 2
    for (; r + 3 < hd; r += 4) {
        s0 += (double)__ldg(pk + (int64_t)r * kd) * (double)__ldg(pq + (int64_t)r * qd);
        s1 += (double)__ldg(pk + (int64_t)(r + 1) * kd) * (double)__ldg(pq + (int64_t)(r + 1) * qd);
        s2 += (double)__ldg(pk + (int64_t)(r + 2) * kd) * (double)__ldg(pq + (int64_t)(r + 2) * qd);
        s3 += (double)__ldg(pk + (int64_t)(r + 3) * kd) * (double)__ldg(pq + (int64_t)(r + 3) * qd);
    }
    for (; r < hd; ++r) s0 += (double)__ldg(pk + (int64_t)r * kd) * (double)__ldg(pq + (int64_t)r * qd);
    mfoldT[idx] = (float)(((s0 + s1) + (s2 + s3)) * scale);
}

__global__ void fold_vo_kernel(const float* __restrict__ wr, const float* __restrict__ wv, int qd, int kd, int H,
                               float* __restrict__ wvoT) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int zw = H * kd;
    if (idx >= (int64_t)qd * zw) return;
    const int o = (int)(idx / zw), c = (int)(idx % zw), h = c / kd, a = c % kd, hd = qd / H;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const float* pr = wr + (int64_t)o * qd + h * hd;
    const float* pv = wv + (int64_t)(h * hd) * kd + a;
    int r = 0;
#pragma unroll 2
    for (; r + 3 < hd; r += 4) {
        s0 += (double)__ldg(pr + r) * (double)__ldg(pv + (int64_t)r * kd);
        s1 += (double)__ldg(pr + r + 1) * (double)__ldg(pv + (int64_t)(r + 1) * kd);
        s2 += (double)__ldg(pr + r + 2) * (double)__ldg(pv + (int64_t)(r + 2) * kd);
        s3 += (double)__ldg(pr + r + 3) * (double)__ldg(pv + (int64_t)(r + 3) * kd);
    }
    for (; r < hd; ++r) s0 += (double)__ldg(pr + r) * (double)__ldg(pv + (int64_t)r * kd);
    wvoT[idx] = (float)((s0 + s1) + (s2 + s3));
}

// te0 = cos(b) (time encoding of dt = 0) and the bounds the stream kernel uses to pick its cosine path
__global__ void te0_kernel(const float* __restrict__ w, const float* __restrict__ b, int T, float* __restrict__ te0,
                           float* __restrict__ bound) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float wm = 0.f, bm = 0.f;
        for (int c = 0; c < T; ++c) wm = fmaxf(wm, fabsf(w[c])), bm = fmaxf(bm, fabsf(b[c]));
        bound[0] = wm, bound[1] = bm;
    }
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < T) te0[c] = time_channel(0.f, w[c], b[c]);
}

__global__ void u0_kernel(const float* __restrict__ mfoldT, const float* __restrict__ te0, int zw, int qd, int dn,
                          int T, float* __restrict__ u0) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= zw) return;
    float s = 0.f;
    for (int c = 0; c < T; ++c) s = fmaf(te0[c], mfoldT[(int64_t)row * qd + dn + c], s);
    u0[row] = s;
}

// LayerNorm fold of fc1: w1g = fc1_w[:, :qd] diag(gamma); c1 = row sums of w1g; c2 = fc1_w[:, :qd] beta + fc1_b
__global__ void ln_fold_prep_kernel(const float* __restrict__ fc1_w, const float* __restrict__ fc1_b,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, int dn, int qd,
                                    float* __restrict__ w1g, float* __restrict__ c1, float* __restrict__ c2) {
    const int c = blockIdx.x;          // one block per output row
    const int ld = qd + dn;
    double s1 = 0.0, s2 = 0.0;
    for (int k = threadIdx.x; k < qd; k += blockDim.x) {
        const float w = fc1_w[(int64_t)c * ld + k];
        const float wg = w * gamma[k];
        w1g[(int64_t)c * qd + k] = wg;
        s1 += (double)wg, s2 += (double)w * (double)beta[k];
    }
    __shared__ double sh1[32], sh2[32];
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(FULL, s1, o), s2 += __shfl_xor_sync(FULL, s2, o);
    if ((threadIdx.x & 31) == 0) sh1[threadIdx.x >> 5] = s1, sh2[threadIdx.x >> 5] = s2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sh1[w], b += sh2[w];
        c1[c] = (float)a;
        c2[c] = (float)(b + (double)fc1_b[c]);
    }
}

// ------------------------------------------------------------------ root conversion
__global__ void roots_kernel(const int64_t* __restrict__ nodes, const double* __restrict__ times, int64_t n,
                             int64_t num_nodes, int32_t* __restrict__ ids, double* __restrict__ t_out,
                             int* __restrict__ bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t v = nodes[i];
    if (v < 0 || v > num_nodes) {
        atomicExch(bad, 1);
        v = 0;
    }
    ids[i] = (int32_t)v;
    t_out[i] = times[i];
}

// ------------------------------------------------------------------ level sampler
// get_historical_neighbors('recent') for one recursion level, in the layout the layer
// kernels consume: int32 ids, float32 dt (with the reference's dtype rule: float64
// subtraction for the first n_f64 targets, float32 below; models/TGAT.py:120-125), and
// the next (lower) level's target list [self targets ; neighbour targets].
// Two queries per warp (16 lanes each): a query needs one 16-way probe search (the same number of dependent
// rounds as a 32-way one up to degree 256, one more beyond) and k = 20 output slots, so a full warp per query left
// most lanes idle; halving the warps per query took the kernel from 1.25 to 0.82 ms per Reddit-shape pass.
constexpr int LS_W = 16;
__global__ void __launch_bounds__(256) level_sample_kernel(
    const int64_t* __restrict__ indptr, const int2* __restrict__ adj, const double* __restrict__ ts,
    const int32_t* __restrict__ ids, const double* __restrict__ times, int64_t n, int64_t n_f64, int k,
    int32_t* __restrict__ nbr, int32_t* __restrict__ eid, float* __restrict__ dt, int32_t* __restrict__ next_ids,
    double* __restrict__ next_times, int32_t* __restrict__ pos, int32_t pad_pos,
    const int32_t* __restrict__ mirror, int32_t* __restrict__ self_pos, unsigned long long* __restrict__ valid_slots,
    int2* __restrict__ win = nullptr) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane & (LS_W - 1), shift = lane & LS_W;   // shift: 0 or 16
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LS_W;
    const bool active = q < n;
    int32_t v = 0;
    double t = 0.0;
    int64_t start = 0, end = 0;
    if (active) {
        v = __ldg(ids + q);
        t = __ldg(times + q);
        start = __ldg(indptr + v);
        end = __ldg(indptr + v + 1);
    }
    // searchsorted(ts[start:end), t, side='left') with 16 probes per round; both halves of the warp iterate together
    int64_t lo = start, hi = end;
    while (__any_sync(FULL, lo < hi)) {
        const int64_t len = hi - lo;
        const int64_t stride = (len + LS_W - 1) / LS_W;
        const int64_t p = lo + (int64_t)sub * stride;
        const bool before = (lo < hi) && (p < hi) && (__ldg(ts + p) < t);
        const int c = __popc((__ballot_sync(FULL, before) >> shift) & 0xFFFFu);
        if (lo < hi) {
            if (c == 0) {
                hi = lo;
            } else {
                const int64_t first_false = lo + (int64_t)c * stride;
                lo = lo + (int64_t)(c - 1) * stride + 1;
                hi = first_false < hi ? first_false : hi;
            }
        }
    }
    if (active) {
        const int64_t cut = lo;
        const int64_t have = cut - start;
        const int cnt = have < (int64_t)k ? (int)have : k;
        if (next_ids && sub == 0) next_ids[q] = v, next_times[q] = t;
        if (win && sub == 0) win[q] = make_int2((int)cut, cnt);  // the valid slots are positions [cut - cnt, cut)
        if (self_pos && sub == 0) {
            // Is (v, t) itself an event of the graph at a float32-exact time?  Then the lower-layer
            // embeddings of this root are already in the layer memo: the row of the event's entry in
            // the other endpoint's list is h_l(v, float32(t)), and with float32(t) == t both the
            // neighbourhood (ts < t) and every float32 dt of the root (models/TGAT.py:120-125:
            // float64 subtraction rounded once == float32 subtraction of the same two values) coincide.
            int32_t sp = -1;
            if (cut < end && __ldg(ts + cut) == t && (double)(float)t == t) sp = __ldg(mirror + cut);
            self_pos[q] = sp;
            if (sp < 0) atomicAdd(valid_slots + 1, 1ull);  // misses
        }
        const bool f64_rule = q < n_f64;
        const float tf = (float)t;
        for (int j = sub; j < k; j += LS_W) {
            int a = 0, e = 0, pp = pad_pos;
            float tsf = 0.f;
            if (j >= k - cnt) {
                const int64_t p = cut - k + j;
                const int2 ne = __ldg(adj + p);
                a = ne.x, e = ne.y, tsf = (float)__ldg(ts + p), pp = (int32_t)p;
            }
            const float d = f64_rule ? (float)(t - (double)tsf) : (tf - tsf);
            const int64_t o = q * k + j;
            nbr[o] = a, eid[o] = e, dt[o] = d;
            if (pos) pos[o] = pp;
            if (next_ids) next_ids[n + o] = a, next_times[n + o] = (double)tsf;
        }
        if (sub == 0) atomicAdd(&s_cnt, cnt);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(valid_slots, (unsigned long long)s_cnt);
}

// ------------------------------------------------------------------ residual + LayerNorm
// out = LayerNorm(O + [h_self | te0])   (models/modules.py:235-238; O already holds
// residual_fc's output incl. bias).  One warp per row, 16-byte accesses, two rows in flight per
// warp (a row is only ~1 KB: with one row per warp the kernel sat at a third of HBM bandwidth).
template <int NV>  // float4 per lane: ceil(qd / 128)
__global__ void __launch_bounds__(256) ln_kernel(const float* __restrict__ O, const float* __restrict__ self_base,
                                                 const int32_t* __restrict__ self_idx, const float* __restrict__ te0,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float* __restrict__ A, int64_t n, int dn, int T) {
    constexpr int R = 2;
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int qd = dn + T, q4 = qd >> 2, d4 = dn >> 2;
    const float inv_qd = 1.0f / (float)qd;
    float4 x[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t i = w * R + r;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int f = lane + 32 * v;
            x[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n && f < q4) {
                const float4 o = __ldg(reinterpret_cast<const float4*>(O + i * qd) + f);
                const float* self = self_base + (self_idx ? (int64_t)__ldg(self_idx + i) : i) * (int64_t)dn;
                const float4 s4 = f < d4 ? __ldg(reinterpret_cast<const float4*>(self) + f)
                                         : __ldg(reinterpret_cast<const float4*>(te0) + (f - d4));
                x[r][v] = make_float4(o.x + s4.x, o.y + s4.y, o.z + s4.z, o.w + s4.w);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t i = w * R + r;
        if (i >= n) break;  // warp-uniform
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) s += (x[r][v].x + x[r][v].y) + (x[r][v].z + x[r][v].w);
        const float mean = warp_sum(s) * inv_qd;
        float q = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            if (lane + 32 * v < q4) {
                const float a = x[r][v].x - mean, b = x[r][v].y - mean, c = x[r][v].z - mean, d = x[r][v].w - mean;
                q = fmaf(a, a, q), q = fmaf(b, b, q), q = fmaf(c, c, q), q = fmaf(d, d, q);
            }
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) * inv_qd + 1e-5f);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int f = lane + 32 * v;
            if (f < q4) {
                const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + f);
                const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + f);
                float4 y;
                y.x = fmaf((x[r][v].x - mean) * rstd, gm.x, bt.x);
                y.y = fmaf((x[r][v].y - mean) * rstd, gm.y, bt.y);
                y.z = fmaf((x[r][v].z - mean) * rstd, gm.z, bt.z);
                y.w = fmaf((x[r][v].w - mean) * rstd, gm.w, bt.w);
                reinterpret_cast<float4*>(A + i * qd)[f] = y;
            }
        }
    }
}

static int launch_ln(const float* O, const float* self_base, const int32_t* self_idx, const float* te0,
                     const float* gamma, const float* beta, float* A, int64_t n, int dn, int T, cudaStream_t st) {
    const int q4 = (dn + T) / 4;
    const unsigned blocks = (unsigned)ceil_div(ceil_div(n, 2) * 32, 256);
    if (q4 <= 96)
        ln_kernel<3><<<blocks, 256, 0, st>>>(O, self_base, self_idx, te0, gamma, beta, A, n, dn, T);
    else
        ln_kernel<4><<<blocks, 256, 0, st>>>(O, self_base, self_idx, te0, gamma, beta, A, n, dn, T);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

// ------------------------------------------------------------------ host pipeline
static int dev_copy(float** dst, const float* src, size_t count, cudaStream_t st) {
    if (!*dst) FLID_CUDA(cudaMalloc((void**)dst, count * sizeof(float)));
    FLID_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return FLID_OK;
}

// u rows for `n` rows of a feature table: U = feat[idx] . mfoldT[:, :dn]^T + u0
static int query_fold(flid_tgat* m, int layer, const float* feat, const int32_t* idx, int64_t n, float* U,
                      cudaStream_t st) {
    const LayerDev& ld = m->layers[layer];
    if (m->use_tc) {
        TcGemmArgs t;
        t.A0 = feat, t.lda0 = m->dn, t.idx0 = idx, t.w0 = m->dn;
        t.C = U, t.ldc = m->zw, t.bias = ld.u0, t.M = n;
        return tc_gemm(t, ld.tc_q, st);
    }
    GemmArgs g{feat, m->dn, idx, ld.mfoldT, m->qd, U, m->zw, ld.u0, n, m->zw, m->dn, 0, 0};
    return launch_gemm(g, st);
}

// everything after the attention stream for one level: out-projection(+residual_fc) ->
// +residual -> LayerNorm -> MergeLayer.  self rows: layer-(l-1) features of the targets.
static int output_chain(flid_tgat* m, int layer, int64_t n, const float* Z, const float* self_base,
                        const int32_t* self_idx, const float* merge_feat, const int32_t* ids, float* O, float* A,
                        float* Hd, float* out, const int32_t* out_idx, cudaStream_t st, bool kv = false) {
    const LayerDev& ld = m->layers[layer];
    if (m->use_tc) {
        TcGemmArgs t1;
        // kv: Z is the projected stream's output [sum a V | sum a te per head] (bulk_kv.cu), qd + H T wide
        const int zw = kv ? m->qd + m->H * m->T : m->zw;
        t1.A0 = Z, t1.lda0 = zw, t1.w0 = zw, t1.C = O, t1.ldc = m->qd, t1.bias = ld.res_b, t1.M = n;
        FLID_TRY(tc_gemm(t1, kv ? ld.tc_o2 : ld.tc_o, st));
        if (m->ln_fold && m->nt_src == merge_feat && ids != nullptr && ld.nt.p != nullptr) {
            // bulk passes: + residual, LayerNorm and fc1 in one GEMM (no LayerNorm kernel, no normalised rows in HBM,
            // fc1's raw-feature block as a per-node table); see TcGemmArgs::ln_*
            TcGemmArgs t2;
            t2.A0 = O, t2.lda0 = m->qd, t2.w0 = m->qd, t2.C = Hd, t2.ldc = m->dn, t2.M = n, t2.relu = 1;
            t2.ln_self = self_base, t2.ln_self_idx = self_idx, t2.ln_self_w = m->dn, t2.ln_tail = m->te0;
            t2.ln_c1 = ld.ln_c1, t2.ln_add = ld.nt.as<float>(), t2.ln_add_idx = ids, t2.ln_add_ld = m->dn;
            FLID_TRY(tc_gemm(t2, ld.tc_f1g, st));
            TcGemmArgs t3;
            t3.A0 = Hd, t3.lda0 = m->dn, t3.w0 = m->dn, t3.C = out, t3.ldc = m->dn, t3.bias = ld.fc2_b, t3.M = n;
            t3.cidx = out_idx;
            return tc_gemm(t3, ld.tc_f2, st);
        }
        FLID_TRY(launch_ln(O, self_base, self_idx, m->te0, ld.ln_w, ld.ln_b, A, n, m->dn, m->T, st));
        TcGemmArgs t2;   // fc1 on [attention output | layer-0 row of the target], ReLU fused
        t2.A0 = A, t2.lda0 = m->qd, t2.w0 = m->qd, t2.A1 = merge_feat, t2.lda1 = m->dn, t2.idx1 = ids, t2.w1 = m->dn;
        t2.C = Hd, t2.ldc = m->dn, t2.bias = ld.fc1_b, t2.M = n, t2.relu = 1;
        FLID_TRY(tc_gemm(t2, ld.tc_f1, st));
        TcGemmArgs t3;
        t3.A0 = Hd, t3.lda0 = m->dn, t3.w0 = m->dn, t3.C = out, t3.ldc = m->dn, t3.bias = ld.fc2_b, t3.M = n;
        t3.cidx = out_idx;
        return tc_gemm(t3, ld.tc_f2, st);
    }
    GemmArgs g1{Z, m->zw, nullptr, ld.wvoT, m->zw, O, m->qd, ld.res_b, n, m->qd, m->zw, 0, 0};
    FLID_TRY(launch_gemm(g1, st));
    FLID_TRY(launch_ln(O, self_base, self_idx, m->te0, ld.ln_w, ld.ln_b, A, n, m->dn, m->T, st));
    const int64_t ld1 = m->qd + m->dn;
    GemmArgs g2{A, m->qd, nullptr, ld.fc1_w, ld1, Hd, m->dn, nullptr, n, m->dn, m->qd, 0, 0};
    FLID_TRY(launch_gemm(g2, st));
    GemmArgs g3{merge_feat, m->dn, ids, ld.fc1_w + m->qd, ld1, Hd, m->dn, ld.fc1_b, n, m->dn, m->dn, 1, 1};
    FLID_TRY(launch_gemm(g3, st));
    GemmArgs g4{Hd, m->dn, nullptr, ld.fc2_w, m->dn, out, m->dn, ld.fc2_b, n, m->dn, m->dn, 0, 0, out_idx};
    return launch_gemm(g4, st);
}

int tgat_embed_ids(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                   const int32_t* ids, const double* times, int64_t n_f64, int64_t n, int k, float* out,
                   cudaStream_t st) {
    const int L = m->L;
    // roots per chunk so that the widest level (level 1) stays below max_l1_targets
    int64_t fan = 1;
    for (int l = 1; l < L; ++l) fan *= (1 + k);
    const int64_t chunk = std::max<int64_t>(1, m->max_l1_targets / fan);
    const int64_t nc_max = std::min(chunk, n);
    // level sizes / offsets for a full chunk (level L first)
    std::vector<int64_t> cnt(L + 1), off(L + 1);
    int64_t total = 0;
    cnt[L] = nc_max;
    for (int l = L; l >= 1; --l) {
        off[l] = total;
        total += cnt[l];
        if (l > 1) cnt[l - 1] = cnt[l] * (1 + k);
    }
    const int64_t n1 = cnt[1];
    FLID_TRY(m->ws_ids.reserve(sizeof(int32_t) * total));
    FLID_TRY(m->ws_times.reserve(sizeof(double) * total));
    FLID_TRY(m->ws_nbr.reserve(sizeof(int32_t) * total * k));
    FLID_TRY(m->ws_eid.reserve(sizeof(int32_t) * total * k));
    FLID_TRY(m->ws_dt.reserve(sizeof(float) * total * k));
    FLID_TRY(m->ws_h.reserve(sizeof(float) * (total - cnt[L] + 1) * m->dn));
    const bool use_table = (m->table_src == node_feat && m->table_rows > 0);
    FLID_TRY(m->ws_u.reserve(sizeof(float) * (use_table ? (L > 1 ? cnt[2] : 1) : n1) * m->zw));
    FLID_TRY(m->ws_z.reserve(sizeof(float) * n1 * m->zw));
    FLID_TRY(m->ws_o.reserve(sizeof(float) * n1 * m->qd));
    FLID_TRY(m->ws_a.reserve(sizeof(float) * n1 * m->qd));
    FLID_TRY(m->ws_hd.reserve(sizeof(float) * n1 * m->dn));
    FLID_TRY(m->ws_misc.reserve(64));
    unsigned long long* d_valid = m->ws_misc.as<unsigned long long>();
    FLID_CUDA(cudaMemsetAsync(d_valid, 0, sizeof(unsigned long long), st));

    int32_t* w_ids = m->ws_ids.as<int32_t>();
    double* w_times = m->ws_times.as<double>();
    int32_t* w_nbr = m->ws_nbr.as<int32_t>();
    int32_t* w_eid = m->ws_eid.as<int32_t>();
    float* w_dt = m->ws_dt.as<float>();
    float* w_h = m->ws_h.as<float>();
    float *U = m->ws_u.as<float>(), *Z = m->ws_z.as<float>(), *O = m->ws_o.as<float>(), *A = m->ws_a.as<float>(),
          *Hd = m->ws_hd.as<float>();

    int64_t evals = 0, queries = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t nc = std::min(chunk, n - r0);
        std::vector<int64_t> c(L + 1), o(L + 1), ho(L + 1);
        int64_t tot = 0, htot = 0;
        c[L] = nc;
        for (int l = L; l >= 1; --l) {
            o[l] = tot;
            tot += c[l];
            if (l > 1) c[l - 1] = c[l] * (1 + k);
        }
        for (int l = L - 1; l >= 1; --l) ho[l] = htot, htot += c[l];  // H_l buffers, l < L
        const int64_t nf = std::max<int64_t>(0, std::min(nc, n_f64 - r0));
        // level L targets = this chunk of roots
        FLID_CUDA(cudaMemcpyAsync(w_ids + o[L], ids + r0, sizeof(int32_t) * nc, cudaMemcpyDeviceToDevice, st));
        FLID_CUDA(cudaMemcpyAsync(w_times + o[L], times + r0, sizeof(double) * nc, cudaMemcpyDeviceToDevice, st));
        // top-down sampling
        for (int l = L; l >= 1; --l) {
            ProfScope prof(m, PROF_SAMPLE, st);
            level_sample_kernel<<<(unsigned)ceil_div(c[l] * LS_W, 256), 256, 0, st>>>(
                g->indptr, g->adj, g->ts, w_ids + o[l], w_times + o[l], c[l], nf, k, w_nbr + o[l] * k,
                w_eid + o[l] * k, w_dt + o[l] * k, l > 1 ? w_ids + o[l - 1] : nullptr,
                l > 1 ? w_times + o[l - 1] : nullptr, nullptr, 0, nullptr, nullptr, d_valid);
            FLID_LAUNCH_CHECK();
            queries += c[l];
        }
        // bottom-up layers
        for (int l = 1; l <= L; ++l) {
            const int64_t nl = c[l];
            const int32_t* lids = w_ids + o[l];
            AttnArgs a;
            a.edge_feat = edge_feat;
            a.nbr = w_nbr + o[l] * k, a.eid = w_eid + o[l] * k, a.dt = w_dt + o[l] * k;
            a.time_w = m->time_w, a.time_b = m->time_b, a.time_bound = m->time_bound;
            a.z = Z, a.n = nl, a.k = k, a.dn = m->dn, a.de = m->de, a.T = m->T;
            const float* self_base;
            const int32_t* self_idx;
            if (l == 1) {
                if (use_table) {
                    a.u_base = m->table.as<float>(), a.u_index = lids;
                } else {
                    ProfScope prof(m, PROF_QFOLD, st);
                    FLID_TRY(query_fold(m, 0, node_feat, lids, nl, U, st));
                    a.u_base = U, a.u_index = nullptr;
                }
                a.hrow_base = node_feat, a.hrow_by_id = 1, a.hrow_offset = 0;
                self_base = node_feat, self_idx = lids;
            } else {
                const float* hprev = w_h + ho[l - 1] * m->dn;  // [c[l-1], dn]: first nl rows = self, then nl*k nbr rows
                {
                    ProfScope prof(m, PROF_QFOLD, st);
                    FLID_TRY(query_fold(m, l - 1, hprev, nullptr, nl, U, st));
                }
                a.u_base = U, a.u_index = nullptr;
                a.hrow_base = hprev, a.hrow_by_id = 0, a.hrow_offset = nl;
                self_base = hprev, self_idx = nullptr;
            }
            {
                ProfScope prof(m, PROF_ATTN, st);
                FLID_TRY(launch_attn(a, m->H, st));
            }
            float* dst = (l == L) ? out + r0 * m->dn : w_h + ho[l] * m->dn;
            {
                ProfScope prof(m, PROF_OUT, st);
                FLID_TRY(output_chain(m, l - 1, nl, Z, self_base, self_idx, node_feat, lids, O, A, Hd, dst, nullptr, st));
            }
            evals += nl;
        }
    }
    m->stats[0] = evals;
    m->stats[2] = queries;
    m->stats[1] = -1, m->valid_mult = 1;  // fetched lazily by flid_tgat_last_stats
    return FLID_OK;
}


// ------------------------------------------------------------------ layer memo
// h_l(a, float32(ts)) of a neighbour slot depends only on the CSR entry it was sampled from
// (owner -> a at ts): the reference recomputes it for every root whose neighbourhood holds
// that entry (models/TGAT.py:108-113), a bulk pass needs it once.  memo_l is a table
// [M + 1, dn] indexed by CSR position (row M = the padded slot's query (node 0, t = 0.0)).
// A target's k neighbour rows are then k consecutive table rows, and the layer-(l-1)
// feature of target p itself is memo_{l-1}[p] (same key).
__global__ void memo_targets_kernel(const int2* __restrict__ adj, const double* __restrict__ ts, int64_t M,
                                    int64_t lo, int64_t n, int32_t* __restrict__ ids, double* __restrict__ times,
                                    const int32_t* __restrict__ mirror, int32_t* __restrict__ rows, int64_t range_hi) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t p = lo + i;
    if (p >= range_hi) p = M;   // the item after an owner range: the padded slot's query
    // Owner-major order: work item q evaluates the table row of q's partner entry, i.e. the query
    // (owner of q, time of q).  Consecutive items then ask for the same node at increasing times and
    // their neighbour windows overlap in all but one slot, so the gathered rows are cache hits.
    if (mirror) {
        if (p < M) p = __ldg(mirror + p);
        rows[i] = (int32_t)p;
    }
    ids[i] = p < M ? __ldg(adj + p).x : 0;
    times[i] = p < M ? (double)(float)__ldg(ts + p) : 0.0;  // the float32 neighbour time the recursion passes down
}

struct LayerCall {
    int64_t n = 0;
    const int32_t* ids = nullptr;  // target node ids [n]
    const int32_t *nbr = nullptr, *eid = nullptr, *pos = nullptr;
    const float* dt = nullptr;
    const float* hrow_base = nullptr;  // layer-(l-1) rows of the neighbours: by node id (pos == nullptr) or by CSR position
    const float* self_base = nullptr;  // layer-(l-1) rows of the targets
    const int32_t* self_idx = nullptr; // nullable: self row i = self_base[self_idx[i]]
    bool u_from_table = false;         // layer 1 with the cached per-node query fold
    float* out = nullptr;              // [n, dn]
    const int32_t* out_idx = nullptr;  // nullable: row i is written to out + out_idx[i] * dn
    bool kv = false;                   // projected bulk path (bulk_kv.cu); needs pos, eid and the level's tables
    const int2* win = nullptr;         // kv: (cut, cnt) per target
    const int2* adj = nullptr;
    int pad_pos = 0;
    bool zero_nbr = false;
};

// one attention layer (1-based `layer`) for n targets: query fold -> stream -> out chain
static int layer_eval(flid_tgat* m, int layer, const LayerCall& c, const float* node_feat, const float* edge_feat,
                      int k, cudaStream_t st) {
    float *U = m->ws_u.as<float>(), *Z = m->ws_z.as<float>(), *O = m->ws_o.as<float>(), *A = m->ws_a.as<float>(),
          *Hd = m->ws_hd.as<float>();
    if (c.kv) {
        KvCall kc;
        kc.level = layer, kc.n = c.n, kc.ids = c.ids, kc.self_base = c.self_base, kc.self_idx = c.self_idx;
        kc.nbr = c.nbr, kc.eid = c.eid, kc.pos = c.pos, kc.dt = c.dt, kc.U = U, kc.Y = Z;
        kc.win = m->kv_windowed ? c.win : nullptr, kc.adj = c.adj, kc.pad_pos = c.pad_pos, kc.graph_zero_nbr = c.zero_nbr;
        FLID_TRY(kv_attention(m, kc, k, st));
        ProfScope prof(m, PROF_OUT, st);
        return output_chain(m, layer - 1, c.n, Z, c.self_base, c.self_idx, node_feat, c.ids, O, A, Hd, c.out, c.out_idx, st,
                            true);
    }
    AttnArgs a;
    a.edge_feat = edge_feat;
    a.nbr = c.nbr, a.eid = c.eid, a.dt = c.dt;
    a.time_w = m->time_w, a.time_b = m->time_b, a.time_bound = m->time_bound;
    a.z = Z, a.n = c.n, a.k = k, a.dn = m->dn, a.de = m->de, a.T = m->T;
    const bool by_pos = c.pos != nullptr && layer > 1;  // level 1 rows are node-table rows even when positions were sampled
    a.hrow_base = c.hrow_base, a.hrow_by_id = by_pos ? 0 : 1, a.hrow_offset = 0, a.hrow_idx = by_pos ? c.pos : nullptr;
    if (c.u_from_table) {
        a.u_base = m->table.as<float>(), a.u_index = c.ids;
    } else {
        ProfScope prof(m, PROF_QFOLD, st);
        FLID_TRY(query_fold(m, layer - 1, c.self_base, c.self_idx, c.n, U, st));
        a.u_base = U, a.u_index = nullptr;
    }
    {
        if (layer > 1 && m->wait_event) {   // memo rows produced by other ranks: their exchange ran beside the query fold
            FLID_CUDA(cudaStreamWaitEvent(st, m->wait_event, 0));
            m->wait_event = nullptr;
        }
        ProfScope prof(m, PROF_ATTN, st);
        FLID_TRY(launch_attn(a, m->H, st));
    }
    ProfScope prof(m, PROF_OUT, st);
    return output_chain(m, layer - 1, c.n, Z, c.self_base, c.self_idx, node_feat, c.ids, O, A, Hd, c.out, c.out_idx, st);
}

static int reserve_layer_ws(flid_tgat* m, int64_t n, int k, bool need_u) {
    FLID_TRY(m->ws_ids.reserve(sizeof(int32_t) * n));
    FLID_TRY(m->ws_times.reserve(sizeof(double) * n));
    FLID_TRY(m->ws_nbr.reserve(sizeof(int32_t) * n * k));
    FLID_TRY(m->ws_eid.reserve(sizeof(int32_t) * n * k));
    FLID_TRY(m->ws_dt.reserve(sizeof(float) * n * k));
    FLID_TRY(m->ws_pos.reserve(sizeof(int32_t) * n * k));
    FLID_TRY(m->ws_win.reserve(sizeof(int2) * n));
    if (need_u) FLID_TRY(m->ws_u.reserve(sizeof(float) * n * m->zw));
    FLID_TRY(m->ws_z.reserve(sizeof(float) * n * m->zw));
    FLID_TRY(m->ws_o.reserve(sizeof(float) * n * m->qd));
    FLID_TRY(m->ws_a.reserve(sizeof(float) * n * m->qd));
    FLID_TRY(m->ws_hd.reserve(sizeof(float) * n * m->dn));
    FLID_TRY(m->ws_misc.reserve(64));
    return FLID_OK;
}

// enough free device memory for the projected tables of `level` (already allocated ones count as fitting)?
static bool kv_tables_fit(flid_tgat* m, const flid_graph* g, int level) {
    size_t need = 0;
    const size_t M1 = (size_t)g->num_entries + 1;
    if (level == 1) {
        // per-node V, per-edge V (at most one row per entry), per-entry scores
        const size_t want = sizeof(float) * ((size_t)(m->table_rows > 0 ? m->table_rows : 1) * m->qd + M1 * m->qd + M1 * m->H);
        const size_t have = m->kv_vn1.cap + m->kv_ve1.cap + m->kv_s1.cap;
        need = want > have ? want - have : 0;
    } else {
        const size_t want = sizeof(float) * M1 * 2 * m->qd;
        const size_t have = (int)m->kv_tab.size() > level - 2 ? m->kv_tab[level - 2].cap : 0;
        need = want > have ? want + want / 8 : 0;
    }
    if (need == 0) return true;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
    return need + (size_t(2) << 30) < free_b;   // keep 2 GiB of headroom for the caller's own tensors
}

// owner_range: [row_lo, row_hi) are WORK ITEMS in owner-major order -- item q evaluates the query (owner of q, time
// of q) and its row is scattered to q's partner position (any row of the table); one rank's share of an
// owner-partitioned pass.  Otherwise [row_lo, row_hi) are table rows.
int tgat_memo_build(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat, int k, int level,
                    const float* memo_prev, int64_t row_lo, int64_t row_hi, float* memo_out, cudaStream_t st,
                    bool owner_range = false, bool with_padded_row = false) {
    const int64_t M = g->num_entries;
    const bool use_table = level == 1 && (m->table_src == node_feat && m->table_rows > 0);
    const int64_t chunk = std::max<int64_t>(1, m->max_l1_targets);
    const int64_t ws_rows = std::max<int64_t>(1, std::min(chunk, row_hi - row_lo)) + (with_padded_row ? 1 : 0);
    FLID_TRY(reserve_layer_ws(m, ws_rows, k, !use_table));
    unsigned long long* d_valid = m->ws_misc.as<unsigned long long>();
    FLID_CUDA(cudaMemsetAsync(d_valid, 0, sizeof(unsigned long long), st));
    int32_t* w_ids = m->ws_ids.as<int32_t>();
    double* w_times = m->ws_times.as<double>();
    // A build of the whole table walks it in owner-major order and scatters the rows (the mirror is a
    // permutation, so every row is written exactly once); a row-range build (one rank's slice of a
    // sharded build) keeps table order so that its output stays one contiguous block.
    const bool owner_major = g->mirror != nullptr && (owner_range || (row_lo == 0 && row_hi == M + 1));
    if (owner_major) FLID_TRY(m->ws_self.reserve(sizeof(int32_t) * ws_rows));
    // projected bulk path (bulk_kv.cu): every entry is a slot of up to k targets of this build, so the per-entry
    // projections pay off; level 1 needs the cached per-node query table
    bool kv = kv_supported(m) && (level > 1 || use_table) && kv_tables_fit(m, g, level);
    if (kv) {
        if (level == 1)
            FLID_TRY(kv_ensure_level1(m, g, node_feat, edge_feat, st));
        else
            FLID_TRY(kv_ensure_level(m, g, level, memo_prev, node_feat, edge_feat, st));
    }
    int32_t* w_rows = owner_major ? m->ws_self.as<int32_t>() : nullptr;
    int64_t evals = 0;
    // chunks of the range, then (owner-range builds) the padded slot's query as an item of its own
    std::vector<std::pair<int64_t, int64_t>> chunks;
    for (int64_t c0 = row_lo; c0 < row_hi; c0 += chunk) chunks.push_back({c0, std::min(chunk, row_hi - c0)});
    if (with_padded_row && row_hi <= M) {
        // the padded slot's query rides along as one more item of the last chunk (memo_targets_kernel maps items past
        // the range to row M) instead of costing a launch chain of its own
        if (!chunks.empty() && chunks.back().second < chunk)
            chunks.back().second += 1;
        else
            chunks.push_back({M, 1});
    }
    for (const auto& ch : chunks) {
        const int64_t c0 = ch.first, n = ch.second;
        {
            ProfScope prof(m, PROF_SAMPLE, st);
            memo_targets_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(g->adj, g->ts, M, c0, n, w_ids, w_times,
                                                                         owner_major ? g->mirror : nullptr, w_rows,
                                                                         with_padded_row ? row_hi : M + 1);
            FLID_LAUNCH_CHECK();
            level_sample_kernel<<<(unsigned)ceil_div(n * LS_W, 256), 256, 0, st>>>(
                g->indptr, g->adj, g->ts, w_ids, w_times, n, 0, k, m->ws_nbr.as<int32_t>(), m->ws_eid.as<int32_t>(),
                m->ws_dt.as<float>(), nullptr, nullptr, (level > 1 || kv) ? m->ws_pos.as<int32_t>() : nullptr, (int32_t)M,
                nullptr, nullptr, d_valid, kv ? m->ws_win.as<int2>() : nullptr);
            FLID_LAUNCH_CHECK();
        }
        LayerCall c;
        c.kv = kv, c.win = m->ws_win.as<int2>(), c.adj = g->adj, c.pad_pos = (int)M, c.zero_nbr = g->zero_nbr != 0;
        c.n = n, c.ids = w_ids;
        c.nbr = m->ws_nbr.as<int32_t>(), c.eid = m->ws_eid.as<int32_t>(), c.dt = m->ws_dt.as<float>();
        if (level == 1) {
            c.hrow_base = node_feat, c.self_base = node_feat, c.self_idx = w_ids, c.u_from_table = use_table;
            if (kv) c.pos = m->ws_pos.as<int32_t>();
        } else {
            c.pos = m->ws_pos.as<int32_t>();
            c.hrow_base = memo_prev;
            if (owner_major)
                c.self_base = memo_prev, c.self_idx = w_rows;
            else
                c.self_base = memo_prev + c0 * m->dn;
        }
        if (owner_major)
            c.out = memo_out, c.out_idx = w_rows;
        else
            c.out = memo_out + c0 * m->dn;
        FLID_TRY(layer_eval(m, level, c, node_feat, edge_feat, k, st));
        evals += n;
    }
    m->stats[0] = evals, m->stats[2] = evals, m->stats[1] = -1, m->valid_mult = 1;
    return FLID_OK;
}

// roots with the lower layers memoised: one sampling pass and, per root, L attention evaluations --
// or a single one (level L) when the root is itself an event of the graph at a float32-exact time,
// because its own lower-layer embeddings are then rows of the memo as well (see level_sample_kernel).
int tgat_embed_memo(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                    const float* const* memo, const int32_t* ids, const double* times, int64_t n_f64, int64_t n, int k,
                    float* out, const int32_t* out_perm, int* bad_ids, cudaStream_t st) {
    const int L = m->L;
    const int64_t M = g->num_entries;
    const bool use_table = (m->table_src == node_feat && m->table_rows > 0);
    const bool try_self = L > 1 && g->mirror != nullptr && m->self_from_memo;
    const int64_t chunk = std::max<int64_t>(1, m->max_l1_targets);
    const int64_t nmax = std::min(chunk, n);
    // projected bulk path: worth building the per-entry tables when the call's slots outnumber the entries
    // (a bulk pass); per-batch calls keep the x-space stream
    const int64_t range_entries = m->bulk_hi < 0 ? M : m->bulk_hi - m->bulk_lo;
    bool kv = kv_supported(m) && use_table && n * (int64_t)k >= range_entries && M > 0;
    for (int l = 2; l <= L && kv; ++l) kv = kv_tables_fit(m, g, l);
    if (kv) {
        for (int l = 2; l <= L; ++l) FLID_TRY(kv_ensure_level(m, g, l, memo[l - 2], node_feat, edge_feat, st));
    }
    bool kv_l1_ready = false;
    FLID_TRY(reserve_layer_ws(m, nmax, k, L > 1 || !use_table));
    if (L > 1) FLID_TRY(m->ws_h.reserve(sizeof(float) * 2 * nmax * m->dn));
    if (try_self) FLID_TRY(m->ws_self.reserve(sizeof(int32_t) * nmax));
    // [0] valid slots, [1] roots without a memo row, [2] "node id outside the graph" flag of the caller's
    // root conversion (already zeroed / set before this function runs; read with the first chunk's counters)
    unsigned long long* d_cnt = m->ws_misc.as<unsigned long long>();
    FLID_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), st));
    int64_t evals = 0, valid_weighted = 0;
    unsigned long long h_prev[2] = {0, 0};
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t nc = std::min(chunk, n - r0);
        const int64_t nf = std::max<int64_t>(0, std::min(nc, n_f64 - r0));
        {
            ProfScope prof(m, PROF_SAMPLE, st);
            level_sample_kernel<<<(unsigned)ceil_div(nc * LS_W, 256), 256, 0, st>>>(
                g->indptr, g->adj, g->ts, ids + r0, times + r0, nc, nf, k, m->ws_nbr.as<int32_t>(),
                m->ws_eid.as<int32_t>(), m->ws_dt.as<float>(), nullptr, nullptr,
                (L > 1 || kv) ? m->ws_pos.as<int32_t>() : nullptr, (int32_t)M, try_self ? g->mirror : nullptr,
                try_self ? m->ws_self.as<int32_t>() : nullptr, d_cnt, kv ? m->ws_win.as<int2>() : nullptr);
            FLID_LAUNCH_CHECK();
        }
        // counters of this chunk (one small synchronous read per chunk of up to 65 536 roots)
        unsigned long long h_cnt[3];
        FLID_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        FLID_CUDA(cudaStreamSynchronize(st));
        if (bad_ids) *bad_ids = h_cnt[2] != 0;
        const int64_t chunk_valid = (int64_t)(h_cnt[0] - h_prev[0]);
        const bool all_in_memo = try_self && h_cnt[1] == h_prev[1];
        h_prev[0] = h_cnt[0], h_prev[1] = h_cnt[1];
        float* hbuf[2] = {m->ws_h.as<float>(), L > 1 ? m->ws_h.as<float>() + nmax * m->dn : nullptr};
        if (kv && !all_in_memo && !kv_l1_ready) {
            FLID_TRY(kv_ensure_level1(m, g, node_feat, edge_feat, st));
            kv_l1_ready = true;
        }
        for (int l = all_in_memo ? L : 1; l <= L; ++l) {
            LayerCall c;
            c.kv = kv, c.win = m->ws_win.as<int2>(), c.adj = g->adj, c.pad_pos = (int)M, c.zero_nbr = g->zero_nbr != 0;
            c.n = nc, c.ids = ids + r0;
            c.nbr = m->ws_nbr.as<int32_t>(), c.eid = m->ws_eid.as<int32_t>(), c.dt = m->ws_dt.as<float>();
            if (l == 1) {
                c.hrow_base = node_feat, c.self_base = node_feat, c.self_idx = ids + r0, c.u_from_table = use_table;
                if (kv) c.pos = m->ws_pos.as<int32_t>();
            } else {
                c.pos = m->ws_pos.as<int32_t>();
                c.hrow_base = memo[l - 2];
                if (all_in_memo)
                    c.self_base = memo[l - 2], c.self_idx = m->ws_self.as<int32_t>();
                else
                    c.self_base = hbuf[l & 1];
            }
            c.out = (l == L) ? out + r0 * m->dn : hbuf[(l + 1) & 1];
            if (l == L && out_perm != nullptr) c.out = out, c.out_idx = out_perm + r0;  // queries were reordered
            FLID_TRY(layer_eval(m, l, c, node_feat, edge_feat, k, st));
            evals += nc;
            valid_weighted += chunk_valid;
        }
    }
    m->stats[0] = evals, m->stats[2] = n, m->stats[1] = valid_weighted, m->valid_mult = 1;
    return FLID_OK;
}

}  // namespace flid

// ====================================================================== C ABI
extern "C" {

int flid_tgat_create(int node_dim, int edge_dim, int time_dim, int num_layers, int num_heads, flid_tgat** out) {
    using namespace flid;
    FLID_REQUIRE(out != nullptr, "flid_tgat_create: out is null");
    FLID_REQUIRE(node_dim > 0 && edge_dim > 0 && time_dim > 0 && num_layers > 0 && num_heads > 0,
                 "flid_tgat_create: dims must be positive");
    FLID_REQUIRE(node_dim % 4 == 0 && edge_dim % 4 == 0 && time_dim % 4 == 0,
                 "flid_tgat_create: node/edge/time dims must be multiples of 4 (16-byte rows)");
    // the reference's own assertion (models/modules.py:149)
    FLID_REQUIRE((node_dim + time_dim) % num_heads == 0,
                 "The sum of node_feat_dim and time_feat_dim should be divided by num_heads!");
    FLID_REQUIRE(num_heads == 1 || num_heads == 2 || num_heads == 4, "flid_tgat_create: num_heads must be 1, 2 or 4");
    FLID_REQUIRE(node_dim + edge_dim <= (num_heads == 4 ? 384 : 768) && time_dim <= 128 && node_dim + time_dim <= 512,
                 "flid_tgat_create: feature widths above kernel limits");
    flid_tgat* m = new flid_tgat();
    m->dn = node_dim, m->de = edge_dim, m->T = time_dim, m->L = num_layers, m->H = num_heads;
    m->qd = node_dim + time_dim, m->kd = node_dim + edge_dim + time_dim, m->hd = m->qd / num_heads;
    m->zw = num_heads * m->kd;
    m->layers.resize(num_layers);
    const char* mode = getenv("FLID_GEMM");
    m->use_tc = !(mode && strcmp(mode, "simt") == 0);
    const char* sm = getenv("FLID_SELF_MEMO");
    m->self_from_memo = !(sm && sm[0] == '0');
    const char* kvv = getenv("FLID_BULK_KV");
    m->kv_enabled = !(kvv && kvv[0] == '0');
    const char* kk = getenv("FLID_KV_KERNEL");
    m->kv_windowed = !(kk && strcmp(kk, "mask") == 0);
    const char* so = getenv("FLID_SORT_QUERIES");
    m->sort_bulk_queries = !(so && so[0] == '0');
    if (const char* ct = getenv("FLID_CHUNK_TARGETS"))   // development knob, same as flid_tgat_set_chunk_targets
        if (atoll(ct) >= 1024) m->max_l1_targets = atoll(ct);
    *out = m;
    return FLID_OK;
}

void flid_tgat_free(flid_tgat* m) {
    if (!m) return;
    cudaFree(m->time_w), cudaFree(m->time_b), cudaFree(m->te0), cudaFree(m->time_bound);
    for (auto& l : m->layers) {
        cudaFree(l.mfoldT), cudaFree(l.u0), cudaFree(l.wvoT), cudaFree(l.res_b), cudaFree(l.ln_w), cudaFree(l.ln_b);
        cudaFree(l.fc1_w), cudaFree(l.fc1_b), cudaFree(l.fc2_w), cudaFree(l.fc2_b);
        flid::tc_free_weight(&l.tc_q), flid::tc_free_weight(&l.tc_o), flid::tc_free_weight(&l.tc_f1);
        flid::tc_free_weight(&l.tc_f2);
        flid::kv_free_layer(l);
        cudaFree(l.lnw), flid::tc_free_weight(&l.tc_f1g), flid::tc_free_weight(&l.tc_f1n), l.nt.release();
    }
    m->kv_vn1.release(), m->kv_ve1.release(), m->kv_s1.release();
    for (auto& t : m->kv_tab) t.release();
    for (auto e : m->prof_ev) cudaEventDestroy(e);
    flid::DevBuf* bufs[] = {&m->raw_q, &m->raw_k, &m->raw_v, &m->raw_r, &m->table, &m->ws_ids, &m->ws_times,
                            &m->ws_nbr, &m->ws_eid, &m->ws_dt, &m->ws_h, &m->ws_u, &m->ws_z, &m->ws_o,
                            &m->ws_a, &m->ws_hd, &m->ws_misc, &m->ws_rid, &m->ws_rt, &m->ws_bad, &m->ws_pos, &m->ws_self, &m->ws_sort,
                            &m->tgn_ids, &m->tgn_times, &m->tgn_eids, &m->tgn_gi, &m->tgn_gh, &m->ws_win, &m->tgn_out, &m->tgn_ctr};
    flid::tc_free_weight(&m->tc_gih), flid::tc_free_weight(&m->tc_ghh);
    for (int i = 0; i < flid_tgat::PREP_STREAMS; ++i) {
        if (m->prep_stream[i]) cudaStreamDestroy(m->prep_stream[i]);
        if (m->prep_join[i]) cudaEventDestroy(m->prep_join[i]);
    }
    if (m->prep_fork) cudaEventDestroy(m->prep_fork);
    if (m->tgn_stream) cudaStreamDestroy(m->tgn_stream);
    if (m->tgn_ev) cudaEventDestroy(m->tgn_ev);
    for (auto* b : bufs) b->release();
    delete m;
}

int flid_tgat_set_weights(flid_tgat* m, const float* time_w, const float* time_b,
                          const flid_tgat_layer_weights* layers_host, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && time_w && time_b && layers_host, "flid_tgat_set_weights: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int qd = m->qd, kd = m->kd, dn = m->dn, T = m->T, H = m->H, zw = m->zw;
    FLID_TRY(dev_copy(&m->time_w, time_w, T, st));
    FLID_TRY(dev_copy(&m->time_b, time_b, T, st));
    if (!m->te0) FLID_CUDA(cudaMalloc((void**)&m->te0, sizeof(float) * T));
    if (!m->time_bound) FLID_CUDA(cudaMalloc((void**)&m->time_bound, sizeof(float) * 2));
    te0_kernel<<<(unsigned)ceil_div(T, 128), 128, 0, st>>>(m->time_w, m->time_b, T, m->te0, m->time_bound);
    FLID_LAUNCH_CHECK();
    // python: head_dim ** -0.5 is a float64; multiplying a float32 tensor by it uses its float32 value
    // the stream kernel evaluates softmax with exp2, so log2(e) is folded into the scale here
    const double scale = (double)(float)pow((double)m->hd, -0.5) * 1.4426950408889634074;
    // Three independent chains per layer (query fold, output fold, MergeLayer weights), each a handful of small
    // launches: they run side by side on the prep streams (an E-step pass follows an M-step, so this upload is part
    // of every pass; serialised it is ~0.4 ms of 4-70 us kernels), joined back into `st` at the end.
    constexpr int NS = flid_tgat::PREP_STREAMS;
    static const bool multi = [] {
        const char* e = getenv("FLID_PREP_STREAMS");
        return !(e && e[0] == '0');
    }();
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    FLID_CUDA(cudaStreamIsCapturing(st, &cap));
    const bool fork = multi && cap == cudaStreamCaptureStatusNone;
    if (fork) {
        if (!m->prep_fork) {
            FLID_CUDA(cudaEventCreateWithFlags(&m->prep_fork, cudaEventDisableTiming));
            for (int i = 0; i < NS; ++i) {
                FLID_CUDA(cudaStreamCreateWithFlags(&m->prep_stream[i], cudaStreamNonBlocking));
                FLID_CUDA(cudaEventCreateWithFlags(&m->prep_join[i], cudaEventDisableTiming));
            }
        }
        FLID_CUDA(cudaEventRecord(m->prep_fork, st));
        for (int i = 0; i < NS; ++i) FLID_CUDA(cudaStreamWaitEvent(m->prep_stream[i], m->prep_fork, 0));
    }
    auto chain = [&](int l, int j) { return fork ? m->prep_stream[(3 * l + j) % NS] : st; };
    const int64_t tot = (int64_t)zw * qd;
    for (int l = 0; l < m->L; ++l) {   // the two long kernels of every layer first: they run while the host enqueues the rest
        const flid_tgat_layer_weights& w = layers_host[l];
        LayerDev& d = m->layers[l];
        FLID_REQUIRE(w.query_w && w.key_w && w.value_w && w.ln_w && w.ln_b && w.res_w && w.res_b && w.fc1_w &&
                         w.fc1_b && w.fc2_w && w.fc2_b,
                     "flid_tgat_set_weights: layer %d has a null weight pointer", l);
        if (!d.mfoldT) FLID_CUDA(cudaMalloc((void**)&d.mfoldT, sizeof(float) * (size_t)zw * qd));
        if (!d.wvoT) FLID_CUDA(cudaMalloc((void**)&d.wvoT, sizeof(float) * (size_t)zw * qd));
        if (!d.u0) FLID_CUDA(cudaMalloc((void**)&d.u0, sizeof(float) * zw));
        fold_vo_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, chain(l, 1)>>>(w.res_w, w.value_w, qd, kd, H, d.wvoT);
        FLID_LAUNCH_CHECK();
        fold_qk_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, chain(l, 0)>>>(w.query_w, w.key_w, qd, kd, H, scale, d.mfoldT);
        FLID_LAUNCH_CHECK();
    }
    for (int l = 0; l < m->L; ++l) {
        const flid_tgat_layer_weights& w = layers_host[l];
        LayerDev& d = m->layers[l];
        cudaStream_t sq = chain(l, 0), so = chain(l, 1), sm = chain(l, 2);
        // chain 1: query fold -> u0 -> its tiled image
        u0_kernel<<<(unsigned)ceil_div(zw, 128), 128, 0, sq>>>(d.mfoldT, m->te0, zw, qd, dn, T, d.u0);
        FLID_LAUNCH_CHECK();
        FLID_TRY(tc_prepare_weight(d.mfoldT, qd, zw, dn, &d.tc_q, sq, m->numeric));
        // chain 2: output fold -> its tiled image
        FLID_TRY(dev_copy(&d.res_b, w.res_b, qd, so));
        FLID_TRY(tc_prepare_weight(d.wvoT, zw, qd, zw, &d.tc_o, so, m->numeric));
        // chain 3: LayerNorm / MergeLayer parameters and their images
        FLID_TRY(dev_copy(&d.ln_w, w.ln_w, qd, sm));
        FLID_TRY(dev_copy(&d.ln_b, w.ln_b, qd, sm));
        FLID_TRY(dev_copy(&d.fc1_w, w.fc1_w, (size_t)dn * (qd + dn), sm));
        FLID_TRY(dev_copy(&d.fc1_b, w.fc1_b, dn, sm));
        FLID_TRY(dev_copy(&d.fc2_w, w.fc2_w, (size_t)dn * dn, sm));
        FLID_TRY(dev_copy(&d.fc2_b, w.fc2_b, dn, sm));
        FLID_TRY(tc_prepare_weight(d.fc1_w, qd + dn, dn, qd + dn, &d.tc_f1, sm, m->numeric));
        FLID_TRY(tc_prepare_weight(d.fc2_w, dn, dn, dn, &d.tc_f2, sm, m->numeric));
        FLID_TRY(kv_fold_layer(m, l, w.query_w, w.key_w, w.value_w, w.res_w, sm));
        if (m->use_tc && qd % 4 == 0 && dn % 4 == 0) {
            if (!d.lnw) {
                FLID_CUDA(cudaMalloc((void**)&d.lnw, sizeof(float) * ((size_t)dn * qd + 2 * dn)));
                d.w1g = d.lnw, d.ln_c1 = d.lnw + (size_t)dn * qd, d.ln_c2 = d.ln_c1 + dn;
            }
            ln_fold_prep_kernel<<<dn, 128, 0, sm>>>(d.fc1_w, d.fc1_b, d.ln_w, d.ln_b, dn, qd, d.w1g, d.ln_c1, d.ln_c2);
            FLID_LAUNCH_CHECK();
            FLID_TRY(tc_prepare_weight(d.w1g, qd, dn, qd, &d.tc_f1g, sm, m->numeric));
            FLID_TRY(tc_prepare_weight(d.fc1_w + qd, qd + dn, dn, dn, &d.tc_f1n, sm, m->numeric));
        }
    }
    if (fork) {
        for (int i = 0; i < NS; ++i) {
            FLID_CUDA(cudaEventRecord(m->prep_join[i], m->prep_stream[i]));
            FLID_CUDA(cudaStreamWaitEvent(st, m->prep_join[i], 0));
        }
    }
    m->nt_src = nullptr;
    m->weights_version += 1;
    m->have_weights = true;
    m->table_src = nullptr;  // cached query folds are stale now
    m->table_rows = 0;
    return FLID_OK;
}

int flid_tgat_cache_node_table(flid_tgat* m, const float* node_feat, int64_t rows, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && node_feat && rows > 0, "flid_tgat_cache_node_table: bad argument");
    FLID_REQUIRE(m->have_weights, "flid_tgat_cache_node_table: weights not set");
    FLID_TRY(m->table.reserve(sizeof(float) * (size_t)rows * m->zw));
    FLID_TRY(query_fold(m, 0, node_feat, nullptr, rows, m->table.as<float>(), (cudaStream_t)stream));
    m->table_src = node_feat;
    m->table_rows = rows;
    // per-layer node part of the LayerNorm-folded fc1: node_feat . fc1_w[:, qd:]^T + (fc1_w[:, :qd] beta + fc1_b)
    m->nt_src = nullptr;
    if (m->use_tc && m->ln_fold) {
        for (auto& ld : m->layers) {
            if (!ld.tc_f1n.buf) return FLID_OK;
            FLID_TRY(ld.nt.reserve(sizeof(float) * (size_t)rows * m->dn));
            TcGemmArgs t;
            t.A0 = node_feat, t.lda0 = m->dn, t.w0 = m->dn, t.C = ld.nt.as<float>(), t.ldc = m->dn, t.bias = ld.ln_c2;
            t.M = rows;
            FLID_TRY(tc_gemm(t, ld.tc_f1n, (cudaStream_t)stream));
        }
        m->nt_src = node_feat;
    }
    return FLID_OK;
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ ids, int64_t n, int w,
                                    float* __restrict__ table) {
    const int64_t i = blockIdx.x;
    for (int c = threadIdx.x; c < w; c += blockDim.x) table[(int64_t)ids[i] * w + c] = src[i * w + c];
}

int flid_tgat_refresh_node_rows(flid_tgat* m, const float* node_feat, const int32_t* row_ids, int64_t n,
                                flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && node_feat && row_ids, "flid_tgat_refresh_node_rows: null argument");
    FLID_REQUIRE(m->table_src == node_feat && m->table_rows > 0, "flid_tgat_refresh_node_rows: table not cached");
    if (n <= 0) return FLID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    FLID_TRY(m->ws_u.reserve(sizeof(float) * (size_t)n * m->zw));
    FLID_TRY(query_fold(m, 0, node_feat, row_ids, n, m->ws_u.as<float>(), st));
    scatter_rows_kernel<<<(unsigned)n, 256, 0, st>>>(m->ws_u.as<float>(), row_ids, n, m->zw, m->table.as<float>());
    FLID_LAUNCH_CHECK();
    m->nt_src = nullptr;   // the per-node fc1 tables of the LayerNorm fold are not refreshed (mutable tables: TGN)
    return FLID_OK;
}

int flid_tgat_embed(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                    const int64_t* nodes, const double* times, int times_are_f32, int64_t n, int k, float* out,
                    flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && g && node_feat && edge_feat, "flid_tgat_embed: null argument");
    FLID_REQUIRE(m->have_weights, "flid_tgat_embed: weights not set");
    FLID_REQUIRE(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    if (n <= 0) return FLID_OK;
    FLID_REQUIRE(nodes && times && out, "flid_tgat_embed: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    FLID_TRY(m->ws_rid.reserve(sizeof(int32_t) * n));
    FLID_TRY(m->ws_rt.reserve(sizeof(double) * n));
    FLID_TRY(m->ws_bad.reserve(sizeof(int)));
    FLID_CUDA(cudaMemsetAsync(m->ws_bad.p, 0, sizeof(int), st));
    roots_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(nodes, times, n, g->num_nodes, m->ws_rid.as<int32_t>(),
                                                            m->ws_rt.as<double>(), m->ws_bad.as<int>());
    FLID_LAUNCH_CHECK();
    FLID_TRY(tgat_embed_ids(m, g, node_feat, edge_feat, m->ws_rid.as<int32_t>(), m->ws_rt.as<double>(),
                            times_are_f32 ? 0 : n, n, k, out, st));
    // out-of-range ids were clamped to the padding node; report them like the reference's IndexError
    int hbad = 0;
    FLID_CUDA(cudaMemcpyAsync(&hbad, m->ws_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLID_CUDA(cudaStreamSynchronize(st));
    if (hbad) {
        set_error("flid_tgat_embed: node id outside the graph");
        return FLID_ERR_RANGE;
    }
    return FLID_OK;
}

int flid_tgat_memo_build(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat, int k,
                         int level, const float* memo_prev, int64_t row_lo, int64_t row_hi, float* memo_out,
                         flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && g && node_feat && edge_feat && memo_out, "flid_tgat_memo_build: null argument");
    FLID_REQUIRE(m->have_weights, "flid_tgat_memo_build: weights not set");
    FLID_REQUIRE(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    FLID_REQUIRE(level >= 1 && level <= m->L, "flid_tgat_memo_build: level %d outside 1..%d", level, m->L);
    FLID_REQUIRE(level == 1 || memo_prev != nullptr, "flid_tgat_memo_build: level %d needs the level-%d table", level,
                 level - 1);
    FLID_REQUIRE(g->num_entries < 0x7fffffffLL, "flid_tgat_memo_build: more than 2^31 adjacency entries");
    FLID_REQUIRE(row_lo >= 0 && row_lo <= row_hi && row_hi <= g->num_entries + 1,
                 "flid_tgat_memo_build: row range outside [0, entries + 1]");
    if (row_lo == row_hi) return FLID_OK;
    return tgat_memo_build(m, g, node_feat, edge_feat, k, level, memo_prev, row_lo, row_hi, memo_out,
                           (cudaStream_t)stream);
}

int flid_tgat_memo_build_owner_range(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                                     int k, int level, const float* memo_prev, int64_t item_lo, int64_t item_hi,
                                     int with_padded_row, float* memo_out, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && g && node_feat && edge_feat && memo_out, "flid_tgat_memo_build_owner_range: null argument");
    FLID_REQUIRE(m->have_weights, "flid_tgat_memo_build_owner_range: weights not set");
    FLID_REQUIRE(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    FLID_REQUIRE(level >= 1 && level <= m->L, "flid_tgat_memo_build_owner_range: level %d outside 1..%d", level, m->L);
    FLID_REQUIRE(level == 1 || memo_prev != nullptr, "flid_tgat_memo_build_owner_range: level %d needs the level-%d table",
                 level, level - 1);
    FLID_REQUIRE(g->mirror != nullptr, "flid_tgat_memo_build_owner_range: the graph has no partner index (build it from events)");
    FLID_REQUIRE(item_lo >= 0 && item_lo <= item_hi && item_hi <= g->num_entries,
                 "flid_tgat_memo_build_owner_range: item range outside [0, entries]");
    if (item_hi == item_lo && !with_padded_row) return FLID_OK;
    return tgat_memo_build(m, g, node_feat, edge_feat, k, level, memo_prev, item_lo, item_hi, memo_out,
                           (cudaStream_t)stream, true, with_padded_row != 0);
}

int flid_graph_export_mirror(const flid_graph* g, int32_t* out_dev, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(g && out_dev, "flid_graph_export_mirror: null argument");
    FLID_REQUIRE(g->mirror != nullptr, "flid_graph_export_mirror: the graph has no partner index (build it from events)");
    FLID_CUDA(cudaMemcpyAsync(out_dev, g->mirror, sizeof(int32_t) * g->num_entries, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream));
    return FLID_OK;
}

int flid_tgat_embed_memo(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                         const float* const* memo_tables_host, const int64_t* nodes, const double* times,
                         int times_are_f32, int64_t n, int k, float* out, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && g && node_feat && edge_feat, "flid_tgat_embed_memo: null argument");
    FLID_REQUIRE(m->have_weights, "flid_tgat_embed_memo: weights not set");
    FLID_REQUIRE(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    FLID_REQUIRE(m->L == 1 || memo_tables_host != nullptr, "flid_tgat_embed_memo: memo tables missing");
    for (int l = 0; l + 1 < m->L; ++l)
        FLID_REQUIRE(memo_tables_host[l] != nullptr, "flid_tgat_embed_memo: memo table of layer %d is null", l + 1);
    if (n <= 0) return FLID_OK;
    FLID_REQUIRE(nodes && times && out, "flid_tgat_embed_memo: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    FLID_TRY(m->ws_rid.reserve(sizeof(int32_t) * n));
    FLID_TRY(m->ws_rt.reserve(sizeof(double) * n));
    FLID_TRY(m->ws_misc.reserve(64));
    // the range flag lives next to the sampling counters so that one small read per chunk (which also
    // orders the host after the kernels that consume the caller's staging buffers) returns all three
    int* d_bad = reinterpret_cast<int*>(m->ws_misc.as<unsigned long long>() + 2);
    FLID_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), st));
    roots_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(nodes, times, n, g->num_nodes, m->ws_rid.as<int32_t>(),
                                                            m->ws_rt.as<double>(), d_bad);
    FLID_LAUNCH_CHECK();
    int hbad = 0;
    const int32_t* q_ids = m->ws_rid.as<int32_t>();
    const double* q_times = m->ws_rt.as<double>();
    const int32_t* perm = nullptr;
    if (n >= 8192 && n < 0x7fffffffLL && m->sort_bulk_queries) {
        // bulk pass: evaluate the roots in (node, time) order, scatter the rows back through the last GEMM
        int32_t *p = nullptr, *si = nullptr;
        double* stm = nullptr;
        FLID_TRY(sort_queries(q_ids, q_times, n, g->num_nodes, m->ws_sort, &p, &si, &stm, st));
        perm = p, q_ids = si, q_times = stm;
    }
    FLID_TRY(tgat_embed_memo(m, g, node_feat, edge_feat, memo_tables_host, q_ids, q_times, times_are_f32 ? 0 : n, n, k,
                             out, perm, &hbad, st));
    if (hbad) {
        FLID_CUDA(cudaStreamSynchronize(st));
        set_error("flid_tgat_embed_memo: node id outside the graph");
        return FLID_ERR_RANGE;
    }
    return FLID_OK;
}

int flid_tgat_set_self_from_memo(flid_tgat* m, int enable) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_set_self_from_memo: null handle");
    m->self_from_memo = enable != 0;
    return FLID_OK;
}

int flid_tgat_set_ln_fold(flid_tgat* m, int enable) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_set_ln_fold: null handle");
    if (m->ln_fold != (enable != 0)) m->table_src = nullptr, m->table_rows = 0, m->nt_src = nullptr;  // rebuild with / without the fc1 tables
    m->ln_fold = enable != 0;
    return FLID_OK;
}

int flid_tgat_set_wait_event(flid_tgat* m, void* cuda_event) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_set_wait_event: null handle");
    m->wait_event = (cudaEvent_t)cuda_event;
    return FLID_OK;
}

int flid_tgat_set_sort_queries(flid_tgat* m, int enable) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_set_sort_queries: null handle");
    const char* so = getenv("FLID_SORT_QUERIES");
    m->sort_bulk_queries = enable != 0 && !(so && so[0] == '0');
    return FLID_OK;
}

int flid_tgat_set_numeric_mode(flid_tgat* m, int mode) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_set_numeric_mode: null handle");
    FLID_REQUIRE(mode == 0 || mode == 1, "flid_tgat_set_numeric_mode: mode must be 0 (f32) or 1 (bf16 projections)");
    FLID_REQUIRE(mode == 0 || m->use_tc, "flid_tgat_set_numeric_mode: the bf16 mode needs the tcgen05 GEMMs (FLID_GEMM=simt is set)");
    if (m->numeric != mode) m->have_weights = false;  // the tiled weight images are per mode
    m->numeric = mode;
    return FLID_OK;
}

int flid_tgat_set_chunk_targets(flid_tgat* m, int64_t max_layer1_targets) {
    using namespace flid;
    FLID_REQUIRE(m && max_layer1_targets > 0, "flid_tgat_set_chunk_targets: bad argument");
    m->max_l1_targets = max_layer1_targets;
    return FLID_OK;
}

int flid_tgat_profile(flid_tgat* m, int enable) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_profile: null handle");
    m->prof_on = enable != 0;
    m->prof_used = 0;
    return FLID_OK;
}

int flid_tgat_profile_read(flid_tgat* m, double ms[4], int64_t launches[4]) {
    using namespace flid;
    FLID_REQUIRE(m && ms && launches, "flid_tgat_profile_read: null argument");
    for (int c = 0; c < PROF_CLASSES; ++c) ms[c] = 0.0, launches[c] = 0;
    FLID_CUDA(cudaDeviceSynchronize());
    for (size_t s = 0; s + 1 < m->prof_used; s += 2) {
        float t = 0.f;
        FLID_CUDA(cudaEventElapsedTime(&t, m->prof_ev[s], m->prof_ev[s + 1]));
        ms[m->prof_cls[s / 2]] += t;
        launches[m->prof_cls[s / 2]] += 1;
    }
    m->prof_used = 0;
    return FLID_OK;
}

int flid_tgat_last_stats(const flid_tgat* m, int64_t stats[4]) {
    using namespace flid;
    FLID_REQUIRE(m && stats, "flid_tgat_last_stats: null argument");
    unsigned long long hv = 0;
    if (m->stats[1] < 0 && m->ws_misc.p) {
        FLID_CUDA(cudaDeviceSynchronize());
        FLID_CUDA(cudaMemcpy(&hv, m->ws_misc.p, sizeof(hv), cudaMemcpyDeviceToHost));
    }
    stats[0] = m->stats[0], stats[2] = m->stats[2];
    stats[1] = m->stats[1] < 0 ? (int64_t)hv * m->valid_mult : m->stats[1];
    const flid::DevBuf* bufs[] = {&m->table, &m->ws_ids, &m->ws_times, &m->ws_nbr, &m->ws_eid, &m->ws_dt, &m->ws_h,
                                  &m->ws_u,  &m->ws_z,   &m->ws_o,     &m->ws_a,   &m->ws_hd, &m->ws_pos};
    int64_t b = 0;
    for (auto* x : bufs) b += (int64_t)x->cap;
    stats[3] = b;
    return FLID_OK;
}

}  // extern "C"
