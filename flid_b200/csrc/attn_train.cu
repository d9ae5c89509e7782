// Training-mode attention stream: forward and hand-written backward.
//
// The M-step batches of the reference (PTCL/M_step.py:196-325, NPL/NPL.py:185-314) run the same
// MultiHeadAttention (models/modules.py:167-245) with autograd and dropout enabled.  After the
// re-association of DESIGN.md the only part of a layer that is not a dense GEMM is
//
//     x_j   = [ table[hrow_j] | edge_feat[eid_j] | cos(fma(dt_j, w, b)) ]          (k rows per target)
//     a_hj  = softmax_j( masked_fill(u_h . x_j, nbr_j == 0, -1e10) )               (modules.py:217-224)
//     z_h   = sum_j dropout(a)_hj x_j
//
// and this file is that function and its vector-Jacobian product: given dz it produces du, the
// scatter-added gradient of the table rows and the gradients of the time-encoder parameters.  The
// dense algebra around it (the query fold, value / residual projections, LayerNorm, MergeLayer)
// stays in torch so that autograd differentiates it (flid_b200/train.py).  Nothing of shape
// [n, k, 444] or [n, k, 272] is ever materialised, forward or backward: rows are gathered again.
//
// One warp per target.  Lane j owns slot j's scalars (score, probability, dropout bit); lane l
// owns float4 chunks l, l+32, ... of the concatenated [node | edge] row and time channels
// l, l+32, ...  Score dropout is counter based (Philox4x32-10 keyed by `seed`, counter = target,
// slot; word h = head h), so the backward pass regenerates the same bits and tests can export them.
#include "attn_train.cuh"

namespace flid {
namespace {

constexpr int MAX_TC = 4;  // time channels per lane: T <= 128

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
        key.x += 0x9E3779B9u, key.y += 0xBB67AE85u;
    }
    return c;
}

// dropout bits of (target, slot): bit h set = score of head h is KEPT
__device__ __forceinline__ unsigned keep_bits(uint64_t seed, int64_t target, int slot, int H, unsigned threshold) {
    if (threshold == 0u) return 0xFu;
    const uint4 r = philox4x32_10(make_uint4((unsigned)target, (unsigned)((uint64_t)target >> 32), (unsigned)slot, 0x464C6944u),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
    unsigned bits = 0;
#pragma unroll
    for (int h = 0; h < 4; ++h)
        if (h < H && w[h] >= threshold) bits |= 1u << h;
    return bits;
}

__host__ __device__ inline unsigned drop_threshold(float p) {  // P(word < threshold) = p
    if (!(p > 0.f)) return 0u;
    const double t = (double)p * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (unsigned)t;
}

// one slot's row chunks owned by this lane: x[r] = chunk lane+32r of [node | edge], te[r] = channel lane+32r
template <int NV>
__device__ __forceinline__ void load_row(const AttnTrainArgs& a, int lane, int64_t hrow, int eid, float dt, int nv4,
                                         int tot4, const float (&tw)[MAX_TC], const float (&tb)[MAX_TC],
                                         float4 (&x)[NV], float (&te)[MAX_TC]) {
    const float4* hp = reinterpret_cast<const float4*>(a.table + hrow * (int64_t)a.dn);
    const float4* ep = reinterpret_cast<const float4*>(a.edge_feat + (int64_t)eid * a.de) - nv4;
#pragma unroll
    for (int r = 0; r < NV; ++r) {
        const int f = lane + 32 * r;
        x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < tot4) x[r] = __ldg((f < nv4 ? hp : ep) + f);
    }
#pragma unroll
    for (int r = 0; r < MAX_TC; ++r) {
        const int c = lane + 32 * r;
        te[r] = (c < a.T) ? time_channel(dt, tw[r], tb[r]) : 0.f;
    }
}

__device__ __forceinline__ float dot4(const float4& p, const float4& q, float acc) {
    acc = fmaf(p.x, q.x, acc);
    acc = fmaf(p.y, q.y, acc);
    acc = fmaf(p.z, q.z, acc);
    return fmaf(p.w, q.w, acc);
}
__device__ __forceinline__ void axpy4(float s, const float4& p, float4& acc) {
    acc.x = fmaf(s, p.x, acc.x);
    acc.y = fmaf(s, p.y, acc.y);
    acc.z = fmaf(s, p.z, acc.z);
    acc.w = fmaf(s, p.w, acc.w);
}

// this lane's chunks of a per-target [H, kd] vector (u, dz)
template <int H, int NV>
__device__ __forceinline__ void load_vec(const float* v, int lane, int nv4, int tot4, int T, int kd, int row_w,
                                         float4 (&vx)[H][NV], float (&vt)[H][MAX_TC]) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            vx[h][r] = (f < tot4) ? __ldg(reinterpret_cast<const float4*>(v + h * kd) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < MAX_TC; ++r) {
            const int c = lane + 32 * r;
            vt[h][r] = (c < T) ? __ldg(v + h * kd + row_w + c) : 0.f;
        }
    }
}
template <int H, int NV>
__device__ __forceinline__ void store_vec(float* v, int lane, int tot4, int T, int kd, int row_w,
                                          const float4 (&vx)[H][NV], const float (&vt)[H][MAX_TC]) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            if (f < tot4) reinterpret_cast<float4*>(v + h * kd)[f] = vx[h][r];
        }
#pragma unroll
        for (int r = 0; r < MAX_TC; ++r) {
            const int c = lane + 32 * r;
            if (c < T) v[h * kd + row_w + c] = vt[h][r];
        }
    }
}

template <int H, int NV>
__global__ void __launch_bounds__(128) attn_train_fwd_kernel(AttnTrainArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= a.n) return;
    const int k = a.k, nv4 = a.dn >> 2, tot4 = (a.dn + a.de) >> 2, row_w = a.dn + a.de, kd = row_w + a.T;

    int64_t hrow_l = 0;
    int e_l = 0;
    float dt_l = 0.f;
    bool masked_l = true;
    if (lane < k) {
        hrow_l = a.hrow ? __ldg(a.hrow + i * k + lane) : a.hrow_offset + i * k + lane;
        e_l = (int)__ldg(a.eid + i * k + lane);
        dt_l = __ldg(a.dt + i * k + lane);
        masked_l = __ldg(a.nbr + i * k + lane) == 0;
    }
    const unsigned in_k = (k >= 32) ? FULL : ((1u << k) - 1u);
    const unsigned valid = __ballot_sync(FULL, !masked_l) & in_k;
    // (a target without any neighbour keeps -1e10 everywhere: the reference's uniform 1/k over its padded rows)

    float tw[MAX_TC], tb[MAX_TC];
#pragma unroll
    for (int r = 0; r < MAX_TC; ++r) {
        const int c = lane + 32 * r;
        tw[r] = c < a.T ? __ldg(a.time_w + c) : 0.f;
        tb[r] = c < a.T ? __ldg(a.time_b + c) : 0.f;
    }
    float4 ux[H][NV];
    float ut[H][MAX_TC];
    load_vec<H, NV>(a.u + i * (int64_t)(H * kd), lane, nv4, tot4, a.T, kd, row_w, ux, ut);

    // ---- scores: lane j keeps u_h . x_j of slot j (padded slots keep the -1e10 fill)
    float s[H];
#pragma unroll
    for (int h = 0; h < H; ++h) s[h] = -1e10f;
    for (unsigned todo = valid; todo; todo &= todo - 1) {
        const int j = __ffs(todo) - 1;
        float4 x[NV];
        float te[MAX_TC];
        load_row<NV>(a, lane, __shfl_sync(FULL, hrow_l, j), __shfl_sync(FULL, e_l, j), __shfl_sync(FULL, dt_l, j), nv4,
                     tot4, tw, tb, x, te);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float p = 0.f;
#pragma unroll
            for (int r = 0; r < NV; ++r) p = dot4(x[r], ux[h][r], p);
#pragma unroll
            for (int r = 0; r < MAX_TC; ++r) p = fmaf(te[r], ut[h][r], p);
            p = warp_sum(p);
            if (lane == j) s[h] = p;
        }
    }
    // ---- masked softmax over the k slots, score dropout (modules.py:217-224)
    const unsigned thr = drop_threshold(a.p_drop);
    const float keep_scale = 1.0f / (1.0f - a.p_drop);
    const unsigned kb = (lane < k) ? keep_bits(a.seed, i, lane, H, thr) : 0u;
    float ad[H];  // probability after dropout
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float sv = lane < k ? s[h] : -INFINITY;
        const float m = warp_max(sv);
        const float e = lane < k ? expf(sv - m) : 0.f;
        const float prob = e / warp_sum(e);
        if (lane < k) a.probs[(i * H + h) * k + lane] = prob;
        ad[h] = ((kb >> h) & 1u) ? prob * keep_scale : 0.f;
    }
    // ---- z_h = sum_j ad_hj x_j over the slots with a non-zero weight
    bool live = false;
#pragma unroll
    for (int h = 0; h < H; ++h) live |= ad[h] != 0.f;
    float4 zx[H][NV];
    float zt[H][MAX_TC];
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int r = 0; r < NV; ++r) zx[h][r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < MAX_TC; ++r) zt[h][r] = 0.f;
    }
    for (unsigned todo = __ballot_sync(FULL, live) & in_k; todo; todo &= todo - 1) {
        const int j = __ffs(todo) - 1;
        float4 x[NV];
        float te[MAX_TC];
        load_row<NV>(a, lane, __shfl_sync(FULL, hrow_l, j), __shfl_sync(FULL, e_l, j), __shfl_sync(FULL, dt_l, j), nv4,
                     tot4, tw, tb, x, te);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float w = __shfl_sync(FULL, ad[h], j);
#pragma unroll
            for (int r = 0; r < NV; ++r) axpy4(w, x[r], zx[h][r]);
#pragma unroll
            for (int r = 0; r < MAX_TC; ++r) zt[h][r] = fmaf(w, te[r], zt[h][r]);
        }
    }
    store_vec<H, NV>(a.z + i * (int64_t)(H * kd), lane, tot4, a.T, kd, row_w, zx, zt);
}

// Backward.  With a = saved probabilities, m = dropout keep / (1 - p), ad = a m:
//   dad_hj = dz_h . x_j                  da_hj = m_hj dad_hj
//   ds_hj  = a_hj (da_hj - sum_j' a_hj' da_hj')      only for unmasked slots: masked_fill passes no gradient
//   du_h   = sum_j ds_hj x_j
//   dx_j   = sum_h (ad_hj dz_h + ds_hj u_h)          -> table rows (atomic add), time encoder (-sin chain rule)
template <int H, int NV>
__global__ void __launch_bounds__(128) attn_train_bwd_kernel(AttnTrainArgs a, AttnTrainGrads g) {
    __shared__ float red[4][2 * 32 * MAX_TC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const bool active = i < a.n;
    const int k = a.k, nv4 = a.dn >> 2, tot4 = (a.dn + a.de) >> 2, row_w = a.dn + a.de, kd = row_w + a.T;
    float gw[MAX_TC], gb[MAX_TC];
#pragma unroll
    for (int r = 0; r < MAX_TC; ++r) gw[r] = gb[r] = 0.f;

    if (active) {
        int64_t hrow_l = 0;
        int e_l = 0;
        float dt_l = 0.f;
        bool masked_l = true;
        if (lane < k) {
            hrow_l = a.hrow ? __ldg(a.hrow + i * k + lane) : a.hrow_offset + i * k + lane;
            e_l = (int)__ldg(a.eid + i * k + lane);
            dt_l = __ldg(a.dt + i * k + lane);
            masked_l = __ldg(a.nbr + i * k + lane) == 0;
        }
        const unsigned in_k = (k >= 32) ? FULL : ((1u << k) - 1u);
        const unsigned valid = __ballot_sync(FULL, !masked_l) & in_k;

        float tw[MAX_TC], tb[MAX_TC];
#pragma unroll
        for (int r = 0; r < MAX_TC; ++r) {
            const int c = lane + 32 * r;
            tw[r] = c < a.T ? __ldg(a.time_w + c) : 0.f;
            tb[r] = c < a.T ? __ldg(a.time_b + c) : 0.f;
        }
        float4 dzx[H][NV];
        float dzt[H][MAX_TC];
        load_vec<H, NV>(g.dz + i * (int64_t)(H * kd), lane, nv4, tot4, a.T, kd, row_w, dzx, dzt);

        const unsigned thr = drop_threshold(a.p_drop);
        const float keep_scale = 1.0f / (1.0f - a.p_drop);
        const unsigned kb = (lane < k) ? keep_bits(a.seed, i, lane, H, thr) : 0u;
        float prob[H], ad[H], ds[H];
        bool live = false;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            prob[h] = lane < k ? __ldg(g.probs + (i * H + h) * k + lane) : 0.f;
            ad[h] = ((kb >> h) & 1u) ? prob[h] * keep_scale : 0.f;
            ds[h] = 0.f;
            live |= ad[h] != 0.f;
        }
        const unsigned live_slots = __ballot_sync(FULL, live) & in_k;

        // ---- pass 1 (only if some slot is unmasked): dad_hj for the live unmasked slots
        if (valid) {
            float da[H];
#pragma unroll
            for (int h = 0; h < H; ++h) da[h] = 0.f;
            for (unsigned todo = live_slots & valid; todo; todo &= todo - 1) {
                const int j = __ffs(todo) - 1;
                float4 x[NV];
                float te[MAX_TC];
                load_row<NV>(a, lane, __shfl_sync(FULL, hrow_l, j), __shfl_sync(FULL, e_l, j), __shfl_sync(FULL, dt_l, j),
                             nv4, tot4, tw, tb, x, te);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float p = 0.f;
#pragma unroll
                    for (int r = 0; r < NV; ++r) p = dot4(x[r], dzx[h][r], p);
#pragma unroll
                    for (int r = 0; r < MAX_TC; ++r) p = fmaf(te[r], dzt[h][r], p);
                    p = warp_sum(p);
                    if (lane == j) da[h] = ((kb >> h) & 1u) ? p * keep_scale : 0.f;
                }
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float dot = warp_sum(prob[h] * da[h]);
                ds[h] = masked_l ? 0.f : prob[h] * (da[h] - dot);
            }
        }
        bool touch = live;
#pragma unroll
        for (int h = 0; h < H; ++h) touch |= ds[h] != 0.f;

        // ---- pass 2: du and the row / time-encoder gradients
        float4 ux[H][NV], dux[H][NV];
        float ut[H][MAX_TC], dut[H][MAX_TC];
        load_vec<H, NV>(a.u + i * (int64_t)(H * kd), lane, nv4, tot4, a.T, kd, row_w, ux, ut);
#pragma unroll
        for (int h = 0; h < H; ++h) {
#pragma unroll
            for (int r = 0; r < NV; ++r) dux[h][r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < MAX_TC; ++r) dut[h][r] = 0.f;
        }
        for (unsigned todo = __ballot_sync(FULL, touch) & in_k; todo; todo &= todo - 1) {
            const int j = __ffs(todo) - 1;
            const int64_t hrow = __shfl_sync(FULL, hrow_l, j);
            const float dt = __shfl_sync(FULL, dt_l, j);
            float4 x[NV];
            float te[MAX_TC];
            load_row<NV>(a, lane, hrow, __shfl_sync(FULL, e_l, j), dt, nv4, tot4, tw, tb, x, te);
            float4 dx[NV];
            float dte[MAX_TC];
#pragma unroll
            for (int r = 0; r < NV; ++r) dx[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < MAX_TC; ++r) dte[r] = 0.f;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float w = __shfl_sync(FULL, ad[h], j), gs = __shfl_sync(FULL, ds[h], j);
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    axpy4(gs, x[r], dux[h][r]);
                    axpy4(w, dzx[h][r], dx[r]);
                    axpy4(gs, ux[h][r], dx[r]);
                }
#pragma unroll
                for (int r = 0; r < MAX_TC; ++r) {
                    dut[h][r] = fmaf(gs, te[r], dut[h][r]);
                    dte[r] = fmaf(w, dzt[h][r], fmaf(gs, ut[h][r], dte[r]));
                }
            }
            if (g.dtable) {
                float4* drow = reinterpret_cast<float4*>(g.dtable + hrow * (int64_t)a.dn);
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    const int f = lane + 32 * r;
                    if (f < nv4) atomicAdd(drow + f, dx[r]);
                }
            }
            if (g.dtime_partial) {
#pragma unroll
                for (int r = 0; r < MAX_TC; ++r) {
                    if (lane + 32 * r < a.T) {
                        const float sn = -sin_accurate(fmaf(dt, tw[r], tb[r])) * dte[r];  // d cos(arg) = -sin(arg) d arg
                        gw[r] = fmaf(sn, dt, gw[r]);
                        gb[r] += sn;
                    }
                }
            }
        }
        store_vec<H, NV>(g.du + i * (int64_t)(H * kd), lane, tot4, a.T, kd, row_w, dux, dut);
    }
    // ---- per-block partial sums of the time-encoder gradients: [gridDim.x][2][T]
    if (g.dtime_partial) {
#pragma unroll
        for (int r = 0; r < MAX_TC; ++r) {
            red[warp][lane + 32 * r] = gw[r];
            red[warp][32 * MAX_TC + lane + 32 * r] = gb[r];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < 2 * a.T; c += blockDim.x) {
            const int q = (c < a.T) ? c : 32 * MAX_TC + (c - a.T);
            g.dtime_partial[(int64_t)blockIdx.x * 2 * a.T + c] = red[0][q] + red[1][q] + red[2][q] + red[3][q];
        }
    }
}

// Backward when no table gradient is wanted (TGAT layer 1: raw node features are constants) -- ONE pass over
// the neighbour rows instead of two.  With S_h = sum_j a_hj da_hj the softmax backward ds_hj = a_hj (da_hj - S_h)
// separates into sums that can be accumulated before S_h is known:
//   du_h = sum_j (a_hj da_hj) x_j - S_h sum_j a_hj x_j
//   d(w,b)_c = sum_j s_jc (dt_j, 1) sum_h (ad_hj dz_hc + a_hj da_hj u_hc)  -  sum_h S_h u_hc sum_j s_jc (dt_j, 1) a_hj
// (s_jc = -sin(arg_jc); a target without neighbours has da = 0, masked slots have a = 0).
template <int H, int NV>
__global__ void __launch_bounds__(128) attn_train_bwd_onepass_kernel(AttnTrainArgs a, AttnTrainGrads g) {
    __shared__ float red[4][2 * 32 * MAX_TC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int k = a.k, nv4 = a.dn >> 2, tot4 = (a.dn + a.de) >> 2, row_w = a.dn + a.de, kd = row_w + a.T;
    float gw[MAX_TC], gb[MAX_TC];
#pragma unroll
    for (int r = 0; r < MAX_TC; ++r) gw[r] = gb[r] = 0.f;

    if (i < a.n) {
        int64_t hrow_l = 0;
        int e_l = 0;
        float dt_l = 0.f;
        bool masked_l = true;
        if (lane < k) {
            hrow_l = a.hrow ? __ldg(a.hrow + i * k + lane) : a.hrow_offset + i * k + lane;
            e_l = (int)__ldg(a.eid + i * k + lane);
            dt_l = __ldg(a.dt + i * k + lane);
            masked_l = __ldg(a.nbr + i * k + lane) == 0;
        }
        const unsigned in_k = (k >= 32) ? FULL : ((1u << k) - 1u);
        const unsigned valid = __ballot_sync(FULL, !masked_l) & in_k;
        const bool all_masked = valid == 0u;
        float tw[MAX_TC], tb[MAX_TC];
#pragma unroll
        for (int r = 0; r < MAX_TC; ++r) {
            const int c = lane + 32 * r;
            tw[r] = c < a.T ? __ldg(a.time_w + c) : 0.f;
            tb[r] = c < a.T ? __ldg(a.time_b + c) : 0.f;
        }
        float4 dzx[H][NV], px[H][NV], qx[H][NV];
        float dzt[H][MAX_TC], pt[H][MAX_TC], qt[H][MAX_TC], ut[H][MAX_TC], rw[H][MAX_TC], rb[H][MAX_TC];
        load_vec<H, NV>(g.dz + i * (int64_t)(H * kd), lane, nv4, tot4, a.T, kd, row_w, dzx, dzt);
        const float* u = a.u + i * (int64_t)(H * kd);
        float S[H], prob[H];
        const unsigned thr = drop_threshold(a.p_drop);
        const float keep_scale = 1.0f / (1.0f - a.p_drop);
        const unsigned kb = (lane < k) ? keep_bits(a.seed, i, lane, H, thr) : 0u;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            S[h] = 0.f;
            prob[h] = lane < k ? __ldg(g.probs + (i * H + h) * k + lane) : 0.f;
#pragma unroll
            for (int r = 0; r < NV; ++r) px[h][r] = qx[h][r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < MAX_TC; ++r) {
                const int c = lane + 32 * r;
                pt[h][r] = qt[h][r] = rw[h][r] = rb[h][r] = 0.f;
                ut[h][r] = c < a.T ? __ldg(u + h * kd + row_w + c) : 0.f;
            }
        }
        for (unsigned todo = all_masked ? in_k : valid; todo; todo &= todo - 1) {
            const int j = __ffs(todo) - 1;
            const float dt = __shfl_sync(FULL, dt_l, j);
            const unsigned kbj = __shfl_sync(FULL, kb, j);
            float4 x[NV];
            float te[MAX_TC];
            load_row<NV>(a, lane, __shfl_sync(FULL, hrow_l, j), __shfl_sync(FULL, e_l, j), dt, nv4, tot4, tw, tb, x, te);
            float e[MAX_TC];
#pragma unroll
            for (int r = 0; r < MAX_TC; ++r) e[r] = 0.f;
            float ah[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float d = 0.f;
#pragma unroll
                for (int r = 0; r < NV; ++r) d = dot4(x[r], dzx[h][r], d);
#pragma unroll
                for (int r = 0; r < MAX_TC; ++r) d = fmaf(te[r], dzt[h][r], d);
                d = warp_sum(d);
                const bool kept = (kbj >> h) & 1u;
                const float av = __shfl_sync(FULL, prob[h], j);
                const float da = (kept && !all_masked) ? d * keep_scale : 0.f;
                const float ad = kept ? av * keep_scale : 0.f;
                const float c1 = av * da;
                ah[h] = av;
                S[h] += c1;
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    axpy4(c1, x[r], px[h][r]);
                    axpy4(av, x[r], qx[h][r]);
                }
#pragma unroll
                for (int r = 0; r < MAX_TC; ++r) {
                    pt[h][r] = fmaf(c1, te[r], pt[h][r]);
                    qt[h][r] = fmaf(av, te[r], qt[h][r]);
                    e[r] = fmaf(ad, dzt[h][r], fmaf(c1, ut[h][r], e[r]));
                }
            }
            if (g.dtime_partial) {
#pragma unroll
                for (int r = 0; r < MAX_TC; ++r) {
                    if (lane + 32 * r < a.T) {
                        const float sn = -sin_accurate(fmaf(dt, tw[r], tb[r]));
                        gw[r] = fmaf(sn * dt, e[r], gw[r]);
                        gb[r] = fmaf(sn, e[r], gb[r]);
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            rw[h][r] = fmaf(sn * dt, ah[h], rw[h][r]);
                            rb[h][r] = fmaf(sn, ah[h], rb[h][r]);
                        }
                    }
                }
            }
        }
        // du = P - S Q ; a target without neighbours passes no score gradient
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float s = all_masked ? 0.f : S[h];
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                px[h][r].x = all_masked ? 0.f : fmaf(-s, qx[h][r].x, px[h][r].x);
                px[h][r].y = all_masked ? 0.f : fmaf(-s, qx[h][r].y, px[h][r].y);
                px[h][r].z = all_masked ? 0.f : fmaf(-s, qx[h][r].z, px[h][r].z);
                px[h][r].w = all_masked ? 0.f : fmaf(-s, qx[h][r].w, px[h][r].w);
            }
#pragma unroll
            for (int r = 0; r < MAX_TC; ++r) {
                pt[h][r] = all_masked ? 0.f : fmaf(-s, qt[h][r], pt[h][r]);
                gw[r] = fmaf(-s * ut[h][r], rw[h][r], gw[r]);
                gb[r] = fmaf(-s * ut[h][r], rb[h][r], gb[r]);
            }
        }
        store_vec<H, NV>(g.du + i * (int64_t)(H * kd), lane, tot4, a.T, kd, row_w, px, pt);
    }
    if (g.dtime_partial) {
#pragma unroll
        for (int r = 0; r < MAX_TC; ++r) {
            red[warp][lane + 32 * r] = gw[r];
            red[warp][32 * MAX_TC + lane + 32 * r] = gb[r];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < 2 * a.T; c += blockDim.x) {
            const int q = (c < a.T) ? c : 32 * MAX_TC + (c - a.T);
            g.dtime_partial[(int64_t)blockIdx.x * 2 * a.T + c] = red[0][q] + red[1][q] + red[2][q] + red[3][q];
        }
    }
}

__global__ void keep_mask_kernel(uint64_t seed, int64_t n, int H, int k, float p, uint8_t* keep) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const int64_t i = t / k;
    const int j = (int)(t % k);
    const unsigned kb = keep_bits(seed, i, j, H, drop_threshold(p));
    for (int h = 0; h < H; ++h) keep[(i * H + h) * k + j] = (kb >> h) & 1u;
}

int check_args(const AttnTrainArgs& a, int H) {
    FLID_REQUIRE(a.n >= 0 && a.k > 0 && a.k <= 32, "attn_train: num_neighbors must be in 1..32 (got %d)", a.k);
    FLID_REQUIRE(H == 1 || H == 2 || H == 4, "attn_train: num_heads must be 1, 2 or 4 (got %d)", H);
    FLID_REQUIRE(a.dn > 0 && a.de > 0 && a.dn % 4 == 0 && a.de % 4 == 0, "attn_train: feature widths must be multiples of 4");
    FLID_REQUIRE(a.T > 0 && a.T <= 32 * MAX_TC, "attn_train: time_feat_dim must be <= %d", 32 * MAX_TC);
    FLID_REQUIRE((a.dn + a.de) / 4 <= 32 * 6, "attn_train: node + edge width must be <= 768");
    FLID_REQUIRE(a.T % 4 == 0, "attn_train: time_feat_dim must be a multiple of 4");
    FLID_REQUIRE(a.p_drop >= 0.f && a.p_drop < 1.f, "attn_train: dropout probability must be in [0, 1)");
    return FLID_OK;
}

}  // namespace

int64_t attn_train_blocks(int64_t n) { return ceil_div(n * 32, 128); }

#define FLID_TRAIN_DISPATCH(KERNEL, ...)                                              \
    do {                                                                              \
        const unsigned blocks = (unsigned)attn_train_blocks(a.n);                     \
        const int nv = (int)ceil_div((a.dn + a.de) / 4, 32);                          \
        if (nv <= 3) {                                                                \
            if (H == 1) KERNEL<1, 3><<<blocks, 128, 0, st>>>(__VA_ARGS__);            \
            if (H == 2) KERNEL<2, 3><<<blocks, 128, 0, st>>>(__VA_ARGS__);            \
            if (H == 4) KERNEL<4, 3><<<blocks, 128, 0, st>>>(__VA_ARGS__);            \
        } else {                                                                      \
            if (H == 1) KERNEL<1, 6><<<blocks, 128, 0, st>>>(__VA_ARGS__);            \
            if (H == 2) KERNEL<2, 6><<<blocks, 128, 0, st>>>(__VA_ARGS__);            \
            if (H == 4) KERNEL<4, 6><<<blocks, 128, 0, st>>>(__VA_ARGS__);            \
        }                                                                             \
        FLID_LAUNCH_CHECK();                                                          \
    } while (0)

int launch_attn_train_fwd(const AttnTrainArgs& a, int H, cudaStream_t st) {
    FLID_TRY(check_args(a, H));
    if (a.n == 0) return FLID_OK;
    FLID_TRAIN_DISPATCH(attn_train_fwd_kernel, a);
    return FLID_OK;
}

int launch_attn_train_bwd(const AttnTrainArgs& a, const AttnTrainGrads& g, int H, cudaStream_t st) {
    FLID_TRY(check_args(a, H));
    if (a.n == 0) return FLID_OK;
    if (g.dtable == nullptr)
        FLID_TRAIN_DISPATCH(attn_train_bwd_onepass_kernel, a, g);
    else
        FLID_TRAIN_DISPATCH(attn_train_bwd_kernel, a, g);
    return FLID_OK;
}

int launch_keep_mask(uint64_t seed, int64_t n, int H, int k, float p, uint8_t* keep, cudaStream_t st) {
    if (n <= 0) return FLID_OK;
    keep_mask_kernel<<<(unsigned)ceil_div(n * k, 256), 256, 0, st>>>(seed, n, H, k, p, keep);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

}  // namespace flid

using namespace flid;

extern "C" int64_t flid_attn_train_partials(int64_t n) { return attn_train_blocks(n); }

extern "C" int flid_attn_train_fwd(const float* u, const float* table, const int64_t* hrow, const int64_t* nbr,
                                   const int64_t* eid, const float* dt, const float* edge_feat, const float* time_w,
                                   const float* time_b, int64_t n, int k, int num_heads, int node_dim, int edge_dim,
                                   int time_dim, float p_drop, uint64_t seed, float* z, float* probs,
                                   flid_stream stream) {
    AttnTrainArgs a{u, table, hrow, nbr, eid, dt, edge_feat, time_w, time_b, n, k, node_dim, edge_dim, time_dim, p_drop, seed, z, probs};
    return launch_attn_train_fwd(a, num_heads, (cudaStream_t)stream);
}

extern "C" int flid_attn_train_bwd(const float* u, const float* table, const int64_t* hrow, const int64_t* nbr,
                                   const int64_t* eid, const float* dt, const float* edge_feat, const float* time_w,
                                   const float* time_b, int64_t n, int k, int num_heads, int node_dim, int edge_dim,
                                   int time_dim, float p_drop, uint64_t seed, const float* probs, const float* dz,
                                   float* du, float* dtable, float* dtime_partial, flid_stream stream) {
    AttnTrainArgs a{u, table, hrow, nbr, eid, dt, edge_feat, time_w, time_b, n, k, node_dim, edge_dim, time_dim, p_drop, seed, nullptr, nullptr};
    AttnTrainGrads g{probs, dz, du, dtable, dtime_partial};
    return launch_attn_train_bwd(a, g, num_heads, (cudaStream_t)stream);
}

extern "C" int flid_attn_train_keep_mask(uint64_t seed, int64_t n, int num_heads, int k, float p_drop, uint8_t* keep,
                                         flid_stream stream) {
    FLID_REQUIRE(num_heads >= 1 && num_heads <= 4 && k > 0 && k <= 32, "keep_mask: unsupported shape");
    return launch_keep_mask(seed, n, num_heads, k, p_drop, keep, (cudaStream_t)stream);
}
