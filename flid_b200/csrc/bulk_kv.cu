// Projected bulk path of the attention layers (models/modules.py:183-231, models/TGAT.py:108-132).
//
// The x-space stream (attn_packed.cu) reads the raw rows [h_nbr | e] of every neighbour SLOT and scores /
// accumulates them against per-target folded vectors: 2 * H * (dn + de) multiply-adds per slot and an
// out-projection over H * kd = 888 inputs per target.  In a bulk pass (layer-memo build over every adjacency
// entry, E-step pass over every event) the same adjacency ENTRY p = (owner -> nbr, edge, ts) is a slot of up
// to k targets, and everything that depends on the entry alone can be projected once per pass:
//
//   K_p = Wk[:, :dn+de] [h_{l-1}(p) | e_p]      V_p = Wv[:, :dn+de] [h_{l-1}(p) | e_p]       (the reference's own
//   key / value projections, models/modules.py:191-197, hoisted from "per slot" to "per entry")
//
//   score_h(target, p) = qs_h . K_p,h + ut_h . te(dt)      qs = scale Wq [h_self | te0],  ut_h = scale Wk_t,h^T qs_h
//   out                = Wr [ sum_j a_hj V_pj,h ]_h + (Wr_h Wv_t,h) sum_j a_hj te_j + br
//
// At level 1 h_0(p) = node_feat[nbr_p] and the query depends on the owner node only, so the [h | e] part of
// the score is a per-entry constant s1[p, h] (one pass over the raw rows) and V_p = Vn[nbr_p] + Ve[eid_p] with a
// per-node and a per-edge table.  Per slot the stream then does qd (level >= 2: 2 qd) multiply-adds on the
// projected rows instead of 2 H (dn + de), the time-encoding part is unchanged, and the out-projection reads
// qd + H T = 472 inputs instead of 888.  Same algebra, different association: results agree with the x-space
// path and the oracle to fp32 rounding (tests), not bit for bit.
//
// Row layout of every projected vector (K, V, qs, the first qd columns of the stream output): position c holds
// projection output kv_perm(c) -- float4 chunk f belongs to head f % H -- so a lane of the stream kernel
// (chunks lane, lane + 32, ...) only ever touches head lane % H: one accumulator set per lane, no per-head
// select in the inner loop.
#include <stdlib.h>

#include <algorithm>

#include "attn.cuh"
#include "tgat.cuh"

namespace flid {

namespace {

typedef unsigned long long u64;

__host__ __device__ __forceinline__ int kv_perm(int c, int qd, int H) {
    const int hd = qd / H, full = (qd / 128) * 128;  // columns covered by whole rounds of 32 float4 chunks
    if (c < full) {
        const int f = c >> 2;
        return (f % H) * hd + (f / H) * 4 + (c & 3);
    }
    const int j = c - full;  // tail: one float per lane, head j % H
    return (j % H) * hd + full / H + j / H;
}

// ------------------------------------------------------------------ weight folds
// what: 0 wqs, 1 cqs, 2 wut, 3 cut, 4 wk2, 5 wv2, 6 wvn, 7 wve, 8 wo2
struct FoldArgs {
    const float *wq, *wk, *wv, *wr, *mfoldT, *wvoT, *te0, *u0;
    float *wqs, *cqs, *wut, *cut, *wk2, *wv2, *wvn, *wve, *wo2;
    int dn, de, T, H, qd, kd;
    double scale;
};

__global__ void kv_fold_kernel(FoldArgs a, int what, int64_t total) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int dn = a.dn, de = a.de, T = a.T, H = a.H, qd = a.qd, kd = a.kd, he = dn + de;
    switch (what) {
        case 0: {  // wqs[r, c] = scale * Wq[perm(r), c], c < dn
            const int r = (int)(idx / dn), c = (int)(idx % dn);
            a.wqs[idx] = (float)((double)a.wq[(int64_t)kv_perm(r, qd, H) * qd + c] * a.scale);
            break;
        }
        case 1: {  // cqs[r] = scale * sum_t Wq[perm(r), dn + t] te0[t]
            const int r = (int)idx;
            double s = 0.0;
            for (int t = 0; t < T; ++t) s += (double)a.wq[(int64_t)kv_perm(r, qd, H) * qd + dn + t] * (double)a.te0[t];
            a.cqs[idx] = (float)(s * a.scale);
            break;
        }
        case 2: {  // wut[h*T + t, c] = mfoldT[h*kd + he + t, c], c < dn
            const int r = (int)(idx / dn), c = (int)(idx % dn), h = r / T, t = r % T;
            a.wut[idx] = a.mfoldT[(int64_t)(h * kd + he + t) * qd + c];
            break;
        }
        case 3: {
            const int r = (int)idx, h = r / T, t = r % T;
            a.cut[idx] = a.u0[h * kd + he + t];
            break;
        }
        case 4:
        case 5: {  // wk2 / wv2 [r, c] = W[perm(r), c], c < dn + de
            const int r = (int)(idx / he), c = (int)(idx % he);
            const float* W = what == 4 ? a.wk : a.wv;
            (what == 4 ? a.wk2 : a.wv2)[idx] = W[(int64_t)kv_perm(r, qd, H) * kd + c];
            break;
        }
        case 6: {  // wvn[r, c] = Wv[perm(r), c], c < dn
            const int r = (int)(idx / dn), c = (int)(idx % dn);
            a.wvn[idx] = a.wv[(int64_t)kv_perm(r, qd, H) * kd + c];
            break;
        }
        case 7: {  // wve[r, c] = Wv[perm(r), dn + c], c < de
            const int r = (int)(idx / de), c = (int)(idx % de);
            a.wve[idx] = a.wv[(int64_t)kv_perm(r, qd, H) * kd + dn + c];
            break;
        }
        default: {  // wo2[o, c]: c < qd: Wr[o, perm(c)];  c = qd + h*T + t: wvoT[o, h*kd + he + t]
            const int P = qd + H * T;
            const int o = (int)(idx / P), c = (int)(idx % P);
            if (c < qd) {
                a.wo2[idx] = a.wr[(int64_t)o * qd + kv_perm(c, qd, H)];
            } else {
                const int h = (c - qd) / T, t = (c - qd) % T;
                a.wo2[idx] = a.wvoT[(int64_t)o * (H * kd) + h * kd + he + t];
            }
            break;
        }
    }
}

// ------------------------------------------------------------------ per-entry helpers
__global__ void kv_owner_kernel(const int64_t* __restrict__ indptr, int64_t num_nodes, int64_t M,
                                int32_t* __restrict__ owner) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= M) return;
    // largest v with indptr[v] <= p  (indptr has num_nodes + 2 entries, indptr[num_nodes + 1] = M)
    int64_t lo = 0, hi = num_nodes + 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(indptr + mid) <= p)
            lo = mid;
        else
            hi = mid - 1;
    }
    owner[p] = (int32_t)lo;
}

// out[0] = largest edge id, out[1] = 1 if some entry's neighbour is node 0 (the reference masks such a slot as
// padding, utils/utils.py:204 + models/modules.py:210: only the slot-mask kernels reproduce that)
__global__ void kv_max_eid_kernel(const int2* __restrict__ adj, int64_t M, int* __restrict__ out) {
    int m = 0, z = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        const int2 ne = __ldg(adj + i);
        m = max(m, ne.y), z |= (ne.x == 0);
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
    z = __any_sync(FULL, z);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, m);
        if (z) atomicMax(out + 1, 1);
    }
}

__global__ void kv_entry_eid_kernel(const int2* __restrict__ adj, int64_t M, int32_t* __restrict__ eid) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p <= M) eid[p] = p < M ? __ldg(adj + p).y : 0;  // row M: the padded slot (edge 0)
}

// level-1 score of entry p against its owner's folded query: s1[p, h] = u_h(owner)[:dn+de] . [nf[nbr_p] | ef[eid_p]]
// (already in the log2 domain).  One warp per entry; consecutive entries share the owner, so the u rows hit L1.
template <int H>
__global__ void __launch_bounds__(256) kv_score1_kernel(const int32_t* __restrict__ owner, const int2* __restrict__ adj,
                                                        int64_t M, int64_t lo, int64_t hi, const float* __restrict__ table,
                                                        const float* __restrict__ nf, const float* __restrict__ ef, int dn,
                                                        int de, int kd, float* __restrict__ s1) {
    const int lane = threadIdx.x & 31;
    int64_t p = lo + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (p > hi) return;
    if (p == hi) p = M;  // the work item after the range: the padded slot's row
    if (p == M) {  // padded slot: never read (padded slots are skipped unless every slot is padded, then scores are 0)
        if (lane < H) s1[p * H + lane] = 0.f;
        return;
    }
    const int2 ne = __ldg(adj + p);
    const float* u = table + (int64_t)__ldg(owner + p) * (H * kd);
    const float4* hr = reinterpret_cast<const float4*>(nf + (int64_t)ne.x * dn);
    const float4* er = reinterpret_cast<const float4*>(ef + (int64_t)ne.y * de);
    const int nv4 = dn >> 2, tot4 = (dn + de) >> 2;
    float acc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = 0.f;
    for (int f = lane; f < tot4; f += 32) {
        const float4 x = f < nv4 ? __ldg(hr + f) : __ldg(er + (f - nv4));
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(u + h * kd) + f);
            acc[h] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[h]))));
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = warp_sum(acc[h]);
    if (lane == 0) {
#pragma unroll
        for (int h = 0; h < H; ++h) s1[p * H + h] = acc[h];
    }
}

// ------------------------------------------------------------------ stream kernel
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ u64 pk1(float v) { return pk(v, v); }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float hsum(u64 v) {
    float lo, hi;
    upk(v, lo, hi);
    return lo + hi;
}
__device__ __forceinline__ float ex2(float x) {  // 2^x, rel. error 2^-22; ex2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed cosine of two arguments, both |x| < COS_FAST_LIMIT (see common.cuh cos_fast)
__device__ __forceinline__ u64 cos2_fast(u64 x) {
    const u64 t = fma2(x, pk1(0.31830987334251404f), pk1(12582912.0f));
    const u64 kf = add2(t, pk1(-12582912.0f));
    u64 r = fma2(kf, pk1(-3.1415927410125732f), x);
    r = fma2(kf, pk1(8.742277657347586e-08f), r);
    const u64 r2 = mul2(r, r);
    u64 p = fma2(r2, pk1(FLID_COS_C4), pk1(FLID_COS_C3));
    p = fma2(p, r2, pk1(FLID_COS_C2));
    p = fma2(p, r2, pk1(FLID_COS_C1));
    p = fma2(p, r2, pk1(FLID_COS_C0));
    float tl, th, pl, ph;
    upk(t, tl, th);
    upk(p, pl, ph);
    pl = __int_as_float(__float_as_int(pl) ^ (__float_as_int(tl) << 31));  // (-1)^k
    ph = __int_as_float(__float_as_int(ph) ^ (__float_as_int(th) << 31));
    return pk(pl, ph);
}
__device__ __forceinline__ u64 cos2_accurate(u64 x) {
    float lo, hi;
    upk(x, lo, hi);
    return pk(cos_accurate(lo), cos_accurate(hi));
}

template <int V>
__device__ __forceinline__ void reduce_bcast(float (&v)[V], int lane) {
    constexpr int LV = (V == 1) ? 0 : (V == 2) ? 1 : (V == 4) ? 2 : (V == 8) ? 3 : 4;
    static_assert(V == 1 || V == 2 || V == 4 || V == 8 || V == 16, "unsupported reduction width");
    int off = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1, off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = hi ? v[i] : v[i + n / 2];
            const float keep = hi ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, off);
        }
    }
#pragma unroll
    for (int o = 16 >> LV; o > 0; o >>= 1) v[0] += __shfl_xor_sync(FULL, v[0], o);
    const float total = v[0];
#pragma unroll
    for (int q = 0; q < V; ++q) v[q] = __shfl_sync(FULL, total, q << (5 - LV));
}

struct Row4 {  // one float4 chunk as two packed pairs
    u64 a, b;
};
__device__ __forceinline__ Row4 ldg_row4(const void* p) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    return Row4{v.x, v.y};
}

struct KvArgs {
    const float* q_base;     // per-target query-side rows
    int64_t q_stride;
    const int32_t* q_index;  // nullable: row of target i = q_base[q_index[i]]
    int ut_off, ut_hstride;  // ut_h = row + ut_off + h * ut_hstride   (qs = row[0 .. qd), MODE 2 only)
    const float* kv;         // MODE 2: [*, 2 qd] = [K | V] by CSR position
    const float *vn, *ve;    // MODE 1: [*, qd] by node id / edge id
    const float* s1;         // MODE 1: [*, H] by CSR position
    const int32_t *nbr, *eid, *pos;
    const float* dt;
    const float *time_w, *time_b, *time_bound;
    float* y;                // [n, qd + H T]
    int64_t n;
    int k, qd, T, tpw;       // tpw: targets per warp (a block walks 4 * tpw consecutive targets)
    // windowed kernel: the valid slots of a target are the `cnt` consecutive CSR positions before `cut`
    const int2* win;         // [n] (cut, cnt)
    const int2* adj;         // [M] (neighbour, edge) by position (MODE 1 row ids)
    int pad_pos;             // position standing for a padded slot (row of the projected tables)
    int ve_by_pos;           // MODE 1: rows of `ve` are indexed by CSR position instead of edge id
};

// MODE 1: level 1 (scores precomputed per entry, V = Vn[nbr] + Ve[eid]); MODE 2: level >= 2 ([K | V] per entry).
// NVF: whole rounds of 32 float4 chunks in a projected row (qd = 128 NVF + tail, tail <= 32 floats); TP: packed
// time-channel pairs per lane.  One warp per target at a time; the four warps of a block take consecutive targets
// (which, in owner-major / (node, time)-sorted order, share all but a few of their rows) and the block walks
// 4 * tpw of them, so the rows a target needs were fetched into this SM's L1 by its predecessors.
template <int H, int MODE, int NVF, int TP>
__global__ void __launch_bounds__(128, 4) attn_kv_kernel(KvArgs a) {
    extern __shared__ __align__(16) unsigned char q_smem[];
    constexpr int G = 2, V = G * H;
    constexpr int WARP_BYTES = NVF * 32 * 16 + 32 * 4 + H * TP * 32 * 8;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, hl = lane & (H - 1);
    const int k = a.k, qd = a.qd, T = a.T;
    const int full = NVF * 128, tail = qd - full;
    unsigned char* ws = q_smem + wib * WARP_BYTES;
    Row4* qf_s = reinterpret_cast<Row4*>(ws) + lane;                            // [NVF][32]
    float* qt_s = reinterpret_cast<float*>(ws + NVF * 32 * 16) + lane;          // [32]
    u64* ut_s = reinterpret_cast<u64*>(ws + NVF * 32 * 16 + 32 * 4) + lane;     // [H][TP][32]
    const float wmax = __ldg(a.time_bound), bmax = __ldg(a.time_bound + 1);
    u64 tw[TP], tb[TP];
#pragma unroll
    for (int r = 0; r < TP; ++r) {
        const int c0 = lane + 64 * r, c1 = c0 + 32;
        tw[r] = pk(c0 < T ? __ldg(a.time_w + c0) : 0.f, c1 < T ? __ldg(a.time_w + c1) : 0.f);
        tb[r] = pk(c0 < T ? __ldg(a.time_b + c0) : 0.f, c1 < T ? __ldg(a.time_b + c1) : 0.f);
    }
    // byte offsets of this lane's pieces inside a projected row (the tail lane index is clamped: lanes beyond the
    // tail read a valid float whose products are never stored)
    const int tl = tail > 0 ? min(lane, tail - 1) : 0;

    for (int jt = 0; jt < a.tpw; ++jt) {
        const int64_t i = (int64_t)blockIdx.x * (4 * a.tpw) + jt * 4 + wib;
        if (i >= a.n) break;  // warp-uniform
        const float* qrow = a.q_base + (a.q_index ? (int64_t)__ldg(a.q_index + i) : i) * a.q_stride;
        if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < NVF; ++r) qf_s[r * 32] = ldg_row4(qrow + 4 * (lane + 32 * r));
            qt_s[0] = (lane < tail) ? __ldg(qrow + full + lane) : 0.f;
        }
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int r = 0; r < TP; ++r) {
                const int c0 = lane + 64 * r, c1 = c0 + 32;
                const float* ut = qrow + a.ut_off + h * a.ut_hstride;
                ut_s[(h * TP + r) * 32] = pk(c0 < T ? __ldg(ut + c0) : 0.f, c1 < T ? __ldg(ut + c1) : 0.f);
            }
        // every lane only reads back what it wrote itself: no barrier needed

        bool all_masked = false;
        if (k > 32) {
            int any = 0;
            for (int base = 0; base < k; base += 32)
                any |= __any_sync(FULL, base + lane < k && __ldg(a.nbr + i * k + base + lane) != 0);
            all_masked = !any;
        }
        Row4 acc[NVF];
        float acc_t = 0.f;
        u64 acct[H][TP];
        float mx[H], den[H];
#pragma unroll
        for (int r = 0; r < NVF; ++r) acc[r] = Row4{0ull, 0ull};
#pragma unroll
        for (int h = 0; h < H; ++h) {
            mx[h] = -INFINITY, den[h] = 0.f;
#pragma unroll
            for (int r = 0; r < TP; ++r) acct[h][r] = 0ull;
        }
        float dt_l = 0.f;
        u64 addr0 = 0ull, addr1 = 0ull;  // MODE 1: Vn row / Ve row of this lane's slot; MODE 2: K row (V follows at + qd)
        int pos_l = 0;
        unsigned todo = 0u;

        struct Rows {
            Row4 x0[NVF], x1[NVF];  // MODE 1: Vn, Ve chunks; MODE 2: K, V chunks
            float t0, t1;           // tails
            float sv;               // MODE 1: s1 of head `lane` (lanes < H)
        };
        auto next_group = [&](int(&j)[G]) {
#pragma unroll
            for (int s = 0; s < G; ++s) {
                j[s] = -1;
                if (todo) {
                    j[s] = __ffs(todo) - 1;
                    todo &= todo - 1;
                }
            }
        };
        // Loads are unconditional: a missing second slot re-reads the first one's rows (its weight is ex2(-inf) = 0).
        auto load_group = [&](const int(&j)[G], Rows(&x)[G]) {
            if (j[0] < 0) return;  // warp-uniform: nothing left to prefetch
#pragma unroll
            for (int s = 0; s < G; ++s) {
                const int js = j[s] < 0 ? j[0] : j[s];
                const u64 p0 = __shfl_sync(FULL, addr0, js), p1 = __shfl_sync(FULL, addr1, js);
#pragma unroll
                for (int r = 0; r < NVF; ++r) {
                    x[s].x0[r] = ldg_row4(reinterpret_cast<const void*>(p0 + (u64)(lane + 32 * r) * 16ull));
                    x[s].x1[r] = ldg_row4(reinterpret_cast<const void*>(p1 + (u64)(lane + 32 * r) * 16ull));
                }
                x[s].t0 = __ldg(reinterpret_cast<const float*>(p0) + full + tl);
                x[s].t1 = __ldg(reinterpret_cast<const float*>(p1) + full + tl);
                if (MODE == 1) {
                    const int ps = __shfl_sync(FULL, pos_l, js);
                    x[s].sv = (lane < H) ? __ldg(a.s1 + (int64_t)ps * H + lane) : 0.f;
                }
            }
        };
        auto process = [&](const int(&j)[G], const Rows(&x)[G]) {
            const bool two = j[1] >= 0;
            float d[G];
            d[0] = __shfl_sync(FULL, dt_l, j[0]);
            d[1] = __shfl_sync(FULL, dt_l, two ? j[1] : j[0]);
            u64 xt[G][TP];
            const float amax = fmaf(fmaxf(fabsf(d[0]), fabsf(d[1])), wmax, bmax);
            if (amax < COS_FAST_LIMIT) {
#pragma unroll
                for (int s = 0; s < G; ++s)
#pragma unroll
                    for (int r = 0; r < TP; ++r) xt[s][r] = cos2_fast(fma2(pk1(d[s]), tw[r], tb[r]));
            } else {
#pragma unroll
                for (int s = 0; s < G; ++s)
#pragma unroll
                    for (int r = 0; r < TP; ++r) xt[s][r] = cos2_accurate(fma2(pk1(d[s]), tw[r], tb[r]));
            }
            float part[V];
            if (!all_masked) {
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    float own;  // this lane's share of the [h | e] score of slot s (head hl), or the precomputed score
                    if (MODE == 2) {
                        u64 p0 = 0ull, p1 = 0ull;
#pragma unroll
                        for (int r = 0; r < NVF; ++r) {
                            const Row4 q = qf_s[r * 32];
                            p0 = fma2(x[s].x0[r].a, q.a, p0);
                            p1 = fma2(x[s].x0[r].b, q.b, p1);
                        }
                        own = fmaf(x[s].t0, qt_s[0], hsum(add2(p0, p1)));
                    } else {
                        own = x[s].sv;
                    }
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        u64 p0 = 0ull;
#pragma unroll
                        for (int r = 0; r < TP; ++r) p0 = fma2(xt[s][r], ut_s[(h * TP + r) * 32], p0);
                        const bool mine = (MODE == 2) ? (h == hl) : (h == lane);
                        part[s * H + h] = hsum(p0) + (mine ? own : 0.f);
                    }
                }
                reduce_bcast<V>(part, lane);
            } else {
#pragma unroll
                for (int q = 0; q < V; ++q) part[q] = 0.f;  // all scores equal the -1e10 fill: uniform weights
            }
            if (!two) {
#pragma unroll
                for (int h = 0; h < H; ++h) part[H + h] = -INFINITY;  // weight ex2(-inf) = 0 on the re-read row
            }
            float w0[H], w1[H], corr_own = 1.f;
            bool raised = false;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float gmax = fmaxf(part[h], part[H + h]);
                if (gmax > mx[h]) {  // warp-uniform: raise the running max, rescale what was accumulated
                    const float corr = ex2(mx[h] - gmax);
                    const u64 c2 = pk1(corr);
                    mx[h] = gmax;
                    den[h] *= corr;
#pragma unroll
                    for (int r = 0; r < TP; ++r) acct[h][r] = mul2(acct[h][r], c2);
                    if (h == hl) corr_own = corr;
                    raised = true;
                }
                w0[h] = ex2(part[h] - mx[h]), w1[h] = ex2(part[H + h] - mx[h]);
                den[h] += w0[h] + w1[h];
                const u64 W0 = pk1(w0[h]), W1 = pk1(w1[h]);
#pragma unroll
                for (int r = 0; r < TP; ++r) acct[h][r] = fma2(W1, xt[1][r], fma2(W0, xt[0][r], acct[h][r]));
            }
            if (raised) {
                const u64 c2 = pk1(corr_own);
#pragma unroll
                for (int r = 0; r < NVF; ++r) acc[r].a = mul2(acc[r].a, c2), acc[r].b = mul2(acc[r].b, c2);
                acc_t *= corr_own;
            }
            float wo0 = w0[0], wo1 = w1[0];
#pragma unroll
            for (int h = 1; h < H; ++h)
                if (hl == h) wo0 = w0[h], wo1 = w1[h];
            const u64 W0 = pk1(wo0), W1 = pk1(wo1);
            if (MODE == 2) {
#pragma unroll
                for (int r = 0; r < NVF; ++r) {
                    acc[r].a = fma2(W1, x[1].x1[r].a, fma2(W0, x[0].x1[r].a, acc[r].a));
                    acc[r].b = fma2(W1, x[1].x1[r].b, fma2(W0, x[0].x1[r].b, acc[r].b));
                }
                acc_t = fmaf(wo1, x[1].t1, fmaf(wo0, x[0].t1, acc_t));
            } else {
#pragma unroll
                for (int r = 0; r < NVF; ++r) {
                    acc[r].a = fma2(W1, add2(x[1].x0[r].a, x[1].x1[r].a), fma2(W0, add2(x[0].x0[r].a, x[0].x1[r].a), acc[r].a));
                    acc[r].b = fma2(W1, add2(x[1].x0[r].b, x[1].x1[r].b), fma2(W0, add2(x[0].x0[r].b, x[0].x1[r].b), acc[r].b));
                }
                acc_t = fmaf(wo1, x[1].t0 + x[1].t1, fmaf(wo0, x[0].t0 + x[0].t1, acc_t));
            }
        };

        for (int base = 0; base < k; base += 32) {
            const int kb = min(32, k - base);
            int nb_l = 0;
            if (lane < kb) {
                nb_l = __ldg(a.nbr + i * k + base + lane);
                pos_l = __ldg(a.pos + i * k + base + lane);
                dt_l = __ldg(a.dt + i * k + base + lane);
                if (MODE == 1) {
                    const int e_l = __ldg(a.eid + i * k + base + lane);
                    addr0 = (u64)(a.vn + (int64_t)nb_l * qd);
                    addr1 = (u64)(a.ve + (int64_t)e_l * qd);
                } else {
                    addr0 = (u64)(a.kv + (int64_t)pos_l * (2 * qd));
                    addr1 = addr0 + (u64)qd * 4ull;
                }
            }
            const unsigned valid = __ballot_sync(FULL, lane < kb && nb_l != 0);
            if (k <= 32) all_masked = (valid == 0u);
            todo = all_masked ? (kb >= 32 ? FULL : ((1u << kb) - 1u)) : valid;
            if (todo == 0u) continue;  // warp-uniform
            int ja[G], jb[G];
            Rows xa[G], xb[G];
            next_group(ja);
            load_group(ja, xa);
            while (true) {
                next_group(jb);
                load_group(jb, xb);
                process(ja, xa);
                if (jb[0] < 0) break;
                next_group(ja);
                load_group(ja, xa);
                process(jb, xb);
                if (ja[0] < 0) break;
            }
        }

        const int P = qd + H * T;
        float* y = a.y + i * (int64_t)P;
        float inv_own = 1.0f / den[0];
#pragma unroll
        for (int h = 1; h < H; ++h)
            if (hl == h) inv_own = 1.0f / den[h];
        const u64 io = pk1(inv_own);
#pragma unroll
        for (int r = 0; r < NVF; ++r)
            reinterpret_cast<ulonglong2*>(y)[lane + 32 * r] = make_ulonglong2(mul2(acc[r].a, io), mul2(acc[r].b, io));
        if (lane < tail) y[full + lane] = acc_t * inv_own;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const u64 inv = pk1(1.0f / den[h]);
#pragma unroll
            for (int r = 0; r < TP; ++r) {
                const int c0 = lane + 64 * r, c1 = c0 + 32;
                float lo, hi;
                upk(mul2(acct[h][r], inv), lo, hi);
                if (c0 < T) y[qd + h * T + c0] = lo;
                if (c1 < T) y[qd + h * T + c1] = hi;
            }
        }
    }
}


// ------------------------------------------------------------------ windowed stream kernel
// Same arithmetic as attn_kv_kernel, organised around the fact that the valid slots of a target are CONSECUTIVE
// adjacency positions [cut - cnt, cut) (the 'recent' sampler takes the last cnt <= k entries before the cut,
// utils/utils.py:188-206): no slot bit mask, no per-slot address shuffles, rows of level >= 2 at a constant
// stride from a running pointer.  ncu on the mask-driven kernel showed ~140 warp instructions per slot of which
// a third was slot bookkeeping and 64-bit address traffic; the kernel is bound by instruction issue and dependent
// latencies (issue slots 51-57 % busy, L1 hit 73-82 %, DRAM 9-17 %), so instructions are what to remove.
// QDC / TC: compile-time row width and time dimension (0 = take them from the arguments).  With the reference's
// shape (qd = 272, T = 100) every row offset becomes an immediate: the generic instance spent ~60 of its ~266
// instructions per slot pair on 64-bit address arithmetic.
template <int H, int MODE, int NVF, int TP, int QDC, int TC, int MINB>
__global__ void __launch_bounds__(128, MINB) attn_win_kernel(KvArgs a) {
    extern __shared__ __align__(16) unsigned char q_smem[];
    constexpr int G = 2, V = G * H;
    constexpr int WARP_BYTES = NVF * 32 * 16 + 32 * 4 + H * TP * 32 * 8;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, hl = lane & (H - 1);
    const int k = a.k, qd = QDC ? QDC : a.qd, T = TC ? TC : a.T;
    const int full = NVF * 128, tail = qd - full;
    unsigned char* ws = q_smem + wib * WARP_BYTES;
    Row4* qf_s = reinterpret_cast<Row4*>(ws) + lane;                            // [NVF][32]
    float* qt_s = reinterpret_cast<float*>(ws + NVF * 32 * 16) + lane;          // [32]
    u64* ut_s = reinterpret_cast<u64*>(ws + NVF * 32 * 16 + 32 * 4) + lane;     // [H][TP][32]
    const float wmax = __ldg(a.time_bound), bmax = __ldg(a.time_bound + 1);
    u64 tw[TP], tb[TP];
#pragma unroll
    for (int r = 0; r < TP; ++r) {
        const int c0 = lane + 64 * r, c1 = c0 + 32;
        tw[r] = pk(c0 < T ? __ldg(a.time_w + c0) : 0.f, c1 < T ? __ldg(a.time_w + c1) : 0.f);
        tb[r] = pk(c0 < T ? __ldg(a.time_b + c0) : 0.f, c1 < T ? __ldg(a.time_b + c1) : 0.f);
    }
    const int tl = tail > 0 ? min(lane, tail - 1) : 0;  // lanes beyond the tail read a valid float; never stored
    const int64_t rowf = MODE == 2 ? 2 * (int64_t)qd : (int64_t)qd;  // floats per table row
    // this lane's view of the tables: chunk `lane` of row 0
    const float* kv_l = (MODE == 2 ? a.kv : a.vn) + 4 * lane;
    const float* ve_l = (MODE == 1 ? a.ve : a.kv) + 4 * lane;

    struct Rows {
        Row4 x0[NVF], x1[NVF];  // MODE 1: Vn, Ve chunks; MODE 2: K, V chunks
        float t0, t1;           // tails
        float sv[MODE == 1 ? H : 1];  // MODE 1: precomputed [h | e] score of the slot per head (same in every lane)
    };

    for (int jt = 0; jt < a.tpw; ++jt) {
        const int64_t i = (int64_t)blockIdx.x * (4 * a.tpw) + jt * 4 + wib;
        if (i >= a.n) break;  // warp-uniform
        const float* qrow = a.q_base + (a.q_index ? (int64_t)__ldg(a.q_index + i) : i) * a.q_stride;
        if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < NVF; ++r) qf_s[r * 32] = ldg_row4(qrow + 4 * (lane + 32 * r));
            qt_s[0] = (lane < tail) ? __ldg(qrow + full + lane) : 0.f;
        }
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int r = 0; r < TP; ++r) {
                const int c0 = lane + 64 * r, c1 = c0 + 32;
                const float* ut = qrow + a.ut_off + h * a.ut_hstride;
                ut_s[(h * TP + r) * 32] = pk(c0 < T ? __ldg(ut + c0) : 0.f, c1 < T ? __ldg(ut + c1) : 0.f);
            }
        const int2 wn = __ldg(a.win + i);
        // a target without neighbours: the reference's uniform weights over k identical padded rows == that row once
        const bool empty = wn.y == 0;
        const int cnt = empty ? 1 : wn.y;
        const int first = empty ? a.pad_pos : wn.x - wn.y;   // position of slot 0
        const int64_t dt0 = i * k + (empty ? k - 1 : k - wn.y);  // index of slot 0's dt

        Row4 acc[NVF];
        float acc_t = 0.f;
        u64 acct[H][TP];
        float mx[H], den[H];
#pragma unroll
        for (int r = 0; r < NVF; ++r) acc[r] = Row4{0ull, 0ull};
#pragma unroll
        for (int h = 0; h < H; ++h) {
            mx[h] = -INFINITY, den[h] = 0.f;
#pragma unroll
            for (int r = 0; r < TP; ++r) acct[h][r] = 0ull;
        }

        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int nb = min(32, cnt - j0);
            // lane-held per-slot scalars of this block of <= 32 slots
            float dt_l = 0.f;
            int nbr_l = 0, eid_l = 0;
            if (lane < nb) {
                dt_l = __ldg(a.dt + dt0 + j0 + lane);
                if (MODE == 1 && !empty) {
                    const int2 ne = __ldg(a.adj + first + j0 + lane);
                    nbr_l = ne.x, eid_l = a.ve_by_pos ? first + j0 + lane : ne.y;
                } else if (MODE == 1 && a.ve_by_pos) {
                    eid_l = a.pad_pos;
                }
            }
            auto load = [&](int g, Rows(&x)[G]) {
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    const int js = min(g + s, nb - 1);  // a missing second slot re-reads the first (weight 0)
                    const float *p0, *p1;
                    if (MODE == 2) {
                        p0 = kv_l + (int64_t)(first + j0 + js) * rowf;
                        p1 = p0 + qd;
                    } else {
                        p0 = kv_l + (int64_t)__shfl_sync(FULL, nbr_l, js) * rowf;
                        p1 = ve_l + (int64_t)__shfl_sync(FULL, eid_l, js) * rowf;
                    }
#pragma unroll
                    for (int r = 0; r < NVF; ++r) {
                        x[s].x0[r] = ldg_row4(p0 + 128 * r);
                        x[s].x1[r] = ldg_row4(p1 + 128 * r);
                    }
                    x[s].t0 = __ldg(p0 - 4 * lane + full + tl);
                    x[s].t1 = __ldg(p1 - 4 * lane + full + tl);
                    if (MODE == 1) {
                        const float* sp = a.s1 + (int64_t)(first + j0 + js) * H;  // uniform address: one wavefront
#pragma unroll
                        for (int h = 0; h < H; ++h) x[s].sv[h] = __ldg(sp + h);
                    }
                }
            };
            auto process = [&](int g, const Rows(&x)[G]) {
                const bool two = g + 1 < nb;
                float d[G];
                d[0] = __shfl_sync(FULL, dt_l, g);
                d[1] = __shfl_sync(FULL, dt_l, two ? g + 1 : g);
                u64 xt[G][TP];
                const float amax = fmaf(fmaxf(fabsf(d[0]), fabsf(d[1])), wmax, bmax);
                if (amax < COS_FAST_LIMIT) {
#pragma unroll
                    for (int s = 0; s < G; ++s)
#pragma unroll
                        for (int r = 0; r < TP; ++r) xt[s][r] = cos2_fast(fma2(pk1(d[s]), tw[r], tb[r]));
                } else {
#pragma unroll
                    for (int s = 0; s < G; ++s)
#pragma unroll
                        for (int r = 0; r < TP; ++r) xt[s][r] = cos2_accurate(fma2(pk1(d[s]), tw[r], tb[r]));
                }
                float part[V];
                if (!empty) {
#pragma unroll
                    for (int s = 0; s < G; ++s) {
                        float own = 0.f;
                        if (MODE == 2) {
                            u64 p0 = 0ull, p1 = 0ull;
#pragma unroll
                            for (int r = 0; r < NVF; ++r) {
                                const Row4 q = qf_s[r * 32];
                                p0 = fma2(x[s].x0[r].a, q.a, p0);
                                p1 = fma2(x[s].x0[r].b, q.b, p1);
                            }
                            own = fmaf(x[s].t0, qt_s[0], hsum(add2(p0, p1)));
                        }
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            u64 p0 = 0ull;
#pragma unroll
                            for (int r = 0; r < TP; ++r) p0 = fma2(xt[s][r], ut_s[(h * TP + r) * 32], p0);
                            part[s * H + h] = (MODE == 2 && h == hl) ? hsum(p0) + own : hsum(p0);
                        }
                    }
                    reduce_bcast<V>(part, lane);
                    if (MODE == 1) {  // + the precomputed [h | e] scores of the two slots
#pragma unroll
                        for (int h = 0; h < H; ++h) part[h] += x[0].sv[h], part[H + h] += x[1].sv[h];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < V; ++q) part[q] = 0.f;
                }
                if (!two) {
#pragma unroll
                    for (int h = 0; h < H; ++h) part[H + h] = -INFINITY;  // weight ex2(-inf) = 0 on the re-read row
                }
                float w0[H], w1[H], corr_own = 1.f;
                bool raised = false;
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float gmax = fmaxf(part[h], part[H + h]);
                    if (gmax > mx[h]) {  // warp-uniform: raise the running max, rescale what was accumulated
                        const float corr = ex2(mx[h] - gmax);
                        const u64 c2 = pk1(corr);
                        mx[h] = gmax;
                        den[h] *= corr;
#pragma unroll
                        for (int r = 0; r < TP; ++r) acct[h][r] = mul2(acct[h][r], c2);
                        if (h == hl) corr_own = corr;
                        raised = true;
                    }
                    w0[h] = ex2(part[h] - mx[h]), w1[h] = ex2(part[H + h] - mx[h]);
                    den[h] += w0[h] + w1[h];
                    const u64 W0 = pk1(w0[h]), W1 = pk1(w1[h]);
#pragma unroll
                    for (int r = 0; r < TP; ++r) acct[h][r] = fma2(W1, xt[1][r], fma2(W0, xt[0][r], acct[h][r]));
                }
                if (raised) {
                    const u64 c2 = pk1(corr_own);
#pragma unroll
                    for (int r = 0; r < NVF; ++r) acc[r].a = mul2(acc[r].a, c2), acc[r].b = mul2(acc[r].b, c2);
                    acc_t *= corr_own;
                }
                float wo0 = w0[0], wo1 = w1[0];
#pragma unroll
                for (int h = 1; h < H; ++h)
                    if (hl == h) wo0 = w0[h], wo1 = w1[h];
                const u64 W0 = pk1(wo0), W1 = pk1(wo1);
                if (MODE == 2) {
#pragma unroll
                    for (int r = 0; r < NVF; ++r) {
                        acc[r].a = fma2(W1, x[1].x1[r].a, fma2(W0, x[0].x1[r].a, acc[r].a));
                        acc[r].b = fma2(W1, x[1].x1[r].b, fma2(W0, x[0].x1[r].b, acc[r].b));
                    }
                    acc_t = fmaf(wo1, x[1].t1, fmaf(wo0, x[0].t1, acc_t));
                } else {
#pragma unroll
                    for (int r = 0; r < NVF; ++r) {
                        acc[r].a = fma2(W1, add2(x[1].x0[r].a, x[1].x1[r].a), fma2(W0, add2(x[0].x0[r].a, x[0].x1[r].a), acc[r].a));
                        acc[r].b = fma2(W1, add2(x[1].x0[r].b, x[1].x1[r].b), fma2(W0, add2(x[0].x0[r].b, x[0].x1[r].b), acc[r].b));
                    }
                    acc_t = fmaf(wo1, x[1].t0 + x[1].t1, fmaf(wo0, x[0].t0 + x[0].t1, acc_t));
                }
            };

            Rows xa[G], xb[G];
            int g = 0;
            load(0, xa);
            while (true) {
                if (g + 2 < nb) load(g + 2, xb);
                process(g, xa);
                g += 2;
                if (g >= nb) break;
                if (g + 2 < nb) load(g + 2, xa);
                process(g, xb);
                g += 2;
                if (g >= nb) break;
            }
        }

        const int P = qd + H * T;
        float* y = a.y + i * (int64_t)P;
        float inv_own = 1.0f / den[0];
#pragma unroll
        for (int h = 1; h < H; ++h)
            if (hl == h) inv_own = 1.0f / den[h];
        const u64 io = pk1(inv_own);
#pragma unroll
        for (int r = 0; r < NVF; ++r)
            reinterpret_cast<ulonglong2*>(y)[lane + 32 * r] = make_ulonglong2(mul2(acc[r].a, io), mul2(acc[r].b, io));
        if (lane < tail) y[full + lane] = acc_t * inv_own;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const u64 inv = pk1(1.0f / den[h]);
#pragma unroll
            for (int r = 0; r < TP; ++r) {
                const int c0 = lane + 64 * r, c1 = c0 + 32;
                float lo, hi;
                upk(mul2(acct[h][r], inv), lo, hi);
                if (c0 < T) y[qd + h * T + c0] = lo;
                if (c1 < T) y[qd + h * T + c1] = hi;
            }
        }
    }
}

template <int H, int MODE, int NVF>
int launch_kv(const KvArgs& a, cudaStream_t st) {
    constexpr int TP = 2;
    constexpr int WARP_BYTES = NVF * 32 * 16 + 32 * 4 + H * TP * 32 * 8;
    const unsigned blocks = (unsigned)ceil_div(a.n, 4 * (int64_t)a.tpw);
    static const int minb = [] {  // development knob: resident blocks per SM the windowed kernel is compiled for
        const char* e = getenv("FLID_KV_MINB");
        return (e && e[0] == '3') ? 3 : 4;
    }();
    if (a.win != nullptr) {
        if (NVF == 2 && a.qd == 272 && a.T == 100) {
            if (minb == 3)
                attn_win_kernel<H, MODE, NVF, TP, 272, 100, 3><<<blocks, 128, 4 * WARP_BYTES, st>>>(a);
            else
                attn_win_kernel<H, MODE, NVF, TP, 272, 100, 4><<<blocks, 128, 4 * WARP_BYTES, st>>>(a);
        } else {
            attn_win_kernel<H, MODE, NVF, TP, 0, 0, 4><<<blocks, 128, 4 * WARP_BYTES, st>>>(a);
        }
    } else
        attn_kv_kernel<H, MODE, NVF, TP><<<blocks, 128, 4 * WARP_BYTES, st>>>(a);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

template <int MODE>
int launch_kv_mode(const KvArgs& a, int H, int nvf, cudaStream_t st) {
#define FLID_KV_CASE(H_, N_) \
    if (H == H_ && nvf == N_) return launch_kv<H_, MODE, N_>(a, st);
    FLID_KV_CASE(1, 2) FLID_KV_CASE(2, 2) FLID_KV_CASE(4, 2) FLID_KV_CASE(1, 3) FLID_KV_CASE(2, 3) FLID_KV_CASE(4, 3)
#undef FLID_KV_CASE
    set_error("projected stream kernel: unsupported shape (heads %d, qd rounds %d)", H, nvf);
    return FLID_ERR_INVALID;
}

}  // namespace

// ---------------------------------------------------------------------------------- host side
bool kv_supported(const flid_tgat* m) {
    if (!m->kv_enabled || !m->use_tc) return false;
    const int nvf = m->qd / 128, tail = m->qd - nvf * 128;
    return (nvf == 2 || nvf == 3) && tail <= 32 && (tail % m->H) == 0 && (m->hd % 4) == 0 && m->T <= 128 &&
           (m->dn % 4) == 0 && (m->de % 4) == 0 && ((m->H * m->T) % 4) == 0;
}

void kv_free_layer(LayerDev& d) {
    cudaFree(d.kvw);
    d.kvw = nullptr;
    TcWeight* ws[] = {&d.tc_qs, &d.tc_ut, &d.tc_k2, &d.tc_v2, &d.tc_vn, &d.tc_ve, &d.tc_o2};
    for (auto* w : ws) tc_free_weight(w);
}

int kv_fold_layer(flid_tgat* m, int l, const float* wq, const float* wk, const float* wv, const float* wr, cudaStream_t st) {
    if (!kv_supported(m)) return FLID_OK;
    LayerDev& d = m->layers[l];
    const int dn = m->dn, de = m->de, T = m->T, H = m->H, qd = m->qd, kd = m->kd, he = dn + de, P = qd + H * T;
    const size_t n_wqs = (size_t)qd * dn, n_cqs = qd, n_wut = (size_t)H * T * dn, n_cut = (size_t)H * T,
                 n_wkv = (size_t)qd * he, n_wvn = (size_t)qd * dn, n_wve = (size_t)qd * de, n_wo2 = (size_t)qd * P;
    if (!d.kvw) {
        const size_t total = n_wqs + n_cqs + n_wut + n_cut + 2 * n_wkv + n_wvn + n_wve + n_wo2;
        FLID_CUDA(cudaMalloc((void**)&d.kvw, sizeof(float) * total));
        float* p = d.kvw;
        d.wqs = p, p += n_wqs, d.cqs = p, p += n_cqs, d.wut = p, p += n_wut, d.cut = p, p += n_cut;
        d.wk2 = p, p += n_wkv, d.wv2 = p, p += n_wkv, d.wvn = p, p += n_wvn, d.wve = p, p += n_wve, d.wo2 = p;
    }
    FoldArgs a;
    a.wq = wq, a.wk = wk, a.wv = wv, a.wr = wr, a.mfoldT = d.mfoldT, a.wvoT = d.wvoT, a.te0 = m->te0, a.u0 = d.u0;
    a.wqs = d.wqs, a.cqs = d.cqs, a.wut = d.wut, a.cut = d.cut, a.wk2 = d.wk2, a.wv2 = d.wv2, a.wvn = d.wvn, a.wve = d.wve,
    a.wo2 = d.wo2;
    a.dn = dn, a.de = de, a.T = T, a.H = H, a.qd = qd, a.kd = kd;
    // python: head_dim ** -0.5 is a float64; multiplying a float32 tensor by it uses its float32 value; log2(e)
    // because the stream evaluates softmax with exp2 (same constant as fold_qk_kernel's caller)
    a.scale = (double)(float)pow((double)m->hd, -0.5) * 1.4426950408889634074;
    const size_t counts[9] = {n_wqs, n_cqs, n_wut, n_cut, n_wkv, n_wkv, n_wvn, n_wve, n_wo2};
    for (int what = 0; what < 9; ++what) {
        kv_fold_kernel<<<(unsigned)ceil_div((int64_t)counts[what], 256), 256, 0, st>>>(a, what, (int64_t)counts[what]);
        FLID_LAUNCH_CHECK();
    }
    const int s = m->numeric;
    FLID_TRY(tc_prepare_weight(d.wqs, dn, qd, dn, &d.tc_qs, st, s));
    FLID_TRY(tc_prepare_weight(d.wut, dn, H * T, dn, &d.tc_ut, st, s));
    FLID_TRY(tc_prepare_weight(d.wk2, he, qd, he, &d.tc_k2, st, s));
    FLID_TRY(tc_prepare_weight(d.wv2, he, qd, he, &d.tc_v2, st, s));
    FLID_TRY(tc_prepare_weight(d.wvn, dn, qd, dn, &d.tc_vn, st, s));
    FLID_TRY(tc_prepare_weight(d.wve, de, qd, de, &d.tc_ve, st, s));
    FLID_TRY(tc_prepare_weight(d.wo2, P, qd, P, &d.tc_o2, st, s));
    return FLID_OK;
}

static int graph_derived(const flid_graph* gc, cudaStream_t st) {
    flid_graph* g = const_cast<flid_graph*>(gc);  // lazily built caches of an otherwise immutable graph
    const int64_t M = g->num_entries;
    if (!g->owner && M > 0) {
        FLID_CUDA(cudaMalloc((void**)&g->owner, sizeof(int32_t) * M));
        kv_owner_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(g->indptr, g->num_nodes, M, g->owner);
        FLID_LAUNCH_CHECK();
    }
    if (!g->ent_eid) {
        FLID_CUDA(cudaMalloc((void**)&g->ent_eid, sizeof(int32_t) * (M + 1)));
        kv_entry_eid_kernel<<<(unsigned)ceil_div(M + 1, 256), 256, 0, st>>>(g->adj, M, g->ent_eid);
        FLID_LAUNCH_CHECK();
    }
    if (g->max_eid < 0) {
        int* d = nullptr;
        int h[2] = {0, 0};
        FLID_CUDA(cudaMalloc((void**)&d, 2 * sizeof(int)));
        FLID_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(int), st));
        if (M > 0) {
            kv_max_eid_kernel<<<(unsigned)std::min<int64_t>(ceil_div(M, 256), 2048), 256, 0, st>>>(g->adj, M, d);
            FLID_LAUNCH_CHECK();
        }
        FLID_CUDA(cudaMemcpyAsync(h, d, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        FLID_CUDA(cudaStreamSynchronize(st));
        cudaFree(d);
        g->max_eid = h[0];
        g->zero_nbr = h[1];
    }
    return FLID_OK;
}

static void bulk_range(const flid_tgat* m, const flid_graph* g, int64_t* lo, int64_t* hi) {
    *lo = m->bulk_hi < 0 ? 0 : m->bulk_lo;
    *hi = m->bulk_hi < 0 ? g->num_entries : m->bulk_hi;
}

int kv_ensure_level1(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat, cudaStream_t st) {
    flid_tgat::KvKey key;
    key.wv = m->weights_version, key.epoch = m->bulk_epoch, key.g = g, key.nf = node_feat, key.ef = edge_feat;
    key.entries = g->num_entries;
    bulk_range(m, g, &key.lo, &key.hi);
    if (m->kv_l1_key == key && m->kv_s1.p) return FLID_OK;
    FLID_REQUIRE(m->table_src == node_feat && m->table_rows > 0, "projected bulk path: node table not cached");
    FLID_TRY(graph_derived(g, st));
    const LayerDev& ld = m->layers[0];
    const int64_t M = g->num_entries, rows_n = m->table_rows, lo = key.lo, hi = key.hi;
    const bool whole = lo == 0 && hi == M;
    // whole adjacency: one projected row per edge; a range (one rank's share): one per entry of the range, so that
    // a rank projects E / W edge rows instead of all E
    const int64_t rows_e = whole ? g->max_eid + 1 : M + 1;
    const int qd = m->qd;
    FLID_TRY(m->kv_vn1.reserve(sizeof(float) * (size_t)rows_n * qd));
    FLID_TRY(m->kv_ve1.reserve(sizeof(float) * (size_t)rows_e * qd));
    FLID_TRY(m->kv_s1.reserve(sizeof(float) * (size_t)(M + 1) * m->H));
    m->kv_ve_by_pos = !whole;
    {
        ProfScope prof(m, PROF_QFOLD, st);
        TcGemmArgs t;
        t.A0 = node_feat, t.lda0 = m->dn, t.w0 = m->dn, t.C = m->kv_vn1.as<float>(), t.ldc = qd, t.M = rows_n;
        FLID_TRY(tc_gemm(t, ld.tc_vn, st));
        TcGemmArgs e;
        e.A0 = edge_feat, e.lda0 = m->de, e.w0 = m->de, e.ldc = qd;
        if (whole) {
            e.C = m->kv_ve1.as<float>(), e.M = rows_e;
            FLID_TRY(tc_gemm(e, ld.tc_ve, st));
        } else {
            e.idx0 = g->ent_eid + lo, e.C = m->kv_ve1.as<float>() + lo * qd, e.M = hi - lo;
            FLID_TRY(tc_gemm(e, ld.tc_ve, st));
            e.idx0 = g->ent_eid + M, e.C = m->kv_ve1.as<float>() + M * qd, e.M = 1;  // the padded slot's row
            FLID_TRY(tc_gemm(e, ld.tc_ve, st));
        }
        const unsigned blocks = (unsigned)ceil_div((hi - lo + 1) * 32, 256);
        const float* table = m->table.as<float>();
        float* s1 = m->kv_s1.as<float>();
        switch (m->H) {
            case 1: kv_score1_kernel<1><<<blocks, 256, 0, st>>>(g->owner, g->adj, M, lo, hi, table, node_feat, edge_feat, m->dn, m->de, m->kd, s1); break;
            case 2: kv_score1_kernel<2><<<blocks, 256, 0, st>>>(g->owner, g->adj, M, lo, hi, table, node_feat, edge_feat, m->dn, m->de, m->kd, s1); break;
            default: kv_score1_kernel<4><<<blocks, 256, 0, st>>>(g->owner, g->adj, M, lo, hi, table, node_feat, edge_feat, m->dn, m->de, m->kd, s1); break;
        }
        FLID_LAUNCH_CHECK();
    }
    m->kv_l1_key = key;
    return FLID_OK;
}

int kv_ensure_level(flid_tgat* m, const flid_graph* g, int level, const float* memo_prev, const float* node_feat,
                    const float* edge_feat, cudaStream_t st) {
    FLID_REQUIRE(level >= 2 && level <= m->L && memo_prev, "projected bulk path: bad level %d", level);
    if ((int)m->kv_tab.size() < m->L - 1) m->kv_tab.resize(m->L - 1), m->kv_tab_key.resize(m->L - 1);
    flid_tgat::KvKey key;
    key.wv = m->weights_version, key.epoch = m->bulk_epoch, key.g = g, key.nf = node_feat, key.ef = edge_feat;
    key.memo = memo_prev, key.entries = g->num_entries;
    bulk_range(m, g, &key.lo, &key.hi);
    DevBuf& tab = m->kv_tab[level - 2];
    if (m->kv_tab_key[level - 2] == key && tab.p) return FLID_OK;
    FLID_TRY(graph_derived(g, st));
    const int64_t M = g->num_entries;
    const int qd = m->qd;
    FLID_TRY(tab.reserve(sizeof(float) * (size_t)(M + 1) * 2 * qd));
    const int32_t* ent_eid = g->ent_eid;  // edge id of every entry: the gathered second A segment
    ProfScope prof(m, PROF_QFOLD, st);
    const LayerDev& ld = m->layers[level - 1];
    // rows in chunks of full tile rounds (the GEMM's work counter is 32-bit)
    const int64_t chunk = (int64_t)148 * 128 * 4096;
    // rows [lo, hi) of the range and the padded slot's row M (contiguous with the range when it reaches the end)
    int64_t seg_lo[2] = {key.lo, M}, seg_n[2] = {key.hi - key.lo, 1};
    int segs = 2;
    if (key.hi == M) seg_n[0] += 1, segs = 1;
    for (int sg = 0; sg < segs; ++sg) {
        for (int64_t r0 = seg_lo[sg]; r0 < seg_lo[sg] + seg_n[sg]; r0 += chunk) {
            const int64_t nr = std::min(chunk, seg_lo[sg] + seg_n[sg] - r0);
            for (int half = 0; half < 2; ++half) {
                TcGemmArgs t;
                t.A0 = memo_prev + r0 * m->dn, t.lda0 = m->dn, t.w0 = m->dn;
                t.A1 = edge_feat, t.lda1 = m->de, t.idx1 = ent_eid + r0, t.w1 = m->de;
                t.C = tab.as<float>() + r0 * 2 * qd + half * qd, t.ldc = 2 * qd, t.M = nr;
                FLID_TRY(tc_gemm(t, half ? ld.tc_v2 : ld.tc_k2, st));
            }
        }
    }
    m->kv_tab_key[level - 2] = key;
    return FLID_OK;
}

int kv_attention(flid_tgat* m, const KvCall& c, int k, cudaStream_t st) {
    const int qd = m->qd, H = m->H, T = m->T, P = qd + H * T;
    KvArgs a;
    a.nbr = c.nbr, a.eid = c.eid, a.pos = c.pos, a.dt = c.dt;
    a.time_w = m->time_w, a.time_b = m->time_b, a.time_bound = m->time_bound;
    a.y = c.Y, a.n = c.n, a.k = k, a.qd = qd, a.T = T;
    a.kv = nullptr, a.vn = nullptr, a.ve = nullptr, a.s1 = nullptr;
    a.win = c.graph_zero_nbr ? nullptr : c.win, a.adj = c.adj, a.pad_pos = c.pad_pos;
    a.ve_by_pos = m->kv_ve_by_pos ? 1 : 0;
    if (c.level == 1 && m->kv_ve_by_pos) a.eid = c.pos;  // slot-mask kernel: edge rows by position as well
    // a block walks 4 * tpw consecutive targets; keep every SM busy on small calls
    const int64_t per = c.n / (148 * 4 * 4);
    a.tpw = per >= 8 ? 8 : (per >= 4 ? 4 : (per >= 2 ? 2 : 1));
    if (c.level == 1) {
        a.q_base = m->table.as<float>(), a.q_stride = (int64_t)H * m->kd, a.q_index = c.ids;
        a.ut_off = m->dn + m->de, a.ut_hstride = m->kd;
        a.vn = m->kv_vn1.as<float>(), a.ve = m->kv_ve1.as<float>(), a.s1 = m->kv_s1.as<float>();
        ProfScope prof(m, PROF_ATTN, st);
        return launch_kv_mode<1>(a, H, qd / 128, st);
    }
    const LayerDev& ld = m->layers[c.level - 1];
    {
        ProfScope prof(m, PROF_QFOLD, st);
        TcGemmArgs q;
        q.A0 = c.self_base, q.lda0 = m->dn, q.idx0 = c.self_idx, q.w0 = m->dn;
        q.C = c.U, q.ldc = P, q.bias = ld.cqs, q.M = c.n;
        FLID_TRY(tc_gemm(q, ld.tc_qs, st));
        TcGemmArgs u = q;
        u.C = c.U + qd, u.bias = ld.cut;
        FLID_TRY(tc_gemm(u, ld.tc_ut, st));
    }
    a.q_base = c.U, a.q_stride = P, a.q_index = nullptr;
    a.ut_off = qd, a.ut_hstride = T;
    a.kv = m->kv_tab[c.level - 2].as<float>();
    ProfScope prof(m, PROF_ATTN, st);
    return launch_kv_mode<2>(a, H, qd / 128, st);
}

}  // namespace flid

extern "C" int flid_tgat_set_bulk_projection(flid_tgat* m, int enable) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_set_bulk_projection: null handle");
    if (m->kv_enabled != (enable != 0)) m->have_weights = false;  // the projected weight images are built by set_weights
    m->kv_enabled = enable != 0;
    return FLID_OK;
}

extern "C" int flid_tgat_set_bulk_range(flid_tgat* m, int64_t pos_lo, int64_t pos_hi) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_set_bulk_range: null handle");
    FLID_REQUIRE(pos_hi < 0 || (pos_lo >= 0 && pos_lo <= pos_hi), "flid_tgat_set_bulk_range: bad range");
    m->bulk_lo = pos_hi < 0 ? 0 : pos_lo;
    m->bulk_hi = pos_hi;
    return FLID_OK;
}

extern "C" int flid_tgat_bulk_invalidate(flid_tgat* m) {
    using namespace flid;
    FLID_REQUIRE(m != nullptr, "flid_tgat_bulk_invalidate: null handle");
    m->bulk_epoch += 1;
    return FLID_OK;
}
