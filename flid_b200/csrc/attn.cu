// Attention stream kernel (models/modules.py:183-231 after the re-association of DESIGN.md):
// for every target, gather its <=k neighbour rows [h_nbr | e] (16-byte loads), compute the
// time encoding cos(fma(dt, w, b)) in registers, score each row against the folded query
// u_h (scores already in the log2 domain), masked softmax, and accumulate
// z_h = sum_j a_hj [h_nbr_j | e_j | te_j].  HBM-bound by design: rows are read exactly once.
//
// One warp per target.  Lane l owns float4 chunks l, l+32, ... of the concatenated row and
// time channels l, l+32, ...  Neighbour slots are processed G (= 2) at a time:
//   * the next group's rows are already in flight (register prefetch) while this group is reduced,
//   * the G*H partial dot products are reduced with a transposing butterfly (10 shuffles per
//     4 sums instead of 20) and broadcast,
//   * the running max is only raised by a warp-uniform branch (lazy rescale), so the common case
//     is one ex2 and one FMA per element (flash-style online softmax, exact in the limit).
// Padded slots (neighbour id 0) contribute exp(-1e10 - max) == 0 in the reference, so they are
// skipped; a target with no neighbour at all gets the reference's uniform 1/k over its padded
// rows (models/modules.py:217-224).
#include "attn.cuh"

namespace flid {

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

template <int H, int NV, int TC>
__global__ void __launch_bounds__(128, (H * NV <= 6) ? 3 : 2) attn_kernel(AttnArgs a) {
    constexpr int G = (H * NV <= 6) ? 2 : (H * NV <= 12 ? 2 : 1), V = G * H;
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= a.n) return;
    const int k = a.k, dn = a.dn, de = a.de, T = a.T;
    const int nv4 = dn >> 2, ev4 = de >> 2, tot4 = nv4 + ev4, kd = dn + de + T;

    int nb_l = 0, e_l = 0;
    float dt_l = 0.f;
    if (lane < k) {
        nb_l = __ldg(a.nbr + i * k + lane);
        e_l = __ldg(a.eid + i * k + lane);
        dt_l = __ldg(a.dt + i * k + lane);
    }
    const unsigned valid = __ballot_sync(FULL, lane < k && nb_l != 0);
    const bool all_masked = (valid == 0u);
    unsigned todo = all_masked ? (k >= 32 ? FULL : ((1u << k) - 1u)) : valid;
    // per-slot row numbers (lane j holds slot j)
    int64_t hrow_l = a.hrow_by_id ? (int64_t)nb_l : a.hrow_offset + i * k + lane;
    if (a.hrow_idx != nullptr && lane < k) hrow_l = (int64_t)__ldg(a.hrow_idx + i * k + lane);
    const int hlo = (int)(hrow_l & 0xffffffff), hhi = (int)(hrow_l >> 32);

    const float* u = a.u_base + (a.u_index ? (int64_t)__ldg(a.u_index + i) : i) * (int64_t)(H * kd);
    float4 uh[H][NV];
    float ut[H][TC], tw[TC], tb[TC];
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            uh[h][r] = f < tot4 ? __ldg(reinterpret_cast<const float4*>(u + h * kd) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < TC; ++r) {
            const int c = lane + 32 * r;
            ut[h][r] = c < T ? __ldg(u + h * kd + dn + de + c) : 0.f;
        }
    }
#pragma unroll
    for (int r = 0; r < TC; ++r) {
        const int c = lane + 32 * r;
        tw[r] = c < T ? __ldg(a.time_w + c) : 0.f;
        tb[r] = c < T ? __ldg(a.time_b + c) : 0.f;
    }

    float4 acc[H][NV];
    float acct[H][TC], mx[H], den[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        mx[h] = -INFINITY, den[h] = 0.f;
#pragma unroll
        for (int r = 0; r < NV; ++r) acc[h][r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < TC; ++r) acct[h][r] = 0.f;
    }

    const float4* hbase = reinterpret_cast<const float4*>(a.hrow_base);
    const float4* ebase = reinterpret_cast<const float4*>(a.edge_feat);

    auto next_group = [&](int (&j)[G]) {
#pragma unroll
        for (int s = 0; s < G; ++s) {
            j[s] = -1;
            if (todo) {
                j[s] = __ffs(todo) - 1;
                todo &= todo - 1;
            }
        }
    };
    auto load_group = [&](const int (&j)[G], float4 (&x)[G][NV]) {
#pragma unroll
        for (int s = 0; s < G; ++s) {
            if (j[s] < 0) continue;  // warp-uniform
            const int lo = __shfl_sync(FULL, hlo, j[s]), hi = __shfl_sync(FULL, hhi, j[s]);
            const int e = __shfl_sync(FULL, e_l, j[s]);
            const int64_t hrow = ((int64_t)hi << 32) | (uint32_t)lo;
            const float4* hp = hbase + hrow * nv4;
            const float4* ep = ebase + (int64_t)e * ev4 - nv4;
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int f = lane + 32 * r;
                const float4* p = (f < nv4) ? hp + f : ep + f;  // select, not branch
                x[s][r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (f < tot4) x[s][r] = __ldg(p);
            }
        }
    };
    auto process = [&](const int (&j)[G], const float4 (&x)[G][NV]) {
        float part[V], xt[G][TC];
#pragma unroll
        for (int s = 0; s < G; ++s) {
#pragma unroll
            for (int h = 0; h < H; ++h) part[s * H + h] = 0.f;
#pragma unroll
            for (int r = 0; r < TC; ++r) xt[s][r] = 0.f;
            if (j[s] < 0) continue;  // warp-uniform
            const float d = __shfl_sync(FULL, dt_l, j[s]);
            // channels beyond T have w = b = 0 and u = 0: they evaluate to cos(0) and are never used
#pragma unroll
            for (int r = 0; r < TC; ++r) xt[s][r] = time_channel(d, tw[r], tb[r]);
            if (!all_masked) {
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float p = 0.f;
#pragma unroll
                    for (int r = 0; r < NV; ++r) {
                        p = fmaf(x[s][r].x, uh[h][r].x, p);
                        p = fmaf(x[s][r].y, uh[h][r].y, p);
                        p = fmaf(x[s][r].z, uh[h][r].z, p);
                        p = fmaf(x[s][r].w, uh[h][r].w, p);
                    }
#pragma unroll
                    for (int r = 0; r < TC; ++r) p = fmaf(xt[s][r], ut[h][r], p);
                    part[s * H + h] = p;
                }
            }
        }
        if (!all_masked) reduce_bcast<V>(part, lane);  // all-masked: every score is the same fill value
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float gmax = -INFINITY;
#pragma unroll
            for (int s = 0; s < G; ++s)
                if (j[s] >= 0) gmax = fmaxf(gmax, part[s * H + h]);
            if (gmax > mx[h]) {  // warp-uniform: raise the running max, rescale what was accumulated
                const float corr = exp2f(mx[h] - gmax);
                mx[h] = gmax;
                den[h] *= corr;
#pragma unroll
                for (int r = 0; r < NV; ++r)
                    acc[h][r].x *= corr, acc[h][r].y *= corr, acc[h][r].z *= corr, acc[h][r].w *= corr;
#pragma unroll
                for (int r = 0; r < TC; ++r) acct[h][r] *= corr;
            }
#pragma unroll
            for (int s = 0; s < G; ++s) {
                if (j[s] < 0) continue;
                const float w = exp2f(part[s * H + h] - mx[h]);
                den[h] += w;
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    acc[h][r].x = fmaf(w, x[s][r].x, acc[h][r].x);
                    acc[h][r].y = fmaf(w, x[s][r].y, acc[h][r].y);
                    acc[h][r].z = fmaf(w, x[s][r].z, acc[h][r].z);
                    acc[h][r].w = fmaf(w, x[s][r].w, acc[h][r].w);
                }
#pragma unroll
                for (int r = 0; r < TC; ++r) acct[h][r] = fmaf(w, xt[s][r], acct[h][r]);
            }
        }
    };

    int ja[G], jb[G];
    float4 xa[G][NV], xb[G][NV];
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

    float* z = a.z + i * (int64_t)(H * kd);
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float inv = 1.0f / den[h];
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            if (f < tot4)
                reinterpret_cast<float4*>(z + h * kd)[f] =
                    make_float4(acc[h][r].x * inv, acc[h][r].y * inv, acc[h][r].z * inv, acc[h][r].w * inv);
        }
#pragma unroll
        for (int r = 0; r < TC; ++r) {
            const int c = lane + 32 * r;
            if (c < T) z[h * kd + dn + de + c] = acct[h][r] * inv;
        }
    }
}

template <int H>
static int launch_attn_h(const AttnArgs& a, int nv, int tc, cudaStream_t st) {
    const unsigned blocks = (unsigned)ceil_div(a.n * 32, 128);
#define FLID_ATTN_CASE(NV_, TC_)                              \
    if (nv <= NV_ && tc <= TC_) {                             \
        attn_kernel<H, NV_, TC_><<<blocks, 128, 0, st>>>(a);  \
        FLID_LAUNCH_CHECK();                                  \
        return FLID_OK;                                       \
    }
    FLID_ATTN_CASE(3, 4)
    if constexpr (H <= 2) {
        FLID_ATTN_CASE(6, 4)
    }
#undef FLID_ATTN_CASE
    set_error("attention kernel: unsupported feature widths for %d heads", H);
    return FLID_ERR_INVALID;
}

int launch_attn_scalar(const AttnArgs& a, int H, cudaStream_t st) {
    if (a.n <= 0) return FLID_OK;
    const int nv = (int)ceil_div((a.dn + a.de) / 4, 32), tc = (int)ceil_div(a.T, 32);
    switch (H) {
        case 1: return launch_attn_h<1>(a, nv, tc, st);
        case 2: return launch_attn_h<2>(a, nv, tc, st);
        case 4: return launch_attn_h<4>(a, nv, tc, st);
        default: set_error("attention kernel: num_heads must be 1, 2 or 4 (got %d)", H); return FLID_ERR_INVALID;
    }
}

}  // namespace flid
