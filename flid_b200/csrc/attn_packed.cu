// Attention stream kernel (models/modules.py:183-231 after the re-association of DESIGN.md):
// for every target, gather its <=k neighbour rows [h_nbr | e] (16-byte loads), compute the
// time encoding cos(fma(dt, w, b)) in registers, score each row against the folded query
// u_h (scores already in the log2 domain), masked softmax, and accumulate
// z_h = sum_j a_hj [h_nbr_j | e_j | te_j].  Rows are read exactly once.
//
// One warp per target.  Lane l owns float4 chunks l, l+32, ... of the concatenated row and
// time channels (l + 64 r, l + 64 r + 32) as packed pairs.  All per-element arithmetic is
// done on packed float32 pairs (fma.rn.f32x2 -> SASS FFMA2, sm_100): the dot products, the
// weighted accumulation and the cosine polynomial each take half the issue slots of scalar
// FFMA, which is what bounds this kernel once the rows come out of L2.
// Neighbour slots are processed G (= 2) at a time:
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
namespace {

typedef unsigned long long u64;

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

// cos of two arguments at once, both |x| < COS_FAST_LIMIT: the packed form of cos_fast (common.cuh)
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

// NV: float4 chunks per lane per row; TP: packed time-channel pairs per lane
// MULTI: more than 32 neighbour slots per target (taken in blocks of 32); the common k <= 32 instance keeps every
// per-slot scalar in one register per lane and has no outer loop
template <int H, int NV, int TP, bool MULTI>
__global__ void __launch_bounds__(128, (H * NV <= 6 && NV <= 3) ? 4 : 2) attn_pk_kernel(AttnArgs a) {
    // the query fold u of the warp's target lives in shared memory (16-byte, lane-contiguous reads): 24-32
    // registers less than holding it, which is what lets a fourth block (16 warps) fit on the SM
    extern __shared__ __align__(16) unsigned char u_smem[];
    constexpr int G = 2, V = G * H;
    const int lane = threadIdx.x & 31;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i >= a.n) return;
    const int k = a.k, dn = a.dn, de = a.de, T = a.T;
    const int nv4 = dn >> 2, ev4 = de >> 2, tot4 = nv4 + ev4, kd = dn + de + T;

    // k > 32: the slots are taken in blocks of 32 (lane j owns slot base + j of the current block); the online
    // softmax state carries over.  Whether the target has any valid neighbour at all (the reference's all-masked
    // case: uniform weights over the padded rows) must be known before the first block is processed.
    bool all_masked = false;
    int nb_l = 0, e_l = 0;
    float dt_l = 0.f;
    int64_t hrow_l = 0;
    unsigned todo = 0u;
    u64 haddr = 0ull, eaddr = 0ull;
    if constexpr (MULTI) {
        int any = 0;
        for (int base = 0; base < k; base += 32)
            any |= __any_sync(FULL, base + lane < k && __ldg(a.nbr + i * k + base + lane) != 0);
        all_masked = !any;
    } else {
        if (lane < k) {
            nb_l = __ldg(a.nbr + i * k + lane);
            e_l = __ldg(a.eid + i * k + lane);
            dt_l = __ldg(a.dt + i * k + lane);
            hrow_l = a.hrow_idx ? (int64_t)__ldg(a.hrow_idx + i * k + lane)
                                : (a.hrow_by_id ? (int64_t)nb_l : a.hrow_offset + i * k + lane);
        }
        const unsigned valid = __ballot_sync(FULL, lane < k && nb_l != 0);
        all_masked = (valid == 0u);
        todo = all_masked ? (k >= 32 ? FULL : ((1u << k) - 1u)) : valid;
        // byte addresses of this lane's slot rows (lane j owns slot j); the edge base is shifted so
        // that chunk index f >= nv4 addresses the edge row directly
        haddr = (u64)(reinterpret_cast<const char*>(a.hrow_base) + hrow_l * (int64_t)dn * 4);
        eaddr = (u64)(reinterpret_cast<const char*>(a.edge_feat) + ((int64_t)e_l * ev4 - nv4) * 16);
    }
    const float wmax = __ldg(a.time_bound), bmax = __ldg(a.time_bound + 1);
    const float* u = a.u_base + (a.u_index ? (int64_t)__ldg(a.u_index + i) : i) * (int64_t)(H * kd);
    // this warp's slice of shared memory: [H][NV][32] Row4 (row part of u) then [H][TP][32] packed pairs (time part)
    Row4* uh_s = reinterpret_cast<Row4*>(u_smem) + (threadIdx.x >> 5) * (H * NV * 32) + lane;
    u64* ut_s = reinterpret_cast<u64*>(u_smem + (size_t)4 * H * NV * 32 * sizeof(Row4)) + (threadIdx.x >> 5) * (H * TP * 32) + lane;
    u64 tw[TP], tb[TP];
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            Row4 v{0ull, 0ull};
            if (f < tot4) v = ldg_row4(u + h * kd + 4 * f);
            uh_s[(h * NV + r) * 32] = v;
        }
#pragma unroll
        for (int r = 0; r < TP; ++r) {
            const int c0 = lane + 64 * r, c1 = c0 + 32;
            ut_s[(h * TP + r) * 32] =
                pk(c0 < T ? __ldg(u + h * kd + dn + de + c0) : 0.f, c1 < T ? __ldg(u + h * kd + dn + de + c1) : 0.f);
        }
    }
    // every lane only reads back what it wrote itself: no barrier needed
#pragma unroll
    for (int r = 0; r < TP; ++r) {
        const int c0 = lane + 64 * r, c1 = c0 + 32;
        tw[r] = pk(c0 < T ? __ldg(a.time_w + c0) : 0.f, c1 < T ? __ldg(a.time_w + c1) : 0.f);
        tb[r] = pk(c0 < T ? __ldg(a.time_b + c0) : 0.f, c1 < T ? __ldg(a.time_b + c1) : 0.f);
    }

    Row4 acc[H][NV];
    u64 acct[H][TP];
    float mx[H], den[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        mx[h] = -INFINITY, den[h] = 0.f;
#pragma unroll
        for (int r = 0; r < NV; ++r) acc[h][r] = Row4{0ull, 0ull};
#pragma unroll
        for (int r = 0; r < TP; ++r) acct[h][r] = 0ull;
    }

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
    // Loads are unconditional: a missing second slot re-reads the first one's row (its weight is
    // ex2(-inf) = 0) and chunk indices past the row end are clamped to the last chunk (their u is 0
    // and their accumulators are never stored), so no register needs zeroing and nothing branches.
    auto load_group = [&](const int (&j)[G], Row4 (&x)[G][NV]) {
        if (j[0] < 0) return;  // warp-uniform: nothing left to prefetch
#pragma unroll
        for (int s = 0; s < G; ++s) {
            const int js = j[s] < 0 ? j[0] : j[s];
            const u64 hp = __shfl_sync(FULL, haddr, js), ep = __shfl_sync(FULL, eaddr, js);
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int f = min(lane + 32 * r, tot4 - 1);
                const u64 p = ((f < nv4) ? hp : ep) + (u64)f * 16ull;  // select, not branch
                x[s][r] = ldg_row4(reinterpret_cast<const void*>(p));
            }
        }
    };
    auto process = [&](const int (&j)[G], const Row4 (&x)[G][NV]) {
        const bool two = j[1] >= 0;
        float d[G];
        d[0] = __shfl_sync(FULL, dt_l, j[0]);
        d[1] = __shfl_sync(FULL, dt_l, two ? j[1] : j[0]);
        // time encoding; channels beyond T have w = b = 0 and u = 0 (cos(0), never used).  The
        // cosine path is chosen once per group from |dt| * max|w| + max|b| (warp-uniform).
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
            for (int s = 0; s < G; ++s)
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    u64 p0 = 0ull, p1 = 0ull;  // two chains of packed partial sums
#pragma unroll
                    for (int r = 0; r < NV; ++r) {
                        const Row4 uu = uh_s[(h * NV + r) * 32];
                        p0 = fma2(x[s][r].a, uu.a, p0);
                        p1 = fma2(x[s][r].b, uu.b, p1);
                    }
#pragma unroll
                    for (int r = 0; r < TP; ++r) {
                        const u64 uu = ut_s[(h * TP + r) * 32];
                        if (r & 1)
                            p1 = fma2(xt[s][r], uu, p1);
                        else
                            p0 = fma2(xt[s][r], uu, p0);
                    }
                    part[s * H + h] = hsum(add2(p0, p1));
                }
            reduce_bcast<V>(part, lane);
        } else {
#pragma unroll
            for (int q = 0; q < V; ++q) part[q] = 0.f;  // all scores equal the -1e10 fill: uniform weights
        }
        if (!two) {
#pragma unroll
            for (int h = 0; h < H; ++h) part[H + h] = -INFINITY;  // weight ex2(-inf) = 0 on the zero row
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float gmax = fmaxf(part[h], part[H + h]);
            if (gmax > mx[h]) {  // warp-uniform: raise the running max, rescale what was accumulated
                const float corr = ex2(mx[h] - gmax);
                const u64 c2 = pk1(corr);
                mx[h] = gmax;
                den[h] *= corr;
#pragma unroll
                for (int r = 0; r < NV; ++r) acc[h][r].a = mul2(acc[h][r].a, c2), acc[h][r].b = mul2(acc[h][r].b, c2);
#pragma unroll
                for (int r = 0; r < TP; ++r) acct[h][r] = mul2(acct[h][r], c2);
            }
            const float w0 = ex2(part[h] - mx[h]), w1 = ex2(part[H + h] - mx[h]);
            den[h] += w0 + w1;
            const u64 W0 = pk1(w0), W1 = pk1(w1);
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                acc[h][r].a = fma2(W1, x[1][r].a, fma2(W0, x[0][r].a, acc[h][r].a));
                acc[h][r].b = fma2(W1, x[1][r].b, fma2(W0, x[0][r].b, acc[h][r].b));
            }
#pragma unroll
            for (int r = 0; r < TP; ++r) acct[h][r] = fma2(W1, xt[1][r], fma2(W0, xt[0][r], acct[h][r]));
        }
    };

    for (int base = 0; base < (MULTI ? k : 1); base += 32) {
      if constexpr (MULTI) {
        const int kb = min(32, k - base);
        if (lane < kb) {
            nb_l = __ldg(a.nbr + i * k + base + lane);
            e_l = __ldg(a.eid + i * k + base + lane);
            dt_l = __ldg(a.dt + i * k + base + lane);
            hrow_l = a.hrow_idx ? (int64_t)__ldg(a.hrow_idx + i * k + base + lane)
                                : (a.hrow_by_id ? (int64_t)nb_l : a.hrow_offset + i * k + base + lane);
        }
        const unsigned valid = __ballot_sync(FULL, lane < kb && nb_l != 0);
        todo = all_masked ? (kb >= 32 ? FULL : ((1u << kb) - 1u)) : valid;
        if (todo == 0u) continue;  // warp-uniform
        // byte addresses of this lane's slot rows; the edge base is shifted so that chunk index f >= nv4
        // addresses the edge row directly
        haddr = (u64)(reinterpret_cast<const char*>(a.hrow_base) + hrow_l * (int64_t)dn * 4);
        eaddr = (u64)(reinterpret_cast<const char*>(a.edge_feat) + ((int64_t)e_l * ev4 - nv4) * 16);
      }
        int ja[G], jb[G];
        Row4 xa[G][NV], xb[G][NV];
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

    float* z = a.z + i * (int64_t)(H * kd);
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const u64 inv = pk1(1.0f / den[h]);
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            const int f = lane + 32 * r;
            if (f < tot4)
                reinterpret_cast<ulonglong2*>(z + h * kd)[f] =
                    make_ulonglong2(mul2(acc[h][r].a, inv), mul2(acc[h][r].b, inv));
        }
#pragma unroll
        for (int r = 0; r < TP; ++r) {
            const int c0 = lane + 64 * r, c1 = c0 + 32;
            float lo, hi;
            upk(mul2(acct[h][r], inv), lo, hi);
            if (c0 < T) z[h * kd + dn + de + c0] = lo;
            if (c1 < T) z[h * kd + dn + de + c1] = hi;
        }
    }
}

template <int H>
int launch_h(const AttnArgs& a, int nv, int tp, cudaStream_t st) {
    const unsigned blocks = (unsigned)ceil_div(a.n * 32, 128);
#define FLID_ATTN_CASE(NV_, TP_)                                 \
    if (nv <= NV_ && tp <= TP_) {                                \
        if (a.k > 32)                                                                                  \
            attn_pk_kernel<H, NV_, TP_, true><<<blocks, 128, 4 * H * 32 * (NV_ * 16 + TP_ * 8), st>>>(a);  \
        else                                                                                           \
            attn_pk_kernel<H, NV_, TP_, false><<<blocks, 128, 4 * H * 32 * (NV_ * 16 + TP_ * 8), st>>>(a); \
        FLID_LAUNCH_CHECK();                                     \
        return FLID_OK;                                          \
    }
    FLID_ATTN_CASE(3, 2)
    if constexpr (H <= 2) {
        FLID_ATTN_CASE(6, 2)
    }
#undef FLID_ATTN_CASE
    set_error("attention kernel: unsupported feature widths for %d heads", H);
    return FLID_ERR_INVALID;
}

}  // namespace

int launch_attn(const AttnArgs& a, int H, cudaStream_t st) {
    if (a.n <= 0) return FLID_OK;
    const int nv = (int)ceil_div((a.dn + a.de) / 4, 32), tp = (int)ceil_div(a.T, 64);
    switch (H) {
        case 1: return launch_h<1>(a, nv, tp, st);
        case 2: return launch_h<2>(a, nv, tp, st);
        case 4: return launch_h<4>(a, nv, tp, st);
        default: set_error("attention kernel: num_heads must be 1, 2 or 4 (got %d)", H); return FLID_ERR_INVALID;
    }
}

}  // namespace flid
