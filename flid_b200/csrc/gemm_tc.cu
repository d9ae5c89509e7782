// tcgen05 / TMEM GEMM (see gemm_tc.cuh): persistent, warp-specialised, L2-traffic aware.
//
//   warps 0-7   producers : two groups of four warps that take alternate pipeline stages:
//                           coalesced fp32 loads of the A rows (row gather fused), hi/lo tf32
//                           split, st.shared into the UMMA canonical K-major layout.  The
//                           per-stage tail (drain stores, fence.proxy.async, mbarrier arrive)
//                           of one group overlaps the loads and stores of the other.
//   warps 8-11  epilogue  : tcgen05.ld their 32 TMEM lanes, bias / ReLU, store C
//   warp  12    issuer    : one thread waits "stage full", issues the tcgen05.mma.kind::tf32
//                           triples {lo.hi, hi.lo, hi.hi} and tcgen05.commit's the stage's "empty"
//                           mbarrier; after the last K chunk it commits "accumulators full"
//   warp  13    loader    : one thread bulk-copies (cp.async.bulk, SASS UBLKCP) the pre-tiled
//                           weight chunk of every stage as soon as the stage is free
//
// What paces this kernel (measurements on the B200; DESIGN.md section 4 has the full list):
// the tensor pipe is only 40-55 % busy.  Per-stage clock64 stamps of one CTA (FLID_GEMM_TRACE)
// showed the producers finishing a stage thousands of cycles before the issuer sees it full -- the
// wait is on the weight copy -- and switching off loads, smem stores, MMAs and C stores one at a
// time left a 107 us barrier-and-copy skeleton of a 250 us launch.  With the 3xTF32 split every
// 8-float K step issues three MMAs that re-read A (4 KB) and B (n_tile x 32 B) from shared
// memory, ~120 B/cycle at n_tile = 144: the operand fetch alone takes the whole shared-memory
// bandwidth, and the producers' stores and the weight copies land in the same memory.
// Consequences kept in this file:
//   * K chunks of 16 floats and as many ring stages as fit; a dedicated loader thread, so a copy
//     is issued the moment its stage is free; two producer groups, so the per-stage
//     drain/fence/arrive tail is off the critical path; no 64-bit divisions per chunk;
//   * 256 < N <= 512 as one pass over A (two MMA column groups, single-buffered accumulator);
//   * C staged through shared memory in the epilogue (row-coalesced stores) where it pays off;
//   * optionally MS sub-tiles per work item sharing every weight stage -- measured slower, off;
//   * gemm_tc_ts.cu moves the A operand into tensor memory and is what single-n-block shapes use.
// Tried without effect: weight-image replicas against L2 slice hot-spotting, a 3-deep register
// prefetch, a CTA pair with tcgen05.mma.cta_group::2 (half the weight bytes per SM; correct, not faster --
// in the history, not in the tree).
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace flid {

// ---------------------------------------------------------------- weight tiling
// image layout: [n_block][k_chunk][half][c4][n_tile] float4
// element (n, k) of the logical weight is W[n * sn + k * sk]: (ldw, 1) for W[N, K], (1, ldw) for a stored W^T[K, N]
// Both images of a weight (the wide tiles and the 32-column tiles of small launches) are written by one launch:
// elements [0, total0) belong to image 0, the rest to image 1.
struct TcPrepImage {
    float4* out;
    int n_tile, n_blocks;
    int64_t total;
};
__global__ void tc_prep_kernel(const float* __restrict__ W, int64_t sn, int64_t sk, int N, int K, int k_chunks, int single,
                               TcPrepImage im0, TcPrepImage im1) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= im0.total + im1.total) return;
    const bool second = idx >= im0.total;
    if (second) idx -= im0.total;
    const int n_tile = second ? im1.n_tile : im0.n_tile;
    float4* out = second ? im1.out : im0.out;
    int64_t r = idx;
    const int nl = (int)(r % n_tile);
    r /= n_tile;
    const int c = (int)(r % C4);
    r /= C4;
    const int half = (int)(r % 2);
    r /= 2;
    const int kc = (int)(r % k_chunks);
    const int nb = (int)(r / k_chunks);
    const int n = nb * n_tile + nl, k = kc * KC + c * 4;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float x = (n < N && k + e < K) ? W[(int64_t)n * sn + (int64_t)(k + e) * sk] : 0.f;
        const float hi = single ? bf16_round(x) : tf32_hi(x);
        v[e] = half ? (single ? 0.f : x - hi) : hi;
    }
    out[idx] = make_float4(v[0], v[1], v[2], v[3]);
}

// ---------------------------------------------------------------- the GEMM
struct TcShape {
    int N, n_tile, n_blocks, k_chunks, stages;
    int acc_bufs;          // 2 when two accumulator sets fit in TMEM (epilogue overlaps the next group)
    uint32_t acc_stride;   // TMEM columns per accumulator set (MS * n_tile)
    int64_t m_groups;      // groups of MS * 128 rows
    int single;            // TcWeight::single: one MMA per product on bf16-rounded operands
    int staged_epilogue;   // stage C through shared memory (pays off for long K loops and scattered rows)
    long long* trace;      // FLID_GEMM_TRACE (development): per-stage clock64 stamps of CTA 0, [6][TRACE_Q]
};
constexpr int TRACE_Q = 512;
#define TRACE(role, q) do { if (sh.trace && blockIdx.x == 0 && (q) < TRACE_Q) sh.trace[(role) * TRACE_Q + (q)] = clock64(); } while (0)

template <int MS, bool SINGLE, bool LATE>   // LATE: residual add / GELU in the store phase of the staged epilogue (dense.cu)
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(TcGemmArgs g, const float4* __restrict__ wbuf, TcShape sh) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b_half = (uint32_t)C4 * sh.n_tile * 16;
    const uint32_t stage_bytes = MS * A_SUB + 2 * b_half;
    const uint32_t nblk = (uint32_t)sh.n_blocks;
    const uint32_t work = (uint32_t)(sh.m_groups * sh.n_blocks);  // host guarantees < 2^31

    if (tid == 0) {
        for (int s = 0; s < sh.stages; ++s) mbar_init(&bar_full[s], NPROD + 1), mbar_init(&bar_empty[s], 1);
        for (int a = 0; a < 2; ++a) mbar_init(&bar_acc_full[a], 1), mbar_init(&bar_acc_empty[a], NEPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int ktot = g.w0 + g.w1;

    if (warp < 8) {
        // ===================================================== producers (two alternating groups)
        // The (row group, K chunk) sequence of this CTA is flattened; producer group `pg` takes
        // chunks pg, pg + 2, ...  The global loads of a group's next chunk are in flight
        // (registers) while its current one is split and stored.  Everything that depends on the
        // work item only (row pointers, validity, n-block) is recomputed when the item changes
        // (once per k_chunks chunks); the per-chunk path is an add and a compare per load.
        constexpr int NL = MS * 4;                  // 16-byte loads per thread per chunk
        const int pg = warp >> 2, pw = warp & 3;
        const int rsub = lane >> 2, c = lane & 3;   // 8 rows x 4 float4 (64 B of K) per warp instruction
        const uint32_t dq = gridDim.x / nblk, dr = gridDim.x % nblk;  // work item stride as (row group, n block)
        struct Cursor {
            uint32_t t, mg, nb;  // work item (>= work when exhausted), its row group and n block
            int kc;
        };
        auto next_item = [&](Cursor& cu) {
            cu.t += gridDim.x, cu.mg += dq, cu.nb += dr;
            if (cu.nb >= nblk) cu.nb -= nblk, cu.mg += 1;
        };
        // ---- load cursor
        Cursor lc{blockIdx.x, blockIdx.x / nblk, blockIdx.x % nblk, pg};
        const float* p0[NL];
        const float* p1[NL];
        uint32_t okmask = 0;
        auto bind_rows = [&]() {  // row pointers of the load cursor's work item
            okmask = 0;
            const int64_t m0 = (int64_t)lc.mg * (MS * 128);
#pragma unroll
            for (int i = 0; i < NL; ++i) {
                const int64_t row = m0 + i * 32 + pw * 8 + rsub;
                p0[i] = g.A0, p1[i] = g.A1;
                if (lc.t < work && row < g.M) {
                    okmask |= 1u << i;
                    p0[i] = g.A0 + (g.idx0 ? (int64_t)__ldg(g.idx0 + row) : row) * g.lda0;
                    if (g.w1 > 0) p1[i] = g.A1 + (g.idx1 ? (int64_t)__ldg(g.idx1 + row) : row) * g.lda1;
                }
            }
        };
        if (lc.kc >= sh.k_chunks) lc.kc -= sh.k_chunks, next_item(lc);  // k_chunks == 1: group 1 starts on the next item
        bind_rows();
        auto load_next = [&](float4 (&v)[NL]) {
            const int k = lc.kc * KC + c * 4;
            const bool seg0 = k < g.w0;
            const int koff = seg0 ? k : k - g.w0;
            const bool kin = k < ktot;
#pragma unroll
            for (int i = 0; i < NL; ++i) {
                const float* src = (seg0 ? p0[i] : p1[i]) + koff;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kin && ((okmask >> i) & 1u)) v[i] = __ldg(reinterpret_cast<const float4*>(src));
            }
            lc.kc += 2;
            if (lc.kc >= sh.k_chunks) {  // next work item (rare path)
                lc.kc -= sh.k_chunks;
                next_item(lc);
                if (lc.kc >= sh.k_chunks) lc.kc -= sh.k_chunks, next_item(lc);  // k_chunks == 1
                bind_rows();
            }
        };
        // ---- store cursor: only the stage ring and the remaining chunk count matter
        const uint32_t my_items = work > blockIdx.x ? (work - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const uint32_t total_q = my_items * (uint32_t)sh.k_chunks;
        uint32_t sq = (uint32_t)pg;  // chunk index of this group within the CTA's sequence
        uint32_t stage = (uint32_t)pg % (uint32_t)sh.stages, phase = 0;
        auto store_next = [&](const float4 (&v)[NL]) {
            if (pw == 0 && lane == 0) TRACE(0, sq);
            mbar_wait(&bar_empty[stage], phase ^ 1);
            if (pw == 0 && lane == 0) TRACE(1, sq);
            uint8_t* st = smem + (size_t)stage * stage_bytes + c * A_CSTRIDE + (pw * 8 + rsub) * 16;
#pragma unroll
            for (int i = 0; i < NL; ++i) {
                const float4 x = v[i];
                uint8_t* dst = st + (i >> 2) * A_SUB + (i & 3) * (32 * 16);  // sub-tile i / 4, rows (i % 4) * 32 + ...
                if (SINGLE) {
                    *reinterpret_cast<float4*>(dst) = make_float4(bf16_round(x.x), bf16_round(x.y), bf16_round(x.z), bf16_round(x.w));
                    continue;
                }
                const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                *reinterpret_cast<float4*>(dst) = hi;
                *reinterpret_cast<float4*>(dst + A_HALF) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
            }
            fence_async_smem();
            mbar_arrive(&bar_full[stage]);
            if (pw == 0 && lane == 0) TRACE(2, sq);
            sq += 2;
            stage += 2;
            if (stage >= (uint32_t)sh.stages) stage -= (uint32_t)sh.stages, phase ^= 1;
        };
        float4 ra[NL], rb[NL];
        load_next(ra);
        while (sq < total_q) {
            load_next(rb);
            store_next(ra);
            if (sq >= total_q) break;
            load_next(ra);
            store_next(rb);
        }
    } else if (warp < 12) {
        // ===================================================== epilogue
        // A TMEM lane is a row, so after tcgen05.ld every thread holds 16 columns of its own row and
        // a direct store would touch 32 rows (32 half-used sectors) per instruction.  Each warp stages
        // its 32 x 16 block through shared memory and stores it 8 rows x 64 B per instruction instead.
        const int ew = warp - 8;  // TMEM lane quarter == warp id % 4
        // staging area: the last EPI_BYTES of the dynamic allocation (only reserved by staged launches)
        float* stg = reinterpret_cast<float*>(smem + (size_t)sh.stages * stage_bytes) + ew * (32 * EPI_LD);
        const bool vec_ok = (g.ldc & 3) == 0 && (sh.N & 3) == 0;
        const bool staged = vec_ok && sh.staged_epilogue;
        const int sr = lane >> 2, sc = (lane & 3) * 4;  // this lane's (row within 8, column) in the store phase
        uint32_t it = 0;
        for (uint32_t t = blockIdx.x; t < work; t += gridDim.x, ++it) {
            const uint32_t mg = t / nblk, nb = t - mg * nblk;
            const uint32_t acc = sh.acc_bufs == 2 ? (it & 1) : 0;
            const uint32_t par = sh.acc_bufs == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(&bar_acc_full[acc], par);
            tc_fence_after();
#pragma unroll
            for (int ms = 0; ms < MS; ++ms) {
                const int64_t row0 = (int64_t)mg * (MS * 128) + ms * 128 + ew * 32;
                const uint32_t taddr = tmem + acc * sh.acc_stride + ms * sh.n_tile + ((uint32_t)(ew * 32) << 16);
                float* crow4[4];  // destination rows of the store phase
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t r = row0 + j * 8 + sr;
                    crow4[j] = (r < g.M) ? g.C + (g.cidx ? (int64_t)__ldg(g.cidx + r) : r) * g.ldc : nullptr;
                }
                const int64_t row = row0 + lane;
                float* crow = (row < g.M) ? g.C + (g.cidx ? (int64_t)__ldg(g.cidx + row) : row) * g.ldc : nullptr;
                for (int c0 = 0; c0 < sh.n_tile; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
                    const int n0 = nb * sh.n_tile + c0;
                    if (n0 >= sh.N) continue;  // warp-uniform: padding columns of the last n block
                    if (vec_ok && !staged) {
                        if (crow != nullptr) {
                            if (n0 + 16 <= sh.N) {
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                                    if (g.bias) {
                                        const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + i));
                                        o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
                                    }
                                    if (g.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
                                    *reinterpret_cast<float4*>(crow + n0 + i) = o;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const int n = n0 + i;
                                    if (n < sh.N) {
                                        float x = v[i];
                                        if (g.bias) x += __ldg(g.bias + n);
                                        if (g.relu) x = fmaxf(x, 0.f);
                                        crow[n] = x;
                                    }
                                }
                            }
                        }
                    } else if (staged) {
                        constexpr bool late = LATE;
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                            if (g.bias && n0 + i < sh.N) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + i));
                                o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
                            }
                            if (g.relu && !late) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
                            *reinterpret_cast<float4*>(stg + lane * EPI_LD + i) = o;
                        }
                        __syncwarp();
                        if (n0 + sc < sh.N) {
                            if (!late) {   // the hot path of the TGAT chain: four independent row stores
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float4 o = *reinterpret_cast<const float4*>(stg + (j * 8 + sr) * EPI_LD + sc);
                                    if (crow4[j] != nullptr) *reinterpret_cast<float4*>(crow4[j] + n0 + sc) = o;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float4 o = *reinterpret_cast<const float4*>(stg + (j * 8 + sr) * EPI_LD + sc);
                                    if (crow4[j] == nullptr) continue;
                                    if (g.resid) {
                                        const float4 r = __ldg(reinterpret_cast<const float4*>(g.resid + (row0 + j * 8 + sr) * g.ldr + n0 + sc));
                                        o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
                                    }
                                    if (g.gelu) {
                                        o.x = 0.5f * o.x * (1.0f + erff(o.x * 0.70710678118654752f));
                                        o.y = 0.5f * o.y * (1.0f + erff(o.y * 0.70710678118654752f));
                                        o.z = 0.5f * o.z * (1.0f + erff(o.z * 0.70710678118654752f));
                                        o.w = 0.5f * o.w * (1.0f + erff(o.w * 0.70710678118654752f));
                                    } else if (g.relu) {
                                        o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
                                    }
                                    *reinterpret_cast<float4*>(crow4[j] + n0 + sc) = o;
                                }
                            }
                        }
                        __syncwarp();
                    } else if (crow != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int n = n0 + i;
                            if (n < sh.N) {
                                float x = v[i];
                                if (g.bias) x += __ldg(g.bias + n);
                                if (g.relu) x = fmaxf(x, 0.f);
                                crow[n] = x;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bar_acc_empty[acc]);
        }
    } else if (warp == 12 && lane == 0) {
        // ===================================================== MMA issuer (one thread)
        // instruction descriptor: D=f32, A=B=tf32, K-major both, N = n_tile, M = 128
        // An MMA covers at most 256 columns: a wider tile (one pass over A for 256 < N <= 512) is issued as
        // two column groups that read different rows of the same weight stage.
        const uint32_t n_a = sh.n_tile > 256 ? (uint32_t)((sh.n_tile / 2 + 15) / 16 * 16) : (uint32_t)sh.n_tile;
        const uint32_t n_b = (uint32_t)sh.n_tile - n_a;
        auto make_idesc = [](uint32_t n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | (8u << 24); };
        const uint32_t idesc = make_idesc(n_a), idesc_b = make_idesc(n_b);
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t b_lbo = (uint32_t)sh.n_tile * 16;
        uint32_t it = 0, s = 0, ph = 0, tq = 0;
        for (uint32_t t = blockIdx.x; t < work; t += gridDim.x, ++it) {
            const uint32_t acc = sh.acc_bufs == 2 ? (it & 1) : 0;
            const uint32_t par = sh.acc_bufs == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(&bar_acc_empty[acc], par ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem + acc * sh.acc_stride;
            for (int kc = 0; kc < sh.k_chunks; ++kc) {
                mbar_wait(&bar_full[s], ph);
                TRACE(3, tq);
                tc_fence_after();
                const uint32_t sa = smem_base + s * stage_bytes;
                const uint32_t sb = sa + MS * A_SUB;
#pragma unroll
                for (int j = 0; j < KC / 8; ++j) {
                    const uint64_t d_bhi = umma_desc(sb + (2 * j) * b_lbo, b_lbo, 128);
                    const uint64_t d_blo = umma_desc(sb + b_half + (2 * j) * b_lbo, b_lbo, 128);
#pragma unroll
                    for (int ms = 0; ms < MS; ++ms) {
                        const uint32_t a0 = sa + ms * A_SUB + (2 * j) * A_CSTRIDE;
                        const uint64_t d_ahi = umma_desc(a0, A_CSTRIDE, 128);
                        const uint64_t d_alo = umma_desc(a0 + A_HALF, A_CSTRIDE, 128);
                        const uint32_t d = d0 + ms * sh.n_tile;
                        const uint64_t row_off = (uint64_t)((n_a * 16u) >> 4);  // start-address field is in 16 B units
                        if (SINGLE) {
                            umma_tf32(d, d_ahi, d_bhi, idesc, (kc | j) ? 1u : 0u);
                            if (n_b) umma_tf32(d + n_a, d_ahi, d_bhi + row_off, idesc_b, (kc | j) ? 1u : 0u);
                            continue;
                        }
                        umma_tf32(d, d_alo, d_bhi, idesc, (kc | j) ? 1u : 0u);  // small terms first
                        umma_tf32(d, d_ahi, d_blo, idesc, 1u);
                        umma_tf32(d, d_ahi, d_bhi, idesc, 1u);
                        if (n_b) {  // second column group: weight rows n_a.. of the same chunk (16 B per row)
                            umma_tf32(d + n_a, d_alo, d_bhi + row_off, idesc_b, (kc | j) ? 1u : 0u);
                            umma_tf32(d + n_a, d_ahi, d_blo + row_off, idesc_b, 1u);
                            umma_tf32(d + n_a, d_ahi, d_bhi + row_off, idesc_b, 1u);
                        }
                    }
                }
                tc_commit(&bar_empty[s]);  // frees the smem stage when these MMAs have read it
                TRACE(4, tq);
                ++tq;
                if (++s == (uint32_t)sh.stages) s = 0, ph ^= 1;
            }
            tc_commit(&bar_acc_full[acc]);  // accumulators complete -> epilogue
        }
    } else if (warp == 13 && lane == 0) {
        // ===================================================== weight loader (one thread)
        uint32_t s = 0, ph = 0, tq = 0;
        const int64_t chunk4 = (int64_t)2 * C4 * sh.n_tile;  // float4 per (n block, K chunk)
        for (uint32_t t = blockIdx.x; t < work; t += gridDim.x) {
            const uint32_t nb = t % nblk;
            const float4* wsrc = wbuf + (int64_t)nb * sh.k_chunks * chunk4;
            for (int kc = 0; kc < sh.k_chunks; ++kc) {
                mbar_wait(&bar_empty[s], ph ^ 1);
                TRACE(5, tq);
                ++tq;
                uint8_t* st = smem + (size_t)s * stage_bytes + MS * A_SUB;
                const uint32_t wbytes = SINGLE ? b_half : 2 * b_half;  // the lo half is not used by a single-MMA product
                mbar_arrive_expect_tx(&bar_full[s], wbytes);
                bulk_g2s(st, wsrc + kc * chunk4, wbytes, &bar_full[s]);
                if (++s == (uint32_t)sh.stages) s = 0, ph ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// ---------------------------------------------------------------- host side
static int pick_n_tile(int N) {
    // 256 < N <= 512: one tile of the full (padded) width -- A is then read once instead of once per
    // n block; the accumulator (<= 512 TMEM columns) is single-buffered and fed by two MMA column groups
    // (out-projection, N = 272: 221 -> 198 us at M = 65 536; FLID_GEMM_WIDE=0 restores the two-block form)
    const char* wide = getenv("FLID_GEMM_WIDE");
    if (N > 256 && (N + 15) / 16 * 16 <= 512 && !(wide && wide[0] == '0')) return (N + 15) / 16 * 16;
    const int blocks = (N + 255) / 256;
    int t = (N + blocks - 1) / blocks;
    t = (t + 15) / 16 * 16;
    return t < 16 ? 16 : t;
}

static int prepare_weight(const float* W, int64_t sn, int64_t sk, int N, int K, TcWeight* w, cudaStream_t st,
                          int single = 0) {
    FLID_REQUIRE(W && w && N > 0 && K > 0, "tc_prepare_weight: bad argument");
    const int n_tile = pick_n_tile(N);
    const int n_blocks = (N + n_tile - 1) / n_tile;
    const int k_chunks = (K + KC - 1) / KC;
    if (w->buf && (w->N != N || w->K != K)) {
        FLID_CUDA(cudaDeviceSynchronize());
        FLID_CUDA(cudaFree(w->buf));
        w->buf = nullptr;
    }
    w->N = N, w->K = K, w->n_tile = n_tile, w->n_blocks = n_blocks, w->k_chunks = k_chunks, w->single = single;
    if (!w->buf) FLID_CUDA(cudaMalloc((void**)&w->buf, w->bytes()));
    TcPrepImage im0{reinterpret_cast<float4*>(w->buf), n_tile, n_blocks, (int64_t)n_blocks * k_chunks * 2 * C4 * n_tile};
    TcPrepImage im1{nullptr, 1, 0, 0};
    // companion image for small launches
    const int st_tile = 32, st_blocks = (N + st_tile - 1) / st_tile;
    if (n_tile > st_tile) {
        if (w->small_buf && (w->small_tile != st_tile || w->small_blocks != st_blocks)) {
            FLID_CUDA(cudaDeviceSynchronize());
            FLID_CUDA(cudaFree(w->small_buf));
            w->small_buf = nullptr;
        }
        const int64_t stotal = (int64_t)st_blocks * k_chunks * 2 * C4 * st_tile;
        if (!w->small_buf) FLID_CUDA(cudaMalloc((void**)&w->small_buf, (size_t)stotal * 16));
        w->small_tile = st_tile, w->small_blocks = st_blocks;
        im1 = TcPrepImage{reinterpret_cast<float4*>(w->small_buf), st_tile, st_blocks, stotal};
    } else {
        w->small_tile = 0, w->small_blocks = 0;
    }
    tc_prep_kernel<<<(unsigned)ceil_div(im0.total + im1.total, 256), 256, 0, st>>>(W, sn, sk, N, K, k_chunks, single, im0, im1);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

int tc_prepare_weight(const float* W, int64_t ldw, int N, int K, TcWeight* w, cudaStream_t st, int single) {
    return prepare_weight(W, ldw, 1, N, K, w, st, single);
}

int tc_prepare_weight_t(const float* Wt, int64_t ldw, int N, int K, TcWeight* w, cudaStream_t st) {
    return prepare_weight(Wt, 1, ldw, N, K, w, st);
}

void tc_free_weight(TcWeight* w) {
    if (w && w->buf) cudaFree(w->buf);
    if (w && w->small_buf) cudaFree(w->small_buf);
    if (w) w->buf = nullptr, w->small_buf = nullptr;
}

template <int MS>
static int launch_ms(const TcGemmArgs& g, const TcWeight& w, TcShape sh, int sm_count, int smem_max, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        FLID_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - STATIC_SMEM));
        FLID_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - STATIC_SMEM));
        FLID_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - STATIC_SMEM));
        attr_set = true;
    }
    const size_t stage = (size_t)MS * A_SUB + 2 * (size_t)C4 * w.n_tile * 16;
    // measured (tools/gemm_probe.py): staging C through shared memory is +3 % on the K = 888 out-projection, a
    // clear win for scattered rows and for the wide query-fold output; the narrow short-K shapes store directly
    sh.staged_epilogue = (g.cidx != nullptr || w.k_chunks >= 48 || w.N >= 512) ? 1 : 0;
    if (const char* e = getenv("FLID_GEMM_EPI")) sh.staged_epilogue = e[0] == '1';  // development knob
    if (g.resid != nullptr || g.gelu) sh.staged_epilogue = 1;   // applied in the staged store phase only
    const size_t ring_bytes = (size_t)(smem_max - STATIC_SMEM) - (sh.staged_epilogue ? EPI_BYTES : 0);
    int stages = (int)(ring_bytes / stage);
    sh.stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    sh.m_groups = ceil_div(g.M, MS * 128);
    sh.acc_stride = (uint32_t)(MS * w.n_tile);
    sh.acc_bufs = 2 * sh.acc_stride <= 512 ? 2 : 1;
    const int64_t work = sh.m_groups * sh.n_blocks;
    FLID_REQUIRE(work < (1LL << 31) - 65536, "tc_gemm: too many tiles for one launch (M = %lld)", (long long)g.M);
    const unsigned grid = (unsigned)(work < sm_count ? work : sm_count);
    const bool late = g.resid != nullptr || g.gelu;
    FLID_REQUIRE(!late || !w.single, "tc_gemm: the residual / GELU epilogue is built for the fp32-grade mode only");
    if (late)
        gemm_tc_kernel<MS, false, true><<<grid, NTHREADS, sh.stages * stage + (sh.staged_epilogue ? EPI_BYTES : 0), st>>>(
            g, reinterpret_cast<const float4*>(w.buf), sh);
    else if (w.single)
        gemm_tc_kernel<MS, true, false><<<grid, NTHREADS, sh.stages * stage + (sh.staged_epilogue ? EPI_BYTES : 0), st>>>(
            g, reinterpret_cast<const float4*>(w.buf), sh);
    else
        gemm_tc_kernel<MS, false, false><<<grid, NTHREADS, sh.stages * stage + (sh.staged_epilogue ? EPI_BYTES : 0), st>>>(
            g, reinterpret_cast<const float4*>(w.buf), sh);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

int tc_gemm(const TcGemmArgs& g, const TcWeight& w_in, cudaStream_t st) {
    if (g.M <= 0) return FLID_OK;
    FLID_REQUIRE(w_in.buf != nullptr, "tc_gemm: weight not prepared");
    // few rows: the 32-column image, so that the 128-row tiles spread over many CTAs (same per-element arithmetic:
    // the K order of the accumulation does not depend on the tile width)
    TcWeight w = w_in;
    static const bool small_ok = [] {
        const char* e = getenv("FLID_GEMM_SMALL");
        return !(e && e[0] == '0');
    }();
    const bool lnf = g.ln_c1 != nullptr;   // LayerNorm fold: A-from-TMEM kernel only (one path whatever the row count)
    if (small_ok && g.M <= TC_SMALL_M && w_in.small_buf != nullptr && !lnf)
        w.buf = w_in.small_buf, w.n_tile = w_in.small_tile, w.n_blocks = w_in.small_blocks;
    FLID_REQUIRE(!lnf || (w.n_blocks == 1 && (512 - w.n_tile) / 32 >= 4), "tc_gemm: the LayerNorm fold needs a single n block");
    FLID_REQUIRE(g.w0 > 0 && g.w0 + g.w1 == w.K, "tc_gemm: A width %d+%d != weight K %d", g.w0, g.w1, w.K);
    FLID_REQUIRE((g.w0 % 4) == 0 && (g.w1 % 4) == 0 && (g.lda0 % 4) == 0 && (g.lda1 % 4) == 0,
                 "tc_gemm: segment widths / row strides must be multiples of 4 floats");
    FLID_REQUIRE((g.resid == nullptr && !g.gelu) || ((g.ldc & 3) == 0 && (w_in.N & 3) == 0 && (g.ldr & 3) == 0),
                 "tc_gemm: residual / GELU epilogue needs 16-byte aligned rows");
    static int sm_count = 0, smem_max = 0, ms_cap = 1;
    if (sm_count == 0) {
        int dev = 0;
        FLID_CUDA(cudaGetDevice(&dev));
        FLID_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        FLID_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        const char* e = getenv("FLID_GEMM_MS");  // development knob: cap the sub-tiles per work item
        if (e && e[0] >= '1' && e[0] <= '2') ms_cap = e[0] - '0';
    }
    {
        // A-from-TMEM kernel (gemm_tc_ts.cu) for single-n-block shapes: measured 5-12 % faster there; the wide
        // query-fold output (4 n blocks, epilogue-heavy) keeps the kernel below, whose accumulators double-buffer
        static int ts = -1;
        if (ts < 0) {
            const char* e = getenv("FLID_GEMM_TS");
            ts = (e && e[0] == '0') ? 0 : 1;
        }
        if ((ts || lnf) && w.n_blocks == 1 && (512 - w.n_tile) / 32 >= 4)  // >= 4 stages of A columns next to the accumulator
            return tc_gemm_ts(g, w, sm_count, smem_max, st);
    }
    TcShape sh;
    sh.trace = nullptr;
    static long long* d_trace = nullptr;
    if (getenv("FLID_GEMM_TRACE")) {
        if (!d_trace) cudaMalloc((void**)&d_trace, sizeof(long long) * 6 * TRACE_Q);
        cudaMemsetAsync(d_trace, 0, sizeof(long long) * 6 * TRACE_Q, st);
        sh.trace = d_trace;
    }
    sh.N = w.N, sh.n_tile = w.n_tile, sh.n_blocks = w.n_blocks, sh.k_chunks = w.k_chunks, sh.single = w.single;
    // Sub-tiles per work item.  Measured on the B200 (tools/gemm_probe.py, M = 65 536): sharing a
    // weight stage between two sub-tiles halves the weight traffic but was 10-25 % slower on every
    // shape (accumulators can no longer be double-buffered, so the epilogue is exposed, and the
    // ring holds fewer stages than the ~4500-cycle L2 latency under load needs).  One sub-tile is
    // the default; FLID_GEMM_MS=2 keeps the other variant reachable for experiments.
    int ms = 1;
    for (int cand = ms_cap; cand > 1; --cand) {
        const size_t stage = (size_t)cand * A_SUB + 2 * (size_t)C4 * w.n_tile * 16;
        if (cand * w.n_tile <= 512 && (size_t)(smem_max - STATIC_SMEM) / stage >= 3 &&
            ceil_div(g.M, cand * 128) * w.n_blocks >= (int64_t)sm_count) {
            ms = cand;
            break;
        }
    }
    int rc;
    switch (ms) {
        case 2: rc = launch_ms<2>(g, w, sh, sm_count, smem_max, st); break;
        default: rc = launch_ms<1>(g, w, sh, sm_count, smem_max, st); break;
    }
    if (sh.trace && rc == FLID_OK) {  // development: dump the stamps of the launch as text
        static long long h[6 * TRACE_Q];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
        FILE* f = fopen(getenv("FLID_GEMM_TRACE"), "w");
        if (f) {
            fprintf(f, "# ms=%d n_tile=%d k_chunks=%d stages? q p_wait0 p_wait1 p_arrive i_full i_commit l_empty\n", ms, w.n_tile, w.k_chunks);
            for (int q = 0; q < TRACE_Q; ++q)
                fprintf(f, "%d %lld %lld %lld %lld %lld %lld\n", q, h[q], h[TRACE_Q + q], h[2 * TRACE_Q + q], h[3 * TRACE_Q + q], h[4 * TRACE_Q + q], h[5 * TRACE_Q + q]);
            fclose(f);
        }
    }
    return rc;
}

}  // namespace flid
