// CTA-pair variant of the tcgen05 3xTF32 GEMM (see gemm_tc.cu for the single-CTA kernel and for
// why the L2 -> SM path, not the tensor pipe, bounds it).
//
// Two CTAs of a cluster (two SMs) work on one 256-row tile with tcgen05.mma.cta_group::2: each CTA
// stages its own 128 rows of A but only HALF of every weight chunk (N/2 columns), the pair's
// tensor cores read both halves, and each CTA ends up with its 128 rows x N accumulator in its
// own TMEM.  Per SM that halves the weight bytes that have to come out of L2 -- the term that
// dominates the single-CTA kernel (8 B per weight element with hi/lo images, re-streamed for
// every row tile).
//
// Roles per CTA are those of gemm_tc.cu (two producer groups, epilogue warps, loader thread); only
// the even CTA of the pair issues MMAs.  Cross-CTA protocol:
//   * stage s may be consumed when it is full in BOTH CTAs: the odd CTA's warp 12 forwards "my
//     stage s is full" to the even CTA's peer_full[s] mbarrier (remote arrive);
//   * tcgen05.commit.cta_group::2 ... multicast::cluster signals "stage s consumed" (empty[s]) and
//     "accumulator complete" (acc_full) to both CTAs at once;
//   * the odd CTA's epilogue threads arrive remotely on the even CTA's acc_empty barrier;
//   * cluster barriers bracket the kernel so that no CTA touches a peer that has not initialised
//     its barriers or has already exited.
// All waits are bounded (trap instead of hang).
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace flid {

namespace {

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Relaxed variant for the epilogue's "accumulator drained" signal: what must be ordered before it are the
// TMEM reads (tcgen05.wait::ld + tcgen05.fence::before_thread_sync), not the C stores; a cluster-scope
// release would make every epilogue thread wait for its global stores to be acknowledged.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {  // arrives on `bar` in both CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                               uint32_t accum) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(z)
        : "memory");
}

}  // namespace

// ---------------------------------------------------------------- weight tiling for the pair
// image layout: [n_block][k_chunk][cta][half][c4][n_tile / 2] float4.  CTA r holds, of every MMA column
// group, the r-th half of its columns (cta_group::2 splits B along N).
__global__ void tc_prep_pair_kernel(const float* __restrict__ W, int64_t ldw, int N, int K, int n_tile, int n_a,
                                    int n_blocks, int k_chunks, float4* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int half_rows = n_tile / 2;
    const int64_t total = (int64_t)n_blocks * k_chunks * 2 * 2 * C4 * half_rows;
    if (idx >= total) return;
    int64_t r = idx;
    const int j = (int)(r % half_rows);
    r /= half_rows;
    const int c = (int)(r % C4);
    r /= C4;
    const int half = (int)(r % 2);
    r /= 2;
    const int cta = (int)(r % 2);
    r /= 2;
    const int kc = (int)(r % k_chunks);
    const int nb = (int)(r / k_chunks);
    const int n_b = n_tile - n_a;
    const int col = j < n_a / 2 ? cta * (n_a / 2) + j : n_a + cta * (n_b / 2) + (j - n_a / 2);
    const int n = nb * n_tile + col, k = kc * KC + c * 4;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float x = (n < N && k + e < K) ? W[(int64_t)n * ldw + k + e] : 0.f;
        const float hi = tf32_hi(x);
        v[e] = half ? (x - hi) : hi;
    }
    out[idx] = make_float4(v[0], v[1], v[2], v[3]);
}

struct PairShape {
    int N, n_tile, n_a, n_blocks, k_chunks, stages;
    int acc_bufs;         // 2 when two accumulator sets fit in the 512 TMEM columns
    int staged_epilogue;
    int64_t m_pairs;      // 256-row tiles
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
    gemm_tc_pair_kernel(TcGemmArgs g, const float4* __restrict__ wbuf, PairShape sh) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_peer_full[MAX_STAGES];
    __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();                 // 0 = MMA issuer of the pair
    const uint32_t cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int half_rows = sh.n_tile / 2;
    const uint32_t b_half = (uint32_t)C4 * half_rows * 16;  // one hi (or lo) half-chunk of this CTA
    const uint32_t stage_bytes = A_SUB + 2 * b_half;
    const uint32_t nblk = (uint32_t)sh.n_blocks;
    const uint32_t work = (uint32_t)(sh.m_pairs * sh.n_blocks);

    if (tid == 0) {
        for (int s = 0; s < sh.stages; ++s)
            mbar_init(&bar_full[s], NPROD + 1), mbar_init(&bar_empty[s], 1), mbar_init(&bar_peer_full[s], 1);
        for (int a = 0; a < 2; ++a) mbar_init(&bar_acc_full[a], 1), mbar_init(&bar_acc_empty[a], 2 * NEPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers exist and its TMEM is allocated before anything crosses over
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int ktot = g.w0 + g.w1;

    if (warp < 8) {
        // ===================================================== producers (as in gemm_tc.cu, MS = 1)
        constexpr int NL = 4;
        const int pg = warp >> 2, pw = warp & 3;
        const int rsub = lane >> 2, c = lane & 3;
        const uint32_t dq = n_clusters / nblk, dr = n_clusters % nblk;
        struct Cursor {
            uint32_t t, mg, nb;
            int kc;
        };
        auto next_item = [&](Cursor& cu) {
            cu.t += n_clusters, cu.mg += dq, cu.nb += dr;
            if (cu.nb >= nblk) cu.nb -= nblk, cu.mg += 1;
        };
        Cursor lc{cid, cid / nblk, cid % nblk, pg};
        const float* p0[NL];
        const float* p1[NL];
        uint32_t okmask = 0;
        auto bind_rows = [&]() {
            okmask = 0;
            const int64_t m0 = (int64_t)lc.mg * 256 + rank * 128;
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
        if (lc.kc >= sh.k_chunks) lc.kc -= sh.k_chunks, next_item(lc);
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
            if (lc.kc >= sh.k_chunks) {
                lc.kc -= sh.k_chunks;
                next_item(lc);
                if (lc.kc >= sh.k_chunks) lc.kc -= sh.k_chunks, next_item(lc);
                bind_rows();
            }
        };
        const uint32_t my_items = work > cid ? (work - cid + n_clusters - 1) / n_clusters : 0;
        const uint32_t total_q = my_items * (uint32_t)sh.k_chunks;
        uint32_t sq = (uint32_t)pg;
        uint32_t stage = (uint32_t)pg % (uint32_t)sh.stages, phase = 0;
        auto store_next = [&](const float4 (&v)[NL]) {
            mbar_wait(&bar_empty[stage], phase ^ 1);
            uint8_t* st = smem + (size_t)stage * stage_bytes + c * A_CSTRIDE + (pw * 8 + rsub) * 16;
#pragma unroll
            for (int i = 0; i < NL; ++i) {
                const float4 x = v[i];
                const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                uint8_t* dst = st + i * (32 * 16);
                *reinterpret_cast<float4*>(dst) = hi;
                *reinterpret_cast<float4*>(dst + A_HALF) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
            }
            fence_async_smem();
            mbar_arrive(&bar_full[stage]);
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
        // ===================================================== epilogue (own 128 rows of the pair's tile)
        const int ew = warp - 8;
        float* stg = reinterpret_cast<float*>(smem + (size_t)sh.stages * stage_bytes) + ew * (32 * EPI_LD);
        const bool vec_ok = (g.ldc & 3) == 0 && (sh.N & 3) == 0;
        const bool staged = vec_ok && sh.staged_epilogue;
        const int sr = lane >> 2, sc = (lane & 3) * 4;
        const uint32_t acc_empty_leader0 = map_to_cta(smem_u32(&bar_acc_empty[0]), 0);
        const uint32_t acc_empty_leader1 = map_to_cta(smem_u32(&bar_acc_empty[1]), 0);
        uint32_t it = 0;
        for (uint32_t t = cid; t < work; t += n_clusters, ++it) {
            const uint32_t mg = t / nblk, nb = t - mg * nblk;
            const uint32_t acc = sh.acc_bufs == 2 ? (it & 1) : 0;
            const uint32_t par = sh.acc_bufs == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(&bar_acc_full[acc], par);
            tc_fence_after();
            const int64_t row0 = (int64_t)mg * 256 + rank * 128 + ew * 32;
            const uint32_t taddr = tmem + acc * (uint32_t)sh.n_tile + ((uint32_t)(ew * 32) << 16);
            float* crow4[4];
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
                if (n0 >= sh.N) continue;
                if (staged) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (g.bias && n0 + i < sh.N) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + i));
                            o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
                        }
                        if (g.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
                        *reinterpret_cast<float4*>(stg + lane * EPI_LD + i) = o;
                    }
                    __syncwarp();
                    if (n0 + sc < sh.N) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 o = *reinterpret_cast<const float4*>(stg + (j * 8 + sr) * EPI_LD + sc);
                            if (crow4[j] != nullptr) *reinterpret_cast<float4*>(crow4[j] + n0 + sc) = o;
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
            tc_fence_before();
            mbar_arrive_remote_relaxed(acc ? acc_empty_leader1 : acc_empty_leader0);  // the even CTA's barrier (own one for rank 0)
        }
    } else if (warp == 12 && lane == 0 && rank == 0) {
        // ===================================================== MMA issuer of the pair
        const uint32_t n_a = (uint32_t)sh.n_a, n_b = (uint32_t)sh.n_tile - n_a;
        auto make_idesc = [](uint32_t n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | (16u << 24); };  // M = 256
        const uint32_t idesc = make_idesc(n_a), idesc_b = n_b ? make_idesc(n_b) : 0u;
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t b_lbo = (uint32_t)half_rows * 16;  // c4 block stride of this CTA's half-chunk
        uint32_t it = 0, s = 0, ph = 0;
        for (uint32_t t = cid; t < work; t += n_clusters, ++it) {
            const uint32_t acc = sh.acc_bufs == 2 ? (it & 1) : 0;
            const uint32_t par = sh.acc_bufs == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(&bar_acc_empty[acc], par ^ 1);
            tc_fence_after();
            const uint32_t d = tmem + acc * (uint32_t)sh.n_tile;
            for (int kc = 0; kc < sh.k_chunks; ++kc) {
                mbar_wait(&bar_full[s], ph);
                mbar_wait(&bar_peer_full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_base + s * stage_bytes;
                const uint32_t sb = sa + A_SUB;
#pragma unroll
                for (int j = 0; j < KC / 8; ++j) {
                    const uint64_t d_bhi = umma_desc(sb + (2 * j) * b_lbo, b_lbo, 128);
                    const uint64_t d_blo = umma_desc(sb + b_half + (2 * j) * b_lbo, b_lbo, 128);
                    const uint32_t a0 = sa + (2 * j) * A_CSTRIDE;
                    const uint64_t d_ahi = umma_desc(a0, A_CSTRIDE, 128);
                    const uint64_t d_alo = umma_desc(a0 + A_HALF, A_CSTRIDE, 128);
                    umma_tf32_pair(d, d_alo, d_bhi, idesc, (kc | j) ? 1u : 0u);
                    umma_tf32_pair(d, d_ahi, d_blo, idesc, 1u);
                    umma_tf32_pair(d, d_ahi, d_bhi, idesc, 1u);
                    if (n_b) {  // second column group: rows n_a / 2 .. of this CTA's half-chunk
                        const uint64_t row_off = (uint64_t)(n_a / 2);  // 16 B rows, start-address field is in 16 B units
                        umma_tf32_pair(d + n_a, d_alo, d_bhi + row_off, idesc_b, (kc | j) ? 1u : 0u);
                        umma_tf32_pair(d + n_a, d_ahi, d_blo + row_off, idesc_b, 1u);
                        umma_tf32_pair(d + n_a, d_ahi, d_bhi + row_off, idesc_b, 1u);
                    }
                }
                tc_commit_pair(&bar_empty[s]);  // both CTAs' stage s is free once these MMAs have read it
                if (++s == (uint32_t)sh.stages) s = 0, ph ^= 1;
            }
            tc_commit_pair(&bar_acc_full[acc]);
        }
    } else if (warp == 12 && lane == 0 && rank == 1) {
        // ===================================================== forwarder: "my stage is full" -> the issuer
        uint32_t s = 0, ph = 0;
        uint32_t remote[MAX_STAGES];
#pragma unroll
        for (int i = 0; i < MAX_STAGES; ++i) remote[i] = map_to_cta(smem_u32(&bar_peer_full[i]), 0);
        for (uint32_t t = cid; t < work; t += n_clusters) {
            for (int kc = 0; kc < sh.k_chunks; ++kc) {
                mbar_wait(&bar_full[s], ph);
                uint32_t dst = remote[0];
#pragma unroll
                for (int i = 1; i < MAX_STAGES; ++i)
                    if ((uint32_t)i == s) dst = remote[i];
                mbar_arrive_remote(dst);
                if (++s == (uint32_t)sh.stages) s = 0, ph ^= 1;
            }
        }
    } else if (warp == 13 && lane == 0) {
        // ===================================================== weight loader: this CTA's half of every chunk
        uint32_t s = 0, ph = 0;
        const int64_t half4 = (int64_t)2 * C4 * half_rows;  // float4 per (n block, K chunk, cta)
        for (uint32_t t = cid; t < work; t += n_clusters) {
            const uint32_t nb = t % nblk;
            const float4* wsrc = wbuf + ((int64_t)nb * sh.k_chunks * 2 + rank) * half4;
            for (int kc = 0; kc < sh.k_chunks; ++kc) {
                mbar_wait(&bar_empty[s], ph ^ 1);
                uint8_t* st = smem + (size_t)s * stage_bytes + A_SUB;
                mbar_arrive_expect_tx(&bar_full[s], 2 * b_half);
                bulk_g2s(st, wsrc + (int64_t)kc * 2 * half4, 2 * b_half, &bar_full[s]);
                if (++s == (uint32_t)sh.stages) s = 0, ph ^= 1;
            }
        }
    }
    // In the even CTA peer_full of rank 0 itself is never signalled by a forwarder: the issuer waits on it, so
    // rank 0 has nothing to do here; rank 1's issuer slot is the forwarder above.
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody exits (or frees TMEM) while the peer may still signal or multicast into it
    if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// ---------------------------------------------------------------- host side
int tc_prepare_weight_pair(const float* W, int64_t ldw, TcWeight* w, cudaStream_t st) {
    const int n_a = w->n_tile > 256 ? (w->n_tile / 2 + 15) / 16 * 16 : w->n_tile;
    const int n_b = w->n_tile - n_a;
    // cta_group::2 splits every MMA column group in two: the halves must be whole 8-row core matrices
    if ((n_a / 2) % 8 != 0 || (n_b / 2) % 8 != 0) {
        w->pair_ok = 0;
        return FLID_OK;
    }
    if (!w->buf_pair) FLID_CUDA(cudaMalloc((void**)&w->buf_pair, w->bytes()));
    const int64_t total = (int64_t)w->n_blocks * w->k_chunks * 2 * 2 * C4 * (w->n_tile / 2);
    tc_prep_pair_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(W, ldw, w->N, w->K, w->n_tile, n_a, w->n_blocks,
                                                                       w->k_chunks, reinterpret_cast<float4*>(w->buf_pair));
    FLID_LAUNCH_CHECK();
    w->pair_ok = 1;
    return FLID_OK;
}

int tc_gemm_pair(const TcGemmArgs& g, const TcWeight& w, int sm_count, int smem_max, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        FLID_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_max - STATIC_SMEM));
        attr_set = true;
    }
    PairShape sh;
    sh.N = w.N, sh.n_tile = w.n_tile, sh.n_blocks = w.n_blocks, sh.k_chunks = w.k_chunks;
    sh.n_a = w.n_tile > 256 ? (w.n_tile / 2 + 15) / 16 * 16 : w.n_tile;
    sh.m_pairs = ceil_div(g.M, 256);
    sh.acc_bufs = 2 * w.n_tile <= 512 ? 2 : 1;
    sh.staged_epilogue = (g.cidx != nullptr || w.k_chunks >= 48 || w.N >= 512) ? 1 : 0;
    const size_t stage = (size_t)A_SUB + 2 * (size_t)C4 * (w.n_tile / 2) * 16;
    const size_t ring_bytes = (size_t)(smem_max - STATIC_SMEM) - EPI_BYTES;
    int stages = (int)(ring_bytes / stage);
    sh.stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    FLID_REQUIRE(sh.stages >= 3, "tc_gemm_pair: tile does not fit in shared memory");
    const int64_t work = sh.m_pairs * sh.n_blocks;
    FLID_REQUIRE(work < (1LL << 31) - 65536, "tc_gemm_pair: too many tiles for one launch");
    const int64_t max_clusters = sm_count / 2;
    const unsigned grid = 2u * (unsigned)(work < max_clusters ? work : max_clusters);
    gemm_tc_pair_kernel<<<grid, NTHREADS, sh.stages * stage + EPI_BYTES, st>>>(
        g, reinterpret_cast<const float4*>(w.buf_pair), sh);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

}  // namespace flid
