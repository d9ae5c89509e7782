// tcgen05 / TMEM GEMM (see gemm_tc.cuh): persistent, warp-specialised.
//
//   warps 0-3  producers : coalesced fp32 loads of the A tile (row gather fused), hi/lo tf32
//                          split, st.shared into the UMMA canonical K-major layout; thread 0
//                          also launches the bulk-async copy (cp.async.bulk, SASS UBLKCP) of the
//                          pre-tiled weight chunk, completing on the stage's mbarrier
//   warps 4-7  epilogue  : tcgen05.ld their 32 TMEM lanes, bias / ReLU, store C
//   warp  8    issuer    : one thread waits "stage full", issues 12 tcgen05.mma.kind::tf32
//                          (4 k-steps x {lo.hi, hi.lo, hi.hi}) and tcgen05.commit's the stage's
//                          "empty" mbarrier; after the last K chunk it commits "accumulator full"
//
// Ring of `stages` shared-memory stages (32 K-floats each) + two TMEM accumulators, so the
// producers, the tensor core and the epilogue of the previous tile all overlap.  One CTA per
// SM, tiles (128 rows x n_tile columns) are dealt round-robin.
#include "gemm_tc.cuh"

namespace flid {

constexpr int KC = TC_KC;                 // 32 floats per stage
constexpr int C4 = KC / 4;                // 16-byte chunks per row per stage
constexpr uint32_t A_CSTRIDE = 129 * 16;  // byte stride between K chunks of A (odd in 16 B units: conflict-free stores)
constexpr uint32_t A_HALF = C4 * A_CSTRIDE;
constexpr int NPROD = 128, NEPI = 128, NTHREADS = 288;

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t b = smem_u32(bar);
    for (uint32_t spin = 0; !mbar_try(b, parity); ++spin)
        if (spin > (1u << 28)) __trap();
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// K-major, no swizzle: 8 rows x 16 B core matrices; LBO = byte distance between the two 16-byte
// K chunks of one MMA, SBO = byte distance between 8-row groups (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// ---------------------------------------------------------------- weight tiling
// image layout: [n_block][k_chunk][half][c4 = 8][n_tile] float4
__global__ void tc_prep_kernel(const float* __restrict__ W, int64_t ldw, int N, int K, int n_tile, int n_blocks,
                               int k_chunks, float4* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)n_blocks * k_chunks * 2 * C4 * n_tile;
    if (idx >= total) return;
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
        const float x = (n < N && k + e < K) ? W[(int64_t)n * ldw + k + e] : 0.f;
        const float hi = tf32_hi(x);
        v[e] = half ? (x - hi) : hi;
    }
    out[idx] = make_float4(v[0], v[1], v[2], v[3]);
}

// ---------------------------------------------------------------- the GEMM
struct TcShape {
    int N, n_tile, n_blocks, k_chunks, stages;
    int64_t m_tiles;
    uint32_t tmem_stride;  // columns between the two accumulators
};

__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(TcGemmArgs g, const float4* __restrict__ wbuf, TcShape sh) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[4], bar_empty[4], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b_half = (uint32_t)C4 * sh.n_tile * 16;
    const uint32_t stage_bytes = 2 * A_HALF + 2 * b_half;
    const int64_t work = sh.m_tiles * sh.n_blocks;

    if (tid == 0) {
        for (int s = 0; s < sh.stages; ++s) mbar_init(&bar_full[s], NPROD), mbar_init(&bar_empty[s], 1);
        for (int a = 0; a < 2; ++a) mbar_init(&bar_acc_full[a], 1), mbar_init(&bar_acc_empty[a], NEPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int ktot = g.w0 + g.w1;

    if (warp < 4) {
        // ===================================================== producers
        const int rsub = lane >> 3, c = lane & 7;  // 4 rows x 8 float4 per warp instruction
        uint32_t q = 0;                             // chunk counter across tiles
        for (int64_t t = blockIdx.x; t < work; t += gridDim.x) {
            const int nb = (int)(t % sh.n_blocks);
            const int64_t m0 = (t / sh.n_blocks) * 128;
            const float* p0[8];
            const float* p1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t row = m0 + warp * 32 + i * 4 + rsub;
                p0[i] = nullptr, p1[i] = nullptr;
                if (row < g.M) {
                    p0[i] = g.A0 + (g.idx0 ? (int64_t)__ldg(g.idx0 + row) : row) * g.lda0;
                    if (g.w1 > 0) p1[i] = g.A1 + (g.idx1 ? (int64_t)__ldg(g.idx1 + row) : row) * g.lda1;
                }
            }
            const float4* wsrc = wbuf + ((int64_t)nb * sh.k_chunks) * (2 * C4 * sh.n_tile);
            auto load_chunk = [&](int kc, float4 (&v)[8]) {
                const int k = kc * KC + c * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p0[i] != nullptr && k < ktot) {
                        if (k < g.w0)
                            v[i] = __ldg(reinterpret_cast<const float4*>(p0[i] + k));
                        else
                            v[i] = __ldg(reinterpret_cast<const float4*>(p1[i] + (k - g.w0)));
                    }
                }
            };
            float4 cur[8], nxt[8];
            load_chunk(0, cur);
            for (int kc = 0; kc < sh.k_chunks; ++kc, ++q) {
                if (kc + 1 < sh.k_chunks) load_chunk(kc + 1, nxt);  // in flight while this chunk is split / stored
                const uint32_t s = q % sh.stages, use = q / sh.stages;
                mbar_wait(&bar_empty[s], (use & 1) ^ 1);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                if (tid == 0) {
                    mbar_expect_tx(&bar_full[s], 2 * b_half);
                    bulk_g2s(st + 2 * A_HALF, wsrc + (int64_t)kc * (2 * C4 * sh.n_tile), 2 * b_half, &bar_full[s]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = warp * 32 + i * 4 + rsub;
                    const float4 v = cur[i];
                    const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    uint8_t* dst = st + c * A_CSTRIDE + r * 16;
                    *reinterpret_cast<float4*>(dst) = hi;
                    *reinterpret_cast<float4*>(dst + A_HALF) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                }
                fence_async_smem();
                mbar_arrive(&bar_full[s]);
#pragma unroll
                for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
            }
        }
    } else if (warp < 8) {
        // ===================================================== epilogue
        const int ew = warp - 4;  // TMEM lane quarter == warp id % 4
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < work; t += gridDim.x, ++it) {
            const int nb = (int)(t % sh.n_blocks);
            const int64_t row = (t / sh.n_blocks) * 128 + ew * 32 + lane;
            const uint32_t acc = it & 1;
            mbar_wait(&bar_acc_full[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + acc * sh.tmem_stride + ((uint32_t)(ew * 32) << 16);
            float* crow = (row < g.M) ? g.C + row * g.ldc : nullptr;
            for (int c0 = 0; c0 < sh.n_tile; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                const int n0 = nb * sh.n_tile + c0;
                if (crow != nullptr) {
                    if (n0 + 16 <= sh.N && (g.ldc & 3) == 0) {
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
            }
            tc_fence_before();
            mbar_arrive(&bar_acc_empty[acc]);
        }
    } else if (lane == 0) {
        // ===================================================== MMA issuer (one thread)
        // instruction descriptor: D=f32, A=B=tf32, K-major both, N = n_tile, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(sh.n_tile >> 3) << 17) | (8u << 24);
        const uint32_t smem_base = smem_u32(smem);
        uint32_t q = 0, it = 0;
        for (int64_t t = blockIdx.x; t < work; t += gridDim.x, ++it) {
            const uint32_t acc = it & 1;
            mbar_wait(&bar_acc_empty[acc], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem + acc * sh.tmem_stride;
            for (int kc = 0; kc < sh.k_chunks; ++kc, ++q) {
                const uint32_t s = q % sh.stages, use = q / sh.stages;
                mbar_wait(&bar_full[s], use & 1);
                tc_fence_after();
                const uint32_t sa = smem_base + s * stage_bytes;
                const uint32_t sb = sa + 2 * A_HALF;
#pragma unroll
                for (int j = 0; j < KC / 8; ++j) {
                    const uint64_t d_ahi = umma_desc(sa + (2 * j) * A_CSTRIDE, A_CSTRIDE, 128);
                    const uint64_t d_alo = umma_desc(sa + A_HALF + (2 * j) * A_CSTRIDE, A_CSTRIDE, 128);
                    const uint64_t d_bhi = umma_desc(sb + (2 * j) * (sh.n_tile * 16), sh.n_tile * 16, 128);
                    const uint64_t d_blo = umma_desc(sb + b_half + (2 * j) * (sh.n_tile * 16), sh.n_tile * 16, 128);
                    umma_tf32(d, d_alo, d_bhi, idesc, (kc | j) ? 1u : 0u);  // small terms first
                    umma_tf32(d, d_ahi, d_blo, idesc, 1u);
                    umma_tf32(d, d_ahi, d_bhi, idesc, 1u);
                }
                tc_commit(&bar_empty[s]);  // frees the smem stage when these MMAs have read it
            }
            tc_commit(&bar_acc_full[acc]);  // accumulator complete -> epilogue
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// ---------------------------------------------------------------- host side
static int pick_n_tile(int N) {
    const int blocks = (N + 255) / 256;
    int t = (N + blocks - 1) / blocks;
    t = (t + 15) / 16 * 16;
    return t < 16 ? 16 : t;
}

int tc_prepare_weight(const float* W, int64_t ldw, int N, int K, TcWeight* w, cudaStream_t st) {
    FLID_REQUIRE(W && w && N > 0 && K > 0, "tc_prepare_weight: bad argument");
    const int n_tile = pick_n_tile(N);
    const int n_blocks = (N + n_tile - 1) / n_tile;
    const int k_chunks = (K + KC - 1) / KC;
    if (w->buf && (w->N != N || w->K != K)) {
        FLID_CUDA(cudaDeviceSynchronize());
        FLID_CUDA(cudaFree(w->buf));
        w->buf = nullptr;
    }
    w->N = N, w->K = K, w->n_tile = n_tile, w->n_blocks = n_blocks, w->k_chunks = k_chunks;
    if (!w->buf) FLID_CUDA(cudaMalloc((void**)&w->buf, w->bytes()));
    const int64_t total = (int64_t)n_blocks * k_chunks * 2 * C4 * n_tile;
    tc_prep_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(W, ldw, N, K, n_tile, n_blocks, k_chunks,
                                                                  reinterpret_cast<float4*>(w->buf));
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

void tc_free_weight(TcWeight* w) {
    if (w && w->buf) cudaFree(w->buf);
    if (w) w->buf = nullptr;
}

int tc_gemm(const TcGemmArgs& g, const TcWeight& w, cudaStream_t st) {
    if (g.M <= 0) return FLID_OK;
    FLID_REQUIRE(w.buf != nullptr, "tc_gemm: weight not prepared");
    FLID_REQUIRE(g.w0 > 0 && g.w0 + g.w1 == w.K, "tc_gemm: A width %d+%d != weight K %d", g.w0, g.w1, w.K);
    FLID_REQUIRE((g.w0 % 4) == 0 && (g.w1 % 4) == 0 && (g.lda0 % 4) == 0 && (g.lda1 % 4) == 0,
                 "tc_gemm: segment widths / row strides must be multiples of 4 floats");
    static int sm_count = 0, smem_max = 0;
    if (sm_count == 0) {
        int dev = 0;
        FLID_CUDA(cudaGetDevice(&dev));
        FLID_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        FLID_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        FLID_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 1024));
    }
    TcShape sh;
    sh.N = w.N, sh.n_tile = w.n_tile, sh.n_blocks = w.n_blocks, sh.k_chunks = w.k_chunks;
    sh.m_tiles = ceil_div(g.M, 128);
    sh.tmem_stride = 256;
    const size_t stage = 2 * (size_t)A_HALF + 2 * (size_t)C4 * w.n_tile * 16;
    int stages = (int)((size_t)(smem_max - 1024) / stage);
    stages = stages > 4 ? 4 : stages;
    FLID_REQUIRE(stages >= 2, "tc_gemm: tile does not fit in shared memory");
    sh.stages = stages;
    const int64_t work = sh.m_tiles * sh.n_blocks;
    const unsigned grid = (unsigned)(work < sm_count ? work : sm_count);
    gemm_tc_kernel<<<grid, NTHREADS, stages * stage, st>>>(g, reinterpret_cast<const float4*>(w.buf), sh);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

}  // namespace flid
