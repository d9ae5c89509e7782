// tcgen05 / TMEM GEMM (see gemm_tc.cuh).  One CTA = 128 rows x n_tile columns:
//   * 128 threads stage A (fp32 -> hi/lo split, gathered rows) and the pre-tiled W chunk into a
//     2-stage shared-memory ring (generic-proxy stores + fence.proxy.async),
//   * thread 0 issues 6 tcgen05.mma.kind::tf32 per stage (2 k-steps x {hi.hi, hi.lo, lo.hi}) and
//     tcgen05.commit's the stage's mbarrier so the ring slot can be refilled while the tensor
//     core works,
//   * the accumulator [128 lanes x n_tile columns] lives in TMEM; the four warps read their 32
//     lanes back with tcgen05.ld for the bias / ReLU epilogue.
#include "gemm_tc.cuh"

namespace flid {

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
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
__global__ void tc_prep_kernel(const float* __restrict__ W, int64_t ldw, int N, int K, int n_tile, int n_blocks,
                               int k_chunks, float4* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)n_blocks * k_chunks * 2 * 4 * n_tile;
    if (idx >= total) return;
    int64_t r = idx;
    const int nl = (int)(r % n_tile);
    r /= n_tile;
    const int c = (int)(r % 4);
    r /= 4;
    const int half = (int)(r % 2);
    r /= 2;
    const int kc = (int)(r % k_chunks);
    const int nb = (int)(r / k_chunks);
    const int n = nb * n_tile + nl, k = kc * TC_KC + c * 4;
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
__global__ void __launch_bounds__(128) gemm_tc_kernel(TcGemmArgs g, const float4* __restrict__ wbuf, int N, int n_tile,
                                                      int n_blocks, int k_chunks, uint32_t tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_stage[2];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int nb = blockIdx.x % n_blocks;
    const int64_t m0 = (int64_t)(blockIdx.x / n_blocks) * 128;
    const int64_t row = m0 + tid;
    const uint32_t a_half = 4 * 128 * 16;          // bytes of one A half (hi or lo) per stage
    const uint32_t b_half = 4 * (uint32_t)n_tile * 16;
    const uint32_t stage_bytes = 2 * a_half + 2 * b_half;

    if (tid == 0) {
        mbar_init(&bar_stage[0], 1), mbar_init(&bar_stage[1], 1), mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const float* a0 = nullptr;
    const float* a1 = nullptr;
    if (row < g.M) {
        a0 = g.A0 + (g.idx0 ? (int64_t)__ldg(g.idx0 + row) : row) * g.lda0;
        if (g.w1 > 0) a1 = g.A1 + (g.idx1 ? (int64_t)__ldg(g.idx1 + row) : row) * g.lda1;
    }
    const int ktot = g.w0 + g.w1;
    // instruction descriptor: D=f32, A=B=tf32, K-major both, N = n_tile, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_tile >> 3) << 17) | (8u << 24);
    const float4* wsrc = wbuf + (int64_t)nb * k_chunks * (2 * 4 * n_tile);
    const uint32_t smem_base = smem_u32(smem);

    for (int kc = 0; kc < k_chunks; ++kc) {
        const int s = kc & 1;
        if (kc >= 2) mbar_wait(&bar_stage[s], (uint32_t)(((kc >> 1) - 1) & 1));  // MMAs that read this slot are done
        uint8_t* st = smem + (size_t)s * stage_bytes;
        float4* a_hi = reinterpret_cast<float4*>(st);
        float4* a_lo = reinterpret_cast<float4*>(st + a_half);
        float4* b_st = reinterpret_cast<float4*>(st + 2 * a_half);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = kc * TC_KC + c * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a0 != nullptr && k < ktot) {
                if (k < g.w0)
                    v = __ldg(reinterpret_cast<const float4*>(a0 + k));
                else
                    v = __ldg(reinterpret_cast<const float4*>(a1 + (k - g.w0)));
            }
            const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            a_hi[c * 128 + tid] = hi;
            a_lo[c * 128 + tid] = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        }
        const float4* wch = wsrc + (int64_t)kc * (2 * 4 * n_tile);
        for (int i = tid; i < 2 * 4 * n_tile; i += 128) b_st[i] = __ldg(wch + i);
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
            const uint32_t sb = sa + 2 * a_half;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint64_t d_ahi = umma_desc(sa + (2 * j) * (128 * 16), 128 * 16, 128);
                const uint64_t d_alo = umma_desc(sa + a_half + (2 * j) * (128 * 16), 128 * 16, 128);
                const uint64_t d_bhi = umma_desc(sb + (2 * j) * (n_tile * 16), n_tile * 16, 128);
                const uint64_t d_blo = umma_desc(sb + b_half + (2 * j) * (n_tile * 16), n_tile * 16, 128);
                umma_tf32(tmem, d_alo, d_bhi, idesc, (kc | j) ? 1u : 0u);   // small terms first
                umma_tf32(tmem, d_ahi, d_blo, idesc, 1u);
                umma_tf32(tmem, d_ahi, d_bhi, idesc, 1u);
            }
            tc_commit(&bar_stage[s]);
        }
    }
    if (tid == 0) tc_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    tc_fence_after();

    // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows m0 + 32w + lane
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    float* crow = (row < g.M) ? g.C + row * g.ldc : nullptr;
    for (int c0 = 0; c0 < n_tile; c0 += 16) {
        float v[16];
        tmem_ld16(lane_base + (uint32_t)c0, v);
        const int n0 = nb * n_tile + c0;
        if (crow != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = n0 + i;
                if (n < N) {
                    float x = v[i];
                    if (g.bias) x += __ldg(g.bias + n);
                    if (g.relu) x = fmaxf(x, 0.f);
                    crow[n] = x;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
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
    const int k_chunks = (K + TC_KC - 1) / TC_KC;
    if (w->buf && (w->N != N || w->K != K)) {
        FLID_CUDA(cudaDeviceSynchronize());
        FLID_CUDA(cudaFree(w->buf));
        w->buf = nullptr;
    }
    w->N = N, w->K = K, w->n_tile = n_tile, w->n_blocks = n_blocks, w->k_chunks = k_chunks;
    if (!w->buf) FLID_CUDA(cudaMalloc((void**)&w->buf, w->bytes()));
    const int64_t total = (int64_t)n_blocks * k_chunks * 2 * 4 * n_tile;
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
    uint32_t cols = 32;
    while (cols < (uint32_t)w.n_tile) cols <<= 1;
    const size_t smem = 2 * (2 * (size_t)4 * 128 * 16 + 2 * (size_t)4 * w.n_tile * 16);
    static size_t configured = 0;
    if (smem > configured) {
        FLID_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int64_t blocks = ceil_div(g.M, 128) * w.n_blocks;
    gemm_tc_kernel<<<(unsigned)blocks, 128, smem, st>>>(g, reinterpret_cast<const float4*>(w.buf), w.N, w.n_tile,
                                                       w.n_blocks, w.k_chunks, cols);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

}  // namespace flid
