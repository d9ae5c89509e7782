// fp32 SIMT NT GEMM (see gemm.cuh).
#include "gemm.cuh"

namespace flid {

constexpr int GBM = 64, GBN = 64, GBK = 16;

__global__ void __launch_bounds__(256) gemm_nt_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[GBK][GBM + 4];
    __shared__ __align__(16) float Ws[GBK][GBN + 4];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * GBM;
    const int n0 = blockIdx.y * GBN;
    const int lrow = t >> 2, lk = (t & 3) * 4;

    const int64_t am = m0 + lrow;
    const float* arow = nullptr;
    if (am < g.M) arow = g.A + (g.a_idx ? (int64_t)__ldg(g.a_idx + am) : am) * g.lda;
    const int wn = n0 + lrow;
    const float* wrow = (wn < g.N) ? g.W + (int64_t)wn * g.ldw : nullptr;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += GBK) {
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), w4 = a4;
        if (arow && k0 + lk < g.K) a4 = __ldg(reinterpret_cast<const float4*>(arow + k0 + lk));
        if (wrow && k0 + lk < g.K) w4 = __ldg(reinterpret_cast<const float4*>(wrow + k0 + lk));
        __syncthreads();
        As[lk + 0][lrow] = a4.x, As[lk + 1][lrow] = a4.y, As[lk + 2][lrow] = a4.z, As[lk + 3][lrow] = a4.w;
        Ws[lk + 0][lrow] = w4.x, Ws[lk + 1][lrow] = w4.y, Ws[lk + 2][lrow] = w4.z, Ws[lk + 3][lrow] = w4.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        const int64_t mrow = g.c_idx ? (int64_t)__ldg(g.c_idx + m) : m;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            float* c = g.C + mrow * g.ldc + n;
            if (g.accumulate) v += *c;
            if (g.bias) v += __ldg(g.bias + n);
            if (g.relu) v = fmaxf(v, 0.f);
            *c = v;
        }
    }
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return FLID_OK;
    FLID_REQUIRE((g.K % 4) == 0 && (g.lda % 4) == 0 && (g.ldw % 4) == 0, "gemm: K/lda/ldw must be multiples of 4");
    dim3 grid((unsigned)ceil_div(g.M, GBM), (unsigned)ceil_div(g.N, GBN));
    gemm_nt_kernel<<<grid, 256, 0, st>>>(g);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

}  // namespace flid
