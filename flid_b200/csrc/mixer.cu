// Node encoder of GraphMixer (models/GraphMixer.py:119-146), another consumer of the time-sorted CSR
// (SURVEY 8(f) rank 4): for every query (node, t) the `time_gap` most recent neighbours before t, their raw
// feature rows averaged as the reference does it --
//     scores = softmax over the time_gap slots of {1 for a real neighbour, -1e10 for a padded slot}
//     agg    = mean_j (x_j * scores_j)          (mean over ALL time_gap slots: divides by time_gap again)
//     out    = agg + node_feat[node]
// i.e. (1 / time_gap) * (1 / n_valid) * sum of the valid rows; with no valid neighbour the softmax is uniform over
// time_gap copies of the padding row node_feat[0].  A pure gather / reduce: one warp per query, the neighbour ids
// of 32 slots fetched with one coalesced load, four feature rows in flight per lane, 16-byte loads.
#include "graph.cuh"

namespace flid {
namespace {

constexpr int MIX_C = 4;  // float4 chunks per lane: node_dim <= 512

template <bool F32>
__global__ void __launch_bounds__(256) neighbor_mean_kernel(const int64_t* __restrict__ indptr, const int2* __restrict__ adj,
                                                            const double* __restrict__ ts,
                                                            const float* __restrict__ node_feat, int dn,
                                                            const int64_t* __restrict__ nodes, const void* __restrict__ times,
                                                            int64_t n, int time_gap, int add_self, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= n) return;
    const int64_t v = __ldg(nodes + q);
    const double t = F32 ? (double)__ldg(reinterpret_cast<const float*>(times) + q) : __ldg(reinterpret_cast<const double*>(times) + q);
    const int64_t start = __ldg(indptr + v);
    const int64_t cut = warp_lower_bound(ts, start, __ldg(indptr + v + 1), t, lane);
    const int64_t have = cut - start;
    const int64_t cnt = have < (int64_t)time_gap ? have : (int64_t)time_gap;
    const int c4 = dn >> 2;
    float4 acc[MIX_C];
#pragma unroll
    for (int r = 0; r < MIX_C; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    int valid = 0;
    for (int64_t p0 = cut - cnt; p0 < cut; p0 += 32) {
        const int m = (int)((cut - p0) < 32 ? (cut - p0) : 32);
        const int id_l = lane < m ? __ldg(adj + p0 + lane).x : 0;
        valid += __popc(__ballot_sync(FULL, id_l > 0));
        for (int j = 0; j < m; j += 4) {
            float4 x[4][MIX_C];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int id = __shfl_sync(FULL, id_l, (j + u) & 31);   // slots past m carry id 0: the zero-weight padding row
                const float4* row = reinterpret_cast<const float4*>(node_feat + (int64_t)id * dn);
#pragma unroll
                for (int r = 0; r < MIX_C; ++r) {
                    const int f = lane + 32 * r;
                    x[u][r] = (f < c4 && id > 0 && j + u < m) ? __ldg(row + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int r = 0; r < MIX_C; ++r)
                    acc[r].x += x[u][r].x, acc[r].y += x[u][r].y, acc[r].z += x[u][r].z, acc[r].w += x[u][r].w;
        }
    }
    const float inv_gap = 1.0f / (float)time_gap;
    const float4* self = reinterpret_cast<const float4*>(node_feat + v * dn);
    const float4* pad = reinterpret_cast<const float4*>(node_feat);
#pragma unroll
    for (int r = 0; r < MIX_C; ++r) {
        const int f = lane + 32 * r;
        if (f >= c4) continue;
        float4 a;
        if (valid > 0) {
            const float s = 1.0f / (float)valid;   // softmax weight of a real neighbour
            a = make_float4(acc[r].x * s * inv_gap, acc[r].y * s * inv_gap, acc[r].z * s * inv_gap, acc[r].w * s * inv_gap);
        } else {                                   // uniform 1 / time_gap over time_gap copies of the padding row
            const float4 z = __ldg(pad + f);
            const float g = (float)time_gap;
            a = make_float4(z.x * inv_gap * g * inv_gap, z.y * inv_gap * g * inv_gap, z.z * inv_gap * g * inv_gap,
                            z.w * inv_gap * g * inv_gap);
        }
        if (add_self) {
            const float4 h = __ldg(self + f);
            a.x += h.x, a.y += h.y, a.z += h.z, a.w += h.w;
        }
        reinterpret_cast<float4*>(out + q * dn)[f] = a;
    }
}

}  // namespace
}  // namespace flid

extern "C" int flid_neighbor_mean(const flid_graph* g, const float* node_feat, int node_dim, const int64_t* nodes,
                                  const void* times, int times_are_f32, int64_t n, int time_gap, int add_self, float* out,
                                  flid_stream stream) {
    using namespace flid;
    if (n <= 0) return FLID_OK;
    FLID_REQUIRE(g && node_feat && nodes && times && out, "flid_neighbor_mean: null argument");
    FLID_REQUIRE(node_dim > 0 && node_dim % 4 == 0 && node_dim <= 128 * MIX_C, "flid_neighbor_mean: node_dim must be a multiple of 4, <= %d", 128 * MIX_C);
    FLID_REQUIRE(time_gap > 0, "flid_neighbor_mean: time_gap must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)ceil_div(n * 32, 256);
    if (times_are_f32)
        neighbor_mean_kernel<true><<<blocks, 256, 0, st>>>(g->indptr, g->adj, g->ts, node_feat, node_dim, nodes, times, n, time_gap, add_self, out);
    else
        neighbor_mean_kernel<false><<<blocks, 256, 0, st>>>(g->indptr, g->adj, g->ts, node_feat, node_dim, nodes, times, n, time_gap, add_self, out);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}
