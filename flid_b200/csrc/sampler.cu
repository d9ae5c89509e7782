// 'recent' neighbour sampler: warp per query.
//   find_neighbors_before   utils/utils.py:130-147  (searchsorted, side='left')
//   get_historical_neighbors utils/utils.py:149-214 ('recent': last <=k, right-aligned)
// Integer / index work, bit-exact.  HBM-bound: ~16 B query + a few timestamp probes +
// 16 B per gathered entry + 20 B per output slot.
#include "graph.cuh"

namespace flid {

template <bool F32>
__device__ __forceinline__ double load_time(const void* times, int64_t i) {
    if (F32) return (double)__ldg(reinterpret_cast<const float*>(times) + i);  // widening is exact
    return __ldg(reinterpret_cast<const double*>(times) + i);
}

// API-facing form: int64 / int64 / float32 [n, k]
template <bool F32>
__global__ void __launch_bounds__(256) sample_recent_kernel(const int64_t* __restrict__ indptr,
                                                            const int2* __restrict__ adj,
                                                            const double* __restrict__ ts, int64_t num_nodes,
                                                            const int64_t* __restrict__ nodes,
                                                            const void* __restrict__ times, int64_t n, int k,
                                                            int64_t* __restrict__ out_nbr,
                                                            int64_t* __restrict__ out_eid, float* __restrict__ out_ts,
                                                            int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= n) return;
    const int64_t v = __ldg(nodes + q);
    int64_t start = 0, cut = 0;
    if (v < 0 || v > num_nodes) {
        if (lane == 0) atomicExch(bad, 1);
    } else {
        start = __ldg(indptr + v);
        cut = warp_lower_bound(ts, start, __ldg(indptr + v + 1), load_time<F32>(times, q), lane);
    }
    const int64_t have = cut - start;
    const int cnt = have < (int64_t)k ? (int)have : k;
    for (int j = lane; j < k; j += 32) {
        int64_t a = 0, e = 0;
        float t = 0.f;
        if (j >= k - cnt) {
            const int64_t p = cut - k + j;
            const int2 ne = __ldg(adj + p);
            a = ne.x, e = ne.y, t = (float)__ldg(ts + p);  // cvt.rn.f32.f64
        }
        out_nbr[q * k + j] = a;
        out_eid[q * k + j] = e;
        out_ts[q * k + j] = t;
    }
}

template <bool F32>
__global__ void __launch_bounds__(256) sample_cut_kernel(const int64_t* __restrict__ indptr,
                                                         const double* __restrict__ ts, int64_t num_nodes,
                                                         const int64_t* __restrict__ nodes,
                                                         const void* __restrict__ times, int64_t n,
                                                         int64_t* __restrict__ out_start,
                                                         int64_t* __restrict__ out_cut, int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= n) return;
    const int64_t v = __ldg(nodes + q);
    int64_t start = 0, cut = 0;
    if (v < 0 || v > num_nodes) {
        if (lane == 0) atomicExch(bad, 1);
    } else {
        start = __ldg(indptr + v);
        cut = warp_lower_bound(ts, start, __ldg(indptr + v + 1), load_time<F32>(times, q), lane);
    }
    if (lane == 0) out_start[q] = start, out_cut[q] = cut;
}

// the graph's status flag is zero between calls: read it back (the host result arrays need the sync anyway) and
// clear it again only on the error path -- no allocation, no memset per call
static int check_bad(int* bad, cudaStream_t st, const char* who) {
    int h = 0;
    FLID_CUDA(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLID_CUDA(cudaStreamSynchronize(st));
    if (h) {
        cudaMemsetAsync(bad, 0, sizeof(int), st);
        set_error("%s: node id outside the graph", who);
        return FLID_ERR_RANGE;
    }
    return FLID_OK;
}

}  // namespace flid

extern "C" {

int flid_sample_recent(const flid_graph* g, const int64_t* nodes, const void* times, int times_are_f32, int64_t n,
                       int k, int64_t* out_nbr, int64_t* out_eid, float* out_ts, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(g != nullptr, "flid_sample_recent: null graph");
    // the reference's message (utils/utils.py:157)
    FLID_REQUIRE(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    FLID_REQUIRE(n >= 0, "flid_sample_recent: negative n");
    if (n == 0) return FLID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int* bad = g->bad_flag;
    const unsigned blocks = (unsigned)ceil_div(n * 32, 256);
    if (times_are_f32)
        sample_recent_kernel<true><<<blocks, 256, 0, st>>>(g->indptr, g->adj, g->ts, g->num_nodes, nodes, times, n, k,
                                                           out_nbr, out_eid, out_ts, bad);
    else
        sample_recent_kernel<false><<<blocks, 256, 0, st>>>(g->indptr, g->adj, g->ts, g->num_nodes, nodes, times, n,
                                                            k, out_nbr, out_eid, out_ts, bad);
    FLID_LAUNCH_CHECK();
    return check_bad(bad, st, "flid_sample_recent");
}

int flid_sample_cut(const flid_graph* g, const int64_t* nodes, const void* times, int times_are_f32, int64_t n,
                    int64_t* out_start, int64_t* out_cut, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(g != nullptr, "flid_sample_cut: null graph");
    if (n <= 0) return FLID_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int* bad = g->bad_flag;
    const unsigned blocks = (unsigned)ceil_div(n * 32, 256);
    if (times_are_f32)
        sample_cut_kernel<true><<<blocks, 256, 0, st>>>(g->indptr, g->ts, g->num_nodes, nodes, times, n, out_start,
                                                        out_cut, bad);
    else
        sample_cut_kernel<false><<<blocks, 256, 0, st>>>(g->indptr, g->ts, g->num_nodes, nodes, times, n, out_start,
                                                         out_cut, bad);
    FLID_LAUNCH_CHECK();
    return check_bad(bad, st, "flid_sample_cut");
}

}  // extern "C"
