// CSR construction on the device.  Replaces the Python loops of get_neighbor_sampler
// (utils/utils.py:283-302) and NeighborSampler.__init__ (utils/utils.py:73-110):
// every event is appended to both endpoints (src's list first), then each node's list
// is *stably* sorted on the timestamp alone, so ties keep insertion order.
//
// Device formulation: one entry per (owner, neighbour, edge, time) in insertion order,
// LSD radix sort -- stable sort on the order-preserving bit image of the float64 time,
// then stable sort on the owner id.  cub::DeviceRadixSort is used for the two sorts
// (library plumbing, one-off per graph; the hot path never sorts).
#include <algorithm>
#include <cub/cub.cuh>

#include "graph.cuh"

namespace flid {

__device__ __forceinline__ unsigned long long sortable_f64(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// events -> 2E entries in the reference's insertion order (src endpoint, then dst endpoint)
__global__ void expand_events_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                     const int64_t* __restrict__ eid, const double* __restrict__ ts, int64_t E,
                                     int64_t num_nodes, int32_t* owner, int2* adj, unsigned long long* tkey,
                                     int* bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= E) return;
    const int64_t s = src[i], d = dst[i], e = eid[i];
    if (s < 0 || s > num_nodes || d < 0 || d > num_nodes || e < 0 || e > 0x7fffffffLL) {
        atomicExch(bad, 1);
        return;
    }
    const unsigned long long tk = sortable_f64(ts[i] + 0.0);  // -0.0 -> +0.0, ties must compare equal
    owner[2 * i] = (int32_t)s;
    adj[2 * i] = make_int2((int)d, (int)e);
    tkey[2 * i] = tk;
    owner[2 * i + 1] = (int32_t)d;
    adj[2 * i + 1] = make_int2((int)s, (int)e);
    tkey[2 * i + 1] = tk;
}

__global__ void pack_entries_kernel(const int64_t* __restrict__ own, const int64_t* __restrict__ nbr,
                                    const int64_t* __restrict__ eid, const double* __restrict__ ts, int64_t M,
                                    int64_t num_nodes, int32_t* owner, int2* adj, unsigned long long* tkey, int* bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t o = own[i], u = nbr[i], e = eid[i];
    if (o < 0 || o > num_nodes || u < 0 || u > 0x7fffffffLL || e < 0 || e > 0x7fffffffLL) {
        atomicExch(bad, 1);
        return;
    }
    owner[i] = (int32_t)o;
    adj[i] = make_int2((int)u, (int)e);
    tkey[i] = sortable_f64(ts[i] + 0.0);
}

__global__ void iota_kernel(uint32_t* v, int64_t M) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < M) v[i] = (uint32_t)i;
}

__global__ void gather_owner_kernel(const int32_t* __restrict__ owner, const uint32_t* __restrict__ perm, int64_t M,
                                    int32_t* out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < M) out[i] = owner[perm[i]];
}

// final placement + degree histogram
__global__ void place_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ owner_sorted,
                             const int2* __restrict__ adj_in, const unsigned long long* __restrict__ tkey, int64_t M,
                             int2* adj_out, double* ts_out, unsigned long long* counts) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint32_t p = perm[i];
    adj_out[i] = adj_in[p];
    unsigned long long b = tkey[p];
    b = (b & 0x8000000000000000ull) ? (b & 0x7fffffffffffffffull) : ~b;
    ts_out[i] = __longlong_as_double((long long)b);
    atomicAdd(counts + owner_sorted[i], 1ull);
}

// Entries 2i and 2i+1 of the insertion order belong to the same event: after the sort, each
// entry learns where its partner went (used by the layer memo: the partner's table row is the
// layer embedding of this entry's owner at the event time).
__global__ void inverse_perm_kernel(const uint32_t* __restrict__ perm, int64_t M, int32_t* __restrict__ inv) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < M) inv[perm[i]] = (int32_t)i;
}
__global__ void mirror_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ inv, int64_t M,
                              int32_t* __restrict__ mirror) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < M) mirror[i] = inv[perm[i] ^ 1u];
}

__global__ void max_kernel(const unsigned long long* __restrict__ counts, int64_t n, unsigned long long* out) {
    unsigned long long m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = counts[i] > m ? counts[i] : m;
    atomicMax(out, m);
}

// ---------------------------------------------------------------- query ordering
// Sort root queries by (node, time): consecutive queries of a bulk pass then ask for the same node at
// increasing times, their neighbour windows overlap in all but one slot and the rows they gather are
// cache hits.  LSD: stable sort on the time image, then stable sort on the node id.
__global__ void query_keys_kernel(const double* __restrict__ times, int64_t n, unsigned long long* __restrict__ tkey,
                                  uint32_t* __restrict__ iota) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    tkey[i] = sortable_f64(times[i] + 0.0);
    iota[i] = (uint32_t)i;
}
__global__ void gather_ids_kernel(const int32_t* __restrict__ ids, const uint32_t* __restrict__ perm, int64_t n,
                                  int32_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = ids[perm[i]];
}
__global__ void gather_queries_kernel(const int32_t* __restrict__ ids, const double* __restrict__ times,
                                      const uint32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ ids_out,
                                      double* __restrict__ times_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = perm[i];
    ids_out[i] = ids[p], times_out[i] = times[p];
}

int sort_queries(const int32_t* ids, const double* times, int64_t n, int64_t num_nodes, DevBuf& scratch,
                 int32_t** perm_out, int32_t** ids_sorted, double** times_sorted, cudaStream_t st) {
    const int T = 256;
    const unsigned B = (unsigned)ceil_div(n, T);
    size_t need_a = 0, need_b = 0;
    FLID_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need_a, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                              (uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 64, st));
    FLID_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need_b, (int32_t*)nullptr, (int32_t*)nullptr, (uint32_t*)nullptr,
                                              (uint32_t*)nullptr, n, 0, 32, st));
    const size_t tmp_bytes = std::max(need_a, need_b);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    // layout: tkey_a | tkey_b | perm_a | perm_b | id_a | id_b | times_sorted | cub temp
    const size_t o_ka = 0, o_kb = o_ka + al(8 * n), o_pa = o_kb + al(8 * n), o_pb = o_pa + al(4 * n), o_ia = o_pb + al(4 * n),
                 o_ib = o_ia + al(4 * n), o_ts = o_ib + al(4 * n), o_tmp = o_ts + al(8 * n);
    FLID_TRY(scratch.reserve(o_tmp + tmp_bytes));
    char* base = scratch.as<char>();
    unsigned long long *ka = (unsigned long long*)(base + o_ka), *kb = (unsigned long long*)(base + o_kb);
    uint32_t *pa = (uint32_t*)(base + o_pa), *pb = (uint32_t*)(base + o_pb);
    int32_t *ia = (int32_t*)(base + o_ia), *ib = (int32_t*)(base + o_ib);
    double* ts = (double*)(base + o_ts);
    void* tmp = base + o_tmp;
    size_t tb = tmp_bytes;
    query_keys_kernel<<<B, T, 0, st>>>(times, n, ka, pa);
    FLID_LAUNCH_CHECK();
    FLID_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, ka, kb, pa, pb, n, 0, 64, st));
    gather_ids_kernel<<<B, T, 0, st>>>(ids, pb, n, ia);
    FLID_LAUNCH_CHECK();
    int bits = 1;
    while (bits < 32 && (num_nodes >> bits) != 0) ++bits;
    tb = tmp_bytes;
    FLID_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, ia, ib, pb, pa, n, 0, bits, st));
    count_launch(8);
    // pa = final permutation, ib = sorted ids; sorted times into ts (ia is free again: reuse for nothing else)
    gather_queries_kernel<<<B, T, 0, st>>>(ids, times, pa, n, ia, ts);
    FLID_LAUNCH_CHECK();
    *perm_out = reinterpret_cast<int32_t*>(pa), *ids_sorted = ia, *times_sorted = ts;
    return FLID_OK;
}

static int build_sorted(flid_graph* g, int32_t* owner, int2* adj_in, unsigned long long* tkey, int64_t M,
                        bool paired, cudaStream_t st) {
    const int64_t N1 = g->num_nodes + 1;
    const int T = 256;
    const unsigned B = (unsigned)ceil_div(M > 0 ? M : 1, T);
    uint32_t *perm_a = nullptr, *perm_b = nullptr;
    unsigned long long *tkey_b = nullptr, *counts = nullptr;
    int32_t *own_a = nullptr, *own_b = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0, need = 0;
    FLID_CUDA(cudaMalloc(&g->indptr, sizeof(int64_t) * (N1 + 1)));
    FLID_CUDA(cudaMalloc(&counts, sizeof(unsigned long long) * (N1 + 1)));
    FLID_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (N1 + 1), st));
    FLID_CUDA(cudaMalloc(&g->adj, sizeof(int2) * (M > 0 ? M : 1)));
    FLID_CUDA(cudaMalloc(&g->ts, sizeof(double) * (M > 0 ? M : 1)));
    if (M > 0) {
        FLID_CUDA(cudaMalloc(&perm_a, sizeof(uint32_t) * M));
        FLID_CUDA(cudaMalloc(&perm_b, sizeof(uint32_t) * M));
        FLID_CUDA(cudaMalloc(&tkey_b, sizeof(unsigned long long) * M));
        FLID_CUDA(cudaMalloc(&own_a, sizeof(int32_t) * M));
        FLID_CUDA(cudaMalloc(&own_b, sizeof(int32_t) * M));
        iota_kernel<<<B, T, 0, st>>>(perm_a, M);
        FLID_LAUNCH_CHECK();
        // pass 1: stable sort on time
        FLID_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, tkey, tkey_b, perm_a, perm_b, M, 0, 64, st));
        tmp_bytes = need;
        FLID_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, own_a, own_b, perm_b, perm_a, M, 0, 32, st));
        tmp_bytes = need > tmp_bytes ? need : tmp_bytes;
        FLID_CUDA(cudaMalloc(&tmp, tmp_bytes));
        FLID_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, tkey, tkey_b, perm_a, perm_b, M, 0, 64, st));
        count_launch(8);
        gather_owner_kernel<<<B, T, 0, st>>>(owner, perm_b, M, own_a);
        FLID_LAUNCH_CHECK();
        // pass 2: stable sort on owner (only as many bits as node ids need)
        int bits = 1;
        while (bits < 32 && (g->num_nodes >> bits) != 0) ++bits;
        FLID_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, own_a, own_b, perm_b, perm_a, M, 0, bits, st));
        count_launch(4);
        place_kernel<<<B, T, 0, st>>>(perm_a, own_b, adj_in, tkey, M, g->adj, g->ts, counts);
        FLID_LAUNCH_CHECK();
        if (paired && M < 0x7fffffffLL) {
            FLID_CUDA(cudaMalloc(&g->mirror, sizeof(int32_t) * M));
            inverse_perm_kernel<<<B, T, 0, st>>>(perm_a, M, own_a);  // own_a is free again: reuse as the inverse
            FLID_LAUNCH_CHECK();
            mirror_kernel<<<B, T, 0, st>>>(perm_a, own_a, M, g->mirror);
            FLID_LAUNCH_CHECK();
        }
    }
    // indptr = exclusive scan of counts over N1 + 1 slots (last slot = M)
    size_t scan_bytes = 0;
    void* scan_tmp = nullptr;
    FLID_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts, (unsigned long long*)g->indptr, N1 + 1, st));
    FLID_CUDA(cudaMalloc(&scan_tmp, scan_bytes > 0 ? scan_bytes : 8));
    FLID_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, counts, (unsigned long long*)g->indptr, N1 + 1, st));
    count_launch(2);
    unsigned long long* dmax = counts + N1;  // slot N1 of counts is zero and unused by the scan result
    FLID_CUDA(cudaMemsetAsync(dmax, 0, sizeof(unsigned long long), st));
    max_kernel<<<(unsigned)std::min<int64_t>(ceil_div(N1, T), 1024), T, 0, st>>>(counts, N1, dmax);
    FLID_LAUNCH_CHECK();
    unsigned long long hmax = 0;
    FLID_CUDA(cudaMemcpyAsync(&hmax, dmax, sizeof(hmax), cudaMemcpyDeviceToHost, st));
    FLID_CUDA(cudaStreamSynchronize(st));
    g->max_degree = (int64_t)hmax;
    cudaFree(perm_a), cudaFree(perm_b), cudaFree(tkey_b), cudaFree(own_a), cudaFree(own_b), cudaFree(tmp);
    cudaFree(scan_tmp), cudaFree(counts);
    return FLID_OK;
}

static int build_common(const int64_t* a, const int64_t* b, const int64_t* eid, const double* ts, int64_t count,
                        int64_t num_nodes, int on_device, bool events, flid_graph** out, cudaStream_t st) {
    FLID_REQUIRE(out != nullptr, "flid_graph_build: out is null");
    FLID_REQUIRE(count >= 0 && num_nodes >= 0 && num_nodes < 0x7fffffffLL, "flid_graph_build: bad sizes");
    const int64_t M = events ? 2 * count : count;
    FLID_REQUIRE(M < 0xffffffffLL, "flid_graph_build: more than 2^32-1 adjacency entries");
    const int64_t *da = a, *db = b, *de = eid;
    const double* dt = ts;
    void* staged[4] = {nullptr, nullptr, nullptr, nullptr};
    if (!on_device && count > 0) {
        const void* hsrc[4] = {a, b, eid, ts};
        for (int i = 0; i < 4; ++i) {
            FLID_CUDA(cudaMalloc(&staged[i], 8 * count));
            FLID_CUDA(cudaMemcpyAsync(staged[i], hsrc[i], 8 * count, cudaMemcpyHostToDevice, st));
        }
        da = (const int64_t*)staged[0], db = (const int64_t*)staged[1], de = (const int64_t*)staged[2];
        dt = (const double*)staged[3];
    }
    flid_graph* g = new flid_graph();
    g->num_nodes = num_nodes;
    g->num_entries = M;
    int32_t* owner = nullptr;
    int2* adj_in = nullptr;
    unsigned long long* tkey = nullptr;
    int* bad = nullptr;
    FLID_CUDA(cudaMalloc(&bad, sizeof(int)));
    FLID_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    if (M > 0) {
        FLID_CUDA(cudaMalloc(&owner, sizeof(int32_t) * M));
        FLID_CUDA(cudaMalloc(&adj_in, sizeof(int2) * M));
        FLID_CUDA(cudaMalloc(&tkey, sizeof(unsigned long long) * M));
        const int T = 256;
        if (events)
            expand_events_kernel<<<(unsigned)ceil_div(count, T), T, 0, st>>>(da, db, de, dt, count, num_nodes, owner,
                                                                            adj_in, tkey, bad);
        else
            pack_entries_kernel<<<(unsigned)ceil_div(count, T), T, 0, st>>>(da, db, de, dt, count, num_nodes, owner,
                                                                           adj_in, tkey, bad);
        FLID_LAUNCH_CHECK();
    }
    int hbad = 0;
    FLID_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLID_CUDA(cudaStreamSynchronize(st));
    int status = FLID_OK;
    if (hbad) {
        set_error("flid_graph_build: node id outside [0, %lld] or edge id outside int32", (long long)num_nodes);
        status = FLID_ERR_RANGE;
    } else {
        status = build_sorted(g, owner, adj_in, tkey, M, events, st);
    }
    for (int i = 0; i < 4; ++i) cudaFree(staged[i]);
    cudaFree(owner), cudaFree(adj_in), cudaFree(tkey), cudaFree(bad);
    if (status != FLID_OK) {
        flid_graph_free(g);
        return status;
    }
    if (cudaMalloc(&g->bad_flag, sizeof(int)) != cudaSuccess || cudaMemsetAsync(g->bad_flag, 0, sizeof(int), st) != cudaSuccess) {
        flid_graph_free(g);
        set_error("flid_graph_build: cannot allocate the status flag");
        return FLID_ERR_CUDA;
    }
    *out = g;
    return FLID_OK;
}

}  // namespace flid

extern "C" {

int flid_graph_build_events(const int64_t* src, const int64_t* dst, const int64_t* eid, const double* ts,
                            int64_t num_events, int64_t num_nodes, int on_device, flid_graph** out,
                            flid_stream stream) {
    return flid::build_common(src, dst, eid, ts, num_events, num_nodes, on_device, true, out, (cudaStream_t)stream);
}

int flid_graph_build_entries(const int64_t* owner, const int64_t* nbr, const int64_t* eid, const double* ts,
                             int64_t num_entries, int64_t num_nodes, int on_device, flid_graph** out,
                             flid_stream stream) {
    return flid::build_common(owner, nbr, eid, ts, num_entries, num_nodes, on_device, false, out,
                              (cudaStream_t)stream);
}

void flid_graph_free(flid_graph* g) {
    if (!g) return;
    cudaFree(g->indptr), cudaFree(g->adj), cudaFree(g->ts), cudaFree(g->mirror), cudaFree(g->bad_flag), cudaFree(g->owner), cudaFree(g->ent_eid);
    delete g;
}

int flid_graph_info(const flid_graph* g, int64_t* num_nodes, int64_t* num_entries, int64_t* max_degree) {
    FLID_REQUIRE(g != nullptr, "flid_graph_info: null graph");
    if (num_nodes) *num_nodes = g->num_nodes;
    if (num_entries) *num_entries = g->num_entries;
    if (max_degree) *max_degree = g->max_degree;
    return FLID_OK;
}

int flid_graph_export_host(const flid_graph* g, int64_t* indptr_host, int64_t* nbr_host, int64_t* eid_host,
                           double* ts_host) {
    FLID_REQUIRE(g != nullptr, "flid_graph_export_host: null graph");
    const int64_t M = g->num_entries;
    FLID_CUDA(cudaDeviceSynchronize());
    FLID_CUDA(cudaMemcpy(indptr_host, g->indptr, sizeof(int64_t) * (g->num_nodes + 2), cudaMemcpyDeviceToHost));
    if (M > 0) {
        FLID_CUDA(cudaMemcpy(ts_host, g->ts, sizeof(double) * M, cudaMemcpyDeviceToHost));
        int2* tmp = (int2*)malloc(sizeof(int2) * M);
        FLID_REQUIRE(tmp != nullptr, "flid_graph_export_host: out of host memory");
        cudaError_t e = cudaMemcpy(tmp, g->adj, sizeof(int2) * M, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            for (int64_t i = 0; i < M; ++i) nbr_host[i] = tmp[i].x, eid_host[i] = tmp[i].y;
        free(tmp);
        FLID_CUDA(e);
    }
    return FLID_OK;
}

}  // extern "C"
