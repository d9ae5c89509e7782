// Device-resident temporal adjacency: CSR over nodes, time-sorted inside each node.
#pragma once
#include "common.cuh"

struct flid_graph {
    int64_t num_nodes = 0;    // largest valid node id
    int64_t num_entries = 0;  // 2 * events
    int64_t max_degree = 0;
    int64_t* indptr = nullptr;  // [num_nodes + 2]
    int2* adj = nullptr;        // [M] (.x = neighbour id, .y = edge id)
    double* ts = nullptr;       // [M]
    int32_t* mirror = nullptr;  // [M] position of the same event's entry in the other endpoint's list
                                // (graphs built from events with M < 2^31 only, else null)
    int* bad_flag = nullptr;    // device int, zero between calls: out-of-range query ids of the sampler entry points
    // derived arrays of the projected bulk path (bulk_kv.cu), built on first use
    int32_t* owner = nullptr;   // [M] node whose list holds the entry
    int32_t* ent_eid = nullptr; // [M + 1] edge id of the entry (row M: the padded slot, edge 0)
    int zero_nbr = 0;           // some entry's neighbour is node 0 (masked like padding by the reference)
    int64_t max_eid = -1;       // largest edge id in the adjacency (rows of the per-edge tables)
};

namespace flid {

// (node, time) ordering of n root queries: perm[i] = original index of the i-th query in sorted order, plus the
// sorted ids / times; all three live in `scratch` (valid until its next use).  n < 2^31.
int sort_queries(const int32_t* ids, const double* times, int64_t n, int64_t num_nodes, DevBuf& scratch,
                 int32_t** perm_out, int32_t** ids_sorted, double** times_sorted, cudaStream_t st);

#ifdef __CUDACC__
// Warp-cooperative searchsorted(ts[lo:hi), t, side='left'): first position whose
// timestamp is >= t.  32 probes per round, so ceil(log32(deg)) + 1 dependent loads.
__device__ __forceinline__ int64_t warp_lower_bound(const double* __restrict__ ts, int64_t lo, int64_t hi, double t,
                                                    int lane) {
    while (lo < hi) {
        const int64_t len = hi - lo;
        const int64_t stride = (len + 31) >> 5;
        const int64_t p = lo + (int64_t)lane * stride;
        const bool before = (p < hi) && (__ldg(ts + p) < t);
        const int c = __popc(__ballot_sync(FULL, before));
        if (c == 0) {
            hi = lo;
        } else {
            const int64_t first_false = lo + (int64_t)c * stride;
            lo = lo + (int64_t)(c - 1) * stride + 1;
            hi = first_false < hi ? first_false : hi;
        }
    }
    return lo;
}
#endif

}  // namespace flid
