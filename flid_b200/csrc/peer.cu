// Multi-GPU plumbing of the owner-partitioned pass: peer-mapped memo tables and the row exchange as one kernel that
// stores straight into the other ranks' HBM over NVLink (no staging buffer, no library collective).
//
// One process per GPU: a rank allocates its memo table here, publishes the CUDA IPC handle, and maps the tables of
// the other ranks (cudaIpcOpenMemHandle with lazy peer access).  After an owner-range build
// (flid_tgat_memo_build_owner_range) rank r holds the rows mirror[q], q in its own position range; those whose
// position belongs to rank s are written to the same row of rank s's table.  Every row has exactly one producer, so
// there are no write conflicts; the caller separates the exchange from its consumers with a cross-rank barrier.
#include "common.cuh"
#include <stdlib.h>

#include "graph.cuh"

namespace flid {

constexpr int MAX_RANKS = 16;
struct PeerArgs {
    float* table[MAX_RANKS];
    int64_t bound[MAX_RANKS + 1];
    int world, rank;
};

// A persistent grid of small CTAs (one 128-thread CTA per SM): each warp walks over groups of XR consecutive work items
// of this rank's range, 16-byte copies.  All loads of a group are issued before its first remote store (2 * XR float4
// in flight per lane): the stores are posted writes over NVLink, the local gather is what needs the memory-level
// parallelism.  The grid is kept this small on purpose -- the kernel runs on a side stream beside the next call's
// query-side GEMM (448 threads, ~48 K registers, most of the shared memory per SM) and must fit next to it.
constexpr int XR = 8;
__global__ void __launch_bounds__(128) memo_exchange_kernel(const int32_t* __restrict__ mirror, const float* __restrict__ local,
                                                            PeerArgs a, int64_t item_lo, int64_t n, int row4) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w0 = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * XR; w0 < n; w0 += warps * XR) {
        int64_t p[XR];
        int dest[XR];
#pragma unroll
        for (int i = 0; i < XR; ++i) {
            dest[i] = a.rank;
            p[i] = 0;
            if (w0 + i < n) {
                p[i] = __ldg(mirror + item_lo + w0 + i);
                int d = 0;
#pragma unroll 1
                for (int r = 1; r < a.world; ++r) d += (p[i] >= a.bound[r]);
                dest[i] = d;
            }
        }
        if (row4 <= 64) {
            float4 v[XR][2];
#pragma unroll
            for (int i = 0; i < XR; ++i) {
                if (dest[i] == a.rank) continue;
                const float4* src = reinterpret_cast<const float4*>(local) + p[i] * row4;
                if (lane < row4) v[i][0] = __ldg(src + lane);
                if (lane + 32 < row4) v[i][1] = __ldg(src + lane + 32);
            }
#pragma unroll
            for (int i = 0; i < XR; ++i) {
                if (dest[i] == a.rank) continue;
                float4* dst = reinterpret_cast<float4*>(a.table[dest[i]]) + p[i] * row4;
                if (lane < row4) dst[lane] = v[i][0];
                if (lane + 32 < row4) dst[lane + 32] = v[i][1];
            }
            continue;
        }
        for (int i = 0; i < XR; ++i) {
            if (dest[i] == a.rank) continue;
            const float4* src = reinterpret_cast<const float4*>(local) + p[i] * row4;
            float4* dst = reinterpret_cast<float4*>(a.table[dest[i]]) + p[i] * row4;
            for (int c = lane; c < row4; c += 32) dst[c] = src[c];
        }
    }
}

}  // namespace flid

extern "C" {

int flid_peer_alloc(int64_t bytes, void** dev_ptr, void* handle_out) {
    using namespace flid;
    FLID_REQUIRE(bytes > 0 && dev_ptr && handle_out, "flid_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    FLID_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return FLID_ERR_CUDA;
    }
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return FLID_OK;
}

int flid_peer_open(const void* handle, void** dev_ptr) {
    using namespace flid;
    FLID_REQUIRE(handle && dev_ptr, "flid_peer_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    FLID_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FLID_OK;
}

int flid_peer_close(void* dev_ptr) {
    using namespace flid;
    if (dev_ptr) FLID_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return FLID_OK;
}

int flid_peer_free(void* dev_ptr) {
    using namespace flid;
    if (dev_ptr) FLID_CUDA(cudaFree(dev_ptr));
    return FLID_OK;
}

int flid_memo_exchange_p2p(const flid_graph* g, const float* table_local, void* const* peer_tables_host,
                           const int64_t* pos_bounds_host, int world, int rank, int row_dim, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(g && table_local && peer_tables_host && pos_bounds_host, "flid_memo_exchange_p2p: null argument");
    FLID_REQUIRE(world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world, "flid_memo_exchange_p2p: bad rank / world");
    FLID_REQUIRE(g->mirror != nullptr, "flid_memo_exchange_p2p: the graph has no partner index (build it from events)");
    FLID_REQUIRE(row_dim > 0 && row_dim % 4 == 0, "flid_memo_exchange_p2p: row width must be a multiple of 4 floats");
    PeerArgs a;
    a.world = world, a.rank = rank;
    for (int r = 0; r < world; ++r) {
        a.table[r] = reinterpret_cast<float*>(peer_tables_host[r]);
        FLID_REQUIRE(r == rank || a.table[r] != nullptr, "flid_memo_exchange_p2p: table of rank %d is not mapped", r);
    }
    for (int r = 0; r <= world; ++r) a.bound[r] = pos_bounds_host[r];
    const int64_t lo = a.bound[rank], n = a.bound[rank + 1] - lo;
    if (n <= 0 || world == 1) return FLID_OK;
    static int sms = 0, per_sm = 1;
    if (sms == 0) {
        int dev = 0;
        FLID_CUDA(cudaGetDevice(&dev));
        FLID_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (const char* e = getenv("FLID_XCHG_CTAS")) per_sm = atoi(e) > 0 ? atoi(e) : 1;   // development knob
    }
    const int64_t want = ceil_div(ceil_div(n, XR) * 32, 128);
    const unsigned grid = (unsigned)(want < (int64_t)sms * per_sm ? want : (int64_t)sms * per_sm);
    memo_exchange_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(g->mirror, table_local, a, lo, n, row_dim / 4);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

}  // extern "C"
