// TGN memory path (models/MemoryModel.py:96-189), one batch at a time, in the reference's
// event order.  Observation that makes it a short kernel chain: the aggregator keeps only
// the LAST raw message per node (:312-330), and a node's (memory, pending message) pair
// only changes in a batch that contains the node.  So
//     next_memories[v] = GRU(pending[v], memories[v])   (memories[v] if nothing pending)
// is exactly what get_updated_memories (:190-212) recomputes for all nodes in every batch
// and what update_memories (:214-231) later persists; it is maintained incrementally for
// the batch nodes only.  layer0 = next_memories + node_raw is the layer-0 / merge-input
// table of GraphAttentionEmbedding (:654-658, :712-713), so the embedding itself is the
// TGAT kernel chain on a mutable table.
#include <math.h>

#include <algorithm>

#include "gemm.cuh"
#include "tgat.cuh"

namespace flid {


// batch_ctr (nullable): device counter of the whole-pass driver (flid_tgn_pass) -- the batch starts at event
// *batch_ctr * B of the arrays, so that one captured CUDA graph serves every batch of the pass
__global__ void tgn_prep_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                const double* __restrict__ times, const int64_t* __restrict__ eids, int64_t B,
                                int64_t id_limit, int32_t* __restrict__ ids, double* __restrict__ t2,
                                int32_t* __restrict__ e32, int32_t* __restrict__ err,
                                const long long* __restrict__ batch_ctr = nullptr) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (batch_ctr) {
        const int64_t lo = (int64_t)(*batch_ctr) * B;
        src += lo, dst += lo, times += lo;
        if (eids) eids += lo;
    }
    // id_limit = min(bank rows, sampler nodes + 1): an id the sampler has no list for is the reference's
    // IndexError in nodes_neighbor_times[node_id] (utils/utils.py:141), one beyond the bank its index error there
    int64_t s = src[i], d = dst[i], e = eids ? eids[i] : 0;
    if (s < 0 || s >= id_limit) s = 0, atomicMax(err, 2);
    if (d < 0 || d >= id_limit) d = 0, atomicMax(err, 2);
    if (e < 0 || e > 0x7fffffffLL) e = 0, atomicMax(err, 2);
    ids[i] = (int32_t)s, ids[B + i] = (int32_t)d;
    t2[i] = times[i], t2[B + i] = times[i];
    e32[i] = (int32_t)e;
}

// update_memories for the batch nodes (models/MemoryModel.py:472-499): persist the pending
// GRU result, last_updated = float32(pending time); one warp per occurrence, duplicate
// occurrences write identical values.  Also elects the message winner of each node:
// occurrences are numbered src role 0..B-1, dst role B..2B-1, so atomicMax keeps the last
// dst-role occurrence if any, else the last src-role one (store order at :177-180).
__global__ void __launch_bounds__(256) tgn_persist_kernel(flid_tgn_state s, const int32_t* __restrict__ ids,
                                                          int64_t n2, int dn, int32_t* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t o = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (o >= n2) return;
    const int32_t v = ids[o];
    if (lane == 0) atomicMax(s.scratch + v, (int32_t)o);
    if (!s.has_pending[v]) return;
    const float tf = (float)s.pending_ts[v];
    if (lane == 0) {
        if (!(s.last_updated[v] <= tf)) atomicMax(err, 1);  // "Trying to update memory to time in the past!"
    }
    __syncwarp();
    for (int c = lane; c < dn; c += 32) s.memories[(int64_t)v * dn + c] = s.next_memories[(int64_t)v * dn + c];
    if (lane == 0) s.last_updated[v] = tf;
}

// compute_new_node_raw_messages (models/MemoryModel.py:233-278) for the winning occurrence
// of each node: cat[mem[a], mem[b], te(float32(t) - last_updated[a]), edge[eid]].
__global__ void __launch_bounds__(256) tgn_message_kernel(flid_tgn_state s, const int32_t* __restrict__ ids,
                                                          const double* __restrict__ t2,
                                                          const int32_t* __restrict__ e32, int64_t B, int dn, int de,
                                                          int T, const float* __restrict__ edge_feat,
                                                          const float* __restrict__ time_w,
                                                          const float* __restrict__ time_b,
                                                          int32_t* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t o = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (o >= 2 * B) return;
    const int32_t a = ids[o];
    if (s.scratch[a] != (int32_t)o) return;  // not this node's last message
    const int64_t ev = o < B ? o : o - B;
    const int32_t b = o < B ? ids[B + ev] : ids[ev];
    const double t = t2[o];
    const float dtf = (float)t - s.last_updated[a];
    // the reference asserts last_updated <= float32(message time) when this message is next aggregated
    if (lane == 0 && !(s.last_updated[a] <= (float)t)) atomicMax(err, 1);
    const int msg = 2 * dn + T + de;
    float* m = s.pending_msg + (int64_t)a * msg;
    for (int c = lane; c < dn; c += 32) {
        m[c] = s.memories[(int64_t)a * dn + c];
        m[dn + c] = s.memories[(int64_t)b * dn + c];
    }
    for (int c = lane; c < T; c += 32) m[2 * dn + c] = time_channel(dtf, __ldg(time_w + c), __ldg(time_b + c));
    const float* er = edge_feat + (int64_t)e32[ev] * de;
    for (int c = lane; c < de; c += 32) m[2 * dn + T + c] = __ldg(er + c);
    if (lane == 0) s.pending_ts[a] = t, s.has_pending[a] = 1;
}


__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// nn.GRUCell gate math (gates r, z, n; h' = (h - n) * z + n) on precomputed
// gi = W_ih x + b_ih, gh = W_hh h + b_hh; writes next_memories and layer0 rows.
// ids == nullptr: row r of gi/gh belongs to node r (full rebuild).
__global__ void __launch_bounds__(256) tgn_gate_kernel(flid_tgn_state s, const int32_t* __restrict__ ids, int64_t rows,
                                                       int dn, const float* __restrict__ gi,
                                                       const float* __restrict__ gh,
                                                       const float* __restrict__ node_raw, int reset_scratch) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= rows * dn) return;
    const int64_t r = idx / dn;
    const int c = (int)(idx % dn);
    const int64_t v = ids ? ids[r] : r;
    const float h = s.memories[v * dn + c];
    float hn = h;
    if (s.has_pending[v]) {
        const float* a = gi + r * 3 * dn;
        const float* b = gh + r * 3 * dn;
        const float rg = sigmoidf_acc(b[c] + a[c]);
        const float zg = sigmoidf_acc(b[dn + c] + a[dn + c]);
        const float ng = tanhf(a[2 * dn + c] + b[2 * dn + c] * rg);
        hn = (h - ng) * zg + ng;
    }
    s.next_memories[v * dn + c] = hn;
    s.layer0[v * dn + c] = hn + node_raw[v * dn + c];
    if (reset_scratch && c == 0) s.scratch[v] = -1;
}

__global__ void tgn_reset_kernel(flid_tgn_state s, const float* __restrict__ node_raw, int dn, int msg) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t rows = s.num_rows;
    if (idx < rows * dn) {
        s.memories[idx] = 0.f, s.next_memories[idx] = 0.f, s.layer0[idx] = node_raw[idx];
    }
    if (idx < rows) {
        s.last_updated[idx] = 0.f, s.pending_ts[idx] = 0.0, s.has_pending[idx] = 0, s.scratch[idx] = -1;
    }
    for (int64_t j = idx; j < rows * msg; j += (int64_t)gridDim.x * blockDim.x) s.pending_msg[j] = 0.f;
}

static int gru_rows(flid_tgat* m, const flid_tgn_state* s, const flid_gru_weights* gru, const int32_t* ids,
                    int64_t rows, const float* node_raw, int reset_scratch, cudaStream_t st) {
    const int dn = m->dn, msg = 2 * m->dn + m->T + m->de;
    FLID_TRY(m->tgn_gi.reserve(sizeof(float) * rows * 3 * dn));
    FLID_TRY(m->tgn_gh.reserve(sizeof(float) * rows * 3 * dn));
    float *gi = m->tgn_gi.as<float>(), *gh = m->tgn_gh.as<float>();
    if (m->use_tc && m->gru_ih_src == gru->weight_ih && m->gru_hh_src == gru->weight_hh && m->tc_gih.buf && m->tc_ghh.buf &&
        msg % 4 == 0 && dn % 4 == 0) {
        // W_ih x + b_ih and W_hh h + b_hh on the tcgen05 GEMM (3xTF32, fp32-grade): two ~19 us launches per batch of
        // 400 rows instead of two ~32 us SIMT ones (ncu launch list of the pass, profiles/r2_tgn_pass_launches.txt)
        TcGemmArgs a;
        a.A0 = s->pending_msg, a.lda0 = msg, a.idx0 = ids, a.w0 = msg, a.C = gi, a.ldc = 3 * dn, a.bias = gru->bias_ih, a.M = rows;
        FLID_TRY(tc_gemm(a, m->tc_gih, st));
        TcGemmArgs b;
        b.A0 = s->memories, b.lda0 = dn, b.idx0 = ids, b.w0 = dn, b.C = gh, b.ldc = 3 * dn, b.bias = gru->bias_hh, b.M = rows;
        FLID_TRY(tc_gemm(b, m->tc_ghh, st));
    } else {
        GemmArgs a{s->pending_msg, msg, ids, gru->weight_ih, msg, gi, 3 * dn, gru->bias_ih, rows, 3 * dn, msg, 0, 0};
        FLID_TRY(launch_gemm(a, st));
        GemmArgs b{s->memories, dn, ids, gru->weight_hh, dn, gh, 3 * dn, gru->bias_hh, rows, 3 * dn, dn, 0, 0};
        FLID_TRY(launch_gemm(b, st));
    }
    tgn_gate_kernel<<<(unsigned)ceil_div(rows * dn, 256), 256, 0, st>>>(*s, ids, rows, dn, gi, gh, node_raw,
                                                                       reset_scratch);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

}  // namespace flid

extern "C" {

int flid_tgn_reset(const flid_tgn_state* s, const float* node_raw, int node_dim, int msg_dim, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(s && node_raw && s->num_rows > 0, "flid_tgn_reset: bad argument");
    const int64_t work = std::max<int64_t>(s->num_rows * node_dim, s->num_rows);
    tgn_reset_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, (cudaStream_t)stream>>>(*s, node_raw, node_dim, msg_dim);
    FLID_LAUNCH_CHECK();
    return FLID_OK;
}

int flid_tgn_rebuild(flid_tgat* m, const flid_tgn_state* s, const flid_gru_weights* gru, const float* node_raw,
                     flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && s && gru && node_raw, "flid_tgn_rebuild: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (m->use_tc) {   // the caller rebuilds whenever the GRU parameters changed: re-tile them here
        const int dn = m->dn, msg = 2 * m->dn + m->T + m->de;
        FLID_TRY(tc_prepare_weight(gru->weight_ih, msg, 3 * dn, msg, &m->tc_gih, st, 0));
        FLID_TRY(tc_prepare_weight(gru->weight_hh, dn, 3 * dn, dn, &m->tc_ghh, st, 0));
        m->gru_ih_src = gru->weight_ih, m->gru_hh_src = gru->weight_hh;
    }
    FLID_TRY(gru_rows(m, s, gru, nullptr, s->num_rows, node_raw, 0, st));
    if (m->have_weights) FLID_TRY(flid_tgat_cache_node_table(m, s->layer0, s->num_rows, stream));
    return FLID_OK;
}

// copy the [2B, dn] embeddings of the batch at *batch_ctr to rows [lo, lo + B) of the per-event outputs
__global__ void tgn_store_kernel(const float* __restrict__ emb, int64_t B, int dn4, const long long* __restrict__ batch_ctr,
                                 float* __restrict__ out_src, float* __restrict__ out_dst) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= 2 * B * dn4) return;
    const int64_t row = idx / dn4, c = idx % dn4, lo = (int64_t)(*batch_ctr) * B;
    const float4 v = reinterpret_cast<const float4*>(emb)[idx];
    float4* dst = reinterpret_cast<float4*>(row < B ? out_src : out_dst) + (lo + (row < B ? row : row - B)) * dn4 + c;
    *dst = v;
}
__global__ void tgn_advance_kernel(long long* batch_ctr) { *batch_ctr += 1; }

static int tgn_step_body(flid_tgat* m, const flid_graph* g, const flid_tgn_state* s, const flid_gru_weights* gru,
                         const float* node_raw, const float* edge_feat, const int64_t* src, const int64_t* dst,
                         const double* times, const int64_t* eids, int64_t batch, int positive, int k, float* out,
                         int32_t* err_flag, cudaStream_t st, const long long* batch_ctr);

int flid_tgn_step(flid_tgat* m, const flid_graph* g, const flid_tgn_state* s, const flid_gru_weights* gru,
                  const float* node_raw, const float* edge_feat, const int64_t* src, const int64_t* dst,
                  const double* times, const int64_t* eids, int64_t batch, int positive, int k, float* out,
                  int32_t* err_flag, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && g && s && gru && node_raw && edge_feat && src && dst && times && err_flag,
                 "flid_tgn_step: null argument");
    FLID_REQUIRE(out != nullptr || positive, "flid_tgn_step: nothing to do (no output buffer and no state update)");
    FLID_REQUIRE(m->have_weights, "flid_tgn_step: weights not set");
    FLID_REQUIRE(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    FLID_REQUIRE(!positive || eids, "flid_tgn_step: edge_ids are required for positive edges");
    // A sampler built from a prefix of the stream (train_neighbor_sampler, PTCL/EM_warmup.py:71) may know fewer
    // nodes than the bank has rows (one per node of the full graph); the other way round the bank could not
    // hold the memories of sampled neighbours.
    FLID_REQUIRE(s->num_rows >= g->num_nodes + 1, "flid_tgn_step: memory bank has %lld rows but the sampler knows %lld nodes",
                 (long long)s->num_rows, (long long)g->num_nodes + 1);
    if (batch <= 0) return FLID_OK;
    return tgn_step_body(m, g, s, gru, node_raw, edge_feat, src, dst, times, eids, batch, positive, k, out, err_flag,
                         (cudaStream_t)stream, nullptr);
}

static int tgn_step_body(flid_tgat* m, const flid_graph* g, const flid_tgn_state* s, const flid_gru_weights* gru,
                         const float* node_raw, const float* edge_feat, const int64_t* src, const int64_t* dst,
                         const double* times, const int64_t* eids, int64_t batch, int positive, int k, float* out,
                         int32_t* err_flag, cudaStream_t st, const long long* batch_ctr) {
    using namespace flid;
    const int64_t B = batch, n2 = 2 * batch;
    FLID_TRY(m->tgn_ids.reserve(sizeof(int32_t) * n2));
    FLID_TRY(m->tgn_times.reserve(sizeof(double) * n2));
    FLID_TRY(m->tgn_eids.reserve(sizeof(int32_t) * B));
    int32_t* ids = m->tgn_ids.as<int32_t>();
    double* t2 = m->tgn_times.as<double>();
    int32_t* e32 = m->tgn_eids.as<int32_t>();
    tgn_prep_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(src, dst, times, eids, B,
                                                               std::min<int64_t>(s->num_rows, g->num_nodes + 1), ids, t2,
                                                               e32, err_flag, batch_ctr);
    FLID_LAUNCH_CHECK();
    // (1)+(2): embeddings on memory' + raw   (models/MemoryModel.py:117-146)
    // out == nullptr: state update only (the training-mode host path computes the embeddings itself)
    if (out != nullptr) FLID_TRY(tgat_embed_ids(m, g, s->layer0, edge_feat, ids, t2, n2, n2, k, out, st));
    if (!positive) return FLID_OK;
    // (3): persist, elect, build messages, refresh the incremental GRU state (:155-180)
    tgn_persist_kernel<<<(unsigned)ceil_div(n2 * 32, 256), 256, 0, st>>>(*s, ids, n2, m->dn, err_flag);
    FLID_LAUNCH_CHECK();
    tgn_message_kernel<<<(unsigned)ceil_div(n2 * 32, 256), 256, 0, st>>>(*s, ids, t2, e32, B, m->dn, m->de, m->T,
                                                                        edge_feat, m->time_w, m->time_b, err_flag);
    FLID_LAUNCH_CHECK();
    FLID_TRY(gru_rows(m, s, gru, ids, n2, node_raw, 1, st));
    if (m->table_src == s->layer0 && m->table_rows > 0)
        FLID_TRY(flid_tgat_refresh_node_rows(m, s->layer0, ids, n2, (flid_stream)st));
    return FLID_OK;
}

int flid_tgn_pass(flid_tgat* m, const flid_graph* g, const flid_tgn_state* s, const flid_gru_weights* gru,
                  const float* node_raw, const float* edge_feat, const int64_t* src, const int64_t* dst,
                  const double* times, const int64_t* eids, int64_t num_events, int64_t batch, int k, float* out_src,
                  float* out_dst, int32_t* err_flag, int use_graph, flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(m && g && s && gru && node_raw && edge_feat && src && dst && times && eids && out_src && out_dst && err_flag,
                 "flid_tgn_pass: null argument");
    FLID_REQUIRE(m->have_weights, "flid_tgn_pass: weights not set");
    FLID_REQUIRE(k > 0, "Number of sampled neighbors for each node should be greater than 0!");
    FLID_REQUIRE(batch > 0, "flid_tgn_pass: batch size must be positive");
    FLID_REQUIRE(s->num_rows >= g->num_nodes + 1, "flid_tgn_pass: memory bank has %lld rows but the sampler knows %lld nodes",
                 (long long)s->num_rows, (long long)g->num_nodes + 1);
    FLID_REQUIRE(m->dn % 4 == 0, "flid_tgn_pass: node dim must be a multiple of 4");
    if (num_events <= 0) return FLID_OK;
    cudaStream_t caller = (cudaStream_t)stream;
    // the legacy default stream cannot be captured: run the pass on a stream of our own, ordered after / before the
    // caller's stream with events
    if (!m->tgn_stream) FLID_CUDA(cudaStreamCreateWithFlags(&m->tgn_stream, cudaStreamNonBlocking));
    if (!m->tgn_ev) FLID_CUDA(cudaEventCreateWithFlags(&m->tgn_ev, cudaEventDisableTiming));
    cudaStream_t st = m->tgn_stream;
    FLID_CUDA(cudaEventRecord(m->tgn_ev, caller));
    FLID_CUDA(cudaStreamWaitEvent(st, m->tgn_ev, 0));
    const int64_t B = batch, full = num_events / B, rest = num_events - full * B;
    FLID_TRY(m->tgn_out.reserve(sizeof(float) * 2 * B * m->dn));
    FLID_TRY(m->tgn_ctr.reserve(sizeof(long long)));
    float* emb = m->tgn_out.as<float>();
    long long* ctr = m->tgn_ctr.as<long long>();
    FLID_CUDA(cudaMemsetAsync(ctr, 0, sizeof(long long), st));
    const int dn4 = m->dn / 4;
    auto one_batch = [&]() -> int {   // the batch at *ctr, then ctr += 1
        FLID_TRY(tgn_step_body(m, g, s, gru, node_raw, edge_feat, src, dst, times, eids, B, 1, k, emb, err_flag, st, ctr));
        tgn_store_kernel<<<(unsigned)ceil_div(2 * B * dn4, 256), 256, 0, st>>>(emb, B, dn4, ctr, out_src, out_dst);
        FLID_LAUNCH_CHECK();
        tgn_advance_kernel<<<1, 1, 0, st>>>(ctr);
        FLID_LAUNCH_CHECK();
        return FLID_OK;
    };
    int status = FLID_OK;
    int64_t done = 0;
    if (full > 0) {
        status = one_batch();   // direct launches: sizes every workspace, sets the kernel attributes
        done = 1;
    }
    if (status == FLID_OK && full > done) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        if (use_graph && !m->prof_on && full - done >= 4) {
            // every full batch runs the same launch sequence on fixed buffers; only the device counter differs
            cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                const int64_t launches0 = g_launches;
                status = one_batch();
                const int64_t per_batch = g_launches - launches0;
                e = cudaStreamEndCapture(st, &graph);
                if (status == FLID_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
                if (status == FLID_OK && e == cudaSuccess && exec) {
                    g_launches -= per_batch;   // the capture itself launched nothing
                    for (; done < full && e == cudaSuccess; ++done) {
                        e = cudaGraphLaunch(exec, st);
                        g_launches += per_batch;
                    }
                }
                if (exec) cudaGraphExecDestroy(exec);
                if (graph) cudaGraphDestroy(graph);
                if (status == FLID_OK && e != cudaSuccess) {
                    set_error("flid_tgn_pass: CUDA graph path failed: %s", cudaGetErrorString(e));
                    status = FLID_ERR_CUDA;
                }
            } else {
                cudaGetLastError();   // capture not available: fall through to direct launches
            }
        }
        for (; status == FLID_OK && done < full; ++done) status = one_batch();
    }
    if (status == FLID_OK && rest > 0) {   // the last, shorter batch
        const int64_t lo = full * B;
        status = tgn_step_body(m, g, s, gru, node_raw, edge_feat, src + lo, dst + lo, times + lo, eids + lo, rest, 1, k, emb,
                               err_flag, st, nullptr);
        if (status == FLID_OK) {
            cudaError_t e1 = cudaMemcpyAsync(out_src + lo * m->dn, emb, sizeof(float) * rest * m->dn, cudaMemcpyDeviceToDevice, st);
            cudaError_t e2 = cudaMemcpyAsync(out_dst + lo * m->dn, emb + rest * m->dn, sizeof(float) * rest * m->dn,
                                             cudaMemcpyDeviceToDevice, st);
            if (e1 != cudaSuccess || e2 != cudaSuccess) {
                set_error("flid_tgn_pass: output copy failed");
                status = FLID_ERR_CUDA;
            }
        }
    }
    // hand the results back to the caller's stream
    cudaEventRecord(m->tgn_ev, st);
    cudaStreamWaitEvent(caller, m->tgn_ev, 0);
    return status;
}

}  // extern "C"
