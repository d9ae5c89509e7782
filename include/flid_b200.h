/*
 * flid_b200 -- C ABI of the B200-native (sm_100a) temporal-embedding hot path of FLiD.
 *
 * The FLiD reference is pure Python and has no FFI layer; its "plugin boundary" for
 * this path is a handful of Python methods.  Each entry point below names the
 * reference interface (file:line under the reference tree) it replaces.  The Python
 * classes in flid_b200/ bind these with ctypes and keep the
 * reference's names, argument meaning and error behaviour.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; flid_last_error()
 *     returns a thread-local, human-readable reason.  No exceptions cross the ABI.
 *   - pointers are DEVICE pointers unless the parameter name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, not synchronised,
 *     unless stated.  Handles are thread-compatible, not thread-safe.
 *   - one CUDA device per process (the torchrun layout: one rank per GPU): handles and the
 *     library's internal scratch buffers belong to the device that was current when they
 *     were first used.
 *   - ids are int64 at the boundary (numpy's default, as the reference passes them)
 *     and int32 internally; node id 0 / edge id 0 are the padding rows.
 */
#ifndef FLID_B200_H
#define FLID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLID_OK 0
#define FLID_ERR_INVALID 1   /* bad argument / unsupported shape            */
#define FLID_ERR_CUDA 2      /* CUDA runtime error (see flid_last_error)    */
#define FLID_ERR_RANGE 3     /* node / edge id outside the graph            */
#define FLID_ERR_STATE 4     /* e.g. TGN "update memory to time in the past" */

typedef struct flid_graph flid_graph; /* device CSR, time-sorted inside each node      */
typedef struct flid_tgat flid_tgat;   /* folded TGAT weights + grow-only workspace      */
typedef void* flid_stream;            /* cudaStream_t                                   */

const char* flid_last_error(void);
int flid_abi_version(void);
/* number of kernel launches issued through this library since load (for bench.py's gpu_launches) */
int64_t flid_launch_count(void);

/* Self-test hook for the two projection-GEMM back ends (used by tests/ only):
 * C[M,N] = [A0 | A1][M, w0+w1] . W[N, w0+w1]^T + bias, optional ReLU, optional int32 row gather
 * of A0.  backend 0 = fp32 SIMT, 1 = tcgen05 3xTF32.  All pointers device.              */
int flid_debug_gemm(int backend, const float* a0, int64_t lda0, const int32_t* idx0, int w0, const float* a1,
                    int64_t lda1, int w1, const float* w, int64_t ldw, const float* bias, float* c, int64_t ldc,
                    int64_t m, int n, int relu, flid_stream stream);

/* same shapes, tcgen05 back end only: mean device milliseconds per launch over `reps` launches
 * (CUDA events on `stream`; tools/gemm_probe.py)                                          */
int flid_debug_gemm_time(const float* a0, int64_t lda0, const int32_t* idx0, int w0, const float* a1, int64_t lda1,
                         int w1, const float* w, int64_t ldw, const float* bias, float* c, int64_t ldc, int64_t m,
                         int n, int relu, int reps, float* ms_per_launch, flid_stream stream);

/* ------------------------------------------------------------------ graph ---------
 * Replaces get_neighbor_sampler (utils/utils.py:283-302) + NeighborSampler.__init__
 * (utils/utils.py:73-110): undirected adjacency, per node stably sorted by timestamp.
 *
 * flid_graph_build_events: src/dst/eid int64[E], ts float64[E] (event order; every
 *   event is appended to src's list then to dst's list, as the reference does).
 * flid_graph_build_entries: the adj_list form -- owner/nbr/eid/ts [M] in insertion order.
 * `on_device` != 0 means the four input arrays are device pointers, else host.
 * num_nodes = largest valid node id (tables have num_nodes+1 rows incl. padding id 0).
 * Both calls synchronise the stream before returning.                                */
int flid_graph_build_events(const int64_t* src, const int64_t* dst, const int64_t* eid, const double* ts,
                            int64_t num_events, int64_t num_nodes, int on_device, flid_graph** out,
                            flid_stream stream);
int flid_graph_build_entries(const int64_t* owner, const int64_t* nbr, const int64_t* eid, const double* ts,
                             int64_t num_entries, int64_t num_nodes, int on_device, flid_graph** out,
                             flid_stream stream);
void flid_graph_free(flid_graph* g);
int flid_graph_info(const flid_graph* g, int64_t* num_nodes, int64_t* num_entries, int64_t* max_degree);
/* copy the CSR to host arrays: indptr int64[num_nodes+2], nbr/eid int64[M], ts float64[M] (synchronous) */
int flid_graph_export_host(const flid_graph* g, int64_t* indptr_host, int64_t* nbr_host, int64_t* eid_host,
                           double* ts_host);

/* ---------------------------------------------------------------- sampler ---------
 * NeighborSampler.get_historical_neighbors, strategy 'recent' (utils/utils.py:149-214)
 * on top of find_neighbors_before (utils/utils.py:130-147, searchsorted side='left').
 * nodes int64[n]; times float64[n] or float32[n] (times_are_f32); outputs int64[n,k],
 * int64[n,k], float32[n,k]: the last <=k strictly-earlier interactions, right-aligned,
 * zero padded on the left, timestamps rounded float64->float32 (RN). Bit-exact.       */
int flid_sample_recent(const flid_graph* g, const int64_t* nodes, const void* times, int times_are_f32,
                       int64_t n, int k, int64_t* out_nbr, int64_t* out_eid, float* out_ts, flid_stream stream);
/* find_neighbors_before / get_all_first_hop_neighbors (utils/utils.py:130-147, 254-273):
 * per query the CSR range [start, cut) of strictly-earlier interactions.              */
int flid_sample_cut(const flid_graph* g, const int64_t* nodes, const void* times, int times_are_f32, int64_t n,
                    int64_t* out_start, int64_t* out_cut, flid_stream stream);

/* ------------------------------------------------------------------- TGAT ---------
 * Weights are handed over in the reference's state_dict layout ([out, in] row-major
 * float32), see models/modules.py:126-165 (MultiHeadAttention), :43-56 (MergeLayer),
 * :7-26 (TimeEncoder).                                                                */
typedef struct {
    const float* query_w; /* [qd, qd]  temporal_conv_layers.L.query_projection.weight, qd = dn + T       */
    const float* key_w;   /* [qd, kd]  ...key_projection.weight,   kd = dn + de + T                      */
    const float* value_w; /* [qd, kd]  ...value_projection.weight                                         */
    const float* ln_w;    /* [qd]      ...layer_norm.weight                                               */
    const float* ln_b;    /* [qd]      ...layer_norm.bias                                                 */
    const float* res_w;   /* [qd, qd]  ...residual_fc.weight                                              */
    const float* res_b;   /* [qd]      ...residual_fc.bias                                                */
    const float* fc1_w;   /* [dn, qd + dn]  merge_layers.L.fc1.weight                                     */
    const float* fc1_b;   /* [dn]                                                                         */
    const float* fc2_w;   /* [dn, dn]       merge_layers.L.fc2.weight                                     */
    const float* fc2_b;   /* [dn]                                                                         */
} flid_tgat_layer_weights;

int flid_tgat_create(int node_dim, int edge_dim, int time_dim, int num_layers, int num_heads, flid_tgat** out);
void flid_tgat_free(flid_tgat* m);
/* (Re)load weights; folds Wq/Wk and Wv/residual_fc.  time_w is
 * time_encoder.w.weight [T,1] (read as T floats), time_b is time_encoder.w.bias [T]. */
int flid_tgat_set_weights(flid_tgat* m, const float* time_w, const float* time_b,
                          const flid_tgat_layer_weights* layers_host, flid_stream stream);
/* Optional: cache the layer-1 query fold of every row of a node-feature table
 * (rows = num_nodes+1).  Valid until the next set_weights or until the table changes;
 * flid_tgat_embed uses it when called with the same node_feat pointer.               */
int flid_tgat_cache_node_table(flid_tgat* m, const float* node_feat, int64_t rows, flid_stream stream);
/* refresh `n` rows (int32 ids) of the cached table after node_feat changed (TGN)     */
int flid_tgat_refresh_node_rows(flid_tgat* m, const float* node_feat, const int32_t* row_ids, int64_t n,
                                flid_stream stream);
/* TGAT.compute_node_temporal_embeddings (models/TGAT.py:68-144) at current_layer_num =
 * num_layers, eval mode, for n root queries (node, time).  times float64[n], or the
 * float32 values widened to float64 with times_are_f32 != 0 (the recursion's dtype rule,
 * models/TGAT.py:110-125).  node_feat [N+1, dn], edge_feat [E+1, de] float32 row-major.
 * out float32 [n, dn].  GraphAttentionEmbedding (models/MemoryModel.py:632-715) is the
 * same call with node_feat = memory' + raw.                                            */
int flid_tgat_embed(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                    const int64_t* nodes, const double* times, int times_are_f32, int64_t n, int k, float* out,
                    flid_stream stream);
/* Layer memo for bulk passes (the E/200-iteration loops of PTCL/E_step.py:305-352,
 * PTCL/M_step.py:454-509): the layer-l embedding of a neighbour slot, h_l(nbr, float32(ts)),
 * is a function of the adjacency entry it was sampled from, yet models/TGAT.py:108-113
 * recomputes it for every root whose neighbourhood contains the entry.  memo_l is a
 * caller-owned float32 table [entries + 1, dn] indexed by CSR position (row `entries` is the
 * padded slot's query (node 0, t = 0.0)); it is valid while weights, graph, feature tables
 * and k are unchanged.  flid_tgat_memo_build fills rows [row_lo, row_hi) of level `level`
 * (1-based, needs the complete level-1 table `memo_prev` when level > 1; null for level 1),
 * so ranks can build disjoint row ranges and all-gather them.  Rows are bit-identical to
 * what flid_tgat_embed computes internally for the same slot.                            */
int flid_tgat_memo_build(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat, int k,
                         int level, const float* memo_prev, int64_t row_lo, int64_t row_hi, float* memo_out,
                         flid_stream stream);
/* flid_tgat_embed with the lower num_layers-1 levels read from memo tables
 * (memo_tables_host = host array of num_layers-1 device pointers, level 1 first): one
 * sampling pass and num_layers attention evaluations per root instead of
 * sum_l (1+k)^(L-l).  Per-batch calls give the same results as flid_tgat_embed bit for bit; bulk calls
 * (see flid_tgat_bulk_invalidate) use the projected formulation and agree to fp32 rounding.        */
int flid_tgat_embed_memo(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                         const float* const* memo_tables_host, const int64_t* nodes, const double* times,
                         int times_are_f32, int64_t n, int k, float* out, flid_stream stream);
/* One rank's share of an owner-partitioned (multi-GPU) memo build: work items [item_lo, item_hi) are adjacency
 * positions in owner-major order -- item q evaluates the query (owner of q, float32 time of q), i.e. the memo row of
 * q's partner entry, and writes it to that row of memo_out (a full-size table; rows of other ranks' partitions are
 * exchanged by the caller).  Every window an item reads lies inside the owner's adjacency list, so with item ranges
 * cut at node boundaries a rank only reads table rows / projected entries of its own position range (see
 * flid_tgat_set_bulk_range).  with_padded_row: also evaluate row `entries` (the padded slot's query (0, 0.0)).
 * Needs a graph built from events (partner index).                                                            */
int flid_tgat_memo_build_owner_range(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                                     int k, int level, const float* memo_prev, int64_t item_lo, int64_t item_hi,
                                     int with_padded_row, float* memo_out, flid_stream stream);
/* Restrict the per-entry tables of the projected bulk path to CSR positions [pos_lo, pos_hi) (+ the padded row):
 * every target this handle is asked for from now on has its neighbour window inside that range (one rank of an
 * owner-partitioned pass).  pos_hi < 0 restores the whole adjacency.                                          */
int flid_tgat_set_bulk_range(flid_tgat* m, int64_t pos_lo, int64_t pos_hi);
/* partner index of the adjacency (position of the same event's entry in the other endpoint's list), int32[entries],
 * copied into a caller-owned device buffer; graphs built from events only                                      */
int flid_graph_export_mirror(const flid_graph* g, int32_t* out_dev, flid_stream stream);
/* Peer-mapped memo tables of an owner-partitioned pass (one process per GPU).  flid_peer_alloc: cudaMalloc + the
 * 64-byte CUDA IPC handle to publish to the other ranks; flid_peer_open maps another rank's table into this process
 * (lazy peer access over NVLink); close / free undo them.                                                        */
int flid_peer_alloc(int64_t bytes, void** dev_ptr, void* handle_out_64_bytes);
int flid_peer_open(const void* handle_64_bytes, void** dev_ptr);
int flid_peer_close(void* dev_ptr);
int flid_peer_free(void* dev_ptr);
/* The per-level row exchange of the owner-partitioned memo build as one kernel: for every work item q of this rank's
 * range [pos_bounds[rank], pos_bounds[rank+1]) whose produced row (position p = partner of q) belongs to another
 * rank's range, copy row p of table_local into row p of that rank's table (peer_tables_host[dest], mapped with
 * flid_peer_open), 16-byte stores over NVLink.  Every row has one producer; the caller puts a cross-rank barrier
 * between this call and the first consumer of the tables.                                                      */
int flid_memo_exchange_p2p(const flid_graph* g, const float* table_local, void* const* peer_tables_host,
                           const int64_t* pos_bounds_host, int world, int rank, int row_dim, flid_stream stream);
/* Bulk calls (flid_tgat_memo_build; flid_tgat_embed_memo with at least entries / k roots) project every
 * adjacency entry once per pass -- the reference's key / value projections (models/modules.py:191-197) hoisted
 * from "per neighbour slot" to "per entry" -- and keep those tables inside the handle, keyed on the weights, the
 * graph and the table pointers.  Call this after rebuilding a memo table in place (same pointer, new contents)
 * or after editing a feature table in place, so that the projected tables are rebuilt too.                  */
int flid_tgat_bulk_invalidate(flid_tgat* m);
/* Switch the projected formulation of bulk calls off (0) or on (1, default): off, every call uses the per-slot
 * stream and memoised results equal flid_tgat_embed's bit for bit.  Takes effect at the next
 * flid_tgat_set_weights.  For tests and A/B timing.                                                        */
int flid_tgat_set_bulk_projection(flid_tgat* m, int enable);
/* A root query (v, t) that is itself an event of the graph (built with flid_graph_build_events)
 * at a float32-exact time finds its own lower-layer embeddings in the memo too (the table row
 * of the event's entry in the other endpoint's list), so flid_tgat_embed_memo evaluates only
 * the top layer for it; other roots take the full chain.  Same bits either way; on by default,
 * this switch exists for tests and A/B timing.                                            */
int flid_tgat_set_self_from_memo(flid_tgat* m, int enable);
/* Numeric mode of the projection GEMMs (BASELINE.json north_star; reference: the Q/K/V/out projections of
 * models/modules.py:186-197,227-241 and the MergeLayer of :66-68).
 *   0 (default): fp32-grade -- every product as three tf32 MMAs on hi/lo-split operands (rel. error ~2^-21),
 *                embeddings within fp32 rel 1e-4 of the reference;
 *   1: "bf16 projections" -- operands rounded to bfloat16 (RNE), one MMA per product, fp32 accumulation;
 *                tolerance rel 2e-2 with identical argmax on >= 99.9 % of nodes.
 * Takes effect at the next flid_tgat_set_weights (the weight images are tiled per mode).  The gather /
 * time-encode / softmax stream and the decoder stay fp32 in both modes.                              */
int flid_tgat_set_numeric_mode(flid_tgat* m, int mode);
/* Bulk passes: fold "+ residual, LayerNorm" (models/modules.py:235-238) into the first MergeLayer product
 * (:66): the GEMM's producers add the residual row and keep each row's sum and sum of squares, its epilogue applies
 * rstd * (acc - mean * rowsum(W diag(gamma))) + (W beta + bias + raw-feature block, one row per node).  No LayerNorm
 * kernel, no normalised rows in HBM, K = qd instead of qd + dn.  Same algebra, different rounding order (fp32
 * tolerance, not bit-identical to the unfolded path); the switch is per handle so that every chunk of a pass takes
 * the same path whatever its size.  Takes effect at the next flid_tgat_cache_node_table.  Static node tables only. */
int flid_tgat_set_ln_fold(flid_tgat* m, int enable);
/* Bulk memoised calls (n >= 8192 roots) evaluate their roots in (node, time) order and scatter the rows back
 * (default on): consecutive roots then share most of their neighbour window.  A caller whose roots already
 * arrive in that order (the owner-partitioned pass keeps its routed roots sorted) switches the sort off.
 * Results do not depend on it.                                                                                 */
int flid_tgat_set_sort_queries(flid_tgat* m, int enable);
/* Owner-partitioned passes: the exchange of the last memo level runs on a side stream beside the next call's
 * sampling and query-side GEMM, which only read rows this rank produced itself.  cuda_event (a cudaEvent_t recorded
 * on that side stream after the exchange and its cross-rank barrier) is waited for by the next attention launch
 * above level 1 -- the first reader of rows other ranks produced -- and then forgotten.  The caller keeps the event
 * alive until then.  NULL clears it.                                                                          */
int flid_tgat_set_wait_event(flid_tgat* m, void* cuda_event);
/* upper bound on layer-1 targets processed per internal chunk (workspace ~7 KB per target;
 * default 606208 = 148 x 4096).  Results do not depend on it.                                   */
int flid_tgat_set_chunk_targets(flid_tgat* m, int64_t max_layer1_targets);
/* Per-kernel-class CUDA-event timing of flid_tgat_embed (events recorded on the launch
 * stream around each launch group).  Classes: 0 level sampler, 1 query-fold GEMM,
 * 2 attention stream, 3 out-projection + LayerNorm + MergeLayer chain.
 * flid_tgat_profile_read synchronises, returns summed milliseconds / event-pair counts
 * since the last read, and resets.                                                      */
int flid_tgat_profile(flid_tgat* m, int enable);
int flid_tgat_profile_read(flid_tgat* m, double ms[4], int64_t launches[4]);
/* bytes / counts of the last flid_tgat_embed call, for the roofline report:
 * stats[0] = attention evaluations, stats[1] = valid (non-padded) neighbour rows gathered by them,
 * stats[2] = sampler queries, stats[3] = workspace bytes currently held.             */
int flid_tgat_last_stats(const flid_tgat* m, int64_t stats[4]);

/* --------------------------------------------------------------- TGN (memory) -----
 * MemoryModel('TGN').compute_src_dst_node_temporal_embeddings (models/MemoryModel.py:96-189).
 * All state lives in caller-owned device arrays so that the Python module can expose
 * them under the reference's state_dict names (memory_bank.node_memories, ...).      */
typedef struct {
    int64_t num_rows;        /* N+1 (incl. padding node 0)                                              */
    float* memories;         /* [rows, dn]   memory_bank.node_memories                                  */
    float* last_updated;     /* [rows]       memory_bank.node_last_updated_times                        */
    float* pending_msg;      /* [rows, 2dn+T+de] last raw message per node (node_raw_messages[v][-1][0]) */
    double* pending_ts;      /* [rows]       its timestamp (node_raw_messages[v][-1][1])                */
    uint8_t* has_pending;    /* [rows]                                                                   */
    float* next_memories;    /* [rows, dn]   GRU(pending, memory) where pending, else memory            */
    float* layer0;           /* [rows, dn]   next_memories + node_raw  (layer-0 / merge input table)    */
    int32_t* scratch;        /* [rows] int32, must be -1 filled before first use                        */
} flid_tgn_state;

typedef struct {
    const float* weight_ih; /* [3dn, 2dn+T+de]  memory_updater.memory_updater.weight_ih */
    const float* weight_hh; /* [3dn, dn]                                                */
    const float* bias_ih;   /* [3dn]                                                    */
    const float* bias_hh;   /* [3dn]                                                    */
} flid_gru_weights;

/* zero the state (MemoryBank.__init_memory_bank__, models/MemoryModel.py:359-366) and
 * set layer0 = node_raw.                                                               */
int flid_tgn_reset(const flid_tgn_state* s, const float* node_raw, int node_dim, int msg_dim, flid_stream stream);
/* recompute next_memories / layer0 for every row from (memories, pending) -- after a
 * state reload or a weight change.                                                     */
int flid_tgn_rebuild(flid_tgat* m, const flid_tgn_state* s, const flid_gru_weights* gru, const float* node_raw,
                     flid_stream stream);
/* one batch: embeddings for cat[src,dst] at cat[t,t], then (if positive) persist the
 * pending updates of the batch nodes, build their new raw messages (dst role stored
 * after src role, so it wins) and refresh next_memories/layer0.  err_flag (device int32)
 * is set to 1 if the reference's "update memory to time in the past" assertion would fire
 * and to 2 if a node / edge id is out of range (the id is then clamped to padding).
 * out may be null with positive != 0: only the state update runs (training-mode calls compute
 * their embeddings through autograd on the host side).                                   */
int flid_tgn_step(flid_tgat* m, const flid_graph* g, const flid_tgn_state* s, const flid_gru_weights* gru,
                  const float* node_raw, const float* edge_feat, const int64_t* src, const int64_t* dst,
                  const double* times, const int64_t* eids, int64_t batch, int positive, int k,
                  float* out /* [2*batch, dn]: src rows then dst rows */, int32_t* err_flag, flid_stream stream);

/* A whole chronological pass (the batch loop of PTCL/M_step.py:454-509 with model_name='TGN') as one call: events
 * [0, num_events) of the device arrays are fed to flid_tgn_step's kernels in batches of `batch` (the batch boundary
 * is part of the semantics), the per-event embeddings go to out_src / out_dst float32[num_events, dn].  use_graph:
 * the launch sequence of a full batch is captured once as a CUDA graph and replayed with a device-side batch
 * counter, so a batch costs one graph launch instead of ~25 kernel launches and a host synchronisation.  The error
 * flag is cumulative and read by the caller after the pass (the reference raises inside the offending batch).   */
int flid_tgn_pass(flid_tgat* m, const flid_graph* g, const flid_tgn_state* s, const flid_gru_weights* gru,
                  const float* node_raw, const float* edge_feat, const int64_t* src, const int64_t* dst,
                  const double* times, const int64_t* eids, int64_t num_events, int64_t batch, int k, float* out_src,
                  float* out_dst, int32_t* err_flag, int use_graph, flid_stream stream);

/* -------------------------------------------------------- pseudo-label scoring ----
 * MLPClassifier.forward (models/modules.py:86-97, eval) + softmax/argmax emission
 * (PTCL/E_step.py:334-335) fused: emb float32[n, in] -> probs float32[n, C], labels int64[n]. */
typedef struct {
    const float* fc1_w; const float* fc1_b; /* [h1, in], [h1]  */
    const float* fc2_w; const float* fc2_b; /* [h2, h1], [h2]  */
    const float* fc3_w; const float* fc3_b; /* [C, h2],  [C]   */
    int input_dim, hidden1, hidden2, num_classes;
} flid_mlp_weights;
int flid_pseudo_label(const flid_mlp_weights* w, const float* emb, int64_t n, float* probs, int64_t* labels,
                      float* logits_or_null, flid_stream stream);
/* entropy_filter (EST, PTCL/utils.py:38-54): probs_store_host = host array of num_iters
 * device pointers to float32[n, C]; labels float32[n] are set to -1 where
 * -sum p*log2(p+1e-10) > threshold, p = softmax(sum_iters probs).                      */
int flid_entropy_filter(const float* const* probs_store_host, int num_iters, int64_t n, int num_classes,
                        float threshold, float* labels, flid_stream stream);
/* prob_filter (CST, PTCL/utils.py:56-67): labels[i] = -1 where max_c probs_last[i,c] < threshold */
int flid_prob_filter(const float* probs_last, int64_t n, int num_classes, float threshold, float* labels,
                     flid_stream stream);

/* ------------------------------------------------------------------ training mode ---
 * Forward and backward of the attention stream for the M-step batches that run with autograd and
 * dropout (PTCL/M_step.py:196-325, NPL/NPL.py:185-314, PTCL/EM_warmup.py:113-238): the part of
 * MultiHeadAttention.forward (models/modules.py:183-231) that is not a dense projection,
 *     x_j = [table[hrow_j] | edge_feat[eid_j] | cos(fma(dt_j, time_w, time_b))]
 *     z_h = sum_j dropout(softmax_j(masked_fill(u_h . x_j, nbr_j == 0, -1e10)))_hj x_j ,
 * with u the query folded through the key projection (scaling included, natural-log domain).
 * u, z, dz, du: float32 [n, num_heads, node_dim + edge_dim + time_dim]; hrow/nbr/eid int64 [n, k];
 * dt float32 [n, k]; probs float32 [n, num_heads, k] (softmax before dropout, written by fwd, read
 * by bwd).  Dropout bits are Philox4x32-10(seed; target, slot), word h = head h, identical in fwd
 * and bwd; flid_attn_train_keep_mask exports them (uint8 [n, num_heads, k]) for tests.
 * bwd: dtable (nullable) [rows, node_dim] is ACCUMULATED into (atomic adds, zero it first);
 * dtime_partial (nullable) float32 [flid_attn_train_partials(n), 2, time_dim] receives per-block
 * sums of (dL/dtime_w, dL/dtime_b), to be summed over the first axis by the caller.         */
int64_t flid_attn_train_partials(int64_t n);
int flid_attn_train_fwd(const float* u, const float* table, const int64_t* hrow, const int64_t* nbr,
                        const int64_t* eid, const float* dt, const float* edge_feat, const float* time_w,
                        const float* time_b, int64_t n, int k, int num_heads, int node_dim, int edge_dim,
                        int time_dim, float p_drop, uint64_t seed, float* z, float* probs, flid_stream stream);
int flid_attn_train_bwd(const float* u, const float* table, const int64_t* hrow, const int64_t* nbr,
                        const int64_t* eid, const float* dt, const float* edge_feat, const float* time_w,
                        const float* time_b, int64_t n, int k, int num_heads, int node_dim, int edge_dim,
                        int time_dim, float p_drop, uint64_t seed, const float* probs, const float* dz, float* du,
                        float* dtable, float* dtime_partial, flid_stream stream);
int flid_attn_train_keep_mask(uint64_t seed, int64_t n, int num_heads, int k, float p_drop, uint8_t* keep,
                              flid_stream stream);

/* Training-mode forward / backward of the whole L-layer attention stack (models/TGAT.py:68-144 and
 * models/MemoryModel.py:632-715 under autograd; each layer is models/modules.py:167-245 + :58-69), level-batched.
 *
 * flid_train_sample_levels: the top-down sampling for n roots.  Level l (1..num_levels, index l-1 in the host
 * arrays of device pointers) holds n*(1+k)^(num_levels-l) targets = [targets of level l+1 ; their neighbours]:
 * ids int64, t64 float64 query times, nbr / eid int64 [n_l, k], dt float32 [n_l, k] with the reference's dtype
 * rules (TGAT.py:120-125).  Root ids must be valid (the caller checks them on the host); times_are_f32 = the
 * roots carry float32 times (widened).
 *
 * Weights per layer (flid_train_weights[l-1]); the caller folds the projections (differentiably) and passes
 *   fold_q [H*kd, qd] = scaling * Wk_h^T Wq_h stacked over heads,   fold_o [qd, H*kd] = residual_fc.weight . blockdiag(Wv_h)
 * (kd = node+edge+time, qd = node+time); the backward returns the gradients of those folded matrices.
 * te0 [time_dim] = cos(time_b), the time encoding of the query (TGAT.py:90); its gradient comes back in d_te0.
 * flid_train_saved[l-1]: caller-allocated tensors written by fwd and read by bwd, n = lv[l-1].n:
 *   q, y, ln [n, qd]; merge_self, hid, out [n, node_dim]; u, z [n, H*kd]; probs [n, H, k].
 * The result is saved[num_layers-1].out.  pre_scratch: [lv[0].n, qd] floats.
 * bwd: every member of flid_train_grads[l-1], d_te0 and d_node_feat (nullable: [rows, node_dim], the gradient of
 * the layer-0 table, e.g. TGN's memories + raw features) are ACCUMULATED into (zero them first);
 * scratch: flid_train_model_scratch_floats(...) floats.  Dropout (scores, modules.py:224, and residual_fc
 * output, :235) is Philox4x32-10 keyed by seeds_host[l-1], regenerated in bwd.                              */
typedef struct {
    const float *fold_q, *fold_o, *res_b, *ln_w, *ln_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b, *time_w, *time_b;
} flid_train_weights;
typedef struct {
    float *fold_q, *fold_o, *res_b, *ln_w, *ln_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b, *time_w, *time_b;
} flid_train_grads;
typedef struct {
    float *q, *merge_self, *u, *probs, *z, *y, *ln, *hid, *out;
} flid_train_saved;
typedef struct {
    const int64_t *ids, *nbr, *eid;
    const float* dt;
    int64_t n;
} flid_train_level;
int flid_train_sample_levels(const flid_graph* g, const int64_t* roots, const double* times, int times_are_f32,
                             int64_t n, int k, int num_levels, int64_t* const* ids_host, double* const* t64_host,
                             int64_t* const* nbr_host, int64_t* const* eid_host, float* const* dt_host,
                             flid_stream stream);
int64_t flid_train_model_scratch_floats(int64_t n_roots, int k, int num_layers, int num_heads, int node_dim,
                                        int edge_dim, int time_dim);
int flid_train_model_fwd(const flid_train_weights* w, const flid_train_level* lv, const flid_train_saved* sv,
                         const float* node_feat, const float* edge_feat, const float* te0, int num_layers, int k,
                         int num_heads, int node_dim, int edge_dim, int time_dim, float p_drop,
                         const uint64_t* seeds_host, float* pre_scratch, flid_stream stream);
int flid_train_model_bwd(const flid_train_weights* w, const flid_train_level* lv, const flid_train_saved* sv,
                         const float* node_feat, const float* edge_feat, int num_layers, int k, int num_heads,
                         int node_dim, int edge_dim, int time_dim, float p_drop, const uint64_t* seeds_host,
                         const float* d_out, const flid_train_grads* grads, float* d_te0, float* d_node_feat,
                         float* scratch, flid_stream stream);
/* test hook: the residual_fc-output dropout bits of a layer evaluated with this seed (uint8 [n, qd], 1 = kept) */
int flid_train_layer_out_keep_mask(uint64_t seed, int64_t n, int qd, float p_drop, uint8_t* keep, flid_stream stream);

/* ------------------------------------------------------------------ GraphMixer node encoder ---
 * models/GraphMixer.py:119-146 (other consumers of the sampler): for n queries
 * (nodes int64, times float64 or float32) the `time_gap` most recent neighbours before the query time,
 * out[i] = mean_j(node_feat[nbr_j] * softmax_j({1 real, -1e10 padded})) (+ node_feat[nodes[i]] when
 * add_self), float32 [n, node_dim].  Node ids must be valid (the caller checks them on the host).   */
int flid_neighbor_mean(const flid_graph* g, const float* node_feat, int node_dim, const int64_t* nodes,
                       const void* times, int times_are_f32, int64_t n, int time_gap, int add_self, float* out,
                       flid_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Dense halves of GraphMixer's link encoder (models/GraphMixer.py:91-117, :172-246) and TCL's encoder
 * (models/TCL.py:108-157, models/modules.py:248-312), evaluation (forward-only) path; csrc/dense.cu.
 * Everything is float32, row-major, on the current device; pointers are device pointers.                        */
typedef struct flid_dense_weight flid_dense_weight;   /* tiled tensor-core image of one nn.Linear weight */

/* weight: [n_out, n_in] with row stride ldw (nn.Linear layout; a row slice of a larger matrix is fine), n_in % 4 == 0.
 * Returns NULL on error (flid_last_error).  _update re-tiles after the parameter changed.                          */
flid_dense_weight* flid_dense_weight_create(const float* weight, int64_t ldw, int n_out, int n_in, flid_stream stream);
int flid_dense_weight_update(flid_dense_weight* h, const float* weight, int64_t ldw, flid_stream stream);
void flid_dense_weight_free(flid_dense_weight* h);

/* c[r, :] = act( [a0[i0(r), :w0] | a1[i1(r), :w1]] @ W^T + bias + resid[r, :] ),  r < m.
 * i(r) = idx[r] when idx is given (int32 row gather: torch's x[ids]), else r; w0 + w1 == n_in; w1 == 0: one segment;
 * bias, resid nullable; act 0 none, 1 ReLU, 2 exact (erf) GELU -- F.relu / nn.GELU() of the reference modules.   */
int flid_dense(const flid_dense_weight* h, const float* a0, const int32_t* idx0, int64_t lda0, int w0, const float* a1,
               const int32_t* idx1, int64_t lda1, int w1, const float* bias, const float* resid, int64_t ldr, int act,
               float* c, int64_t ldc, int64_t m, flid_stream stream);

/* out[r, :] = cos(dt[r] * w + b) (TimeEncoder, modules.py:25-38), zeros where ids[r] == 0 (ids nullable)
 * -- GraphMixer.py:103-108, TCL.py:118.                                                                            */
int flid_time_rows(const float* dt, const int64_t* ids, const float* w, const float* b, int time_dim, float* out, int64_t n,
                   flid_stream stream);

/* MLPMixer token mixing (GraphMixer.py:226-236): x, out [m, num_tokens, channels];
 * out = x + (Linear(hidden, tokens) o GELU o Linear(tokens, hidden) o LayerNorm_tokens)(x^T)^T.                    */
int flid_token_mix(const float* x, int num_tokens, int channels, const float* ln_w, const float* ln_b, float eps,
                   const float* w1, const float* b1, const float* w2, const float* b2, int hidden, float* out, int64_t m,
                   flid_stream stream);
/* out[q, :channels] = mean over the tokens of x[q] (GraphMixer.py:117); out row stride ldo.                        */
int flid_token_mean(const float* x, int num_tokens, int channels, float* out, int64_t ldo, int64_t m, flid_stream stream);
/* nn.LayerNorm over the last dimension (dim % 4 == 0, <= 1024).  y may alias x.                                    */
int flid_row_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, float* y, int64_t ldy,
                       int64_t m, int dim, flid_stream stream);
/* x[r, :] += table[r % period, :] -- the depth embedding rows of TCL.py:127-131 over [batch * period, dim].       */
int flid_add_periodic_rows(float* x, const float* table, int period, int dim, int64_t m, flid_stream stream);
/* Core of nn.MultiheadAttention as TransformerEncoder calls it (modules.py:287-300): per sequence of seq_len tokens
 * softmax(q k^T / sqrt(head_dim) with keys whose key_ids == 0 masked out) v for every head; q, k, v are the
 * in-projected rows [num_seqs * seq_len, num_heads * head_dim] with their own row strides; only the first q_rows
 * query tokens of each sequence are evaluated (seq_len: all of them).                                              */
int flid_seq_attention(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                       const int64_t* key_ids, int seq_len, int num_heads, int head_dim, float* out, int64_t ldo,
                       int q_rows, int64_t num_seqs, flid_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* FLID_B200_H */
