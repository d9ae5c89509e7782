// TGAT handle: folded weights + grow-only workspace.  See DESIGN.md for the algebra.
#pragma once
#include <vector>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "graph.cuh"

namespace flid {
struct LayerDev {
    float* mfoldT = nullptr;  // [H*kd, qd]  scale * Wk_h^T Wq_h   (u_h = mfoldT_h . [h_self | te0])
    float* u0 = nullptr;      // [H*kd]      mfoldT[:, dn:] . te0   (the constant time part of the query)
    float* wvoT = nullptr;    // [qd, H*kd]  Wr[:, head h] Wv_h     (residual_fc folded into the value projection)
    float *res_b = nullptr, *ln_w = nullptr, *ln_b = nullptr;              // [qd]
    float *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;  // merge layer, reference layout
    TcWeight tc_q, tc_o, tc_f1, tc_f2;  // hi/lo-split, tiled images for the tcgen05 GEMM
    // LayerNorm folded into fc1 (bulk passes; gemm_tc.cuh TcGemmArgs::ln_*)
    float* lnw = nullptr;                  // one allocation: w1g [dn, qd] = fc1_w[:, :qd] diag(gamma), c1 [dn] its row sums,
    float *w1g = nullptr, *ln_c1 = nullptr, *ln_c2 = nullptr;   // c2 [dn] = fc1_w[:, :qd] beta + fc1_b
    TcWeight tc_f1g, tc_f1n;               // W diag(gamma) (K = qd) and the raw-feature block fc1_w[:, qd:] (K = dn)
    DevBuf nt;                             // [table rows, dn] = node_feat . fc1_w[:, qd:]^T + c2, per cached node table
    // projected bulk path (bulk_kv.cu): per-entry K/V rows instead of per-slot raw rows.  Row layouts are
    // head-interleaved (kv_perm) so that a lane of the stream kernel only ever touches its own head.
    float* kvw = nullptr;     // one allocation holding the small fp32 matrices below
    float *wqs = nullptr, *cqs = nullptr;  // [qd, dn], [qd]   qs = scale * Wq [h | te0]      (level >= 2 targets)
    float *wut = nullptr, *cut = nullptr;  // [H*T, dn], [H*T] time part of the folded query
    float *wk2 = nullptr, *wv2 = nullptr;  // [qd, dn+de]      K / V of [h_{l-1} | e]          (level >= 2 entries)
    float *wvn = nullptr, *wve = nullptr;  // [qd, dn], [qd, de]  V of node / edge rows          (level 1)
    float* wo2 = nullptr;                  // [qd, qd + H*T]   residual_fc on [sum a V | Wv_t sum a te]
    TcWeight tc_qs, tc_ut, tc_k2, tc_v2, tc_vn, tc_ve, tc_o2;
};
}  // namespace flid

struct flid_tgat {
    int dn = 0, de = 0, T = 0, L = 0, H = 0;
    int qd = 0, kd = 0, hd = 0, zw = 0;  // zw = H * kd
    bool have_weights = false;
    bool ln_fold = false;   // flid_tgat_set_ln_fold: bulk passes fold LayerNorm into fc1 (one path whatever the chunk size)
    const float* nt_src = nullptr;   // node table the per-layer `nt` tables were built from
    int numeric = 0;     // flid_tgat_set_numeric_mode: 0 fp32-grade (3xTF32), 1 bf16-rounded operands / one MMA per product
    bool use_tc = true;  // projection GEMMs on tcgen05 (3xTF32); false = fp32 SIMT (FLID_GEMM=simt)
    float *time_w = nullptr, *time_b = nullptr, *te0 = nullptr, *time_bound = nullptr;
    std::vector<flid::LayerDev> layers;
    flid::DevBuf raw_q, raw_k, raw_v, raw_r;  // staging for the fold
    // cached layer-1 query fold per node-table row
    const float* table_src = nullptr;
    int64_t table_rows = 0;
    flid::DevBuf table;
    // workspace
    // 148 SMs x 4096: every 128-row GEMM tile round and every 4-target attention block round is full, and the launch /
    // tail cost of a level is paid once per 0.6 M targets (measured per Reddit-shape pass: 75 776 -> 33.0 ms, 151 552 -> 31.4,
    // 303 104 -> 30.6, 606 208 -> 30.2, 1 363 968 -> 29.9; smaller is worse: 37 888 -> 36.0).  ~7 GB of workspace at L = 2.
    int64_t max_l1_targets = 606208;
    bool sort_bulk_queries = true;  // bulk memoised calls evaluate their roots in (node, time) order (FLID_SORT_QUERIES=0 disables)
    bool self_from_memo = true;  // roots that are graph events read their own lower layers from the memo (FLID_SELF_MEMO=0 disables)
    flid::DevBuf ws_ids, ws_times, ws_nbr, ws_eid, ws_dt, ws_h, ws_u, ws_z, ws_o, ws_a, ws_hd, ws_misc, ws_pos, ws_self, ws_sort;
    flid::DevBuf ws_rid, ws_rt, ws_bad;  // root conversion staging
    flid::DevBuf ws_win;                 // (cut, cnt) per target of the current chunk (windowed stream kernel)
    bool kv_windowed = true;             // FLID_KV_KERNEL=mask selects the slot-mask kernel (A/B timing)
    // projected bulk path: tables derived from (weights, graph, feature tables[, memo of the level below])
    bool kv_enabled = true;             // FLID_BULK_KV=0 disables (A/B timing)
    uint64_t weights_version = 0;       // bumped by flid_tgat_set_weights
    uint64_t bulk_epoch = 0;            // bumped by flid_tgat_bulk_invalidate (the caller rebuilt a memo table)
    struct KvKey {
        uint64_t wv = 0, epoch = 0;
        const void *g = nullptr, *nf = nullptr, *ef = nullptr, *memo = nullptr;
        int64_t entries = -1, lo = 0, hi = -1;
        bool operator==(const KvKey& o) const {
            return wv == o.wv && epoch == o.epoch && g == o.g && nf == o.nf && ef == o.ef && memo == o.memo &&
                   entries == o.entries && lo == o.lo && hi == o.hi;
        }
    };
    // owner-partitioned passes (one rank of a multi-GPU pass): every target this handle sees has its window inside
    // the CSR position range [bulk_lo, bulk_hi), so the per-entry tables are only filled there (+ the padded row);
    // bulk_hi < 0 = the whole adjacency
    int64_t bulk_lo = 0, bulk_hi = -1;
    bool kv_ve_by_pos = false;           // level-1 edge projections indexed by CSR position (range builds) instead of edge id
    KvKey kv_l1_key;
    flid::DevBuf kv_vn1, kv_ve1, kv_s1;             // level 1: V of node rows [N+1, qd], V of edge rows [max_eid+1, qd], scores [M+1, H]
    std::vector<flid::DevBuf> kv_tab;               // level l >= 2 (index l-2): [M+1, 2*qd] = [K | V] of entry p
    std::vector<KvKey> kv_tab_key;
    flid::DevBuf tgn_ids, tgn_times, tgn_eids, tgn_gi, tgn_gh;  // TGN step scratch (tgn.cu): per handle, hence per device
    flid::DevBuf tgn_out, tgn_ctr;       // whole-pass driver (flid_tgn_pass): batch embeddings, device batch counter
    flid::TcWeight tc_gih, tc_ghh;       // GRU weights tiled for the tcgen05 GEMM (flid_tgn_rebuild), fp32-grade in both numeric modes
    const float *gru_ih_src = nullptr, *gru_hh_src = nullptr;
    // set by flid_tgat_set_wait_event: the next attention launch above level 1 (the first reader of exchanged memo
    // rows) waits for it; consumed once
    cudaEvent_t wait_event = nullptr;
    // weight upload: the independent fold / tiling chains of all layers run side by side on these streams
    static constexpr int PREP_STREAMS = 6;
    cudaStream_t prep_stream[PREP_STREAMS] = {};
    cudaEvent_t prep_fork = nullptr, prep_join[PREP_STREAMS] = {};
    cudaStream_t tgn_stream = nullptr;   // capturable stream of the whole-pass driver
    cudaEvent_t tgn_ev = nullptr;
    int64_t stats[4] = {0, 0, 0, 0};
    int64_t valid_mult = 1;  // attention evaluations that consume each sampled neighbour list
    // optional per-kernel-class CUDA-event timing (bench.py's roofline numbers)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;  // pairs (begin, end)
    std::vector<int> prof_cls;
    size_t prof_used = 0;
};

namespace flid {
// run the full L-layer embedding for n roots given as int32 ids (device) -- shared by the
// TGAT entry point and the TGN step.
int tgat_embed_ids(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat,
                   const int32_t* ids, const double* times, int64_t n_f64, int64_t n, int k, float* out,
                   cudaStream_t st);

// ---- projected bulk path (bulk_kv.cu)
bool kv_supported(const flid_tgat* m);
int kv_fold_layer(flid_tgat* m, int l, const float* wq, const float* wk, const float* wv, const float* wr, cudaStream_t st);
void kv_free_layer(LayerDev& d);
// level-1 tables (V of node / edge rows, per-entry scores against the owner's folded query); needs the cached node table
int kv_ensure_level1(flid_tgat* m, const flid_graph* g, const float* node_feat, const float* edge_feat, cudaStream_t st);
// level >= 2: K/V rows of every adjacency entry from the memo of the level below
int kv_ensure_level(flid_tgat* m, const flid_graph* g, int level, const float* memo_prev, const float* node_feat,
                    const float* edge_feat, cudaStream_t st);
struct KvCall {
    int level = 1;
    int64_t n = 0;
    const int32_t* ids = nullptr;       // level 1: node ids of the targets (rows of the cached query table)
    const float* self_base = nullptr;   // level >= 2: layer-(l-1) rows of the targets
    const int32_t* self_idx = nullptr;
    const int32_t *nbr = nullptr, *eid = nullptr, *pos = nullptr;
    const float* dt = nullptr;
    const int2* win = nullptr;          // non-null: (cut, cnt) per target -> the windowed stream kernel
    const int2* adj = nullptr;          // graph adjacency by position (windowed kernel, level 1)
    int pad_pos = 0;                    // position of the padded slot's row in the projected tables
    bool graph_zero_nbr = false;        // an entry with neighbour id 0 exists: only the slot-mask kernel handles it
    float *U = nullptr, *Y = nullptr;   // workspaces: [n, qd + H*T] each
};
// query side (level >= 2: two GEMMs) + stream kernel; Y = [sum a V (interleaved) | sum a te per head]
int kv_attention(flid_tgat* m, const KvCall& c, int k, cudaStream_t st);

// classes for the event timer
enum { PROF_SAMPLE = 0, PROF_QFOLD = 1, PROF_ATTN = 2, PROF_OUT = 3, PROF_CLASSES = 4 };
struct ProfScope {
    flid_tgat* m;
    cudaStream_t st;
    size_t slot = 0;
    bool on;
    ProfScope(flid_tgat* m_, int cls, cudaStream_t st_) : m(m_), st(st_), on(m_->prof_on) {
        if (!on) return;
        if (m->prof_used + 2 > m->prof_ev.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a), cudaEventCreate(&b);
            m->prof_ev.push_back(a), m->prof_ev.push_back(b);
            m->prof_cls.push_back(cls);
        }
        slot = m->prof_used;
        m->prof_cls[slot / 2] = cls;
        m->prof_used += 2;
        cudaEventRecord(m->prof_ev[slot], st);
    }
    ~ProfScope() {
        if (on) cudaEventRecord(m->prof_ev[slot + 1], st);
    }
};
}  // namespace flid
