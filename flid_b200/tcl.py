"""Drop-in for the reference's ``models.TCL.TCL`` (``models/TCL.py``; TransformerEncoder
``models/modules.py:248-312``): same constructor, method names and ``state_dict`` keys.  SURVEY.md
section 8(f) rank 4 -- another consumer of the time-sorted device CSR: the two
``get_historical_neighbors`` calls per batch and the sequence assembly ([node ; its k recent
neighbours], edge id 0 and time difference 0 in front) stay on the device.  The encoder itself
(three input projections + depth embedding, nn.MultiheadAttention self / cross attention over
21-token sequences, feed-forward, LayerNorm) runs, in evaluation (``model.eval()`` under
``torch.no_grad()``), on the library's own kernels (csrc/dense.cu): every Linear -- including the
attention in / out projections -- on the tcgen05 3xTF32 GEMM with the feature-row gathers, bias, ReLU
and residual fused, the 21 x 21 attention as one CTA per sequence.  In training it is composed from
torch CUDA modules with the reference's parameter names, so checkpoints load unchanged and it trains
with autograd as the reference does.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, dense
from .sampler import NeighborSampler
from .tgat import TimeEncoder


class TransformerEncoder(nn.Module):
    """Post-norm encoder block over nn.MultiheadAttention (models/modules.py:248-312)."""

    def __init__(self, attention_dim: int, num_heads: int, dropout: float = 0.1):
        super().__init__()
        self.multi_head_attention = nn.MultiheadAttention(embed_dim=attention_dim, num_heads=num_heads, dropout=dropout)
        self.dropout = nn.Dropout(dropout)
        self.linear_layers = nn.ModuleList([nn.Linear(attention_dim, 4 * attention_dim),
                                            nn.Linear(4 * attention_dim, attention_dim)])
        self.norm_layers = nn.ModuleList([nn.LayerNorm(attention_dim), nn.LayerNorm(attention_dim)])

    def forward(self, inputs_query, inputs_key=None, inputs_value=None, neighbor_masks=None):
        """inputs [B, S, d]; neighbor_masks: ids [B, S_key] (device tensor or ndarray), 0 = padded key."""
        if inputs_key is None or inputs_value is None:
            assert inputs_key is None and inputs_value is None
            inputs_key = inputs_value = inputs_query
        if neighbor_masks is not None:
            if not torch.is_tensor(neighbor_masks):
                neighbor_masks = torch.from_numpy(neighbor_masks).to(inputs_query.device)
            neighbor_masks = neighbor_masks == 0
        hidden = self.multi_head_attention(query=inputs_query.transpose(0, 1), key=inputs_key.transpose(0, 1),
                                           value=inputs_value.transpose(0, 1),
                                           key_padding_mask=neighbor_masks)[0].transpose(0, 1)
        out = self.norm_layers[0](inputs_query + self.dropout(hidden))
        hidden = self.linear_layers[1](self.dropout(F.relu(self.linear_layers[0](out))))
        return self.norm_layers[1](out + self.dropout(hidden))


class TCL(nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, num_layers: int = 2, num_heads: int = 2, num_depths: int = 20, dropout: float = 0.1,
                 device: str = 'cpu'):
        super().__init__()
        self.device = device                  # compute calls require CUDA (no CPU fallback); construction does not
        self.node_raw_features = torch.from_numpy(np.ascontiguousarray(node_raw_features, dtype=np.float32)).to(device)
        self.edge_raw_features = torch.from_numpy(np.ascontiguousarray(edge_raw_features, dtype=np.float32)).to(device)
        self.neighbor_sampler = neighbor_sampler
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.edge_feat_dim = self.edge_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.num_depths = num_depths
        self.dropout = dropout
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim)
        self.depth_embedding = nn.Embedding(num_embeddings=num_depths, embedding_dim=self.node_feat_dim)
        self.projection_layer = nn.ModuleDict({
            'node': nn.Linear(self.node_feat_dim, self.node_feat_dim, bias=True),
            'edge': nn.Linear(self.edge_feat_dim, self.node_feat_dim, bias=True),
            'time': nn.Linear(self.time_feat_dim, self.node_feat_dim, bias=True)})
        self.transformers = nn.ModuleList([TransformerEncoder(self.node_feat_dim, num_heads, dropout)
                                           for _ in range(num_layers)])
        self.output_layer = nn.Linear(self.node_feat_dim, self.node_feat_dim, bias=True)
        self.chunk_events = 16384           # bulk calls are processed in chunks of events
        self._dense = dense.DenseWeights()  # tiled weight images of the evaluation path
        self._cat = {}                      # derived evaluation-path parameters, keyed by the versions of their sources

    # ------------------------------------------------------------------ evaluation path on csrc/dense.cu
    def _features_eval(self, ids, eids, dt):
        """TCL.py:108-131 as two GEMMs: [node[ids] | edge[eids]] @ [Wn | We]^T + (bn + be + bt), then the time
        projection with that as its residual, then the depth rows.  Returns [m * S, d]."""
        pn, pe, pt = self.projection_layer['node'], self.projection_layer['edge'], self.projection_layer['time']
        ver = (pn.weight._version, pe.weight._version, pn.bias._version, pe.bias._version, pt.bias._version,
               pn.weight.data_ptr(), pe.weight.data_ptr())
        if self._cat.get('ver') != ver:
            self._cat = {'ver': ver, 'w': torch.cat([pn.weight, pe.weight], dim=1).contiguous(),
                         'b': (pn.bias + pe.bias + pt.bias).contiguous()}
        m, s = ids.shape
        assert s == self.depth_embedding.weight.shape[0]
        i32, e32 = ids.reshape(-1).to(torch.int32), eids.reshape(-1).to(torch.int32)
        base = dense.linear(self._dense, self.node_raw_features, self._cat['w'], self._cat['b'], idx=i32,
                            x2=self.edge_raw_features, idx2=e32)
        te = dense.time_rows(dt, None, self.time_encoder)
        x = dense.linear(self._dense, te, pt.weight, None, resid=base)
        _lib.check(_lib.lib().flid_add_periodic_rows(_lib.ptr(x), _lib.ptr(self.depth_embedding.weight), s, x.shape[1],
                                                     x.shape[0], _lib.stream()))
        return x

    def _encoder_eval(self, tr, xq, xkv, key_ids, s, first_only=False):
        """TransformerEncoder.forward (modules.py:270-312) on [m * S, d] rows; ``xkv is xq``: self-attention.
        ``first_only``: only token 0 of every query sequence is evaluated and returned ([m, d]) -- the last block's
        cross-attention outputs are read at token 0 alone (TCL.py:153-157), and everything after the attention
        is per token."""
        mha, d = tr.multi_head_attention, xq.shape[1]
        w_in, b_in, heads = mha.in_proj_weight, mha.in_proj_bias, mha.num_heads
        m = xq.shape[0] // s
        if first_only:
            assert xkv is not xq
            xq = xq.view(m, s * d)[:, :d]                                   # token-0 rows, row stride S * d
            q = torch.empty((m, s * d), dtype=torch.float32, device=xq.device)[:, :d]
            dense.linear(self._dense, xq, w_in[:d], b_in[:d], out=q)
            ctx = torch.empty((m, s * d), dtype=torch.float32, device=xq.device)[:, :d]
        else:
            ctx = torch.empty((m * s, d), dtype=torch.float32, device=xq.device)
        if xkv is xq:
            qkv = dense.linear(self._dense, xq, w_in, b_in)
            q, kk, vv, ldq = qkv, qkv[:, d:], qkv[:, 2 * d:], 3 * d
        else:
            if not first_only:
                q = dense.linear(self._dense, xq, w_in[:d], b_in[:d])
            kv = dense.linear(self._dense, xkv, w_in[d:], b_in[d:])
            kk, vv, ldq = kv, kv[:, d:], d                                  # the strided q / ctx are [m * S, d] with holes
        _lib.check(_lib.lib().flid_seq_attention(_lib.ptr(q), ldq, _lib.ptr(kk), kk.stride(0), _lib.ptr(vv), vv.stride(0),
                                                 _lib.ptr(key_ids), s, heads, d // heads, _lib.ptr(ctx), d,
                                                 1 if first_only else s, m, _lib.stream()))
        y = dense.linear(self._dense, ctx, mha.out_proj.weight, mha.out_proj.bias, resid=xq)
        out = dense.layernorm(y, tr.norm_layers[0], out=y)
        h = dense.linear(self._dense, out, tr.linear_layers[0].weight, tr.linear_layers[0].bias, act=1)
        y2 = dense.linear(self._dense, h, tr.linear_layers[1].weight, tr.linear_layers[1].bias, resid=out)
        return dense.layernorm(y2, tr.norm_layers[1], out=y2)

    def _sequences(self, d_ids, d_t, f32_times, k):
        """[node ; k recent neighbours] on the device: ids / edge ids int64 [m, k+1], time differences float32 (TCL.py:75-106, :178-180)."""
        nbr, eid, ts = self.neighbor_sampler.get_historical_neighbors_device(d_ids, d_t, k)
        ids = torch.cat([d_ids[:, None], nbr], dim=1)
        eids = torch.cat([torch.zeros_like(d_ids[:, None]), eid], dim=1)
        if f32_times:                        # numpy: float32 - float32 stays float32
            t32 = d_t.to(torch.float32)
            dt = t32[:, None] - torch.cat([t32[:, None], ts], dim=1)
        else:                                # float64 minus the concatenated float64 matrix, then .float()
            dt = (d_t[:, None] - torch.cat([d_t[:, None], ts.to(torch.float64)], dim=1)).to(torch.float32)
        return ids, eids, dt

    def _features(self, ids, eids, dt):
        """TCL.py:108-131: projected node + edge + time features plus the depth embedding."""
        w_t, b_t = self.time_encoder.w.weight.reshape(-1), self.time_encoder.w.bias
        te = torch.cos(torch.addcmul(b_t, dt.unsqueeze(-1), w_t))        # single-rounded fma, as nn.Linear(1, T)
        assert ids.shape[1] == self.depth_embedding.weight.shape[0]
        depth = self.depth_embedding(torch.arange(ids.shape[1], device=ids.device))
        return (self.projection_layer['node'](self.node_raw_features[ids])
                + self.projection_layer['edge'](self.edge_raw_features[eids])
                + self.projection_layer['time'](te) + depth)

    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray,
                                                 node_interact_times: np.ndarray, num_neighbors: int = 20):
        """TCL.py:60-157: two float32 [B, node_feat_dim] device tensors."""
        sampler = self.neighbor_sampler
        if not isinstance(sampler, NeighborSampler):
            raise TypeError(f"flid_b200 models need a flid_b200.NeighborSampler (device CSR); got {type(sampler).__name__}")
        k = int(num_neighbors)
        assert k > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        src = np.ascontiguousarray(src_node_ids, dtype=np.int64)
        dst = np.ascontiguousarray(dst_node_ids, dtype=np.int64)
        t_np = np.asarray(node_interact_times)
        b = src.shape[0]
        for a in (src, dst):
            if b and (int(a.min()) < 0 or int(a.max()) > sampler.num_nodes):
                raise IndexError("flid_b200.TCL: node id outside the graph")
        dev, f32 = _lib.require_cuda(self.node_raw_features.device), t_np.dtype == np.float32
        outs_s, outs_d = [], []
        with torch.cuda.device(dev):
            d_src = _lib.to_device(src, np.int64, dev, "tcl_src")
            d_dst = _lib.to_device(dst, np.int64, dev, "tcl_dst")
            d_t = _lib.to_device(t_np, np.float64, dev, "tcl_times")        # float32 -> float64 is exact
            for lo in range(0, max(b, 1), self.chunk_events):
                hi = lo + self.chunk_events
                ids_s, eid_s, dt_s = self._sequences(d_src[lo:hi], d_t[lo:hi], f32, k)
                ids_d, eid_d, dt_d = self._sequences(d_dst[lo:hi], d_t[lo:hi], f32, k)
                seq, d_model = ids_s.shape[1], self.node_feat_dim
                # one CTA holds the Q, K, V rows of a sequence in shared memory (csrc/dense.cu); longer sequences keep the modules
                fits = 4 * (3 * seq * (d_model + 1) + self.num_heads * seq * (seq + 1) + seq) <= 200 * 1024 and d_model % 4 == 0
                if dense.fast_path(self) and ids_s.shape[0] > 0 and fits:
                    m, s = ids_s.shape
                    ids_s, ids_d = ids_s.contiguous(), ids_d.contiguous()
                    xs, xd = self._features_eval(ids_s, eid_s, dt_s), self._features_eval(ids_d, eid_d, dt_d)
                    es = ed = None
                    for li, transformer in enumerate(self.transformers):    # TCL.py:133-151
                        last = li == len(self.transformers) - 1
                        xs = self._encoder_eval(transformer, xs, xs, ids_s, s)
                        xd = self._encoder_eval(transformer, xd, xd, ids_d, s)
                        es = self._encoder_eval(transformer, xs, xd, ids_d, s, first_only=last)
                        ed = self._encoder_eval(transformer, xd, xs, ids_s, s, first_only=last)
                        xs, xd = es, ed
                    outs_s.append(dense.linear(self._dense, es, self.output_layer.weight, self.output_layer.bias))   # token-0 rows
                    outs_d.append(dense.linear(self._dense, ed, self.output_layer.weight, self.output_layer.bias))
                    continue
                xs, xd = self._features(ids_s, eid_s, dt_s), self._features(ids_d, eid_d, dt_d)
                es = ed = None
                for transformer in self.transformers:                       # TCL.py:133-151
                    xs = transformer(xs, xs, xs, ids_s)
                    xd = transformer(xd, xd, xd, ids_d)
                    es = transformer(xs, xd, xd, ids_d)
                    ed = transformer(xd, xs, xs, ids_s)
                    xs, xd = es, ed
                outs_s.append(self.output_layer(es[:, 0, :]))
                outs_d.append(self.output_layer(ed[:, 0, :]))
        if len(outs_s) == 1:
            return outs_s[0], outs_d[0]
        return torch.cat(outs_s, dim=0), torch.cat(outs_d, dim=0)

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """TCL.py:192-202."""
        self.neighbor_sampler = neighbor_sampler
        if self.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.neighbor_sampler.seed is not None
            self.neighbor_sampler.reset_random_state()
