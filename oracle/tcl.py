"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's TCL forward (``models/TCL.py``,
``models/modules.py:248-312`` TransformerEncoder) in its literal operation order, as pure functions over a
parameter dict with the reference's ``state_dict`` names.  Pinned against the live reference by
``tests/golden/make_golden.py`` (``tests/golden/tcl.npz``).  Nothing in ``flid_b200/`` imports this module.

  sequences   TCL.py:75-106   [node itself ; its k recent neighbours]: ids, edge ids (0 first), times
  features    TCL.py:108-131, :167-190   Linear(raw node) + Linear(raw edge) + Linear(time encoding) + depth embedding
  encoder     TCL.py:133-151  per layer: self-attention on each side, then cross-attention src<-dst and dst<-src
  output      TCL.py:153-157  Linear on position 0
"""
import numpy as np
import torch
import torch.nn.functional as F


def default_params(node_dim, edge_dim, time_dim, num_layers=2, num_depths=21, seed=0, time_bias_scale=0.0):
    g = torch.Generator().manual_seed(seed)

    def linear(out_f, in_f):
        bound = 1.0 / np.sqrt(in_f)
        return ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound, (torch.rand(out_f, generator=g) * 2 - 1) * bound)

    p = {}
    p["time_encoder.w.weight"] = torch.from_numpy(1 / 10 ** np.linspace(0, 9, time_dim, dtype=np.float32)).reshape(time_dim, 1)
    p["time_encoder.w.bias"] = time_bias_scale * torch.randn(time_dim, generator=g)
    p["depth_embedding.weight"] = torch.randn(num_depths, node_dim, generator=g)
    for name, width in (("node", node_dim), ("edge", edge_dim), ("time", time_dim)):
        p[f"projection_layer.{name}.weight"], p[f"projection_layer.{name}.bias"] = linear(node_dim, width)
    for l in range(num_layers):
        pre = f"transformers.{l}."
        p[pre + "multi_head_attention.in_proj_weight"] = (torch.rand(3 * node_dim, node_dim, generator=g) * 2 - 1) * np.sqrt(6.0 / (4 * node_dim))
        p[pre + "multi_head_attention.in_proj_bias"] = 0.05 * torch.randn(3 * node_dim, generator=g)
        p[pre + "multi_head_attention.out_proj.weight"], p[pre + "multi_head_attention.out_proj.bias"] = linear(node_dim, node_dim)
        p[pre + "linear_layers.0.weight"], p[pre + "linear_layers.0.bias"] = linear(4 * node_dim, node_dim)
        p[pre + "linear_layers.1.weight"], p[pre + "linear_layers.1.bias"] = linear(node_dim, 4 * node_dim)
        for j in (0, 1):
            p[pre + f"norm_layers.{j}.weight"] = 1.0 + 0.1 * torch.randn(node_dim, generator=g)
            p[pre + f"norm_layers.{j}.bias"] = 0.1 * torch.randn(node_dim, generator=g)
    p["output_layer.weight"], p["output_layer.bias"] = linear(node_dim, node_dim)
    return p


def transformer(p, l, num_heads, q, k, v, key_ids):
    """models/modules.py:276-312 (dropout = identity in eval mode); inputs [B, S, d], key_ids ndarray [B, S]."""
    pre = f"transformers.{l}."
    d = q.shape[2]
    mask = torch.from_numpy(key_ids == 0)
    h, _ = F.multi_head_attention_forward(
        q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), d, num_heads,
        p[pre + "multi_head_attention.in_proj_weight"], p[pre + "multi_head_attention.in_proj_bias"], None, None, False, 0.0,
        p[pre + "multi_head_attention.out_proj.weight"], p[pre + "multi_head_attention.out_proj.bias"], training=False,
        key_padding_mask=mask, need_weights=False)
    out = F.layer_norm(q + h.transpose(0, 1), (d,), p[pre + "norm_layers.0.weight"], p[pre + "norm_layers.0.bias"])
    h = F.linear(F.relu(F.linear(out, p[pre + "linear_layers.0.weight"], p[pre + "linear_layers.0.bias"])),
                 p[pre + "linear_layers.1.weight"], p[pre + "linear_layers.1.bias"])
    return F.layer_norm(out + h, (d,), p[pre + "norm_layers.1.weight"], p[pre + "norm_layers.1.bias"])


def _side(p, node_feat, edge_feat, sampler, ids, times, k):
    nbr, eid, nts = sampler.get_historical_neighbors(ids, times, k)
    seq_ids = np.concatenate((ids[:, np.newaxis], nbr), axis=1)
    seq_eid = np.concatenate((np.zeros((len(ids), 1)).astype(np.longlong), eid), axis=1)
    seq_t = np.concatenate((times[:, np.newaxis], nts), axis=1)
    dt = torch.from_numpy(times[:, np.newaxis] - seq_t).float()
    te = torch.cos(F.linear(dt.unsqueeze(2), p["time_encoder.w.weight"], p["time_encoder.w.bias"]))
    assert seq_ids.shape[1] == p["depth_embedding.weight"].shape[0]
    x = (F.linear(node_feat[torch.from_numpy(seq_ids)], p["projection_layer.node.weight"], p["projection_layer.node.bias"])
         + F.linear(edge_feat[torch.from_numpy(seq_eid)], p["projection_layer.edge.weight"], p["projection_layer.edge.bias"])
         + F.linear(te, p["projection_layer.time.weight"], p["projection_layer.time.bias"])
         + p["depth_embedding.weight"][torch.arange(seq_ids.shape[1])])
    return x, seq_ids


def embed_src_dst(p, node_feat, edge_feat, sampler, src, dst, times, num_layers, num_heads=2, k=20):
    """TCL.compute_src_dst_node_temporal_embeddings (TCL.py:60-157)."""
    src, dst, times = np.asarray(src), np.asarray(dst), np.asarray(times)
    xs, ids_s = _side(p, node_feat, edge_feat, sampler, src, times, k)
    xd, ids_d = _side(p, node_feat, edge_feat, sampler, dst, times, k)
    es = ed = None
    for l in range(num_layers):
        xs = transformer(p, l, num_heads, xs, xs, xs, ids_s)
        xd = transformer(p, l, num_heads, xd, xd, xd, ids_d)
        es = transformer(p, l, num_heads, xs, xd, xd, ids_d)
        ed = transformer(p, l, num_heads, xd, xs, xs, ids_s)
        xs, xd = es, ed
    out = lambda e: F.linear(e[:, 0, :], p["output_layer.weight"], p["output_layer.bias"])
    return out(es), out(ed)
