"""TEST INFRASTRUCTURE ONLY -- torch-fp32 CPU restatement of FLiD's TGAT forward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this; the product never does.

Parity status: PINNED against the reference itself (``models/TGAT.py`` and
``models/modules.py`` imported in the build container by
``tests/golden/make_golden.py``; outputs committed under ``tests/golden/``; the
live reference is also compared when present).  The reference has no tests.

Floating-point kernel => the restatement is a plain torch fp32 reference, in
the *literal* operation order of the reference (per-neighbour K/V projections,
einsum scores, masked_fill(-1e10), softmax, residual_fc, LayerNorm, MergeLayer):

* ``time_encode``   -- ``models/modules.py:28-40``  cos(Linear(1->T)(dt))
* ``attention``     -- ``models/modules.py:167-245`` (eval mode: dropout = id)
* ``merge``         -- ``models/modules.py:58-69``
* ``embed``         -- ``models/TGAT.py:68-144`` recursion, incl. its dtype rules:
  root times are float64, recursion times are the sampler's float32 output;
  ``dt = t[:, None] - nbr_t`` is float64 at the root hop and float32 below
  (numpy promotion), then ``.float()``.
* ``embed_src_dst`` -- ``models/TGAT.py:50-66``

Weights are a flat dict with the reference's ``state_dict`` key names
(``time_encoder.w.weight`` ...), so a reference checkpoint drops in.
"""
import numpy as np
import torch
import torch.nn.functional as F


def default_params(node_dim, edge_dim, time_dim, num_layers, num_heads=2, seed=0, time_bias_scale=0.0):
    """Weights with the reference's shapes and init laws (nn.Linear default init,
    TimeEncoder w = 1/10^linspace(0,9,T), LayerNorm 1/0).  ``time_bias_scale`` > 0
    perturbs the time-encoder bias so the fma rounding path is exercised."""
    g = torch.Generator().manual_seed(seed)
    qd, kd = node_dim + time_dim, node_dim + edge_dim + time_dim

    def lin(out_f, in_f, bias=True):
        bound = 1.0 / np.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound if bias else None
        return w, b

    p = {}
    p["time_encoder.w.weight"] = torch.from_numpy(
        1 / 10 ** np.linspace(0, 9, time_dim, dtype=np.float32)).reshape(time_dim, 1)
    p["time_encoder.w.bias"] = (torch.rand(time_dim, generator=g) * 2 - 1) * time_bias_scale
    for l in range(num_layers):
        a = f"temporal_conv_layers.{l}."
        p[a + "query_projection.weight"], _ = lin(qd, qd, False)
        p[a + "key_projection.weight"], _ = lin(qd, kd, False)
        p[a + "value_projection.weight"], _ = lin(qd, kd, False)
        p[a + "layer_norm.weight"] = 1.0 + 0.1 * (torch.rand(qd, generator=g) * 2 - 1)
        p[a + "layer_norm.bias"] = 0.1 * (torch.rand(qd, generator=g) * 2 - 1)
        p[a + "residual_fc.weight"], p[a + "residual_fc.bias"] = lin(qd, qd)
        m = f"merge_layers.{l}."
        p[m + "fc1.weight"], p[m + "fc1.bias"] = lin(node_dim, qd + node_dim)
        p[m + "fc2.weight"], p[m + "fc2.bias"] = lin(node_dim, node_dim)
    p["_num_heads"] = num_heads
    return p


def time_encode(p, dt):
    """dt float32 [n, s] -> float32 [n, s, T]; the K=1 Linear is a fused multiply-add."""
    return torch.cos(F.linear(dt.unsqueeze(2), p["time_encoder.w.weight"], p["time_encoder.w.bias"]))


def attention(p, layer, h_self, te_self, h_nbr, te_nbr, e_nbr, nbr_ids, prefix=""):
    a = f"{prefix}temporal_conv_layers.{layer}."
    heads = p["_num_heads"]
    n, k = h_nbr.shape[0], h_nbr.shape[1]
    q_in = torch.cat([h_self.unsqueeze(1), te_self], dim=2)              # [n,1,qd]
    kv_in = torch.cat([h_nbr, e_nbr, te_nbr], dim=2)                      # [n,k,kd]
    hd = q_in.shape[2] // heads
    q = F.linear(q_in, p[a + "query_projection.weight"]).reshape(n, 1, heads, hd).permute(0, 2, 1, 3)
    key = F.linear(kv_in, p[a + "key_projection.weight"]).reshape(n, k, heads, hd).permute(0, 2, 1, 3)
    val = F.linear(kv_in, p[a + "value_projection.weight"]).reshape(n, k, heads, hd).permute(0, 2, 1, 3)
    s = torch.einsum("bhld,bhnd->bhln", q, key) * (hd ** -0.5)
    pad = (torch.from_numpy(np.ascontiguousarray(nbr_ids)) == 0).reshape(n, 1, 1, k).expand(n, heads, 1, k)
    s = s.masked_fill(pad, -1e10)
    w = torch.softmax(s, dim=-1)
    o = torch.einsum("bhln,bhnd->bhld", w, val).permute(0, 2, 1, 3).flatten(start_dim=2)
    o = F.linear(o, p[a + "residual_fc.weight"], p[a + "residual_fc.bias"])
    o = F.layer_norm(o + q_in, (q_in.shape[2],), p[a + "layer_norm.weight"], p[a + "layer_norm.bias"], 1e-5)
    return o.squeeze(1), w.squeeze(2)


def merge(p, layer, x1, x2, prefix=""):
    m = f"{prefix}merge_layers.{layer}."
    h = F.relu(F.linear(torch.cat([x1, x2], dim=1), p[m + "fc1.weight"], p[m + "fc1.bias"]))
    return F.linear(h, p[m + "fc2.weight"], p[m + "fc2.bias"])


def embed(p, node_feat, edge_feat, sampler, node_ids, times, layer, k, prefix="", layer0=None):
    """models/TGAT.py:68-144.  ``layer0`` (optional) overrides the layer-0 / merge
    second input table (TGN uses memory + raw, models/MemoryModel.py:654-658)."""
    base = node_feat if layer0 is None else layer0
    ids_t = torch.from_numpy(np.ascontiguousarray(node_ids).astype(np.int64))
    raw = base[ids_t]
    if layer == 0:
        return raw
    n = len(node_ids)
    te_self = time_encode(p, torch.zeros(n, 1))
    h_self = embed(p, node_feat, edge_feat, sampler, node_ids, times, layer - 1, k, prefix, layer0)
    nbr, eid, nts = sampler.get_historical_neighbors(node_ids, times, k)
    h_nbr = embed(p, node_feat, edge_feat, sampler, nbr.flatten(), nts.flatten(), layer - 1, k, prefix, layer0)
    h_nbr = h_nbr.reshape(n, k, -1)
    dt = np.asarray(times)[:, None] - nts            # float64 at the root hop, float32 below
    te_nbr = time_encode(p, torch.from_numpy(dt).float())
    e_nbr = edge_feat[torch.from_numpy(eid)]
    out, _ = attention(p, layer - 1, h_self, te_self, h_nbr, te_nbr, e_nbr, nbr, prefix)
    return merge(p, layer - 1, out, raw, prefix)


def embed_src_dst(p, node_feat, edge_feat, sampler, src, dst, times, num_layers, k):
    with torch.no_grad():
        return (embed(p, node_feat, edge_feat, sampler, src, times, num_layers, k),
                embed(p, node_feat, edge_feat, sampler, dst, times, num_layers, k))
