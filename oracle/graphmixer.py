"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's GraphMixer forward
(``models/GraphMixer.py``), in its literal operation order, as pure functions over a parameter
dict with the reference's ``state_dict`` names.  Pinned against the live reference by
``tests/golden/make_golden.py`` (``tests/golden/graphmixer.npz``).  Nothing in ``flid_b200/``
imports this module.

  link encoder   GraphMixer.py:91-117   k recent neighbours -> time encodings (padded slots zeroed)
                                        -> Linear(T, 100) -> MLPMixer x L -> mean over tokens
  node encoder   GraphMixer.py:119-146  ``time_gap`` recent neighbours -> softmax over a {1, -1e10} mask
                                        -> mean_j(x_j * score_j) (divides by time_gap AGAIN) + own raw features
  output         GraphMixer.py:148-151  Linear(100 + dn, dn)
  MLPMixer       GraphMixer.py:199-246  token mixing (LayerNorm over tokens, FFN x0.5) + channel mixing (FFN x4)
"""
import numpy as np
import torch
import torch.nn.functional as F


def default_params(node_dim, time_dim, num_tokens, num_layers=2, token_factor=0.5, channel_factor=4.0, seed=0,
                   time_bias_scale=0.0):
    """Reference-shaped parameters: default nn.Linear / nn.LayerNorm initialisation after manual_seed."""
    g = torch.Generator().manual_seed(seed)
    C = 100

    def linear(out_f, in_f):
        bound = 1.0 / np.sqrt(in_f)
        return ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound, (torch.rand(out_f, generator=g) * 2 - 1) * bound)

    p = {}
    p["time_encoder.w.weight"] = torch.from_numpy(1 / 10 ** np.linspace(0, 9, time_dim, dtype=np.float32)).reshape(time_dim, 1)
    p["time_encoder.w.bias"] = time_bias_scale * torch.randn(time_dim, generator=g)
    p["projection_layer.weight"], p["projection_layer.bias"] = linear(C, time_dim)
    for l in range(num_layers):
        pre = f"mlp_mixers.{l}."
        for norm, width in (("token_norm", num_tokens), ("channel_norm", C)):
            p[pre + norm + ".weight"] = 1.0 + 0.1 * torch.randn(width, generator=g)
            p[pre + norm + ".bias"] = 0.1 * torch.randn(width, generator=g)
        for ffn, width, factor in (("token_feedforward", num_tokens, token_factor),
                                   ("channel_feedforward", C, channel_factor)):
            hidden = int(factor * width)
            p[pre + ffn + ".ffn.0.weight"], p[pre + ffn + ".ffn.0.bias"] = linear(hidden, width)
            p[pre + ffn + ".ffn.3.weight"], p[pre + ffn + ".ffn.3.bias"] = linear(width, hidden)
    p["output_layer.weight"], p["output_layer.bias"] = linear(node_dim, C + node_dim)
    return p


def _ffn(p, pre, x):
    x = F.gelu(F.linear(x, p[pre + ".ffn.0.weight"], p[pre + ".ffn.0.bias"]))
    return F.linear(x, p[pre + ".ffn.3.weight"], p[pre + ".ffn.3.bias"])


def mlp_mixer(p, l, x):
    """GraphMixer.py:227-246 (dropout = identity in eval mode)."""
    pre = f"mlp_mixers.{l}."
    h = F.layer_norm(x.permute(0, 2, 1), (x.shape[1],), p[pre + "token_norm.weight"], p[pre + "token_norm.bias"])
    h = _ffn(p, pre + "token_feedforward", h).permute(0, 2, 1)
    out = h + x
    h = F.layer_norm(out, (out.shape[2],), p[pre + "channel_norm.weight"], p[pre + "channel_norm.bias"])
    return _ffn(p, pre + "channel_feedforward", h) + out


def embed(p, node_feat, sampler, node_ids, times, num_layers, num_neighbors=20, time_gap=2000):
    """GraphMixer.compute_node_temporal_embeddings (GraphMixer.py:80-153).  ``sampler``: anything with the
    reference's ``get_historical_neighbors``; node_feat torch float32 [N+1, dn]."""
    node_ids = np.asarray(node_ids)
    times = np.asarray(times)
    nbr, _, nts = sampler.get_historical_neighbors(node_ids, times, num_neighbors)
    dt = torch.from_numpy(times[:, np.newaxis] - nts).float()
    te = torch.cos(F.linear(dt.unsqueeze(2), p["time_encoder.w.weight"], p["time_encoder.w.bias"]))
    te[torch.from_numpy(nbr == 0)] = 0.0
    x = F.linear(te, p["projection_layer.weight"], p["projection_layer.bias"])
    for l in range(num_layers):
        x = mlp_mixer(p, l, x)
    link = torch.mean(x, dim=1)
    gap_nbr, _, _ = sampler.get_historical_neighbors(node_ids, times, time_gap)
    feats = node_feat[torch.from_numpy(gap_nbr)]
    mask = torch.from_numpy((gap_nbr > 0).astype(np.float32))
    mask[mask == 0] = -1e10
    scores = torch.softmax(mask, dim=1)
    agg = torch.mean(feats * scores.unsqueeze(-1), dim=1)
    node_part = agg + node_feat[torch.from_numpy(node_ids)]
    return F.linear(torch.cat([link, node_part], dim=1), p["output_layer.weight"], p["output_layer.bias"])
