"""Drop-in for the reference's ``models.GraphMixer.GraphMixer`` (``models/GraphMixer.py``): same
constructor, method names and ``state_dict`` keys.  SURVEY.md section 8(f) rank 4 -- another
consumer of the time-sorted device CSR:

* the two ``get_historical_neighbors`` calls per batch (k recent neighbours for the link encoder,
  ``time_gap`` = 2000 recent neighbours for the node encoder, GraphMixer.py:91-96, :119-123) never
  leave the device;
* the node encoder (gather of up to 2000 feature rows per query, masked softmax, mean;
  GraphMixer.py:119-146) is one kernel, ``flid_neighbor_mean`` (csrc/mixer.cu) -- the reference
  materialises a ``[B, 2000, dn]`` tensor for it;
* the link encoder's dense part (Linear(T, 100), MLP-Mixer blocks, output layer) runs, in evaluation
  (``model.eval()`` under ``torch.no_grad()``), on the library's own kernels (csrc/dense.cu): every Linear on the
  tcgen05 3xTF32 GEMM with bias / GELU / residual in its epilogue, token mixing as one kernel per block;
* in training the same modules are composed from torch CUDA ops, so it trains with autograd as the reference
  does (the time encoder is frozen in GraphMixer, the sampler outputs and raw node features are constants:
  nothing on the kernel side needs a gradient).
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib, dense
from .sampler import NeighborSampler
from .tgat import TimeEncoder


class FeedForwardNet(nn.Module):
    """Linear -> GELU -> Dropout -> Linear -> Dropout, parameters under ``ffn.0`` / ``ffn.3`` (GraphMixer.py:172-196)."""

    def __init__(self, input_dim: int, dim_expansion_factor: float, dropout: float = 0.0):
        super().__init__()
        hidden = int(dim_expansion_factor * input_dim)
        self.ffn = nn.Sequential(nn.Linear(input_dim, hidden), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden, input_dim), nn.Dropout(dropout))

    def forward(self, x):
        return self.ffn(x)


class MLPMixer(nn.Module):
    """Token mixing then channel mixing with residuals (GraphMixer.py:199-246)."""

    def __init__(self, num_tokens: int, num_channels: int, token_dim_expansion_factor: float = 0.5,
                 channel_dim_expansion_factor: float = 4.0, dropout: float = 0.0):
        super().__init__()
        self.token_norm = nn.LayerNorm(num_tokens)
        self.token_feedforward = FeedForwardNet(num_tokens, token_dim_expansion_factor, dropout)
        self.channel_norm = nn.LayerNorm(num_channels)
        self.channel_feedforward = FeedForwardNet(num_channels, channel_dim_expansion_factor, dropout)

    def forward(self, input_tensor):
        hidden = self.token_feedforward(self.token_norm(input_tensor.permute(0, 2, 1))).permute(0, 2, 1)
        out = hidden + input_tensor
        return self.channel_feedforward(self.channel_norm(out)) + out


class GraphMixer(nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, num_tokens: int, num_layers: int = 2, token_dim_expansion_factor: float = 0.5,
                 channel_dim_expansion_factor: float = 4.0, dropout: float = 0.1, device: str = 'cpu'):
        super().__init__()
        self.device = device                  # compute calls require CUDA (no CPU fallback); construction does not
        self.node_raw_features = torch.from_numpy(np.ascontiguousarray(node_raw_features, dtype=np.float32)).to(device)
        self.neighbor_sampler = neighbor_sampler
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.num_tokens = num_tokens
        self.num_layers = num_layers
        self.token_dim_expansion_factor = token_dim_expansion_factor
        self.channel_dim_expansion_factor = channel_dim_expansion_factor
        self.dropout = dropout
        self.num_channels = 100
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim, parameter_requires_grad=False)   # frozen, GraphMixer.py:43-45
        self.projection_layer = nn.Linear(time_feat_dim, self.num_channels)
        self.mlp_mixers = nn.ModuleList([MLPMixer(num_tokens, self.num_channels, token_dim_expansion_factor,
                                                  channel_dim_expansion_factor, dropout) for _ in range(num_layers)])
        self.output_layer = nn.Linear(self.num_channels + self.node_feat_dim, self.node_feat_dim, bias=True)
        self.chunk_queries = 65536          # bulk calls are processed in chunks: [chunk, k, 100] activations
        self._dense = dense.DenseWeights()  # tiled weight images of the evaluation path

    def _link_encoder_eval(self, dt, nbr, m, k):
        """GraphMixer.py:103-117 on csrc/dense.cu: [m, k] time differences / neighbour ids -> [m, 100]."""
        lib, st, c = _lib.lib(), _lib.stream(), self.num_channels
        te = dense.time_rows(dt, nbr, self.time_encoder)                                   # [m * k, T], padded rows zero
        x = dense.linear(self._dense, te, self.projection_layer.weight, self.projection_layer.bias)
        for mixer in self.mlp_mixers:
            tf, cf = mixer.token_feedforward.ffn, mixer.channel_feedforward.ffn
            x1 = torch.empty_like(x)
            _lib.check(lib.flid_token_mix(_lib.ptr(x), k, c, _lib.ptr(mixer.token_norm.weight), _lib.ptr(mixer.token_norm.bias),
                                          float(mixer.token_norm.eps), _lib.ptr(tf[0].weight), _lib.ptr(tf[0].bias),
                                          _lib.ptr(tf[3].weight), _lib.ptr(tf[3].bias), tf[0].weight.shape[0], _lib.ptr(x1), m, st))
            y = dense.layernorm(x1, mixer.channel_norm)
            h = dense.linear(self._dense, y, cf[0].weight, cf[0].bias, act=2)
            x = dense.linear(self._dense, h, cf[3].weight, cf[3].bias, resid=x1)
        link = torch.empty((m, c), dtype=torch.float32, device=x.device)
        _lib.check(lib.flid_token_mean(_lib.ptr(x), k, c, _lib.ptr(link), c, m, st))
        return link

    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray,
                                                 node_interact_times: np.ndarray, num_neighbors: int = 20,
                                                 time_gap: int = 2000):
        """GraphMixer.py:60-78."""
        b = len(src_node_ids)
        both = self.compute_node_temporal_embeddings(
            np.concatenate([np.asarray(src_node_ids), np.asarray(dst_node_ids)]),
            np.concatenate([np.asarray(node_interact_times), np.asarray(node_interact_times)]), num_neighbors, time_gap)
        return both[:b], both[b:]

    def compute_node_temporal_embeddings(self, node_ids: np.ndarray, node_interact_times: np.ndarray,
                                         num_neighbors: int = 20, time_gap: int = 2000):
        """GraphMixer.py:80-153: float32 [n, node_feat_dim] on the device."""
        sampler = self.neighbor_sampler
        if not isinstance(sampler, NeighborSampler):
            raise TypeError(f"flid_b200 models need a flid_b200.NeighborSampler (device CSR); got {type(sampler).__name__}")
        k = int(num_neighbors)
        assert k > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        assert int(time_gap) > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        ids = np.ascontiguousarray(node_ids, dtype=np.int64)
        t_np = np.asarray(node_interact_times)
        n = ids.shape[0]
        if n and (int(ids.min()) < 0 or int(ids.max()) > sampler.num_nodes):
            raise IndexError("flid_b200.GraphMixer: node id outside the graph")
        dev = _lib.require_cuda(self.node_raw_features.device)
        w_t, b_t = self.time_encoder.w.weight.reshape(-1), self.time_encoder.w.bias
        outs = []
        with torch.cuda.device(dev):
            d_ids = _lib.to_device(ids, np.int64, dev, "gm_ids")
            d_t = _lib.to_device(t_np, np.float64, dev, "gm_times")          # float32 -> float64 is exact
            for lo in range(0, max(n, 1), self.chunk_queries):
                c_ids, c_t = d_ids[lo:lo + self.chunk_queries], d_t[lo:lo + self.chunk_queries]
                m = c_ids.shape[0]
                # link encoder (GraphMixer.py:91-117)
                nbr, _, ts = sampler.get_historical_neighbors_device(c_ids, c_t, k)
                if t_np.dtype == np.float32:      # numpy: float32 - float32 stays float32 (GraphMixer.py:103-104)
                    dt = c_t.to(torch.float32)[:, None] - ts
                else:                             # float64 minus float32 in float64, then .float()
                    dt = (c_t[:, None] - ts.to(torch.float64)).to(torch.float32)
                # the forward-only kernels cover up to 64 tokens / 256 token-mixing hidden units; wider mixers keep the modules
                fast = (dense.fast_path(self) and m > 0 and k <= 64
                        and all(mx.token_feedforward.ffn[0].weight.shape[0] <= 256 for mx in self.mlp_mixers))
                if fast:
                    link = self._link_encoder_eval(dt, nbr, m, k)
                else:
                    te = torch.cos(torch.addcmul(b_t, dt.unsqueeze(-1), w_t))           # single-rounded fma, as nn.Linear(1, T)
                    te = te.masked_fill((nbr == 0).unsqueeze(-1), 0.0)
                    x = self.projection_layer(te)
                    for mixer in self.mlp_mixers:
                        x = mixer(x)
                    link = torch.mean(x, dim=1)
                # node encoder (GraphMixer.py:119-146): one gather / reduce kernel
                node_part = torch.empty((m, self.node_feat_dim), dtype=torch.float32, device=dev)
                _lib.check(_lib.lib().flid_neighbor_mean(sampler.handle, _lib.ptr(self.node_raw_features), self.node_feat_dim,
                                                         _lib.ptr(c_ids), _lib.ptr(c_t), 0, m, int(time_gap), 1,
                                                         _lib.ptr(node_part), _lib.stream()))
                if fast:
                    outs.append(dense.linear(self._dense, link, self.output_layer.weight, self.output_layer.bias, x2=node_part))
                else:
                    outs.append(self.output_layer(torch.cat([link, node_part], dim=1)))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """GraphMixer.py:155-165."""
        self.neighbor_sampler = neighbor_sampler
        if self.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.neighbor_sampler.seed is not None
            self.neighbor_sampler.reset_random_state()
