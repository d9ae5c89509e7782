"""Synthetic continuous-time dynamic graphs of the shapes named in BASELINE.json.

Host-side numpy only; nothing here is on the hot path.  The generators follow
SURVEY.md section 8(d): ``np.random.RandomState(seed)``, events sorted by time,
node 0 / edge 0 reserved as the zero padding rows, ``edge_ids = 1..E``.  The
event record mirrors the reference's ``Data`` (``utils/DataLoader.py:46-65``).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class TemporalGraph:
    """src/dst/eid int64[E], ts float64[E]; feature tables include padding row 0."""
    name: str
    src_node_ids: np.ndarray
    dst_node_ids: np.ndarray
    node_interact_times: np.ndarray
    edge_ids: np.ndarray
    node_raw_features: np.ndarray  # float32 [N+1, dn]
    edge_raw_features: np.ndarray  # float32 [E+1, de]
    num_neighbors: int = 20

    @property
    def num_interactions(self):
        return len(self.src_node_ids)

    @property
    def num_nodes(self):
        return self.node_raw_features.shape[0] - 1


def _bounded_zipf(rs, n_items, exponent, size):
    """Draw ``size`` ranks in [0, n_items) with P(i) ~ 1/(i+1)^exponent."""
    p = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), exponent)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rs.random_sample(size), side="left").astype(np.int64)


def _features(rs, rows, dim, zeros):
    if zeros:
        return np.zeros((rows, dim), dtype=np.float32)
    f = rs.standard_normal((rows, dim)).astype(np.float32)
    f[0] = 0.0
    return f


def bipartite_graph(name, num_users, num_items, num_edges, t_max, seed=0, dim=172,
                    node_zeros=True, user_exp=0.8, item_exp=1.0, num_neighbors=20,
                    integral_times=True):
    """Wikipedia/Reddit-shaped interaction stream (users 1..U, items U+1..U+I)."""
    rs = np.random.RandomState(seed)
    users = 1 + rs.permutation(num_users)[_bounded_zipf(rs, num_users, user_exp, num_edges)]
    items = 1 + num_users + rs.permutation(num_items)[_bounded_zipf(rs, num_items, item_exp, num_edges)]
    ts = rs.uniform(0.0, float(t_max), num_edges)
    if integral_times:
        ts = np.floor(ts)
    ts = np.sort(ts).astype(np.float64)
    node_feat = _features(rs, num_users + num_items + 1, dim, node_zeros)
    edge_feat = _features(rs, num_edges + 1, dim, False)
    return TemporalGraph(name, users.astype(np.int64), items.astype(np.int64), ts,
                         np.arange(1, num_edges + 1, dtype=np.int64), node_feat, edge_feat, num_neighbors)


def general_graph(name, num_nodes, num_edges, t_max, seed=0, dim=172, node_zeros=False,
                  exponent=0.0, num_neighbors=20, integral_times=True, with_features=True):
    """Non-bipartite stream; ``exponent`` 0 gives uniform endpoints."""
    rs = np.random.RandomState(seed)
    if exponent > 0:
        perm = rs.permutation(num_nodes)
        src = 1 + perm[_bounded_zipf(rs, num_nodes, exponent, num_edges)]
        dst = 1 + perm[_bounded_zipf(rs, num_nodes, exponent, num_edges)]
    else:
        src = rs.randint(1, num_nodes + 1, num_edges)
        dst = rs.randint(1, num_nodes + 1, num_edges)
    ts = rs.uniform(0.0, float(t_max), num_edges)
    if integral_times:
        ts = np.floor(ts)
    ts = np.sort(ts).astype(np.float64)
    if with_features:
        node_feat = _features(rs, num_nodes + 1, dim, node_zeros)
        edge_feat = _features(rs, num_edges + 1, dim, False)
    else:  # caller fills the tables on the device (scaling config)
        node_feat = np.zeros((num_nodes + 1, 0), dtype=np.float32)
        edge_feat = np.zeros((num_edges + 1, 0), dtype=np.float32)
    return TemporalGraph(name, src.astype(np.int64), dst.astype(np.int64), ts,
                         np.arange(1, num_edges + 1, dtype=np.int64), node_feat, edge_feat, num_neighbors)


def wikipedia_shape(seed=0, scale=1.0):
    """BASELINE.json configs[0]/[1]: 9 227 nodes / 157 474 edges, node feats all zero."""
    e = max(200, int(157474 * scale))
    return bipartite_graph("wikipedia-shape", max(8, int(8227 * scale)), max(4, int(1000 * scale)), e,
                           2678373, seed=seed)


def reddit_shape(seed=0, scale=1.0):
    """BASELINE.json configs[2]: 10 984 nodes / 672 447 edges, d=172."""
    e = max(200, int(672447 * scale))
    return bipartite_graph("reddit-shape", max(8, int(10000 * scale)), max(4, int(984 * scale)), e,
                           2678390, seed=seed)


def dsub_shape(seed=0, scale=1.0):
    """BASELINE.json configs[3]: 150 000 nodes / 168 154 edges, k=30, one year of seconds
    (3.15e7 > 2^24, so float32 rounding of hop-1 times is exercised)."""
    return general_graph("dsub-shape", max(16, int(150000 * scale)), max(200, int(168154 * scale)),
                         31536000, seed=seed, node_zeros=False, num_neighbors=30)


def scaling_shape(seed=0, num_nodes=1_000_000, num_edges=50_000_000):
    """BASELINE.json configs[4] topology only (features are filled on the device)."""
    return general_graph("scaling-shape", num_nodes, num_edges, 1e8, seed=seed, exponent=0.8,
                         integral_times=False, with_features=False)
