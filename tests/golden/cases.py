"""Deterministic inputs shared by ``make_golden.py`` (which runs the real FLiD
reference on them) and by the tests (which run the oracle / the CUDA path on
them).  Everything is derived from ``np.random.RandomState`` (frozen stream) or
``torch.Generator`` seeds, so the ``.npz`` fixtures only need to hold the
reference's *outputs* plus a checksum of the inputs.
"""
import numpy as np


def adversarial_events():
    """Small stream with: time ties (incl. across both endpoints), unsorted input
    order, a self-loop, isolated ids, deg << k and deg >> k nodes, timestamps that
    are not float32-representable (16777217, 3.15e7+0.3) and fractional ones."""
    rs = np.random.RandomState(1234)
    src, dst, ts = [], [], []
    # hub node 1 talks to everyone many times (deg >> k), with heavy ties
    for i in range(120):
        src.append(1), dst.append(2 + (i % 9)), ts.append(float(10 + i // 4))
    # sparse nodes (deg 1..3)
    for v in range(12, 20):
        for j in range(v % 3 + 1):
            src.append(v), dst.append(v + 10), ts.append(float(100 + 7 * j + v))
    # self loop and repeated pair at identical time
    src += [5, 5, 6, 6]; dst += [5, 6, 5, 6]; ts += [55.0, 55.0, 55.0, 55.0]
    # large / non-f32-representable times
    big = [16777216.0, 16777217.0, 16777218.0, 16777219.0, 31536000.3, 31536001.7, 99999999.0]
    for i, t in enumerate(big):
        src.append(40 + (i % 2)), dst.append(42), ts.append(t)
        src.append(42), dst.append(43), ts.append(t)
    # fractional
    for i in range(30):
        src.append(50 + rs.randint(0, 5)), dst.append(56 + rs.randint(0, 3)), ts.append(float(rs.uniform(0, 50)))
    src, dst, ts = np.array(src, dtype=np.int64), np.array(dst, dtype=np.int64), np.array(ts, dtype=np.float64)
    perm = rs.permutation(len(src))          # unsorted input order
    src, dst, ts = src[perm], dst[perm], ts[perm]
    eid = np.arange(1, len(src) + 1, dtype=np.int64)
    num_nodes = 64                             # ids 59..64 never appear (isolated)
    return src, dst, eid, ts, num_nodes


def adversarial_queries():
    src, dst, eid, ts, n = adversarial_events()
    rs = np.random.RandomState(99)
    nodes = np.concatenate([src, dst, np.arange(0, n + 1), np.full(40, 1), np.full(20, 42)]).astype(np.int64)
    times = np.concatenate([ts, ts, np.full(n + 1, 1e9), rs.uniform(0, 45, 40),
                            np.array([16777216.0, 16777217.0, 16777218.0, 16777219.0, 16777220.0] * 4)])
    # exact tie queries, just-above queries and t=0
    nodes = np.concatenate([nodes, src[:50], src[:50], src[:50]])
    times = np.concatenate([times, ts[:50], np.nextafter(ts[:50], np.inf), np.zeros(50)])
    return nodes, times.astype(np.float64)


def small_stream(num_nodes=40, num_edges=500, dim=172, seed=7, t_max=3.0e7, node_zeros=False):
    """Chronological Zipf-ish stream used for the TGAT / TGN golden cases."""
    rs = np.random.RandomState(seed)
    p = 1.0 / np.arange(1, num_nodes + 1) ** 0.9
    p /= p.sum()
    src = 1 + rs.choice(num_nodes, num_edges, p=p)
    dst = 1 + rs.choice(num_nodes, num_edges, p=p)
    ts = np.sort(rs.uniform(0, t_max, num_edges))
    ts[::3] = np.floor(ts[::3])                # mix of integral and fractional times
    ts = np.sort(ts)
    eid = np.arange(1, num_edges + 1, dtype=np.int64)
    node_feat = np.zeros((num_nodes + 1, dim), np.float32) if node_zeros else \
        rs.standard_normal((num_nodes + 1, dim)).astype(np.float32)
    node_feat[0] = 0
    edge_feat = rs.standard_normal((num_edges + 1, dim)).astype(np.float32)
    edge_feat[0] = 0
    return src.astype(np.int64), dst.astype(np.int64), eid, ts.astype(np.float64), node_feat, edge_feat


def checksum(*arrays):
    """Order-sensitive float64 checksum used to make sure inputs were regenerated identically."""
    acc = 0.0
    for a in arrays:
        a = np.asarray(a, dtype=np.float64).ravel()
        acc += float(np.dot(a, np.cos(np.arange(a.size) * 0.37)))
    return np.float64(acc)
