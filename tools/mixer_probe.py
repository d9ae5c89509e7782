"""GraphMixer on the Reddit-shape graph: the node-encoder kernel (flid_neighbor_mean, time_gap = 2000) alone --
queries/s and gathered bytes/s -- and one full bulk pass (both endpoints of every event) through the drop-in class."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flid_b200                                   # noqa: E402
from flid_b200 import _lib, synth                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--gap", type=int, default=2000)
    a = ap.parse_args()
    dev = "cuda:0"
    g = synth.reddit_shape(seed=0, scale=a.scale)
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    rs = np.random.RandomState(0)
    nf = rs.standard_normal(g.node_raw_features.shape).astype(np.float32)
    nf[0] = 0
    nodes = np.concatenate([g.src_node_ids, g.dst_node_ids])
    times = np.concatenate([g.node_interact_times, g.node_interact_times])
    n = len(nodes)
    d_nf, d_ids, d_t = torch.from_numpy(nf).to(dev), torch.from_numpy(nodes).to(dev), torch.from_numpy(times).to(dev)
    out = torch.empty((n, nf.shape[1]), dtype=torch.float32, device=dev)

    def run():
        _lib.check(_lib.lib().flid_neighbor_mean(s.handle, _lib.ptr(d_nf), nf.shape[1], _lib.ptr(d_ids), _lib.ptr(d_t), 0, n,
                                                 a.gap, 1, _lib.ptr(out), _lib.stream()))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    # rows gathered: min(gap, history length) per query, from the host CSR
    indptr, _, _, ts = s._host_csr()
    order = np.argsort(nodes, kind="stable")
    rows = 0
    for v in np.unique(nodes):
        q = times[nodes == v]
        cut = np.searchsorted(ts[indptr[v]:indptr[v + 1]], q, side="left")
        rows += int(np.minimum(cut, a.gap).sum())
    gb = rows * nf.shape[1] * 4 / 1e9
    print(f"neighbor_mean gap={a.gap}: {n} queries, {rows / n:.1f} rows/query, {ms:.2f} ms, {n / ms / 1e3:.2f} M queries/s, "
          f"{gb / (ms / 1e3):.0f} GB/s of gathered rows (feature table {nf.nbytes / 1e6:.1f} MB: L2-resident)")
    m = flid_b200.GraphMixer(nf, g.edge_raw_features, s, 100, 20, 2, device=dev).to(dev)
    m.eval()
    with torch.no_grad():
        m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[:4000], g.dst_node_ids[:4000], g.node_interact_times[:4000])
        torch.cuda.synchronize()
        e0.record()
        m.compute_src_dst_node_temporal_embeddings(g.src_node_ids, g.dst_node_ids, g.node_interact_times, 20, a.gap)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"GraphMixer L=2 k=20 gap={a.gap} bulk pass: {n} root queries in {ms:.1f} ms, {n / ms / 1e3:.2f} M root queries/s")


if __name__ == "__main__":
    main()
