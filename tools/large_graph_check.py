#!/usr/bin/env python
"""Large-graph check (BASELINE.json configs[4] topology at a reduced size): tables of several GB, so every
row offset beyond 2^31 bytes is exercised.  The fused path (plain recursion and memoised bulk pass) is
compared with the differentiable torch-op path (torch indexing is 64-bit safe) and with itself.

    python tools/large_graph_check.py [--nodes 200000] [--edges 10000000] [--roots 200000]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flid_b200  # noqa: E402
from flid_b200 import passes, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=200_000)
    ap.add_argument("--edges", type=int, default=10_000_000)
    ap.add_argument("--roots", type=int, default=200_000)
    args = ap.parse_args()
    dev = "cuda:0"
    t0 = time.perf_counter()
    g = synth.scaling_shape(seed=0, num_nodes=args.nodes, num_edges=args.edges)
    print(f"graph: {args.nodes} nodes / {args.edges} edges generated in {time.perf_counter() - t0:.1f} s", flush=True)
    t0 = time.perf_counter()
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    torch.cuda.synchronize()
    print(f"device CSR build: {time.perf_counter() - t0:.2f} s, {s.num_entries} entries, max degree {s.max_degree}", flush=True)
    torch.manual_seed(0)
    m = flid_b200.TGAT(np.zeros((2, 172), np.float32), np.zeros((2, 172), np.float32), s, 100, 2, 2, 0.0, dev).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    m.node_raw_features = torch.randn((args.nodes + 1, 172), device=dev, generator=gen)
    m.edge_raw_features = torch.randn((args.edges + 1, 172), device=dev, generator=gen)
    m.node_raw_features[0] = 0
    m.edge_raw_features[0] = 0
    print(f"edge table {m.edge_raw_features.numel() * 4 / 2**30:.1f} GiB, memo table "
          f"{(s.num_entries + 1) * 172 * 4 / 2**30:.1f} GiB", flush=True)
    m.eval()
    e = g.num_interactions
    sel = np.arange(e - args.roots // 2, e)
    nodes = np.concatenate([g.src_node_ids[sel], g.dst_node_ids[sel]])
    times = np.concatenate([g.node_interact_times[sel], g.node_interact_times[sel]])
    with torch.no_grad():
        for rep in range(2):
            m._engine.memo.clear()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.build_layer_memo(20)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            out = m.compute_node_temporal_embeddings(nodes, times, 2, 20)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        print(f"memo build ({s.num_entries + 1} rows): {(t1 - t0) * 1e3:.1f} ms = {(s.num_entries + 1) / (t1 - t0) / 1e6:.1f} M evaluations/s; "
              f"{len(nodes)} roots: {(t2 - t1) * 1e3:.1f} ms", flush=True)
        assert torch.isfinite(out).all()
        sub = np.random.RandomState(0).choice(len(nodes), 1500, replace=False)
        m.set_layer_memo(False)
        plain = m.compute_node_temporal_embeddings(nodes[sub], times[sub], 2, 20)
        m.set_layer_memo("auto")
    diff = float((plain - out[sub]).abs().max())
    same = diff <= 2e-5 * max(1.0, float(plain.abs().max()))
    print(f"memoised (projected) pass vs recursion: max |diff| = {diff:.2e}", flush=True)
    # independent arithmetic: the differentiable torch-op path (64-bit safe indexing)
    few = sub[:256]
    m.train()
    ref = m.compute_node_temporal_embeddings(nodes[few], times[few], 2, 20).detach()
    m.eval()
    err = float((ref - out[few]).abs().max())
    scale = float(ref.abs().max())
    print(f"fused kernels vs torch-op path on {len(few)} roots: max |diff| = {err:.2e} (max |value| {scale:.2f})", flush=True)
    assert same and err <= 1e-4 * max(1.0, scale)
    print("large-graph check ok")


if __name__ == "__main__":
    main()
