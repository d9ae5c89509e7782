#!/usr/bin/env python
"""Timings of the reference-shaped call patterns (development tool, not bench.py):

  * TGN (configs[1]): chronological pass over the Wikipedia-shape graph in batches of 200
    through MemoryModel.compute_src_dst_node_temporal_embeddings (sequential memory updates);
  * TGAT per-batch drop-in loop: the reference's E/200-iteration loop (PTCL/M_step.py:454-509)
    calling compute_src_dst_node_temporal_embeddings with host numpy batches of 200, layer memo
    in "auto" mode, on the Reddit-shape graph.

    python tools/pass_probe.py [--scale 1.0]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flid_b200  # noqa: E402
from flid_b200 import passes, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    dev = "cuda:0"
    # ---------------- TGN
    g = synth.wikipedia_shape(seed=0, scale=args.scale)
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    for layers in (1, 2):
        torch.manual_seed(0)
        m = flid_b200.MemoryModel(g.node_raw_features, g.edge_raw_features, s, 100, "TGN", layers, 2, 0.1, device=dev).to(dev)
        m.eval()
        passes.tgn_pass(m, g.src_node_ids[:2000], g.dst_node_ids[:2000], g.node_interact_times[:2000], g.edge_ids[:2000])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        passes.tgn_pass(m, g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e = g.num_interactions
        print(f"TGN L={layers} k=20 B=200 Wikipedia-shape ({e} events, {-(-e // 200)} batches): {dt * 1e3:8.1f} ms, "
              f"{dt / (-(-e // 200)) * 1e6:7.1f} us/batch, {2 * e / dt / 1e6:6.3f} M root queries/s", flush=True)
        del m
    # ---------------- TGAT drop-in loop
    g = synth.reddit_shape(seed=0, scale=args.scale)
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    torch.manual_seed(0)
    m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 2, 2, 0.1, dev).to(dev)
    m.eval()
    e = g.num_interactions
    for mode in ("auto", False, "auto"):
        m.set_layer_memo(mode)
        m._engine.memo.clear(), m._engine.served.clear()
        nb = -(-e // 200) if mode else 400
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            for b in range(nb):
                sl = slice(b * 200, min((b + 1) * 200, e))
                a, c = m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sl], g.dst_node_ids[sl],
                                                                   g.node_interact_times[sl], 20)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        roots = 2 * min(nb * 200, e)
        print(f"TGAT L=2 k=20 drop-in loop, B=200, memo={mode}: {nb} calls in {dt * 1e3:8.1f} ms, {dt / nb * 1e6:7.1f} us/call, "
              f"{roots / dt / 1e6:6.3f} M root queries/s (memo builds so far: {m._engine.memo_builds})", flush=True)


def dsub():
    """configs[3]: TGAT L=2, k=30 on the Dsub-shape graph (150 000 nodes / 168 154 edges, mostly padded
    neighbourhoods, one year of seconds: times beyond 2^24 are not float32-exact), double-way bulk pass."""
    dev = "cuda:0"
    g = synth.dsub_shape(seed=0, scale=1.0)
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    torch.manual_seed(0)
    m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 2, 2, 0.1, dev).to(dev)
    m.eval()
    for rep in range(3):
        m._engine.memo.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a, b = passes.embed_events(m, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 30)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    e = g.num_interactions
    st = m.last_stats()
    print(f"TGAT L=2 k=30 Dsub-shape bulk pass ({2 * e} root queries, {st[0]} root-phase evaluations): {dt * 1e3:7.2f} ms, "
          f"{2 * e / dt / 1e6:6.2f} M root queries/s", flush=True)


if __name__ == "__main__":
    dsub()
    main()
