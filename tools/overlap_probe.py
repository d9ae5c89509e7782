#!/usr/bin/env python
"""Does the attention stream of one chunk overlap the projection GEMMs of another (development probe)?

Two TGAT models (L = 1, same shapes) embed the same number of roots, each from its own Python thread on its own
CUDA stream, so that one model's attention kernel can run beside the other's tcgen05 GEMM chain.  Compared with the
two calls issued back to back on one stream.  The speed-up bounds what a pipelined chunk loop inside one pass
could gain.

    python tools/overlap_probe.py [--roots 600000] [--reps 5]
"""
import argparse
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flid_b200  # noqa: E402
from flid_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--roots", type=int, default=600000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = synth.reddit_shape(seed=0, scale=1.0)
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    models = []
    for i in range(2):
        torch.manual_seed(i)
        m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 1, 2, 0.1, dev).to(dev)
        m.eval()
        models.append(m)
    e = g.num_interactions
    n = min(args.roots, e)
    nodes = torch.from_numpy(g.src_node_ids[e - n:]).to(dev)
    times = torch.from_numpy(g.node_interact_times[e - n:]).to(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def work(i, reps, stream):
        with torch.no_grad(), torch.cuda.stream(stream):
            for _ in range(reps):
                models[i].compute_node_temporal_embeddings(nodes, times, 1, 20)

    for i in range(2):                      # warm-up: allocations, weight upload
        work(i, 2, streams[i])
    torch.cuda.synchronize()
    # sequential: both models on one stream, one thread
    t0 = time.perf_counter()
    work(0, args.reps, streams[0])
    work(1, args.reps, streams[0])
    torch.cuda.synchronize()
    seq = time.perf_counter() - t0
    # concurrent: one thread and one stream per model
    bar = threading.Barrier(2)

    def runner(i):
        bar.wait()
        work(i, args.reps, streams[i])

    t0 = time.perf_counter()
    th = [threading.Thread(target=runner, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    con = time.perf_counter() - t0
    per = 2 * args.reps
    print(f"L=1 embedding of {n} roots: sequential {1e3 * seq / per:.3f} ms per call, two streams {1e3 * con / per:.3f} ms "
          f"per call, speed-up {seq / con:.3f}", flush=True)


if __name__ == "__main__":
    main()
