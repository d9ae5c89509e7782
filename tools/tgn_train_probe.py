"""Device + host time of one training-mode TGN batch (MemoryModel.compute_src_dst_node_temporal_embeddings with
grad enabled, forward + backward), next to the eval-mode (no-grad kernels) batch.  Wikipedia shape."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flid_b200                                   # noqa: E402
from flid_b200 import synth                        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.3)
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--batches", type=int, default=60)
    a = ap.parse_args()
    dev = "cuda:0"
    g = synth.wikipedia_shape(seed=0, scale=a.scale)
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    torch.manual_seed(0)
    for mode in ("eval", "train"):
        m = flid_b200.MemoryModel(g.node_raw_features, g.edge_raw_features, s, 100, "TGN", a.layers, 2, 0.1,
                                  device=dev).to(dev)
        m.memory_bank.__init_memory_bank__()
        m.train(mode == "train")
        B = 200

        def batch(i):
            sl = slice(i * B, (i + 1) * B)
            args = (g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl], g.edge_ids[sl], True, 20)
            if mode == "eval":
                with torch.no_grad():
                    m.compute_src_dst_node_temporal_embeddings(*args)
            else:
                m.zero_grad(set_to_none=True)
                x, y = m.compute_src_dst_node_temporal_embeddings(*args)
                (x.square().mean() + y.square().mean()).backward()
                m.memory_bank.detach_memory_bank()

        for i in range(5):
            batch(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(5, 5 + a.batches):
            batch(i)
        torch.cuda.synchronize()
        print(f"TGN L={a.layers} {mode:5s}: {(time.perf_counter() - t0) / a.batches * 1e3:7.3f} ms per batch of {B} events")


if __name__ == "__main__":
    main()
