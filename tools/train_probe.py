"""Device time of one training-mode TGAT step (forward + backward, dropout on) for an M-step batch:
the kernel path (flid_b200/train.py) next to the same layers composed from torch ops in the
reference's literal order (tests/test_gpu_train.reference_layers).  Reddit-shape graph by default."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import flid_b200                                   # noqa: E402
from flid_b200 import synth, train                 # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--batch", type=int, nargs="+", default=[200, 2000])
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--skip-torch", action="store_true", help="time the kernel path only (ncu captures)")
    a = ap.parse_args()
    from test_gpu_train import reference_layers  # torch ops, nn.Dropout for the output dropout
    dev = "cuda:0"
    d = synth.reddit_shape(scale=a.scale)
    src, dst, eid, ts = d.src_node_ids, d.dst_node_ids, d.edge_ids, d.node_interact_times
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=dev, _events=(src, dst, eid, ts, d.num_nodes))
    torch.manual_seed(0)
    m = flid_b200.TGAT(d.node_raw_features, d.edge_raw_features, s, 100, a.layers, 2, 0.1, dev).to(dev)
    m.train()
    E = src.shape[0]

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    for B in a.batch:
        lo = E // 2
        nodes = np.concatenate([src[lo:lo + B], dst[lo:lo + B]])
        times = np.concatenate([ts[lo:lo + B], ts[lo:lo + B]])
        levels = train.sample_levels(s, nodes, times, a.layers, a.k, torch.device(dev))
        keeps = [train.score_keep_mask(l, levels[l][0].shape[0], 2, a.k, 0.1, dev) for l in range(1, a.layers + 1)]

        def kernel_step():
            m.zero_grad(set_to_none=True)
            out = train.autograd_forward(m.time_encoder, m.temporal_conv_layers, m.merge_layers, s,
                                         m.node_raw_features, m.edge_raw_features, nodes, times, a.layers, a.k, True)
            out.square().mean().backward()

        def torch_step():
            m.zero_grad(set_to_none=True)
            lv = train.sample_levels(s, nodes, times, a.layers, a.k, torch.device(dev))
            out = reference_layers(m.time_encoder, m.temporal_conv_layers, m.merge_layers, m.node_raw_features,
                                   m.edge_raw_features, lv, a.layers, a.k, keeps, 0.1)
            out.square().mean().backward()

        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        tk = timed(kernel_step)
        mk = torch.cuda.max_memory_allocated() - base
        if a.skip_torch:
            print(f"B={B:6d} roots={2 * B:6d} L={a.layers} k={a.k}: kernel path {tk:8.3f} ms ({mk / 2**20:8.1f} MiB peak)")
            continue
        torch.cuda.reset_peak_memory_stats()
        tt = timed(torch_step)
        mt = torch.cuda.max_memory_allocated() - base
        print(f"B={B:6d} roots={2 * B:6d} L={a.layers} k={a.k}: kernel path {tk:8.3f} ms ({mk / 2**20:8.1f} MiB peak)   "
              f"torch-op path {tt:8.3f} ms ({mt / 2**20:8.1f} MiB peak)   x{tt / tk:.2f}")


if __name__ == "__main__":
    main()
