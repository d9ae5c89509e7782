#!/usr/bin/env python
"""Where do the ~150 us of a B = 200 drop-in call go (development probe)?  Times, per call of 400 root queries with
the layer memo built: the public numpy-in call, the same with device-resident inputs, and the bare C entry point
(flid_tgat_embed_memo) in a loop with everything pre-staged.

    python tools/call_probe.py
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flid_b200  # noqa: E402
from flid_b200 import _lib, synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    g = synth.reddit_shape(seed=0, scale=1.0)
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    torch.manual_seed(0)
    m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 2, 2, 0.1, dev).to(dev)
    m.eval()
    m.set_layer_memo(True)
    e = g.num_interactions
    nb = 500
    with torch.no_grad():
        m.build_layer_memo(20)
        sl = slice(e - 200, e)
        for _ in range(20):
            m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl], 20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(nb):
            sl = slice(e - 200 * (i + 1), e - 200 * i)
            m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl], 20)
        torch.cuda.synchronize()
        t_api = (time.perf_counter() - t0) / nb
        nodes = torch.from_numpy(np.concatenate([g.src_node_ids[-200:], g.dst_node_ids[-200:]])).to(dev)
        times = torch.from_numpy(np.concatenate([g.node_interact_times[-200:]] * 2)).to(dev)
        for _ in range(20):
            m.compute_node_temporal_embeddings(nodes, times, 2, 20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(nb):
            m.compute_node_temporal_embeddings(nodes, times, 2, 20)
        torch.cuda.synchronize()
        t_dev = (time.perf_counter() - t0) / nb
        # bare C call
        lib = _lib.lib()
        h = m._engine.handles[2]
        memo = m._engine.memo[2][1]
        tabs = (C.c_void_p * len(memo))(*[t.data_ptr() for t in memo])
        out = torch.empty((400, 172), dtype=torch.float32, device=dev)
        args = (h, s.handle, _lib.ptr(m.node_raw_features), _lib.ptr(m.edge_raw_features), tabs, _lib.ptr(nodes),
                _lib.ptr(times), 0, 400, 20, _lib.ptr(out), _lib.stream())
        for _ in range(20):
            _lib.check(lib.flid_tgat_embed_memo(*args))
        torch.cuda.synchronize()
        l0 = lib.flid_launch_count()
        t0 = time.perf_counter()
        for i in range(nb):
            lib.flid_tgat_embed_memo(*args)
        torch.cuda.synchronize()
        t_c = (time.perf_counter() - t0) / nb
        launches = (lib.flid_launch_count() - l0) / nb
        # device time of the same chain
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(nb):
            lib.flid_tgat_embed_memo(*args)
        ev1.record()
        torch.cuda.synchronize()
        t_gpu = ev0.elapsed_time(ev1) / nb * 1e-3
    print(f"per call of 400 roots (L=2, k=20, memo): public numpy API {t_api * 1e6:.1f} us, device-tensor API {t_dev * 1e6:.1f} us, "
          f"bare C entry {t_c * 1e6:.1f} us wall / {t_gpu * 1e6:.1f} us device-timed, {launches:.1f} launches per call", flush=True)


if __name__ == "__main__":
    main()
