#!/usr/bin/env python
"""Does processing root queries in (node, time) order (sliding neighbour windows -> cache reuse) speed up
the attention stream?  Development experiment."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flid_b200  # noqa: E402
from flid_b200 import _lib, synth  # noqa: E402

dev = "cuda:0"
g = synth.reddit_shape(seed=0, scale=1.0)
s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
torch.manual_seed(0)
m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 2, 2, 0.1, dev).to(dev)
m.eval()
nodes = np.concatenate([g.src_node_ids, g.dst_node_ids])
times = np.concatenate([g.node_interact_times, g.node_interact_times])
order = np.lexsort((times, nodes))
lib = _lib.lib()
with torch.no_grad():
    m.build_layer_memo(20)
    for name, (nn, tt) in {"event order": (nodes, times), "(node, time) order": (nodes[order], times[order])}.items():
        nd, td = torch.from_numpy(nn).to(dev), torch.from_numpy(tt).to(dev)
        for rep in range(3):
            out = m.compute_node_temporal_embeddings(nd, td, 2, 20)
        h = m._engine.handles[2]
        _lib.check(lib.flid_tgat_profile(h, 1))
        torch.cuda.synchronize()
        out = m.compute_node_temporal_embeddings(nd, td, 2, 20)
        ms = (ctypes.c_double * 4)()
        cnt = (ctypes.c_int64 * 4)()
        _lib.check(lib.flid_tgat_profile_read(h, ms, cnt))
        _lib.check(lib.flid_tgat_profile(h, 0))
        print(f"{name:22s}: sample {ms[0]:.2f}  qfold {ms[1]:.2f}  attention {ms[2]:.2f}  chain {ms[3]:.2f} ms", flush=True)
