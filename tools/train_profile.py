"""torch.profiler breakdown (CPU and CUDA time by op / kernel) of one training-mode TGAT step at B = 200."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flid_b200
from flid_b200 import synth, train
from torch.profiler import profile, ProfilerActivity
dev='cuda:0'
d = synth.reddit_shape(scale=0.3)
src, dst, eid, ts = d.src_node_ids, d.dst_node_ids, d.edge_ids, d.node_interact_times
s = flid_b200.NeighborSampler(None, "recent", seed=1, device=dev, _events=(src, dst, eid, ts, d.num_nodes))
m = flid_b200.TGAT(d.node_raw_features, d.edge_raw_features, s, 100, 2, 2, 0.1, dev).to(dev); m.train()
for B in (200,):
    lo = len(src)//2
    nodes = np.concatenate([src[lo:lo+B], dst[lo:lo+B]]); times = np.concatenate([ts[lo:lo+B]]*2)
    def step():
        m.zero_grad(set_to_none=True)
        out = train.autograd_forward(m.time_encoder, m.temporal_conv_layers, m.merge_layers, s, m.node_raw_features, m.edge_raw_features, nodes, times, 2, 20, True)
        out.square().mean().backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    print("B", B)
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=50))
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=50))
