#!/usr/bin/env python
"""Multi-GPU plumbing probe (development tool): the collectives the owner-partitioned pass relies on, each with a
watchdog that dumps the Python stacks and exits instead of hanging.

    torchrun --nproc-per-node 2 tools/dist_probe.py
"""
import faulthandler
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    faulthandler.dump_traceback_later(int(os.environ.get("FLID_DEBUG_HANG", "120")), exit=True)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def log(msg):
        print(f"[rank {rank}] {time.strftime('%H:%M:%S')} {msg}", flush=True)

    x = torch.ones(4, device=dev)
    dist.all_reduce(x)
    torch.cuda.synchronize()
    log(f"all_reduce ok {x.tolist()}")
    cnt = torch.arange(world, device=dev, dtype=torch.int64) + rank
    rc = torch.empty_like(cnt)
    dist.all_to_all_single(rc, cnt)
    torch.cuda.synchronize()
    log(f"all_to_all_single (equal) ok {rc.tolist()}")
    send_splits = [(rank + 1) * (j + 1) for j in range(world)]
    recv_splits = [(j + 1) * (rank + 1) for j in range(world)]
    send = torch.full((sum(send_splits), 3), float(rank), device=dev)
    recv = torch.empty((sum(recv_splits), 3), device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_splits, input_split_sizes=send_splits)
    torch.cuda.synchronize()
    log(f"all_to_all_single (ragged) ok {recv[:, 0].tolist()[:6]}")

    import flid_b200
    from flid_b200 import passes, synth
    g = synth.reddit_shape(seed=0, scale=0.03)
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    torch.manual_seed(0)
    m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 2, 2, 0.1, dev).to(dev)
    dec = flid_b200.MLPClassifier(172, 0.1, 2).to(dev)
    m.eval(), dec.eval()
    m.set_layer_memo(True)
    log("model ready")
    from flid_b200.tgat import shard_plan
    plan = shard_plan(m._engine, s, dev)
    torch.cuda.synchronize()
    log(f"plan: nodes [{plan.node_lo}, {plan.node_hi}) positions [{plan.pos_lo}, {plan.pos_hi}) send {plan.send_splits} recv {plan.recv_splits}")
    with torch.no_grad():
        m.build_layer_memo(20, sharded=True)
    torch.cuda.synchronize()
    log(f"sharded memo built; peer-mapped exchange: {plan.p2p}")
    for two in (False, True):
        m.invalidate_caches()
        p_sh, pr_sh, emb_sh = passes.e_step_pass(m, dec, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 20, [],
                                                 "entropy", 0.9, sharded=True, return_embeddings=True, double_way=two)
        torch.cuda.synchronize()
        log(f"sharded pass ok (double_way={two})")
        if rank == 0:
            m.invalidate_caches()
            p_1, pr_1, emb_1 = passes.e_step_pass(m, dec, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 20, [],
                                                  "entropy", 0.9, sharded=False, return_embeddings=True, double_way=two)
            torch.cuda.synchronize()
            log(f"equal: labels {torch.equal(p_sh, p_1)} probs {torch.equal(pr_sh, pr_1)} emb "
                f"{torch.equal(emb_sh[0], emb_1[0])} {torch.equal(emb_sh[1], emb_1[1])} "
                f"max|diff| {float((emb_sh[0] - emb_1[0]).abs().max()):.3e}")
        dist.barrier()
    log("done")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
