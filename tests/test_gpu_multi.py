"""Multi-GPU parity on hardware ("sharded == single-GPU bit for bit").

Runs only when at least two CUDA devices are visible: two ranks (NCCL) run the sharded E-step pass of
``flid_b200.passes`` on a small Reddit-shaped graph; rank 0 repeats it unsharded and the gathered labels,
probabilities and embeddings must be identical.  (bench.py performs the same check at full size for every
N > 1 and prints ``"sharded_equals_single"``.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)      # a wedged collective must not hang the suite
    import torch.distributed as dist
    import flid_b200
    from flid_b200 import passes, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        g = synth.reddit_shape(seed=0, scale=0.03)
        s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
        torch.manual_seed(0)
        m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 2, 2, 0.1, dev).to(dev)
        dec = flid_b200.MLPClassifier(172, 0.1, 2).to(dev)
        m.eval(), dec.eval()
        m.set_layer_memo(True)
        out = {}
        for two in (False, True):
            m.invalidate_caches()
            p_sh, pr_sh, emb_sh = passes.e_step_pass(m, dec, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 20,
                                                     [], "entropy", 0.9, sharded=True, return_embeddings=True,
                                                     double_way=two)
            if rank == 0:
                m.invalidate_caches()
                p_1, pr_1, emb_1 = passes.e_step_pass(m, dec, g.src_node_ids, g.dst_node_ids, g.node_interact_times,
                                                      20, [], "entropy", 0.9, sharded=False, return_embeddings=True,
                                                      double_way=two)
                out[two] = (bool(torch.equal(p_sh, p_1)), bool(torch.equal(pr_sh, pr_1)),
                            bool(torch.equal(emb_sh[0], emb_1[0]) and torch.equal(emb_sh[1], emb_1[1])))
            dist.barrier()
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_sharded_pass_equals_single_gpu_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        out = q.get(timeout=300)
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    for p in procs:
        assert p.exitcode == 0
    for two, flags in out.items():
        assert all(flags), (two, flags)
