"""Host logic of the owner-partitioned multi-GPU pass (flid_b200/shard.py) on CPU: ownership bounds, the
per-level row-exchange index lists (all ranks emulated in one process, then for real on gloo with two
ranks), root routing and the scatter + all-reduce that combines per-root results."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flid_b200 import passes, shard


def _csr_with_mirror(num_nodes, num_events, seed):
    """numpy restatement of the device CSR build: entries sorted by (owner, time, insertion order) + partner index."""
    rs = np.random.RandomState(seed)
    p = 1.0 / np.arange(1, num_nodes + 1) ** 0.9
    p /= p.sum()
    src = 1 + rs.choice(num_nodes, num_events, p=p)
    dst = 1 + rs.choice(num_nodes, num_events, p=p)
    ts = np.sort(np.floor(rs.uniform(0, 1000, num_events)))
    owner = np.stack([src, dst], axis=1).reshape(-1)
    order = np.lexsort((np.arange(2 * num_events), np.repeat(ts, 2), owner))
    inv = np.empty_like(order)
    inv[order] = np.arange(2 * num_events)
    mirror = inv[order ^ 1]                 # partner of entry at sorted position i: original index order[i] ^ 1
    indptr = np.zeros(num_nodes + 2, dtype=np.int64)
    np.add.at(indptr, owner + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, mirror.astype(np.int64), owner[order]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_owner_bounds_and_exchange_lists(world):
    indptr, mirror, owner_sorted = _csr_with_mirror(40, 600, seed=world)
    m = int(indptr[-1])
    assert np.array_equal(mirror[mirror], np.arange(m)), "partner index must be an involution"
    node_b, pos_b = shard.owner_bounds(indptr, world)
    assert node_b[0] == 0 and node_b[-1] == len(indptr) - 1 and pos_b[0] == 0 and pos_b[-1] == m
    assert np.all(np.diff(node_b) >= 0) and np.array_equal(pos_b, indptr[node_b])
    sizes = np.diff(pos_b)
    assert sizes.max() <= m / world + np.diff(indptr).max(), "ranges are balanced up to one adjacency list"
    # every rank's range holds whole adjacency lists
    for r in range(world):
        own = owner_sorted[pos_b[r]:pos_b[r + 1]]
        assert own.size == 0 or (own.min() >= node_b[r] and own.max() < node_b[r + 1])
    # emulate the exchange: rank r produces the rows mirror[q], q in its range (row value = position)
    lists = [shard.exchange_lists(torch.from_numpy(mirror[pos_b[r]:pos_b[r + 1]]), pos_b, r) for r in range(world)]
    tables = []
    for r in range(world):
        t = torch.full((m, 1), float("nan"))
        produced = torch.from_numpy(mirror[pos_b[r]:pos_b[r + 1]])
        t[produced, 0] = produced.to(torch.float32)
        tables.append(t)
    for r in range(world):
        send_idx, send_splits, recv_idx, recv_splits = lists[r]
        assert sum(send_splits) == send_idx.numel() and sum(recv_splits) == recv_idx.numel() and send_splits[r] == 0
        off = 0
        for s in range(world):                     # what rank s sends to rank r
            s_idx, s_splits = lists[s][0], lists[s][1]
            start = sum(s_splits[:r])
            chunk = s_idx[start:start + s_splits[r]]
            assert recv_splits[s] == chunk.numel()
            mine = recv_idx[off:off + recv_splits[s]]
            assert torch.equal(mine, chunk), "sender and receiver must enumerate the rows in the same order"
            tables[r][mine] = tables[s][chunk]
            off += recv_splits[s]
    for r in range(world):
        own = tables[r][pos_b[r]:pos_b[r + 1], 0]
        assert torch.equal(own, torch.arange(pos_b[r], pos_b[r + 1], dtype=torch.float32)), "range incomplete"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        indptr, mirror, _ = _csr_with_mirror(30, 400, seed=5)
        m = int(indptr[-1])
        node_b, pos_b = shard.owner_bounds(indptr, world)
        # row exchange through the real collective
        plan = shard.ShardPlan.__new__(shard.ShardPlan)
        lo, hi = int(pos_b[rank]), int(pos_b[rank + 1])
        plan.send_idx, plan.send_splits, plan.recv_idx, plan.recv_splits = shard.exchange_lists(
            torch.from_numpy(mirror[lo:hi]), pos_b, rank)
        table = torch.full((m + 1, 3), float("nan"))
        produced = torch.from_numpy(mirror[lo:hi])
        table[produced] = produced.to(torch.float32).unsqueeze(1) * torch.tensor([1.0, 2.0, 3.0])
        plan.exchange_rows(table, dist)
        want = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1) * torch.tensor([1.0, 2.0, 3.0])
        ok_rows = torch.equal(table[lo:hi], want)
        # root routing + scatter / all-reduce
        rs = np.random.RandomState(9)
        e = 101
        src, dst = rs.randint(1, 31, e), rs.randint(1, 31, e)
        t = np.sort(rs.uniform(0, 1000, e))
        elo, ehi, _ = passes.shard_bounds(e, rank, world)
        ev = torch.arange(elo, ehi)
        nodes = torch.cat([torch.from_numpy(src[elo:ehi]), torch.from_numpy(dst[elo:ehi])])
        times = torch.cat([torch.from_numpy(t[elo:ehi])] * 2)
        gidx = torch.cat([ev, ev + e])
        n_own, t_own, g_own = shard.route_roots(nodes, times, gidx, torch.from_numpy(node_b[1:-1]), world, dist)
        all_nodes, all_t = np.concatenate([src, dst]), np.concatenate([t, t])
        ok_route = bool(((n_own >= node_b[rank]) & (n_own < node_b[rank + 1])).all())
        ok_route = ok_route and np.array_equal(all_nodes[g_own.numpy()], n_own.numpy())
        ok_route = ok_route and np.array_equal(all_t[g_own.numpy()], t_own.numpy())
        vals = torch.stack([n_own.to(torch.float32), t_own.to(torch.float32)], dim=1)
        full = passes._scatter_all_reduce(vals, g_own, 2 * e, dist)
        ok_full = np.array_equal(full[:, 0].numpy(), all_nodes.astype(np.float32)) and \
            np.array_equal(full[:, 1].numpy(), all_t.astype(np.float32))
        # routing-cache agreement over the host control group: a hit only when every rank has one
        plan.ctl_group = dist.new_group(backend="gloo")
        ok_agree = plan.all_agree(True, dist) and not plan.all_agree(rank == 0, dist) and not plan.all_agree(False, dist)
        q.put((rank, ok_rows, ok_route, ok_full and ok_agree))
    finally:
        dist.destroy_process_group()


def test_row_exchange_and_root_routing_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_rows, ok_route, ok_full in results:
        assert ok_rows and ok_route and ok_full, (rank, ok_rows, ok_route, ok_full)
