"""World-size-2 checks of the sharding / gather host logic on the gloo backend (no GPU).

The device work of a sharded pass is "same kernel, different slice" (bit-identity is pinned by
tests/test_gpu_parity.py::test_tgat_layer_memo_is_bit_identical and
::test_tgat_chunking_and_table_do_not_change_bits); what is exercised here is the part that
only exists with more than one rank: contiguous order-preserving shard bounds, the padded
row all-gather, and the in-place all-gather layout the layer-memo build uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flid_b200 import passes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, width, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_items * width, dtype=torch.float32).reshape(n_items, width) * 0.5 + 1.0
        lo, hi, per = passes.shard_bounds(n_items, rank, world)
        got = passes.all_gather_rows(full[lo:hi].clone(), n_items, per, dist)
        ok_rows = torch.equal(got, full)
        # the (label, probs) packing of e_step_pass
        labels = (torch.arange(n_items) % 3)[lo:hi]
        probs = full[lo:hi, :2]
        packed = torch.cat([labels.to(torch.float32).unsqueeze(1), probs], dim=1)
        g2 = passes.all_gather_rows(packed, n_items, per, dist)
        ok_pack = torch.equal(g2[:, 0].to(torch.int64), torch.arange(n_items) % 3) and torch.equal(g2[:, 1:], full[:, :2])
        # in-place all-gather of a row-sharded table (flid_b200.tgat.build_layer_memo)
        rows = n_items + 1
        per_t = -(-rows // world)
        table = torch.full((per_t * world, width), float("nan"))
        tlo, thi = min(rank * per_t, rows), min((rank + 1) * per_t, rows)
        ref = torch.arange(per_t * world * width, dtype=torch.float32).reshape(per_t * world, width)
        table[tlo:thi] = ref[tlo:thi]
        dist.all_gather_into_tensor(table, table[rank * per_t:(rank + 1) * per_t])
        ok_table = torch.equal(table[:rows], ref[:rows])
        out_q.put((rank, ok_rows, ok_pack, ok_table))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items,width", [(11, 3), (8, 172), (1, 2)])
def test_sharded_gathers_world2(n_items, width):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, width, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_rows, ok_pack, ok_table in results:
        assert ok_rows and ok_pack and ok_table, (rank, ok_rows, ok_pack, ok_table)


def test_shard_bounds_properties():
    for n in (0, 1, 7, 200, 672447):
        for w in (1, 2, 3, 4, 8):
            spans = [passes.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo <= per for lo, hi, per in spans)
            assert len({per for _, _, per in spans}) == 1


def test_memo_piece_bounds_cover_every_row_once():
    """Sharded memo build: the (rank, piece) row ranges tile [0, rows) exactly and every super-block is one
    contiguous all-gather output whose r-th slice is rank r's range."""
    from flid_b200.tgat import memo_piece_bounds
    for rows in (1, 7, 1000, 1344895):
        for world in (1, 2, 3, 8):
            for pieces in (1, 2, 3):
                seen = np.zeros(rows, dtype=np.int32)
                for rank in range(world):
                    per, bounds = memo_piece_bounds(rows, rank, world, pieces)
                    assert per * world * pieces >= rows and len(bounds) == pieces
                    for p, (base, lo, hi) in enumerate(bounds):
                        assert base == p * per * world
                        assert lo == min(base + rank * per, rows) and hi - lo <= per
                        seen[lo:hi] += 1
                assert (seen == 1).all(), (rows, world, pieces)
