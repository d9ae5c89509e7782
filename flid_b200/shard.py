"""Owner partition of a bulk pass across the GPUs of one box.

The graph, the feature tables and the weights are replicated; what is split is the WORK, by
the node that owns it: the CSR positions [0, M) are cut at node boundaries into ``world``
ranges of (nearly) equal size, rank r owning the adjacency lists of nodes
[node_lo, node_hi) = positions [pos_lo, pos_hi).  Everything a query (v, t) reads -- its
neighbour window, the per-entry projections of that window, its own lower-layer rows --
lives in v's list, so a rank that answers exactly the queries of its own nodes only ever
needs tables for its own position range:

* layer-memo build: rank r evaluates the work items of its range in owner-major order
  (``flid_tgat_memo_build_owner_range``); item q produces the row of q's partner entry
  ``mirror[q]``, which belongs to whoever owns that position.  Rows whose position is
  outside the producer's range are exchanged once per level: 1/W of the table per rank
  instead of the whole table (the all-gather this replaces);
* root queries are routed to the owner of their node (one small all-to-all of
  (node, time, index) triples) and the per-root results come back through an all-reduce
  of a zero-initialised result table (12 B per event for C = 2).

Host-side logic only; the collectives are NCCL over NVLink (gloo in the CPU tests).
"""
import os

import numpy as np
import torch


def owner_bounds(indptr: np.ndarray, world: int):
    """Cut the CSR at node boundaries into ``world`` position ranges of nearly equal size.
    Returns (node_bounds int64[world+1], pos_bounds int64[world+1]); rank r owns nodes
    [node_bounds[r], node_bounds[r+1]) = positions [pos_bounds[r], pos_bounds[r+1])."""
    indptr = np.asarray(indptr, dtype=np.int64)
    num_lists = len(indptr) - 1               # node ids 0 .. num_nodes
    m = int(indptr[-1])
    targets = (m * np.arange(world + 1, dtype=np.int64)) // world
    node_bounds = np.searchsorted(indptr, targets, side="left").astype(np.int64)
    node_bounds[0], node_bounds[-1] = 0, num_lists
    node_bounds = np.maximum.accumulate(np.minimum(node_bounds, num_lists))
    return node_bounds, indptr[node_bounds].astype(np.int64)


def exchange_lists(mirror_local: torch.Tensor, pos_bounds, rank: int):
    """Index lists of the per-level row exchange of rank ``rank``.

    ``mirror_local`` = mirror[pos_lo:pos_hi] (int64, any device).  For an item q of this
    rank's range the produced row lives at position mirror[q]; for a position p of this
    rank's range the producing item is mirror[p] (the partner relation is an involution),
    so both directions are bucketisations of the same array.
    Returns (send_idx, send_splits, recv_idx, recv_splits): positions to send grouped by
    destination rank (ascending position inside a group) and the positions to receive grouped
    by source rank in the same order."""
    world = len(pos_bounds) - 1
    dev = mirror_local.device
    inner = torch.as_tensor(np.asarray(pos_bounds[1:-1], dtype=np.int64), device=dev)
    other = torch.bucketize(mirror_local, inner, right=True)              # owner rank of the partner position
    lo = int(pos_bounds[rank])
    remote = other != rank
    # send: rows at positions mirror[q], grouped by their owner, ascending position
    sp, sd = mirror_local[remote], other[remote]
    order = torch.argsort(sd * (int(pos_bounds[-1]) + 1) + sp)
    send_idx = sp[order]
    send_splits = torch.bincount(sd, minlength=world).tolist()
    # receive: positions p of this range whose producer mirror[p] is remote, grouped by producer rank; the sender
    # orders its group by ascending destination position, i.e. by p
    rp = torch.arange(lo, lo + mirror_local.numel(), device=dev, dtype=torch.int64)[remote]
    order = torch.argsort(sd * (int(pos_bounds[-1]) + 1) + rp)
    recv_idx = rp[order]
    recv_splits = list(send_splits)   # |{q in R_r : mirror[q] in R_s}| == |{p in R_r : mirror[p] in R_s}| (same set of pairs)
    return send_idx, send_splits, recv_idx, recv_splits


class ShardPlan:
    """Per (sampler, world size, rank): ownership ranges and the exchange index lists."""

    def __init__(self, sampler, rank: int, world: int, device):
        from . import _lib
        indptr = sampler._host_csr()[0]
        self.rank, self.world = rank, world
        self.node_bounds, self.pos_bounds = owner_bounds(indptr, world)
        self.node_lo, self.node_hi = int(self.node_bounds[rank]), int(self.node_bounds[rank + 1])
        self.pos_lo, self.pos_hi = int(self.pos_bounds[rank]), int(self.pos_bounds[rank + 1])
        self.num_entries = int(indptr[-1])
        self.node_inner = torch.as_tensor(self.node_bounds[1:-1], device=device)
        mirror = torch.empty(max(self.num_entries, 1), dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().flid_graph_export_mirror(sampler.handle, _lib.ptr(mirror), _lib.stream()))
        local = mirror[self.pos_lo:self.pos_hi].to(torch.int64)
        self.send_idx, self.send_splits, self.recv_idx, self.recv_splits = exchange_lists(local, self.pos_bounds, rank)
        self.routing = {}     # last routed root set of a device-resident pass (flid_b200.passes._owned_roots)
        self.peers = {}       # (slot, rows, dn) -> (table, peer pointers, bounds, keep-alive)
        self.p2p = None if os.environ.get("FLID_P2P", "1") != "0" else False   # None: not tried yet
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.flag_dev = torch.zeros(1, dtype=torch.int32, device=device)      # routing-cache agreement (passes._owned_roots)
        self.flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._side = None
        # host-side control group: one-integer agreements between the ranks that must not touch the GPU queue (the
        # device is busy with the memo build while the hosts agree).  Collective: every rank builds its plan at the
        # same point of the pass.
        self.ctl_group = None
        if os.environ.get("FLID_CTL_GLOO", "1") != "0":
            import torch.distributed as dist
            try:
                self.ctl_group = dist.new_group(backend="gloo")
            except Exception:
                self.ctl_group = None
        del mirror

    def side_stream(self):
        """Stream for small control collectives that must not queue behind the pass's kernels."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.flag_dev.device)
        return self._side

    def all_agree(self, ok: bool, dist) -> bool:
        """True when every rank passed ok=True.  Called after the memo build has been enqueued: over the host group
        the exchange overlaps the device work; without one it is an NCCL all-reduce read back on a side stream."""
        if self.ctl_group is not None:
            t = torch.tensor([1 if ok else 0], dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.ctl_group)
            return int(t[0]) == 1
        side = self.side_stream()
        with torch.cuda.stream(side):
            self.flag_dev.fill_(1 if ok else 0)
            dist.all_reduce(self.flag_dev, op=dist.ReduceOp.MIN)
            self.flag_host.copy_(self.flag_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(side)
        done.synchronize()
        return int(self.flag_host[0]) == 1

    # ---- peer-mapped tables (CUDA IPC): the exchange as one kernel storing into the other ranks' HBM
    def peer_table(self, slot: int, rows: int, dn: int, device, dist):
        """A [rows, dn] float32 table of this rank that every other rank has mapped, or None when CUDA IPC is not
        available on this box (the exchange then goes through all_to_all_single).  Allocated once per
        (slot, shape) and reused by every pass; collective on first use."""
        import ctypes as C
        from . import _lib
        key = (slot, rows, dn)
        ent = self.peers.get(key)
        if ent is not None:
            return ent
        if self.p2p is False:
            return None
        lib = _lib.lib()
        ok, ptr, mapped = 1, C.c_void_p(None), []
        handle = (C.c_ubyte * 64)()
        try:
            with torch.cuda.device(device):
                _lib.check(lib.flid_peer_alloc(rows * dn * 4, C.byref(ptr), handle))
        except Exception:
            ok = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if ok else None)
        peers = (C.c_void_p * self.world)()
        if ok and all(h is not None for h in handles):
            try:
                with torch.cuda.device(device):
                    for r in range(self.world):
                        if r == self.rank:
                            peers[r] = ptr.value
                        else:
                            pp = C.c_void_p(None)
                            buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                            _lib.check(lib.flid_peer_open(buf, C.byref(pp)))
                            mapped.append(pp)
                            peers[r] = pp.value
            except Exception:
                ok = 0
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:           # some rank could not export / map: everyone uses the collective path
            for pp in mapped:
                lib.flid_peer_close(pp)
            if ptr.value:
                lib.flid_peer_free(ptr)
            self.p2p = False
            return None
        self.p2p = True

        class _DevArray:       # zero-copy torch view of the cudaMalloc'ed block
            pass
        arr = _DevArray()
        arr.__cuda_array_interface__ = {"shape": (rows, dn), "typestr": "<f4", "data": (int(ptr.value), False),
                                        "version": 3, "strides": None}
        table = torch.as_tensor(arr, device=device)
        bounds = (C.c_int64 * (self.world + 1))(*[int(x) for x in self.pos_bounds])
        ent = (table, peers, bounds, arr)
        self.peers[key] = ent
        return ent

    def exchange_rows_p2p(self, sampler, ent, dist):
        """flid_memo_exchange_p2p + a stream-ordered cross-rank barrier (a one-element all-reduce): when it completes
        on this rank, every rank's exchange kernel has finished storing into this rank's table."""
        from . import _lib
        table, peers, bounds, _ = ent
        _lib.check(_lib.lib().flid_memo_exchange_p2p(sampler.handle, _lib.ptr(table), peers, bounds, self.world, self.rank,
                                                     table.shape[1], _lib.stream()))
        dist.all_reduce(self._flag)

    def exchange_rows(self, table: torch.Tensor, dist):
        """Send the rows this rank produced for other ranks' positions, receive the rows of this rank's positions
        that other ranks produced (in place on the full-size ``table``)."""
        send = table.index_select(0, self.send_idx)
        recv = table.new_empty((int(self.recv_idx.numel()), table.shape[1]))
        dist.all_to_all_single(recv, send, output_split_sizes=self.recv_splits, input_split_sizes=self.send_splits)
        table.index_copy_(0, self.recv_idx, recv)


def route_roots(nodes: torch.Tensor, times: torch.Tensor, gidx: torch.Tensor, node_inner: torch.Tensor, world: int,
                dist):
    """Send every root query (node, time, global index) to the rank that owns its node.
    Inputs are this rank's slice of the pass (device tensors: int64, float64, int64); returns the
    (nodes, times, gidx) this rank owns."""
    dest = torch.bucketize(nodes, node_inner, right=True)
    order = torch.argsort(dest, stable=True)
    packed = torch.stack([nodes, times.view(torch.int64), gidx], dim=1)[order].contiguous()
    send_counts = torch.bincount(dest, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    send_splits, recv_splits = send_counts.tolist(), recv_counts.tolist()
    out = packed.new_empty((int(sum(recv_splits)), 3))
    dist.all_to_all_single(out, packed, output_split_sizes=recv_splits, input_split_sizes=send_splits)
    return out[:, 0].contiguous(), out[:, 1].contiguous().view(torch.float64), out[:, 2].contiguous()
