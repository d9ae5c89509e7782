"""Whole-pass drivers: the E/200-iteration Python loops of the reference
(``PTCL/M_step.py:454-509`` embedding pass, ``PTCL/E_step.py:305-352`` pseudo-label pass,
``PTCL/utils.py:69-123`` filter) as one call each, plus query sharding across the GPUs of
one box (graph / features / weights replicated, contiguous event ranges per rank, results
all-gathered with NCCL -- SURVEY.md section 8(e)).  TGN is sequential in time and is not
sharded ("replicas only").
"""
import numpy as np
import torch

from .pseudo_label import MLPClassifier, emit_pseudo_labels, entropy_filter, prob_filter


import os
import time

# whole-pass calls with at least this many root queries (a global property of the call, so that every rank and the
# unsharded run decide alike) use the LayerNorm-folded fc1 (flid_tgat_set_ln_fold)
BULK_MIN_ROOTS = 1
_TRACE = os.environ.get("FLID_PASS_TRACE") == "1"      # development: synchronising phase timer of the sharded pass
_trace_acc = {}


def _mark(name, t0):
    """Phase timer (FLID_PASS_TRACE=1 only): synchronises, accumulates wall time per phase."""
    if not _TRACE:
        return t0
    torch.cuda.synchronize()
    now = time.perf_counter()
    _trace_acc[name] = _trace_acc.get(name, 0.0) + (now - t0)
    return now


def trace_report(reset=True):
    out = {k: round(1000.0 * v, 3) for k, v in _trace_acc.items()}
    if reset:
        _trace_acc.clear()
    return out


def shard_bounds(num_items: int, rank: int, world_size: int):
    """Contiguous, order-preserving split; every rank gets ceil(n / W) items except the tail."""
    per = -(-num_items // world_size) if num_items else 0
    lo = min(rank * per, num_items)
    return lo, min(lo + per, num_items), per


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def all_gather_rows(local: torch.Tensor, num_items: int, per: int, dist_mod=None):
    """Gather row-sharded [<=per, ...] tensors into [num_items, ...] (pads the tail shard)."""
    dist = dist_mod or _dist()[0]
    world = dist.get_world_size()
    pad = per - local.shape[0]
    if pad:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], dim=0)
    out = local.new_empty((per * world,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out[:num_items]


def _require_eval(*modules):
    """The reference's pass loops run after ``model.eval()`` (PTCL/E_step.py:300, PTCL/M_step.py:456); a module
    left in training mode would silently take the dropout / autograd path for the whole pass."""
    for m in modules:
        if m is not None and m.training:
            raise RuntimeError(f"flid_b200.passes: {type(m).__name__} is in training mode; call .eval() first "
                               "(bulk passes are inference passes)")


def prepare_layer_memo(model, num_roots: int, num_neighbors: int, sharded: bool):
    """Build the layer memo (flid_tgat_memo_build) ahead of a bulk pass when it pays off:
    a pass over ``num_roots`` root queries costs sum_l (1+k)^(L-l) attention evaluations per
    root without it and L per root plus (L-1)(entries+1) with it.  When ``sharded`` every rank
    builds a contiguous slice of the table rows and the slices are all-gathered (NCCL)."""
    build = getattr(model, "build_layer_memo", None)
    depth = getattr(model, "num_layers", 1)
    if build is None or depth < 2 or not model._engine.memo_mode or num_neighbors <= 0:
        return False
    plain = sum((1 + num_neighbors) ** (depth - l) for l in range(1, depth + 1))
    cost = (depth - 1) * (model.neighbor_sampler.num_entries + 1)
    if num_roots * (plain - depth) < cost:
        return False
    build(num_neighbors, sharded=sharded)
    return True


def _is_tensor(x):
    return isinstance(x, torch.Tensor)


def _embed_src_dst(model, src, dst, t, num_neighbors):
    """Both endpoints of every event through one launch chain.  Host numpy arrays take the drop-in call
    (``compute_src_dst_node_temporal_embeddings``: pinned staging + H2D inside); device tensors (int64 ids,
    float64 / float32 times already resident in HBM) skip the copies."""
    if _is_tensor(src):
        b = src.shape[0]
        both = model.compute_node_temporal_embeddings(torch.cat([src, dst]), torch.cat([t, t]), model.num_layers,
                                                      num_neighbors)
        return both[:b], both[b:]
    return model.compute_src_dst_node_temporal_embeddings(src, dst, t, num_neighbors)


def _scatter_all_reduce(rows: torch.Tensor, index: torch.Tensor, total: int, dist):
    """Every row of the [total, w] result is produced by exactly one rank: write the local rows into a zeroed
    table and sum the tables (x + 0 is exact)."""
    full = rows.new_zeros((total,) + tuple(rows.shape[1:]))
    full.index_copy_(0, index, rows)
    dist.all_reduce(full)
    return full


def _owned_roots(model, src, dst, t, num_neighbors, dist, rank, world):
    """Owner-partitioned pass, first half: build the sharded layer memo, route the root queries of this rank's
    contiguous event slice to the ranks that own their nodes, embed the roots this rank owns.
    Returns (embeddings [n_own, dn], global root index [n_own]: event i's source is i, its destination E + i,
    n_src): the first ``n_src`` owned roots are source endpoints, the rest destinations."""
    from . import _lib
    from .shard import route_roots
    from .tgat import shard_plan
    e = len(src)
    dev = model.node_raw_features.device
    lo, hi, _ = shard_bounds(e, rank, world)
    with torch.no_grad(), torch.cuda.device(dev):
        t0 = _mark("start", time.perf_counter()) if _TRACE else 0.0
        plan = shard_plan(model._engine, model.neighbor_sampler, dev)
        # Inputs that did not change since the last pass (the E-step embeds the same event list in every EM iteration)
        # keep their routing: it depends on the events and the partition only.  Device tensors are recognised by
        # (pointer, version), host arrays by (pointer, shape) plus a sum / xor fingerprint of this rank's slice; the
        # ranks agree on a hit with one small all-reduce, since a miss anywhere re-routes everywhere.  The agreement
        # runs on the hosts (ShardPlan.all_agree) after the memo build has been enqueued, so it overlaps device work.
        model._engine.overlap_exchange = not _TRACE      # an embedding call follows: it waits for the exchange itself
        try:
            prepare_layer_memo(model, 2 * e, num_neighbors, True)
        finally:
            model._engine.overlap_exchange = False
        t0 = _mark("memo_build+exchange", t0)
        if _is_tensor(src):
            key = tuple((x.data_ptr(), x._version, tuple(x.shape), x.dtype) for x in (src, dst, t))
        else:
            def fp(a):
                v = np.ascontiguousarray(a[lo:hi]).view(np.uint64)
                return (a.ctypes.data, a.shape, a.dtype.str, int(v.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(v)) if v.size else 0)
            key = tuple(fp(np.asarray(x)) for x in (src, dst, t)) if all(np.asarray(x).dtype.itemsize == 8 for x in (src, dst, t)) else None
        routed = plan.routing.get(key) if key is not None else None
        if not plan.all_agree(routed is not None, dist):
            routed = None
        if routed is None:
            if _is_tensor(src):
                s_loc, d_loc, t_loc = src[lo:hi], dst[lo:hi], t[lo:hi]
                is32 = t.dtype == torch.float32
            else:
                is32 = np.asarray(t).dtype == np.float32
                s_loc = _lib.to_device(src[lo:hi], np.int64, dev, "p_src")
                d_loc = _lib.to_device(dst[lo:hi], np.int64, dev, "p_dst")
                t_loc = _lib.to_device(t[lo:hi], np.float64, dev, "p_t")      # float32 -> float64 is exact
            ev = torch.arange(lo, hi, device=dev, dtype=torch.int64)
            t64 = t_loc.to(torch.float64)
            nodes, times, gidx = route_roots(torch.cat([s_loc.to(torch.int64), d_loc.to(torch.int64)]),
                                             torch.cat([t64, t64]), torch.cat([ev, ev + e]), plan.node_inner, world, dist)
            # source endpoints first: the single-way decoder then reads a prefix instead of a gather; inside each
            # half the roots are kept in (node, time) order, the order the bulk kernels want (windows of one node's
            # adjacency list stay in cache), so the per-pass sort of the embedding call is skipped for them
            is_dst = gidx >= e
            order = torch.argsort(times, stable=True)
            order = order[torch.argsort(nodes[order], stable=True)]
            order = order[torch.argsort(is_dst[order], stable=True)]
            nodes, times, gidx = nodes[order].contiguous(), times[order].contiguous(), gidx[order].contiguous()
            n_src = int(gidx.numel() - int(is_dst.sum()))
            if is32:
                times = times.to(torch.float32)       # the recursion's dtype rule follows the caller's dtype
            routed = (nodes, times, gidx, n_src)
            if key is not None:
                plan.routing.clear()
                plan.routing[key] = routed
        nodes, times, gidx, n_src = routed
        t0 = _mark("route_roots", t0)
        model._engine.presorted = True
        try:
            emb = model.compute_node_temporal_embeddings(nodes, times, model.num_layers, num_neighbors)
        finally:
            model._engine.presorted = False
        t0 = _mark("embed_roots", t0)
    return emb, gidx, n_src


def embed_events(model, src_node_ids, dst_node_ids, node_interact_times, num_neighbors: int = 20, sharded=None):
    """Embeddings of every event's source and destination node at the event time:
    two float32 [E, dn] device tensors, rows in event order (what the reference's full pass
    accumulates batch by batch and copies into ``src_node_embeddings`` / ``dst_node_embeddings``).
    Inputs are host numpy arrays (the reference's types) or device tensors.
    With torch.distributed initialised (or ``sharded=True``) the pass is owner-partitioned
    (flid_b200/shard.py) and the rows are combined on every rank."""
    _require_eval(model)
    dist, rank, world = _dist()
    if sharded is None:
        sharded = world > 1
    if _is_tensor(src_node_ids):
        src, dst, t = src_node_ids, dst_node_ids, node_interact_times
    else:
        src, dst, t = np.asarray(src_node_ids), np.asarray(dst_node_ids), np.asarray(node_interact_times)
    e = len(src)
    eng = getattr(model, "_engine", None)
    if eng is None:                 # a backbone without the TGAT engine (GraphMixer, TCL): plain full-size call
        return model.compute_src_dst_node_temporal_embeddings(np.asarray(src), np.asarray(dst), np.asarray(t), num_neighbors)
    bulk = 2 * e >= BULK_MIN_ROOTS
    try:
        eng.ln_fold = bulk          # whole-pass calls fold LayerNorm into fc1: one path for every chunk of the pass
        if not sharded or world == 1:
            with torch.no_grad():
                prepare_layer_memo(model, 2 * e, num_neighbors, False)
                return _embed_src_dst(model, src, dst, t, num_neighbors)
        emb, gidx, _ = _owned_roots(model, src, dst, t, num_neighbors, dist, rank, world)
        full = _scatter_all_reduce(emb, gidx, 2 * e, dist)
    finally:
        eng.shard_tag = None
        eng.ln_fold = False
    return full[:e], full[e:]


def e_step_pass(model, decoder: MLPClassifier, src_node_ids, dst_node_ids, node_interact_times,
                num_neighbors: int = 20, pseudo_labels_store=None, ps_filter: str = 'entropy', threshold: float = 0.9,
                sharded=None, return_embeddings: bool = False, double_way: bool = False):
    """Embedding pass + pseudo-label emission + EST/CST filter.

    Single-way datasets (``double_way=False``): returns (pseudo_labels float32 [1, E] with -1 marks,
    probabilities float32 [E, C], (src_emb, dst_emb) or None).  Double-way datasets (PTCL/E_step.py:318-343:
    labels for both endpoints): pseudo_labels [2, E], probabilities [2, E, C] (row 0 = sources).
    ``pseudo_labels_store`` is the reference's list of earlier iterations' probabilities; this pass's
    probabilities are appended to it before filtering (PTCL/E_step.py:351 then PTCL/utils.py:80-83).
    When sharded, only (label, probs) rows travel between the ranks (12 B/event/endpoint for C=2) unless
    the embeddings are requested."""
    _require_eval(model, decoder)
    dist, rank, world = _dist()
    if sharded is None:
        sharded = world > 1
    if _is_tensor(src_node_ids):
        src, dst, t = src_node_ids, dst_node_ids, node_interact_times
    else:
        src, dst, t = np.asarray(src_node_ids), np.asarray(dst_node_ids), np.asarray(node_interact_times)
    e = len(src)
    ways = 2 if double_way else 1
    store = pseudo_labels_store if pseudo_labels_store is not None else []
    if not sharded or world == 1:
        src_emb, dst_emb = embed_events(model, src, dst, t, num_neighbors, sharded=False)
        l_all, p_all = emit_pseudo_labels(decoder, torch.cat([src_emb, dst_emb]) if double_way else src_emb)
        labels = l_all.to(torch.float32).reshape(ways, e)
        probs = p_all.reshape(ways, e, -1)
        emb = (src_emb, dst_emb)
    else:
        try:
            model._engine.ln_fold = 2 * e >= BULK_MIN_ROOTS
            own, gidx, n_src = _owned_roots(model, src, dst, t, num_neighbors, dist, rank, world)
            if not double_way:                      # only the source endpoints are scored (PTCL/E_step.py:327-331)
                own_s, gidx_s = own[:n_src], gidx[:n_src]
            else:
                own_s, gidx_s = own, gidx
            t0 = time.perf_counter()
            l_loc, p_loc = emit_pseudo_labels(decoder, own_s)
            packed = torch.cat([l_loc.to(torch.float32).unsqueeze(1), p_loc], dim=1)      # [n_own, 1 + C]
            full = _scatter_all_reduce(packed, gidx_s, ways * e, dist).reshape(ways, e, -1)
            _mark("decode+combine", t0)
            labels, probs = full[:, :, 0].contiguous(), full[:, :, 1:].contiguous()
            emb = None
            if return_embeddings:
                both = _scatter_all_reduce(own, gidx, 2 * e, dist)
                emb = (both[:e], both[e:])
        finally:
            model._engine.shard_tag = None
            model._engine.ln_fold = False
    if not double_way:
        probs = probs[0]
    store.append(probs)
    pseudo = labels.contiguous()
    t0 = time.perf_counter()
    if ps_filter == 'entropy':
        pseudo = entropy_filter(pseudo, store, threshold)
    elif ps_filter == 'probability':
        pseudo = prob_filter(pseudo, store, threshold)
    _mark("filter", t0)
    return pseudo, probs, (emb if return_embeddings or not sharded else None)


def tgn_pass(model, src_node_ids, dst_node_ids, node_interact_times, edge_ids, batch_size: int = 200,
             num_neighbors: int = 20, per_batch_calls: bool = False):
    """Full chronological TGN pass (PTCL/M_step.py:454-509 with model_name='TGN'): the bank is
    reset, events are fed in batches of ``batch_size`` (the batch boundary is part of the
    semantics) and the per-event embeddings are returned as [E, dn] x 2.
    One C call with a CUDA graph per batch (``MemoryModel.embed_pass``); ``per_batch_calls=True`` issues the
    reference's own loop of ``compute_src_dst_node_temporal_embeddings`` calls instead (same results, bit for bit)."""
    src, dst = np.asarray(src_node_ids), np.asarray(dst_node_ids)
    t, eid = np.asarray(node_interact_times), np.asarray(edge_ids)
    e = len(src)
    dev = model.node_raw_features.device
    _require_eval(model)
    model.memory_bank.__init_memory_bank__()
    if not per_batch_calls:
        return model.embed_pass(src, dst, t, eid, batch_size, num_neighbors)
    out_s = torch.empty((e, model.node_feat_dim), dtype=torch.float32, device=dev)
    out_d = torch.empty_like(out_s)
    with torch.no_grad():
        for lo in range(0, e, batch_size):
            hi = min(lo + batch_size, e)
            a, b = model.compute_src_dst_node_temporal_embeddings(src[lo:hi], dst[lo:hi], t[lo:hi], eid[lo:hi], True,
                                                                  num_neighbors)
            out_s[lo:hi], out_d[lo:hi] = a, b
    return out_s, out_d
