"""Drop-in for the reference's ``models.TGAT.TGAT`` (``models/TGAT.py``) and the building
blocks it owns (``models/modules.py``: TimeEncoder :7-40, MultiHeadAttention :126-245,
MergeLayer :43-69), with the same constructor, method names, error behaviour and
``state_dict`` keys / shapes, so reference checkpoints load unchanged.

Inference (``torch.no_grad()`` calls in ``.eval()`` mode: every E-step pass, the full
embedding passes, evaluation) runs the sm_100a kernels through the C ABI
(``flid_tgat_embed``).  The parameter-holder modules below keep the reference's parameter
names and default initialisation; they have no forward of their own on that path.
Training-mode calls (dropout + backward) use ``train.autograd_forward``: device-sampled
neighbourhoods (the same bit-exact kernel), the attention stream as a CUDA kernel with a
hand-written backward kernel, and torch matmuls for the dense projections.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .sampler import NeighborSampler
from .train import autograd_forward


NUMERIC_MODES = {"f32": 0, "bf16": 1}
_numeric_mode = os.environ.get("FLID_NUMERIC", "f32")
if _numeric_mode not in NUMERIC_MODES:
    raise ValueError(f"FLID_NUMERIC must be one of {sorted(NUMERIC_MODES)}, got {_numeric_mode!r}")


def set_numeric_mode(mode: str = "f32"):
    """Numeric mode of the projection GEMMs of every flid_b200 model in this process (BASELINE.json north_star):
    "f32" (default): fp32-grade products, embeddings within fp32 rel 1e-4 of the reference;
    "bf16": operands rounded to bfloat16, fp32 accumulation (rel 2e-2, identical argmax on >= 99.9 % of nodes).
    Models pick the change up at their next call (weights are re-tiled, derived caches dropped)."""
    global _numeric_mode
    if mode not in NUMERIC_MODES:
        raise ValueError(f"numeric mode must be one of {sorted(NUMERIC_MODES)}, got {mode!r}")
    _numeric_mode = mode


def get_numeric_mode() -> str:
    return _numeric_mode


class TimeEncoder(nn.Module):
    """cos(w * t + b); w initialised to 1 / 10^linspace(0, 9, T), b to 0 (models/modules.py:7-26)."""

    def __init__(self, time_dim: int, parameter_requires_grad: bool = True):
        super().__init__()
        self.time_dim = time_dim
        self.w = nn.Linear(1, time_dim)
        self.w.weight = nn.Parameter(torch.from_numpy(
            1 / 10 ** np.linspace(0, 9, time_dim, dtype=np.float32)).reshape(time_dim, -1))
        self.w.bias = nn.Parameter(torch.zeros(time_dim))
        if not parameter_requires_grad:
            self.w.weight.requires_grad = False
            self.w.bias.requires_grad = False


class MultiHeadAttention(nn.Module):
    """Parameter holder with the reference's names/shapes (models/modules.py:128-165)."""

    def __init__(self, node_feat_dim: int, edge_feat_dim: int, time_feat_dim: int, num_heads: int = 2,
                 dropout: float = 0.1):
        super().__init__()
        self.node_feat_dim, self.edge_feat_dim, self.time_feat_dim = node_feat_dim, edge_feat_dim, time_feat_dim
        self.num_heads = num_heads
        self.query_dim = node_feat_dim + time_feat_dim
        self.key_dim = node_feat_dim + edge_feat_dim + time_feat_dim
        assert self.query_dim % num_heads == 0, \
            "The sum of node_feat_dim and time_feat_dim should be divided by num_heads!"
        self.head_dim = self.query_dim // num_heads
        self.query_projection = nn.Linear(self.query_dim, num_heads * self.head_dim, bias=False)
        self.key_projection = nn.Linear(self.key_dim, num_heads * self.head_dim, bias=False)
        self.value_projection = nn.Linear(self.key_dim, num_heads * self.head_dim, bias=False)
        self.scaling_factor = self.head_dim ** -0.5
        self.layer_norm = nn.LayerNorm(self.query_dim)
        self.residual_fc = nn.Linear(num_heads * self.head_dim, self.query_dim)
        self.dropout = nn.Dropout(dropout)


class MergeLayer(nn.Module):
    """fc2(relu(fc1([x1 | x2]))) parameter holder (models/modules.py:45-56)."""

    def __init__(self, input_dim1: int, input_dim2: int, hidden_dim: int, output_dim: int):
        super().__init__()
        self.fc1 = nn.Linear(input_dim1 + input_dim2, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        self.act = nn.ReLU()


class _Engine:
    """C-ABI TGAT handles (one per recursion depth actually requested) kept in sync with
    the owning module's parameters."""

    def __init__(self, node_dim, edge_dim, time_dim, num_heads):
        self.dims = (node_dim, edge_dim, time_dim, num_heads)
        self.handles = {}    # depth -> c_void_p
        self.versions = {}   # depth -> parameter fingerprint
        self.tables = {}     # depth -> (data_ptr, rows) of the cached node table
        self.param_cache = {}   # depth -> [Parameter, ...] in the order flid_tgat_set_weights expects
        self.memo = {}       # depth -> (key, [level tables])   layer memo of bulk passes
        self.memo_mode = "auto"   # False | True | "auto"
        self.memo_builds = 0      # how many times a memo was (re)built (diagnostics)
        self.build_stats = None   # set to [0, 0, 0] to accumulate (evaluations, valid slots, queries) over memo builds
        self.served = {}     # depth -> (key, root queries answered without a memo)
        self.epoch = 0       # bumped by invalidate(): part of every cache key
        # projected (per-entry K/V) formulation of bulk calls, flid_tgat_set_bulk_projection.  Off by default: measured
        # on the B200 it removes a third of the out-projection work but pays it back in per-entry
        # GEMMs, 31.0 vs 30.8 ms per Reddit-shape pass in the fp32 mode; FLID_BULK_KV=1 / set_bulk_projection(True)
        self.bulk_projection = os.environ.get("FLID_BULK_KV", "0") == "1"
        self.shard_tag = None         # (rank, world) while an owner-partitioned pass is running (flid_b200.passes)
        self.ln_fold = False          # bulk passes fold LayerNorm into fc1 (flid_tgat_set_ln_fold); set by flid_b200.passes
        self.ln_state = {}
        self.overlap_exchange = False   # set by passes._owned_roots around the sharded memo build
        self.pending_exchange = None
        self.presorted = False     # the caller's bulk roots already arrive in (node, time) order (passes._owned_roots)
        self.sort_state = {}            # depth -> value last handed to the C handle
        self.shard_plans = {}         # (sampler generation, rank, world) -> shard.ShardPlan

    def __deepcopy__(self, memo):
        """copy.deepcopy(model): the copy gets its own (empty) engine -- C handles and device caches are rebuilt on its
        first call; they are never shared between module instances."""
        new = _Engine(*self.dims)
        new.memo_mode, new.bulk_projection = self.memo_mode, self.bulk_projection
        return new

    def invalidate(self):
        """Forget the uploaded weights, the cached node table and the layer memo.  Needed after writes that
        autograd's version counter does not see (``param.data.copy_()``, ``.data`` mutation, feature tables
        edited in place through ``.data``); ``load_state_dict`` / ``.to()`` / optimizer steps are detected."""
        self.epoch += 1
        self.shard_tag = None
        self.versions.clear()
        self.tables.clear()
        self.memo.clear()
        self.served.clear()

    def close(self):
        for h in self.handles.values():
            try:
                _lib.lib().flid_tgat_free(h)
            except Exception:
                pass
        self.handles.clear()
        self.ln_state.clear()

    def handle(self, depth, time_encoder, conv_layers, merge_layers, device):
        lib = _lib.lib()
        # nn.Module attribute access is slow on a ~170-us call path, so the (owning dict, name, Parameter) triples
        # are collected once per depth; a replaced Parameter object (rare) is caught by the identity check
        cached = self.param_cache.get(depth)
        if cached is None or any(d[k] is not p for d, k, p in cached):
            mods = [(time_encoder.w, ("weight", "bias"))]
            for l in range(depth):
                a, m = conv_layers[l], merge_layers[l]
                mods += [(a.query_projection, ("weight",)), (a.key_projection, ("weight",)), (a.value_projection, ("weight",)),
                         (a.layer_norm, ("weight", "bias")), (a.residual_fc, ("weight", "bias")),
                         (m.fc1, ("weight", "bias")), (m.fc2, ("weight", "bias"))]
            cached = [(mod._parameters, k, mod._parameters[k]) for mod, keys in mods for k in keys]
            self.param_cache[depth] = cached
        params = [p for _, _, p in cached]
        fp = tuple([(p.data_ptr(), p._version) for p in params]) + (_numeric_mode, self.bulk_projection)
        if self.versions.get(depth) != fp:
            for p in params:
                if p.device != device:
                    raise RuntimeError(f"flid_b200: parameters live on {p.device} but the feature tables are on "
                                       f"{device}; move the model with .to(device) first (no CPU fallback)")
                if p.dtype != torch.float32:
                    raise RuntimeError("flid_b200: parameters must be float32")
        if depth not in self.handles:
            h = C.c_void_p(None)
            dn, de, T, H = self.dims
            _lib.check(lib.flid_tgat_create(dn, de, T, depth, H, C.byref(h)))
            self.handles[depth] = h
        h = self.handles[depth]
        if self.versions.get(depth) != fp:
            layers = (_lib.LayerWeights * depth)()
            keep = []
            for l in range(depth):
                chunk = [p.detach().contiguous() for p in params[2 + 11 * l: 2 + 11 * (l + 1)]]
                keep += chunk
                for name, t in zip([f[0] for f in _lib.LayerWeights._fields_], chunk):
                    setattr(layers[l], name, t.data_ptr())
            tw, tb = params[0].detach().contiguous(), params[1].detach().contiguous()
            _lib.check(lib.flid_tgat_set_numeric_mode(h, NUMERIC_MODES[_numeric_mode]))
            _lib.check(lib.flid_tgat_set_bulk_projection(h, 1 if self.bulk_projection else 0))
            _lib.check(lib.flid_tgat_set_weights(h, _lib.ptr(tw), _lib.ptr(tb), layers, _lib.stream()))
            self.versions[depth] = fp
            self.tables.pop(depth, None)
            self.memo.pop(depth, None)
        want = bool(self.ln_fold) and os.environ.get("FLID_LN_FOLD", "1") != "0"
        if self.ln_state.get(depth) != want:
            _lib.check(lib.flid_tgat_set_ln_fold(h, 1 if want else 0))
            self.ln_state[depth] = want
            self.tables.pop(depth, None)      # the node table is re-cached with / without the per-node fc1 tables
        return h

    def ensure_table(self, depth, h, node_feat):
        """Cache the layer-1 query fold of the (static) node feature table once per weight version."""
        key = (node_feat.data_ptr(), node_feat.shape[0], node_feat._version)
        if self.tables.get(depth) != key:
            _lib.check(_lib.lib().flid_tgat_cache_node_table(h, _lib.ptr(node_feat), node_feat.shape[0],
                                                             _lib.stream()))
            self.tables[depth] = key


def _memo_key(engine, depth, sampler, node_feat, edge_feat, k):
    return (engine.versions.get(depth), engine.epoch, sampler.generation, node_feat.data_ptr(), node_feat._version,
            edge_feat.data_ptr(), edge_feat._version, int(k), engine.shard_tag)


def memo_piece_bounds(rows, rank, world, pieces):
    """Row ranges of a table-order sharded memo build (kept for callers that split a build by table rows, e.g. the
    tests of flid_tgat_memo_build's row-range form).  The table (``per * world * pieces`` rows, ``per`` =
    ceil(rows / (world * pieces))) is cut into ``pieces`` super-blocks of ``world * per`` rows; inside
    super-block p rank r owns rows [p*world*per + r*per, ... + per), clipped to ``rows``.
    Returns (per, [(block_start, lo, hi), ...])."""
    per = -(-rows // (world * pieces))
    out = []
    for piece in range(pieces):
        base = piece * per * world
        out.append((base, min(base + rank * per, rows), min(base + (rank + 1) * per, rows)))
    return per, out


def shard_plan(engine, sampler, device):
    """The owner partition of this rank (flid_b200.shard.ShardPlan), cached per sampler."""
    import torch.distributed as dist
    from .shard import ShardPlan
    rank, world = dist.get_rank(), dist.get_world_size()
    key = (sampler.generation, rank, world)
    plan = engine.shard_plans.get(key)
    if plan is None:
        engine.shard_plans.clear()
        plan = ShardPlan(sampler, rank, world, device)
        engine.shard_plans[key] = plan
    return plan


def build_layer_memo(engine, depth, time_encoder, conv_layers, merge_layers, sampler, node_feat, edge_feat,
                     num_neighbors, sharded=False):
    """Fill the layer memo (include/flid_b200.h, flid_tgat_memo_build) for levels 1..depth-1.
    ``sharded`` (torch.distributed initialised, called by every rank): owner-partitioned build -- each rank
    evaluates the work items of its own position range in owner-major order and the rows that belong to other
    ranks' ranges are exchanged once per level (flid_b200/shard.py); the resulting table is complete for this
    rank's range only and is used through ``engine.shard_tag``."""
    if depth < 2:
        return None
    device = node_feat.device
    lib = _lib.lib()
    with torch.cuda.device(device):
        h = engine.handle(depth, time_encoder, conv_layers, merge_layers, device)
        engine.ensure_table(depth, h, node_feat)
        plan = None
        if sharded:
            import torch.distributed as dist
            if dist.get_world_size() > 1:
                plan = shard_plan(engine, sampler, device)
        engine.shard_tag = (plan.rank, plan.world) if plan is not None else None
        _lib.check(lib.flid_tgat_set_bulk_range(h, plan.pos_lo if plan else 0, plan.pos_hi if plan else -1))
        key = _memo_key(engine, depth, sampler, node_feat, edge_feat, num_neighbors)
        have = engine.memo.get(depth)
        if have is not None and have[0] == key:
            return have[1]
        engine.memo.pop(depth, None)
        if engine.pending_exchange is not None:          # a side-stream exchange nobody waited for (no layer-2 call followed)
            torch.cuda.current_stream().wait_event(engine.pending_exchange)
            _lib.check(lib.flid_tgat_set_wait_event(h, None))
            engine.pending_exchange = None
        _lib.check(lib.flid_tgat_bulk_invalidate(h))     # new memo contents: the projected per-entry tables follow
        rows = sampler.num_entries + 1
        dn = node_feat.shape[1]
        tables = []
        prev = None
        for level in range(1, depth):
            ent = plan.peer_table(level, rows, dn, device, dist) if plan is not None else None
            t = ent[0] if ent is not None else torch.empty((rows, dn), dtype=torch.float32, device=device)
            if plan is None:
                _lib.check(lib.flid_tgat_memo_build(h, sampler.handle, _lib.ptr(node_feat), _lib.ptr(edge_feat),
                                                    int(num_neighbors), level, _lib.ptr(prev), 0, rows, _lib.ptr(t),
                                                    _lib.stream()))
            else:
                _lib.check(lib.flid_tgat_memo_build_owner_range(h, sampler.handle, _lib.ptr(node_feat),
                                                                _lib.ptr(edge_feat), int(num_neighbors), level,
                                                                _lib.ptr(prev), plan.pos_lo, plan.pos_hi, 1, _lib.ptr(t),
                                                                _lib.stream()))
            if engine.build_stats is not None:      # measurement runs only: reading the counters synchronises
                st = (C.c_int64 * 4)()
                _lib.check(lib.flid_tgat_last_stats(h, st))
                engine.build_stats = [a + int(b) for a, b in zip(engine.build_stats, st[:3])]
            if plan is not None:
                if os.environ.get("FLID_PASS_TRACE") == "1":
                    from . import passes as _p
                    import time as _time
                    torch.cuda.synchronize()
                    _t0 = _time.perf_counter()
                    plan.exchange_rows_p2p(sampler, ent, dist) if ent is not None else plan.exchange_rows(t, dist)
                    _p._mark("exchange", _t0)
                elif ent is not None and engine.overlap_exchange and level == depth - 1 and os.environ.get("FLID_XCHG_OVERLAP", "1") != "0":
                    # last level inside an owner-partitioned pass: exchange + barrier on a side stream; the embedding
                    # call that follows waits for it before its first read of exchanged rows (flid_tgat_set_wait_event)
                    side = plan.side_stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        plan.exchange_rows_p2p(sampler, ent, dist)
                        done = torch.cuda.Event()
                        done.record(side)
                    _lib.check(lib.flid_tgat_set_wait_event(h, C.c_void_p(done.cuda_event)))
                    engine.pending_exchange = done          # kept alive until the next one replaces it
                elif ent is not None:
                    plan.exchange_rows_p2p(sampler, ent, dist)
                else:
                    plan.exchange_rows(t, dist)
            tables.append(t)
            prev = t
        engine.memo[depth] = (key, tables)
        engine.memo_builds += 1
        return tables


def _memo_for_call(engine, depth, time_encoder, conv_layers, merge_layers, sampler, node_feat, edge_feat, k, n):
    """The valid memo tables for this call, or None.  In "auto" mode the memo is built once
    the root queries answered at this weight version would have paid for it: a plain root
    costs sum_l (1+k)^(depth-l) attention evaluations, a memoised one ``depth``, the build
    (depth-1) * (entries+1) (break-even rule, at most 2x the optimum)."""
    mode = engine.memo_mode
    if depth < 2 or not mode:
        return None
    key = _memo_key(engine, depth, sampler, node_feat, edge_feat, k)    # includes engine.shard_tag
    have = engine.memo.get(depth)
    if have is not None and have[0] == key:
        return have[1]
    if mode == "auto":
        skey, served = engine.served.get(depth, (None, 0))
        if skey != key:
            served = 0
        plain = sum((1 + k) ** (depth - l) for l in range(1, depth + 1))
        build_cost = (depth - 1) * (sampler.num_entries + 1)
        if (served + n) * (plain - depth) < build_cost:
            engine.served[depth] = (key, served + n)
            return None
        need = build_cost * node_feat.shape[1] * 4
        if need > torch.cuda.mem_get_info(node_feat.device)[0] // 2:     # only queried once the memo would pay off
            engine.served[depth] = (key, served + n)
            return None
    return build_layer_memo(engine, depth, time_encoder, conv_layers, merge_layers, sampler, node_feat, edge_feat, k)


def embed_roots(engine, depth, time_encoder, conv_layers, merge_layers, sampler, node_feat, edge_feat,
                node_ids, node_interact_times, num_neighbors, use_table=True, use_memo=False):
    """Shared by TGAT and MemoryModel: n root queries -> float32 [n, dn] on the device."""
    device = node_feat.device
    _lib.require_cuda(device)
    if not isinstance(sampler, NeighborSampler):
        raise TypeError("flid_b200 models need a flid_b200.NeighborSampler (device CSR); got "
                        f"{type(sampler).__name__}")
    if sampler.device != device:
        raise RuntimeError(f"neighbor sampler lives on {sampler.device}, model on {device}")
    with torch.cuda.device(device):
        h = engine.handle(depth, time_encoder, conv_layers, merge_layers, device)
        if use_table:
            engine.ensure_table(depth, h, node_feat)
        if isinstance(node_ids, torch.Tensor):
            d_nodes = node_ids.to(device, torch.int64).contiguous()
        else:
            d_nodes = _lib.to_device(node_ids, np.int64, device, "e_nodes", sync_follows=True)
        if isinstance(node_interact_times, torch.Tensor):
            is32 = 1 if node_interact_times.dtype == torch.float32 else 0
            d_times = node_interact_times.to(device, torch.float64).contiguous()
        else:
            t = np.asarray(node_interact_times)
            is32 = 1 if t.dtype == np.float32 else 0
            d_times = _lib.to_device(t, np.float64, device, "e_times", sync_follows=True)   # float32 -> float64 is exact
        n = d_nodes.shape[0]
        out = torch.empty((n, node_feat.shape[1]), dtype=torch.float32, device=device)
        memo = None
        if use_memo and int(num_neighbors) > 0:
            memo = _memo_for_call(engine, depth, time_encoder, conv_layers, merge_layers, sampler, node_feat,
                                  edge_feat, int(num_neighbors), n)
        if memo is not None:
            tabs = (C.c_void_p * len(memo))(*[t.data_ptr() for t in memo])
            want_sort = not engine.presorted
            if engine.sort_state.get(depth, True) != want_sort:
                _lib.check(_lib.lib().flid_tgat_set_sort_queries(h, 1 if want_sort else 0))
                engine.sort_state[depth] = want_sort
            _lib.check(_lib.lib().flid_tgat_embed_memo(h, sampler.handle, _lib.ptr(node_feat), _lib.ptr(edge_feat),
                                                       tabs, _lib.ptr(d_nodes), _lib.ptr(d_times), is32, n,
                                                       int(num_neighbors), _lib.ptr(out), _lib.stream()))
        else:
            _lib.check(_lib.lib().flid_tgat_embed(h, sampler.handle, _lib.ptr(node_feat), _lib.ptr(edge_feat),
                                                  _lib.ptr(d_nodes), _lib.ptr(d_times), is32, n, int(num_neighbors),
                                                  _lib.ptr(out), _lib.stream()))
    return out


class TGAT(nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, num_layers: int = 2, num_heads: int = 2, dropout: float = 0.1, device: str = 'cpu'):
        """Same arguments as models/TGAT.py:11-48.  ``device`` must be a CUDA device."""
        super().__init__()
        self.node_raw_features = torch.from_numpy(np.ascontiguousarray(node_raw_features.astype(np.float32))).to(device)
        self.edge_raw_features = torch.from_numpy(np.ascontiguousarray(edge_raw_features.astype(np.float32))).to(device)
        self.neighbor_sampler = neighbor_sampler
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.edge_feat_dim = self.edge_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.dropout = dropout
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim)
        self.temporal_conv_layers = nn.ModuleList([
            MultiHeadAttention(self.node_feat_dim, self.edge_feat_dim, self.time_feat_dim, self.num_heads, self.dropout)
            for _ in range(num_layers)])
        self.merge_layers = nn.ModuleList([
            MergeLayer(self.node_feat_dim + self.time_feat_dim, self.node_feat_dim, self.node_feat_dim,
                       self.node_feat_dim) for _ in range(num_layers)])
        self._engine = _Engine(self.node_feat_dim, self.edge_feat_dim, self.time_feat_dim, self.num_heads)

    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.close()

    def invalidate_caches(self):
        """Drop every derived device cache (see ``_Engine.invalidate``)."""
        self._engine.invalidate()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.invalidate()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._engine.invalidate()
        return out

    def _needs_autograd(self):
        """Training-mode calls need dropout and a backward pass; eval-mode calls with grad enabled need the
        graph too if any parameter requires grad (the fused kernels are forward-only)."""
        if self.training:
            return True
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray,
                                                 node_interact_times: np.ndarray, num_neighbors: int = 20):
        """models/TGAT.py:50-66 -> (Tensor[B, dn], Tensor[B, dn]) on the model's device.
        src and dst roots are independent, so both halves go through one launch chain."""
        b = len(src_node_ids)
        dev = self.node_raw_features.device
        t = np.asarray(node_interact_times)
        if self._needs_autograd() or dev.type != "cuda" or t.dtype == np.float32:
            both = self.compute_node_temporal_embeddings(
                np.concatenate([np.asarray(src_node_ids), np.asarray(dst_node_ids)]), np.concatenate([t, t]),
                self.num_layers, num_neighbors)
        else:
            # stage [src ; dst] straight into pinned memory (no host-side concatenation), send the times once
            with torch.cuda.device(dev):
                d_nodes = _lib.to_device_concat([src_node_ids, dst_node_ids], np.int64, dev, "e_nodes", sync_follows=True)
                d_t = _lib.to_device(t, np.float64, dev, "e_times", sync_follows=True)
                d_times = torch.cat([d_t, d_t])
            both = self.compute_node_temporal_embeddings(d_nodes, d_times, self.num_layers, num_neighbors)
        return both[:b], both[b:]

    def compute_node_temporal_embeddings(self, node_ids: np.ndarray, node_interact_times: np.ndarray,
                                         current_layer_num: int, num_neighbors: int = 20):
        """models/TGAT.py:68-144."""
        assert current_layer_num >= 0
        if current_layer_num == 0:
            idx = torch.as_tensor(np.asarray(node_ids), dtype=torch.int64, device=self.node_raw_features.device)
            return self.node_raw_features[idx]
        if current_layer_num > self.num_layers:
            raise IndexError("current_layer_num exceeds num_layers")
        if self._needs_autograd():
            _lib.require_cuda(self.node_raw_features.device)
            return autograd_forward(self.time_encoder, self.temporal_conv_layers, self.merge_layers,
                                    self.neighbor_sampler, self.node_raw_features, self.edge_raw_features,
                                    node_ids, node_interact_times, current_layer_num, num_neighbors, self.training)
        return embed_roots(self._engine, current_layer_num, self.time_encoder, self.temporal_conv_layers,
                           self.merge_layers, self.neighbor_sampler, self.node_raw_features, self.edge_raw_features,
                           node_ids, node_interact_times, num_neighbors, use_memo=True)

    # ---- layer memo of bulk passes (not in the reference API; see include/flid_b200.h) ----
    def set_layer_memo(self, mode="auto"):
        """False: never memoise; True: build on the next call; "auto" (default): build once the
        root queries answered at the current weights would have paid for the build."""
        assert mode in (False, True, "auto")
        self._engine.memo_mode = mode
        if not mode:
            self._engine.memo.clear()

    def set_bulk_projection(self, enable: bool = True):
        """``True``: bulk calls (memo build, whole-pass embedding) project every adjacency entry once per pass and
        stream the projected rows (csrc/bulk_kv.cu); results agree with the per-slot path to fp32 rounding.
        ``False`` (default) keeps the per-slot stream everywhere: memoised results equal the recursion bit for bit."""
        self._engine.bulk_projection = bool(enable)
        self._engine.memo.clear()

    def build_layer_memo(self, num_neighbors: int = 20, sharded: bool = False):
        """Explicitly (re)build the memo for full-depth calls; collective when ``sharded``."""
        return build_layer_memo(self._engine, self.num_layers, self.time_encoder, self.temporal_conv_layers,
                                self.merge_layers, self.neighbor_sampler, self.node_raw_features,
                                self.edge_raw_features, num_neighbors, sharded)

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """models/TGAT.py:146-155."""
        self.neighbor_sampler = neighbor_sampler
        if self.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.neighbor_sampler.seed is not None
            self.neighbor_sampler.reset_random_state()

    def last_stats(self, depth=None):
        """(attention evaluations, valid neighbour slots, sampler queries, workspace bytes) of the last call."""
        depth = depth or self.num_layers
        out = (C.c_int64 * 4)()
        _lib.check(_lib.lib().flid_tgat_last_stats(self._engine.handles[depth], out))
        return tuple(out)
