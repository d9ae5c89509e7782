"""Drop-in for the reference's ``models.MemoryModel.MemoryModel`` with ``model_name='TGN'``
(``models/MemoryModel.py``): same constructor, same ``compute_src_dst_node_temporal_embeddings``
signature, same ``memory_bank`` API used by the EM drivers (``__init_memory_bank__``,
``backup_memory_bank``, ``reload_memory_bank``, ``detach_memory_bank``, ``node_raw_messages``),
same ``state_dict`` keys (incl. the aliased ``memory_updater.memory_bank.*`` and
``embedding_module.time_encoder.*`` entries).

State lives in device tensors owned by ``MemoryBank``; every batch is one C-ABI call
(``flid_tgn_step``).  DyRep / JODIE / RNN updaters cannot be constructed by any reference
driver (PTCL/EM_init.py:30) and are not provided.
"""
import ctypes as C
import weakref
from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .sampler import NeighborSampler
from .tgat import MergeLayer, MultiHeadAttention, TimeEncoder, _Engine, autograd_forward


class MessageAggregator(nn.Module):
    """Last-message aggregation (models/MemoryModel.py:295-330) happens inside flid_tgn_step."""


class MemoryBank(nn.Module):
    """models/MemoryModel.py:334-459.  ``node_memories`` / ``node_last_updated_times`` are
    Parameters (requires_grad=False) so they are saved with the model, as in the reference.
    The per-node message *lists* of the reference are held as "last raw message per node"
    (the only element the TGN aggregator ever reads, :319-322)."""

    def __init__(self, num_nodes: int, memory_dim: int, message_dim: int = 0):
        super().__init__()
        self.num_nodes = num_nodes
        self.memory_dim = memory_dim
        self.message_dim = message_dim
        self.node_memories = nn.Parameter(torch.zeros((num_nodes, memory_dim)), requires_grad=False)
        self.node_last_updated_times = nn.Parameter(torch.zeros(num_nodes), requires_grad=False)
        self._alloc_state(self.node_memories.device)
        self._owner_ref = None   # weakref to the owning MemoryModel (a plain attribute, not a submodule)
        self._dirty = True       # derived state (next_memories / layer0 / query table) needs a rebuild

    def _alloc_state(self, device):
        n, d, m = self.num_nodes, self.memory_dim, self.message_dim
        self._pending_msg = torch.zeros((n, m), dtype=torch.float32, device=device)
        self._pending_ts = torch.zeros(n, dtype=torch.float64, device=device)
        self._has_pending = torch.zeros(n, dtype=torch.uint8, device=device)
        self._next_memories = torch.zeros((n, d), dtype=torch.float32, device=device)
        self._layer0 = torch.zeros((n, d), dtype=torch.float32, device=device)
        self._scratch = torch.full((n,), -1, dtype=torch.int32, device=device)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        dev = self.node_memories.device
        if self._pending_msg.device != dev:
            for name in ("_pending_msg", "_pending_ts", "_has_pending", "_next_memories", "_layer0", "_scratch"):
                setattr(self, name, getattr(self, name).to(dev))
        self._dirty = True
        return out

    def _c_state(self):
        s = _lib.TgnState()
        s.num_rows = self.num_nodes
        s.memories, s.last_updated = self.node_memories.data_ptr(), self.node_last_updated_times.data_ptr()
        s.pending_msg, s.pending_ts = self._pending_msg.data_ptr(), self._pending_ts.data_ptr()
        s.has_pending, s.next_memories = self._has_pending.data_ptr(), self._next_memories.data_ptr()
        s.layer0, s.scratch = self._layer0.data_ptr(), self._scratch.data_ptr()
        return s

    def _fingerprint(self):
        return (self.node_memories.data_ptr(), self.node_memories._version,
                self.node_last_updated_times.data_ptr(), self.node_last_updated_times._version)

    # ---- reference API
    def __init_memory_bank__(self):
        """models/MemoryModel.py:359-366: zero memories / times, drop all raw messages."""
        dev = self.node_memories.device
        if dev.type != "cuda":
            # construction time on CPU (before .to(device)): plain zero fill, nothing derived yet
            self.node_memories.data.zero_()
            self.node_last_updated_times.data.zero_()
            self._has_pending.zero_()
            self._dirty = True
            return
        owner = self._owner_ref()
        with torch.cuda.device(dev):
            s = self._c_state()
            _lib.check(_lib.lib().flid_tgn_reset(C.byref(s), _lib.ptr(owner.node_raw_features), self.memory_dim,
                                                 self.message_dim, _lib.stream()))
        self._dirty = True   # the cached query table must follow layer0

    def get_memories(self, node_ids: np.ndarray):
        return self.node_memories[torch.from_numpy(np.asarray(node_ids)).to(self.node_memories.device)]

    def set_memories(self, node_ids: np.ndarray, updated_node_memories: torch.Tensor):
        self.node_memories[torch.from_numpy(np.asarray(node_ids)).to(self.node_memories.device)] = updated_node_memories
        self._dirty = True

    def get_node_last_updated_times(self, unique_node_ids: np.ndarray):
        return self.node_last_updated_times[torch.from_numpy(np.asarray(unique_node_ids)).to(self.node_memories.device)]

    @property
    def node_raw_messages(self):
        """{node_id: [(message Tensor[msg_dim], timestamp)]} for nodes with a stored raw message."""
        out = defaultdict(list)
        ids = torch.nonzero(self._has_pending).reshape(-1)
        if ids.numel():
            msgs = self._pending_msg[ids]
            ts = self._pending_ts[ids].cpu().numpy()
            for i, v in enumerate(ids.cpu().numpy().tolist()):
                out[v].append((msgs[i], np.float64(ts[i])))
        return out

    @node_raw_messages.setter
    def node_raw_messages(self, messages):
        self._has_pending.zero_()
        ids, rows, ts = [], [], []
        for v, lst in messages.items():
            if len(lst) > 0:
                ids.append(int(v)), rows.append(torch.as_tensor(lst[-1][0])), ts.append(float(lst[-1][1]))
        if ids:
            dev = self._pending_msg.device
            idx = torch.tensor(ids, dtype=torch.int64, device=dev)
            self._pending_msg[idx] = torch.stack([r.to(dev, torch.float32) for r in rows])
            self._pending_ts[idx] = torch.tensor(ts, dtype=torch.float64, device=dev)
            self._has_pending[idx] = 1
        self._dirty = True

    def backup_memory_bank(self):
        """models/MemoryModel.py:386-396 (third element: our compact message store)."""
        return (self.node_memories.data.clone(), self.node_last_updated_times.data.clone(),
                (self._pending_msg.clone(), self._pending_ts.clone(), self._has_pending.clone()))

    def reload_memory_bank(self, backup_memory_bank: tuple):
        """models/MemoryModel.py:398-410."""
        self.node_memories.data.copy_(backup_memory_bank[0])
        self.node_last_updated_times.data.copy_(backup_memory_bank[1])
        third = backup_memory_bank[2]
        if isinstance(third, dict):
            self.node_raw_messages = third
        else:
            self._pending_msg.copy_(third[0]), self._pending_ts.copy_(third[1]), self._has_pending.copy_(third[2])
        self._dirty = True

    def detach_memory_bank(self):
        """models/MemoryModel.py:412-427; state here never carries autograd history."""
        self.node_memories.detach_()

    def store_node_raw_messages(self, node_ids, new_node_raw_messages):
        cur = self.node_raw_messages
        for v in node_ids:
            cur[int(v)].extend(new_node_raw_messages[v])
        self.node_raw_messages = cur

    def clear_node_raw_messages(self, node_ids):
        idx = torch.as_tensor(np.asarray(node_ids), dtype=torch.int64, device=self._has_pending.device)
        self._has_pending[idx] = 0
        self._dirty = True

    def extra_repr(self):
        return 'num_nodes={}, memory_dim={}'.format(self.node_memories.shape[0], self.node_memories.shape[1])


class GRUMemoryUpdater(nn.Module):
    """models/MemoryModel.py:531-543: holder of nn.GRUCell(message_dim -> memory_dim) and the bank alias."""

    def __init__(self, memory_bank: MemoryBank, message_dim: int, memory_dim: int):
        super().__init__()
        self.memory_bank = memory_bank
        self.memory_updater = nn.GRUCell(input_size=message_dim, hidden_size=memory_dim)


class GraphAttentionEmbedding(nn.Module):
    """models/MemoryModel.py:592-630: parameter holder; shares the model's TimeEncoder."""

    def __init__(self, node_raw_features, edge_raw_features, neighbor_sampler, time_encoder, node_feat_dim,
                 edge_feat_dim, time_feat_dim, num_layers=2, num_heads=2, dropout=0.1):
        super().__init__()
        self.node_raw_features, self.edge_raw_features = node_raw_features, edge_raw_features
        self.neighbor_sampler = neighbor_sampler
        self.time_encoder = time_encoder
        self.node_feat_dim, self.edge_feat_dim, self.time_feat_dim = node_feat_dim, edge_feat_dim, time_feat_dim
        self.num_layers, self.num_heads, self.dropout = num_layers, num_heads, dropout
        self.temporal_conv_layers = nn.ModuleList([
            MultiHeadAttention(node_feat_dim, edge_feat_dim, time_feat_dim, num_heads, dropout) for _ in range(num_layers)])
        self.merge_layers = nn.ModuleList([
            MergeLayer(node_feat_dim + time_feat_dim, node_feat_dim, node_feat_dim, node_feat_dim)
            for _ in range(num_layers)])


class MemoryModel(nn.Module):

    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler: NeighborSampler,
                 time_feat_dim: int, model_name: str = 'TGN', num_layers: int = 2, num_heads: int = 2, dropout: float = 0.1,
                 src_node_mean_time_shift: float = 0.0, src_node_std_time_shift: float = 1.0,
                 dst_node_mean_time_shift_dst: float = 0.0, dst_node_std_time_shift: float = 1.0, device: str = 'cpu'):
        """Same arguments as models/MemoryModel.py:12-94; only model_name='TGN' is accelerated."""
        super().__init__()
        if model_name != 'TGN':
            if model_name in ('DyRep', 'JODIE'):
                raise NotImplementedError(f"flid_b200.MemoryModel implements model_name='TGN' only (got {model_name})")
            raise ValueError(f'Not implemented error for model_name {model_name}!')
        self.node_raw_features = torch.from_numpy(np.ascontiguousarray(node_raw_features.astype(np.float32))).to(device)
        self.edge_raw_features = torch.from_numpy(np.ascontiguousarray(edge_raw_features.astype(np.float32))).to(device)
        self.node_feat_dim = self.node_raw_features.shape[1]
        self.edge_feat_dim = self.edge_raw_features.shape[1]
        self.time_feat_dim = time_feat_dim
        self.num_layers, self.num_heads, self.dropout = num_layers, num_heads, dropout
        self.device = device
        self.src_node_mean_time_shift, self.src_node_std_time_shift = src_node_mean_time_shift, src_node_std_time_shift
        self.dst_node_mean_time_shift_dst, self.dst_node_std_time_shift = dst_node_mean_time_shift_dst, dst_node_std_time_shift
        self.model_name = model_name
        self.num_nodes = self.node_raw_features.shape[0]
        self.memory_dim = self.node_feat_dim
        self.message_dim = self.memory_dim + self.memory_dim + self.time_feat_dim + self.edge_feat_dim
        self.time_encoder = TimeEncoder(time_dim=time_feat_dim)
        self.message_aggregator = MessageAggregator()
        self.memory_bank = MemoryBank(num_nodes=self.num_nodes, memory_dim=self.memory_dim, message_dim=self.message_dim)
        self.memory_bank._owner_ref = weakref.ref(self)
        self.memory_updater = GRUMemoryUpdater(self.memory_bank, self.message_dim, self.memory_dim)
        self.embedding_module = GraphAttentionEmbedding(self.node_raw_features, self.edge_raw_features, neighbor_sampler,
                                                        self.time_encoder, self.node_feat_dim, self.edge_feat_dim,
                                                        self.time_feat_dim, self.num_layers, self.num_heads, self.dropout)
        self._engine = _Engine(self.node_feat_dim, self.edge_feat_dim, self.time_feat_dim, self.num_heads)
        self._synced = None
        self._err = None

    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.close()

    def invalidate_caches(self):
        """Drop every derived device cache (uploaded weights, query table, incremental GRU state); needed after
        ``.data`` writes that autograd's version counters do not see."""
        self._engine.invalidate()
        self.memory_bank._dirty = True
        self._synced = None

    # ---- internals
    def _gru(self):
        cell = self.memory_updater.memory_updater
        g = _lib.GruWeights()
        g.weight_ih, g.weight_hh = cell.weight_ih.data_ptr(), cell.weight_hh.data_ptr()
        g.bias_ih, g.bias_hh = cell.bias_ih.data_ptr(), cell.bias_hh.data_ptr()
        return g

    def _gru_fingerprint(self):
        cell = self.memory_updater.memory_updater
        return tuple((p.data_ptr(), p._version) for p in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh))

    def _sync(self, handle):
        """Rebuild next_memories / layer0 / the query table if memories, messages or weights
        were changed from outside (load_state_dict, reload_memory_bank, optimizer step, ...)."""
        bank = self.memory_bank
        key = (bank._fingerprint(), self._gru_fingerprint(), self._engine.versions.get(self.num_layers), handle.value)
        if bank._dirty or key != self._synced:
            s, g = bank._c_state(), self._gru()
            _lib.check(_lib.lib().flid_tgn_rebuild(handle, C.byref(s), C.byref(g), _lib.ptr(self.node_raw_features),
                                                   _lib.stream()))
            bank._dirty = False
            self._synced = key

    # ---- reference API
    def compute_src_dst_node_temporal_embeddings(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray,
                                                 node_interact_times: np.ndarray, edge_ids: np.ndarray,
                                                 edges_are_positive: bool = True, num_neighbors: int = 20):
        """models/MemoryModel.py:96-189."""
        dev = self.node_raw_features.device
        _lib.require_cuda(dev)
        sampler = self.embedding_module.neighbor_sampler
        if not isinstance(sampler, NeighborSampler):
            raise TypeError("flid_b200.MemoryModel needs a flid_b200.NeighborSampler")
        if edges_are_positive:
            assert edge_ids is not None
        b = len(src_node_ids)
        emb = self.embedding_module
        if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            return self._autograd_step(sampler, src_node_ids, dst_node_ids, node_interact_times, edge_ids,
                                       edges_are_positive, num_neighbors)
        with torch.cuda.device(dev):
            h = self._engine.handle(self.num_layers, self.time_encoder, emb.temporal_conv_layers, emb.merge_layers, dev)
            self._sync(h)
            if self._err is None or self._err.device != dev:
                self._err = torch.zeros(1, dtype=torch.int32, device=dev)
            d_src = _lib.to_device(src_node_ids, np.int64, dev, "g_src", sync_follows=True)
            d_dst = _lib.to_device(dst_node_ids, np.int64, dev, "g_dst", sync_follows=True)
            d_t = _lib.to_device(node_interact_times, np.float64, dev, "g_t", sync_follows=True)
            d_e = _lib.to_device(edge_ids, np.int64, dev, "g_e", sync_follows=True) if edge_ids is not None else None
            out = torch.empty((2 * b, self.node_feat_dim), dtype=torch.float32, device=dev)
            s, g = self.memory_bank._c_state(), self._gru()
            _lib.check(_lib.lib().flid_tgn_step(h, sampler.handle, C.byref(s), C.byref(g), _lib.ptr(self.node_raw_features),
                                                _lib.ptr(self.edge_raw_features), _lib.ptr(d_src), _lib.ptr(d_dst),
                                                _lib.ptr(d_t), _lib.ptr(d_e), b, 1 if edges_are_positive else 0,
                                                int(num_neighbors), _lib.ptr(out), _lib.ptr(self._err), _lib.stream()))
            err = int(self._err.item())   # also the per-batch sync point, as .item()-style reads are in the callers
            if err:
                self._err.zero_()
                if err == 1:
                    raise AssertionError("Trying to update memory to time in the past!")
                raise IndexError("flid_b200.MemoryModel: node or edge id out of range")
        return out[:b], out[b:]

    def embed_pass(self, src_node_ids, dst_node_ids, node_interact_times, edge_ids, batch_size: int = 200,
                   num_neighbors: int = 20, use_graph: bool = True):
        """The whole chronological loop of ``compute_src_dst_node_temporal_embeddings(..., edges_are_positive=True)``
        over consecutive batches of ``batch_size`` events as ONE C call (``flid_tgn_pass``): same kernels, same
        batch boundaries and state updates as calling the method per batch, but the launch sequence of a batch is a
        CUDA graph replayed with a device-side batch counter -- no per-batch Python, staging or synchronisation.
        Eval / no-grad only.  Returns (src embeddings, dst embeddings) float32 [E, dn] on the device.  The
        monotone-time assertion and the id range check are reported after the pass."""
        dev = self.node_raw_features.device
        _lib.require_cuda(dev)
        if self.training:
            raise RuntimeError("flid_b200.MemoryModel.embed_pass is an inference pass; call .eval() first")
        sampler = self.embedding_module.neighbor_sampler
        if not isinstance(sampler, NeighborSampler):
            raise TypeError("flid_b200.MemoryModel needs a flid_b200.NeighborSampler")
        emb = self.embedding_module
        e = len(src_node_ids)
        with torch.no_grad(), torch.cuda.device(dev):
            h = self._engine.handle(self.num_layers, self.time_encoder, emb.temporal_conv_layers, emb.merge_layers, dev)
            self._sync(h)
            if self._err is None or self._err.device != dev:
                self._err = torch.zeros(1, dtype=torch.int32, device=dev)
            d_src = _lib.to_device(src_node_ids, np.int64, dev, "gp_src")
            d_dst = _lib.to_device(dst_node_ids, np.int64, dev, "gp_dst")
            d_t = _lib.to_device(node_interact_times, np.float64, dev, "gp_t")
            d_e = _lib.to_device(edge_ids, np.int64, dev, "gp_e")
            out_s = torch.empty((e, self.node_feat_dim), dtype=torch.float32, device=dev)
            out_d = torch.empty_like(out_s)
            s, g = self.memory_bank._c_state(), self._gru()
            _lib.check(_lib.lib().flid_tgn_pass(h, sampler.handle, C.byref(s), C.byref(g), _lib.ptr(self.node_raw_features),
                                                _lib.ptr(self.edge_raw_features), _lib.ptr(d_src), _lib.ptr(d_dst),
                                                _lib.ptr(d_t), _lib.ptr(d_e), e, int(batch_size), int(num_neighbors),
                                                _lib.ptr(out_s), _lib.ptr(out_d), _lib.ptr(self._err),
                                                1 if use_graph else 0, _lib.stream()))
            err = int(self._err.item())
            if err:
                self._err.zero_()
                if err == 1:
                    raise AssertionError("Trying to update memory to time in the past!")
                raise IndexError("flid_b200.MemoryModel: node or edge id out of range")
        return out_s, out_d

    def _autograd_step(self, sampler, src_node_ids, dst_node_ids, node_interact_times, edge_ids, positive, k):
        """Training-mode batch (models/MemoryModel.py:96-189 with dropout and a backward pass).
        Differentiable part in torch CUDA ops: the GRU update of every node with a pending message
        (get_updated_memories, :190-212 -- the path the memory updater's gradients take) and the
        embedding on ``memory' + raw`` over device-sampled neighbourhoods.  The stored messages are
        constants, as they are in the reference after ``detach_memory_bank``.  The state update
        (persist, last-message election, new raw messages, :155-180) is the same C call as in eval
        mode, run without an output buffer."""
        dev = self.node_raw_features.device
        bank, emb = self.memory_bank, self.embedding_module
        b = len(src_node_ids)
        with torch.cuda.device(dev):
            h = self._engine.handle(self.num_layers, self.time_encoder, emb.temporal_conv_layers, emb.merge_layers, dev)
            self._sync(h)
            mem = bank.node_memories.data
            idx = torch.nonzero(bank._has_pending).reshape(-1)
            mem_view = mem
            if idx.numel():
                tf = bank._pending_ts[idx].to(torch.float32)
                assert bool((bank.node_last_updated_times.data[idx] <= tf).all()), \
                    "Trying to update memory to time in the past!"
                upd = self.memory_updater.memory_updater(bank._pending_msg[idx], mem[idx])
                mem_view = mem.index_put((idx,), upd)
            layer0 = mem_view + self.node_raw_features
            ids = np.concatenate([np.asarray(src_node_ids), np.asarray(dst_node_ids)])
            tt = np.concatenate([np.asarray(node_interact_times), np.asarray(node_interact_times)])
            out = autograd_forward(self.time_encoder, emb.temporal_conv_layers, emb.merge_layers, sampler, layer0,
                                   self.edge_raw_features, ids, tt, self.num_layers, num_neighbors=k,
                                   training=self.training)
            if positive:
                if self._err is None or self._err.device != dev:
                    self._err = torch.zeros(1, dtype=torch.int32, device=dev)
                d_src = _lib.to_device(src_node_ids, np.int64, dev, "g_src", sync_follows=True)
                d_dst = _lib.to_device(dst_node_ids, np.int64, dev, "g_dst", sync_follows=True)
                d_t = _lib.to_device(node_interact_times, np.float64, dev, "g_t", sync_follows=True)
                d_e = _lib.to_device(edge_ids, np.int64, dev, "g_e", sync_follows=True)
                s, g = bank._c_state(), self._gru()
                _lib.check(_lib.lib().flid_tgn_step(h, sampler.handle, C.byref(s), C.byref(g),
                                                    _lib.ptr(self.node_raw_features), _lib.ptr(self.edge_raw_features),
                                                    _lib.ptr(d_src), _lib.ptr(d_dst), _lib.ptr(d_t), _lib.ptr(d_e), b, 1,
                                                    int(k), None, _lib.ptr(self._err), _lib.stream()))
                err = int(self._err.item())
                if err:
                    self._err.zero_()
                    if err == 1:
                        raise AssertionError("Trying to update memory to time in the past!")
                    raise IndexError("flid_b200.MemoryModel: node or edge id out of range")
        return out[:b], out[b:]

    def set_neighbor_sampler(self, neighbor_sampler: NeighborSampler):
        """models/MemoryModel.py:280-291."""
        assert self.model_name in ['TGN', 'DyRep'], f'Neighbor sampler is not defined in model {self.model_name}!'
        self.embedding_module.neighbor_sampler = neighbor_sampler
        if neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert neighbor_sampler.seed is not None
            neighbor_sampler.reset_random_state()
