"""TEST INFRASTRUCTURE ONLY -- torch-fp32 CPU restatement of FLiD's TGN (MemoryModel).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this; the product never does.

Parity status: PINNED against the reference itself (``models/MemoryModel.py``
imported in the build container by ``tests/golden/make_golden.py``, outputs under
``tests/golden/``; the live reference is also compared when present).

Restated, in the reference's literal order, with its dict-of-lists message store:

* ``aggregate_last``      -- ``models/MemoryModel.py:303-330`` keep ``msgs[node][-1]``
* ``gru_cell``            -- ``nn.GRUCell`` as constructed at ``:531-543``
  (gates r,z,n; ``h' = (h - n) * z + n``)
* ``get_updated_memories``-- ``:190-212`` + ``:501-528`` clone of the bank, not persisted
* ``update_memories``     -- ``:214-231`` + ``:472-499`` persist + ``last_updated = t.float()``,
  with the reference's monotone-time assertion
* ``new_raw_messages``    -- ``:233-278`` ``cat[mem[a], mem[b], te(t - last_upd[a]), edge[eid]]``
* ``step``                -- ``:96-189``; the src-role messages are stored before the
  dst-role ones (``:177-180``) so a node seen in both roles keeps the dst-role one last.
* embedding               -- ``:632-715`` == TGAT recursion with layer-0 / merge input
  ``memory' + raw`` (delegated to ``oracle.tgat.embed(..., layer0=...)``).
"""
from collections import defaultdict

import contextlib

import numpy as np
import torch
import torch.nn.functional as F

from . import tgat as otgat


def default_params(node_dim, edge_dim, time_dim, num_layers, num_heads=2, seed=0, time_bias_scale=0.0):
    base = otgat.default_params(node_dim, edge_dim, time_dim, num_layers, num_heads, seed, time_bias_scale)
    p = {"_num_heads": num_heads}
    for key, v in base.items():
        if key.startswith("time_encoder."):
            p[key] = v
        elif not key.startswith("_"):
            p["embedding_module." + key] = v
    g = torch.Generator().manual_seed(seed + 1000)
    msg_dim = 2 * node_dim + time_dim + edge_dim
    bound = 1.0 / np.sqrt(node_dim)
    u = "memory_updater.memory_updater."
    p[u + "weight_ih"] = (torch.rand(3 * node_dim, msg_dim, generator=g) * 2 - 1) * bound
    p[u + "weight_hh"] = (torch.rand(3 * node_dim, node_dim, generator=g) * 2 - 1) * bound
    p[u + "bias_ih"] = (torch.rand(3 * node_dim, generator=g) * 2 - 1) * bound
    p[u + "bias_hh"] = (torch.rand(3 * node_dim, generator=g) * 2 - 1) * bound
    return p


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    gi = F.linear(x, w_ih, b_ih).chunk(3, dim=1)
    gh = F.linear(h, w_hh, b_hh).chunk(3, dim=1)
    r = torch.sigmoid(gh[0] + gi[0])
    z = torch.sigmoid(gh[1] + gi[1])
    n = torch.tanh(gi[2] + gh[2] * r)
    return (h - n) * z + n


class OracleTGN:
    def __init__(self, params, node_feat, edge_feat, sampler, num_layers, num_neighbors=20):
        self.p = params
        self.node_feat = node_feat            # torch float32 [N+1, d]
        self.edge_feat = edge_feat
        self.sampler = sampler
        self.num_layers = num_layers
        self.k = num_neighbors
        self.num_nodes = node_feat.shape[0]
        self.reset()

    def reset(self):
        """MemoryBank.__init_memory_bank__ (models/MemoryModel.py:359-366)."""
        self.mem = torch.zeros(self.num_nodes, self.node_feat.shape[1])
        self.last_upd = torch.zeros(self.num_nodes)
        self.msgs = defaultdict(list)

    # -- pieces ---------------------------------------------------------------
    def _gru(self, x, h):
        u = "memory_updater.memory_updater."
        return gru_cell(x, h, self.p[u + "weight_ih"], self.p[u + "weight_hh"], self.p[u + "bias_ih"], self.p[u + "bias_hh"])

    def aggregate_last(self, node_ids):
        ids, ms, ts = [], [], []
        for v in np.unique(node_ids):
            lst = self.msgs.get(int(v), [])
            if len(lst) > 0:
                ids.append(int(v)), ms.append(lst[-1][0]), ts.append(lst[-1][1])
        return np.array(ids, dtype=np.int64), (torch.stack(ms) if ms else torch.zeros(0)), np.array(ts)

    def get_updated_memories(self):
        ids, ms, ts = self.aggregate_last(np.arange(self.num_nodes))
        mem, lu = self.mem.clone(), self.last_upd.clone()
        if len(ids) == 0:
            return mem, lu
        idx = torch.from_numpy(ids)
        tf = torch.from_numpy(ts).float()
        assert (self.last_upd[idx] <= tf).all().item(), "Trying to update memory to time in the past!"
        mem[idx] = self._gru(ms, mem[idx])
        lu[idx] = tf
        return mem, lu

    def update_memories(self, node_ids):
        ids, ms, ts = self.aggregate_last(node_ids)
        if len(ids) == 0:
            return
        idx = torch.from_numpy(ids)
        tf = torch.from_numpy(ts).float()
        assert (self.last_upd[idx] <= tf).all().item(), "Trying to update memory to time in the past!"
        self.mem[idx] = self._gru(ms, self.mem[idx]).detach()
        self.last_upd[idx] = tf

    def new_raw_messages(self, a_ids, b_ids, times, eids):
        a, b = torch.from_numpy(a_ids), torch.from_numpy(b_ids)
        dt = torch.from_numpy(times).float() - self.last_upd[a]
        te = otgat.time_encode(self.p, dt.unsqueeze(1)).reshape(len(a_ids), -1)
        rows = torch.cat([self.mem[a], self.mem[b], te, self.edge_feat[torch.from_numpy(eids)]], dim=1).detach()
        out = defaultdict(list)
        for i in range(len(a_ids)):
            out[int(a_ids[i])].append((rows[i], times[i]))
        return np.unique(a_ids), out

    # -- the call -------------------------------------------------------------
    def step(self, src, dst, times, eids, positive=True, grad=False):
        """``grad=True``: the embeddings keep their autograd graph (training-mode parity tests; the parameters
        must then be leaf tensors).  The state written back is detached either way, as the reference's is after
        ``detach_memory_bank`` (models/MemoryModel.py:440-445)."""
        with (contextlib.nullcontext() if grad else torch.no_grad()):
            node_ids = np.concatenate([src, dst])
            mem2, _ = self.get_updated_memories()
            emb = otgat.embed(self.p, self.node_feat, self.edge_feat, self.sampler, node_ids,
                              np.concatenate([times, times]), self.num_layers, self.k,
                              prefix="embedding_module.", layer0=mem2 + self.node_feat)
            src_emb, dst_emb = emb[:len(src)], emb[len(src):]
            if positive:
                self.update_memories(node_ids)
                for v in node_ids:
                    self.msgs[int(v)] = []
                us, ms = self.new_raw_messages(src, dst, times, eids)
                ud, md = self.new_raw_messages(dst, src, times, eids)
                for v in us:
                    self.msgs[int(v)].extend(ms[int(v)])
                for v in ud:
                    self.msgs[int(v)].extend(md[int(v)])
            return src_emb, dst_emb
