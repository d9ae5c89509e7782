"""Drop-in for the reference's ``utils.utils.NeighborSampler`` / ``get_neighbor_sampler``
(``utils/utils.py:71-302``) backed by a device-resident, time-sorted CSR.

Same constructor and method signatures, same return dtypes (``int64``, ``int64``,
``float32`` host numpy arrays), same error behaviour (``assert num_neighbors > 0``,
``IndexError`` for an unknown node id).  Only ``sample_neighbor_strategy='recent'`` is
implemented: the 'uniform' / 'time_interval_aware' strategies draw from numpy's MT19937
stream and are outside the hot path named by BASELINE.json.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


import itertools

_generation = itertools.count(1)


class NeighborSampler:
    def __init__(self, adj_list=None, sample_neighbor_strategy: str = 'uniform', time_scaling_factor: float = 0.0,
                 seed: int = None, device=None, _events=None):
        """adj_list[v] = [(neighbor_id, edge_id, timestamp), ...] in insertion order, adj_list[0] empty
        (utils/utils.py:73-103)."""
        if sample_neighbor_strategy != 'recent':
            if sample_neighbor_strategy in ('uniform', 'time_interval_aware'):
                raise NotImplementedError(
                    f"flid_b200 implements sample_neighbor_strategy='recent' only (got {sample_neighbor_strategy!r}); "
                    "the RNG-driven strategies are out of the accelerated path")
            raise ValueError(f'Not implemented error for sample_neighbor_strategy {sample_neighbor_strategy}!')
        self.sample_neighbor_strategy = sample_neighbor_strategy
        self.seed = seed
        self.time_scaling_factor = time_scaling_factor
        self.device = _lib.require_cuda(device)
        self._handle = C.c_void_p(None)
        self._host = None
        # process-unique id of this CSR: caches keyed on it (the layer memo) cannot be served to a later sampler
        # whose C handle happens to be allocated at the same address
        self.generation = next(_generation)
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            if _events is not None:
                src, dst, eid, ts, num_nodes = _events
                src = np.ascontiguousarray(src, dtype=np.int64)
                dst = np.ascontiguousarray(dst, dtype=np.int64)
                eid = np.ascontiguousarray(eid, dtype=np.int64)
                ts = np.ascontiguousarray(ts, dtype=np.float64)
                _lib.check(lib.flid_graph_build_events(src.ctypes.data, dst.ctypes.data, eid.ctypes.data,
                                                       ts.ctypes.data, len(src), int(num_nodes), 0,
                                                       C.byref(self._handle), _lib.stream()))
            else:
                owners, nbrs, eids, tss = [], [], [], []
                for v, lst in enumerate(adj_list):
                    if len(lst):
                        owners.append(np.full(len(lst), v, dtype=np.int64))
                        nbrs.append(np.array([x[0] for x in lst], dtype=np.int64))
                        eids.append(np.array([x[1] for x in lst], dtype=np.int64))
                        tss.append(np.array([x[2] for x in lst], dtype=np.float64))
                cat = (lambda xs, dt: np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(0, dtype=dt))
                owner, nbr, eid, ts = cat(owners, np.int64), cat(nbrs, np.int64), cat(eids, np.int64), cat(tss, np.float64)
                _lib.check(lib.flid_graph_build_entries(owner.ctypes.data, nbr.ctypes.data, eid.ctypes.data,
                                                        ts.ctypes.data, len(owner), max(len(adj_list) - 1, 0), 0,
                                                        C.byref(self._handle), _lib.stream()))
        n, m, d = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(lib.flid_graph_info(self._handle, C.byref(n), C.byref(m), C.byref(d)))
        self.num_nodes, self.num_entries, self.max_degree = n.value, m.value, d.value

    def __deepcopy__(self, memo):
        return self                   # the device CSR is immutable: copies of a model share it

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                _lib.lib().flid_graph_free(self._handle)
                self._handle = C.c_void_p(None)
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    @property
    def handle(self):
        return self._handle

    def _host_csr(self):
        """Host mirror of the CSR (exported once) for the ragged-return APIs."""
        if self._host is None:
            indptr = np.empty(self.num_nodes + 2, dtype=np.int64)
            nbr = np.empty(self.num_entries, dtype=np.int64)
            eid = np.empty(self.num_entries, dtype=np.int64)
            ts = np.empty(self.num_entries, dtype=np.float64)
            _lib.check(_lib.lib().flid_graph_export_host(self._handle, indptr.ctypes.data, nbr.ctypes.data,
                                                         eid.ctypes.data, ts.ctypes.data))
            self._host = (indptr, nbr, eid, ts)
        return self._host

    def _times(self, node_interact_times):
        t = np.asarray(node_interact_times)
        if t.dtype == np.float32:
            return t, 1
        return t.astype(np.float64, copy=False), 0

    def _cuts(self, node_ids, node_interact_times):
        t, is32 = self._times(node_interact_times)
        n = len(node_ids)
        with torch.cuda.device(self.device):
            d_nodes = _lib.to_device(node_ids, np.int64, self.device, "s_nodes")
            d_times = _lib.to_device(t, t.dtype, self.device, "s_times")
            out = torch.empty((2, n), dtype=torch.int64, device=self.device)
            _lib.check(_lib.lib().flid_sample_cut(self._handle, _lib.ptr(d_nodes), _lib.ptr(d_times), is32, n,
                                                  _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.stream()))
            h = _lib.to_host(out, "s_cut")
        return h[0], h[1]

    # ------------------------------------------------------------------ reference API
    def find_neighbors_before(self, node_id: int, interact_time: float, return_sampled_probabilities: bool = False):
        """utils/utils.py:130-147: interactions of node_id strictly before interact_time, time-sorted."""
        start, cut = self._cuts(np.array([node_id], dtype=np.int64), np.array([interact_time]))
        _, nbr, eid, ts = self._host_csr()
        s, c = int(start[0]), int(cut[0])
        return nbr[s:c], eid[s:c], ts[s:c], None

    def get_historical_neighbors_device(self, node_ids, node_interact_times, num_neighbors: int = 20):
        """Same as get_historical_neighbors but inputs may be device tensors and the three
        results stay on the device (torch int64 / int64 / float32 [n, k])."""
        n, k = len(node_ids), int(num_neighbors)
        with torch.cuda.device(self.device):
            if isinstance(node_ids, torch.Tensor):
                d_nodes = node_ids.to(self.device, torch.int64).contiguous()
            else:
                d_nodes = _lib.to_device(node_ids, np.int64, self.device, "s_nodes")
            if isinstance(node_interact_times, torch.Tensor):
                d_times = node_interact_times.to(self.device).contiguous()
                is32 = 1 if d_times.dtype == torch.float32 else 0
                if not is32:
                    d_times = d_times.to(torch.float64)
            else:
                t, is32 = self._times(node_interact_times)
                d_times = _lib.to_device(t, t.dtype, self.device, "s_times")
            nbr = torch.empty((n, max(k, 0)), dtype=torch.int64, device=self.device)
            eid = torch.empty((n, max(k, 0)), dtype=torch.int64, device=self.device)
            ts = torch.empty((n, max(k, 0)), dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib().flid_sample_recent(self._handle, _lib.ptr(d_nodes), _lib.ptr(d_times), is32, n, k,
                                                     _lib.ptr(nbr), _lib.ptr(eid), _lib.ptr(ts), _lib.stream()))
        return nbr, eid, ts

    def get_historical_neighbors(self, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 20):
        """utils/utils.py:149-214 ('recent'): three host arrays [n, k] int64 / int64 / float32."""
        nbr, eid, ts = self.get_historical_neighbors_device(np.asarray(node_ids), node_interact_times, num_neighbors)
        with torch.cuda.device(self.device):
            return _lib.to_host(nbr, "s_nbr"), _lib.to_host(eid, "s_eid"), _lib.to_host(ts, "s_ts")

    def get_multi_hop_neighbors(self, num_hops: int, node_ids: np.ndarray, node_interact_times: np.ndarray,
                                num_neighbors: int = 20):
        """utils/utils.py:216-252: hop h samples the flattened hop h-1 frontier with its float32 times."""
        assert num_hops > 0, 'Number of sampled hops should be greater than 0!'
        n = len(node_ids)
        nbr, eid, ts = self.get_historical_neighbors_device(np.asarray(node_ids), node_interact_times, num_neighbors)
        d_lists = [(nbr, eid, ts)]
        for _ in range(1, num_hops):
            p_nbr, _, p_ts = d_lists[-1]
            nbr, eid, ts = self.get_historical_neighbors_device(p_nbr.reshape(-1), p_ts.reshape(-1), num_neighbors)
            d_lists.append((nbr.reshape(n, -1), eid.reshape(n, -1), ts.reshape(n, -1)))
        with torch.cuda.device(self.device):
            host = [tuple(_lib.to_host(x, f"s_mh{i}") for i, x in enumerate(trip)) for trip in d_lists]
        return [h[0] for h in host], [h[1] for h in host], [h[2] for h in host]

    def get_all_first_hop_neighbors(self, node_ids: np.ndarray, node_interact_times: np.ndarray):
        """utils/utils.py:254-273: ragged lists of all strictly-earlier interactions per query."""
        start, cut = self._cuts(np.asarray(node_ids), node_interact_times)
        _, nbr, eid, ts = self._host_csr()
        return ([nbr[s:c] for s, c in zip(start, cut)], [eid[s:c] for s, c in zip(start, cut)],
                [ts[s:c] for s, c in zip(start, cut)])

    def reset_random_state(self):
        """utils/utils.py:275-280; 'recent' consumes no random numbers."""
        self.random_state = np.random.RandomState(self.seed)


def get_neighbor_sampler(data, sample_neighbor_strategy: str = 'uniform', time_scaling_factor: float = 0.0,
                         seed: int = None, device=None):
    """utils/utils.py:283-302: undirected adjacency over ``data`` (a reference ``Data`` record or anything
    with src_node_ids / dst_node_ids / edge_ids / node_interact_times), built on the device."""
    max_node_id = int(max(data.src_node_ids.max(), data.dst_node_ids.max()))
    return NeighborSampler(None, sample_neighbor_strategy, time_scaling_factor, seed, device,
                           _events=(data.src_node_ids, data.dst_node_ids, data.edge_ids, data.node_interact_times,
                                    max_node_id))
