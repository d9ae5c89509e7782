"""TEST INFRASTRUCTURE ONLY -- CPU restatement of FLiD's 'recent' neighbour sampler.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product (``flid_b200/``)
never does.

Parity status: PINNED against the reference itself.  ``tests/golden/make_golden.py``
imports ``/root/reference/utils/utils.py`` in the build container and writes the
reference's outputs to ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this restatement against those files (and against the live reference
when the tree is present).  The reference ships no tests or golden vectors of
its own.

What is restated (numpy, integer + one float64 compare + one f64->f32 cast):

* adjacency construction -- ``utils/utils.py:283-302`` (every event appended to
  both endpoints, src endpoint first) and ``utils/utils.py:96-103`` (per node a
  *stable* sort on the timestamp only, so ties keep event order);
* ``find_neighbors_before`` -- ``utils/utils.py:130-147``:
  ``i = searchsorted(times[node], t)`` with the default ``side='left'`` i.e. the
  strictly-earlier prefix; a float32 query time is widened to float64 exactly;
* ``get_historical_neighbors`` with ``sample_neighbor_strategy='recent'`` --
  ``utils/utils.py:149-214``: last <=k entries of the prefix, right-aligned in
  zero-initialised ``int64 / int64 / float32`` ``[n, k]`` arrays;
* ``get_multi_hop_neighbors`` -- ``utils/utils.py:216-252``: hop h re-samples the
  flattened hop h-1 frontier with its **float32** times;
* ``get_all_first_hop_neighbors`` -- ``utils/utils.py:254-273``.

The adjacency is held as one CSR (owner-sorted, time-sorted inside a node)
instead of three numpy arrays per node; the per-query search is a vectorised
bisection.  ``get_historical_neighbors_loop`` keeps the reference's literal
one-query-at-a-time form for cross-checking on small inputs.
"""
import numpy as np


class OracleSampler:
    def __init__(self, owners, nbrs, eids, ts, num_nodes):
        """owners/nbrs/eids int64[M], ts float64[M] in *insertion order*; ids in [0, num_nodes]."""
        owners = np.asarray(owners, dtype=np.int64)
        ts = np.asarray(ts, dtype=np.float64)
        # stable sort on time, then stable sort on owner == per-node stable sort on time
        by_time = np.argsort(ts, kind="stable")
        order = by_time[np.argsort(owners[by_time], kind="stable")]
        self.nbr = np.asarray(nbrs, dtype=np.int64)[order]
        self.eid = np.asarray(eids, dtype=np.int64)[order]
        self.ts = ts[order]
        counts = np.bincount(owners, minlength=num_nodes + 1)
        self.indptr = np.zeros(num_nodes + 2, dtype=np.int64)
        np.cumsum(counts, out=self.indptr[1:])
        self.num_nodes = num_nodes
        self.sample_neighbor_strategy = "recent"
        self.seed = None

    # -- construction -------------------------------------------------------
    @classmethod
    def from_events(cls, src, dst, eid, ts, num_nodes=None):
        """utils/utils.py:283-302: undirected, src endpoint appended before dst endpoint."""
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        eid = np.asarray(eid, dtype=np.int64)
        ts = np.asarray(ts, dtype=np.float64)
        if num_nodes is None:
            num_nodes = int(max(src.max(), dst.max())) if len(src) else 0
        e = len(src)
        owners = np.empty(2 * e, dtype=np.int64)
        nbrs = np.empty(2 * e, dtype=np.int64)
        owners[0::2], owners[1::2] = src, dst
        nbrs[0::2], nbrs[1::2] = dst, src
        return cls(owners, nbrs, np.repeat(eid, 2), np.repeat(ts, 2), num_nodes)

    @classmethod
    def from_adj_list(cls, adj_list):
        """utils/utils.py:73-103: adj_list[v] = [(nbr, edge_id, ts), ...] in insertion order."""
        owners, nbrs, eids, ts = [], [], [], []
        for v, lst in enumerate(adj_list):
            for (u, e, t) in lst:
                owners.append(v), nbrs.append(u), eids.append(e), ts.append(t)
        return cls(np.array(owners, dtype=np.int64), np.array(nbrs, dtype=np.int64),
                   np.array(eids, dtype=np.int64), np.array(ts, dtype=np.float64), len(adj_list) - 1)

    # -- queries ------------------------------------------------------------
    def _cut(self, node_ids, times):
        """Vectorised searchsorted(side='left') inside each node's segment."""
        node_ids = np.asarray(node_ids, dtype=np.int64)
        t = np.asarray(times).astype(np.float64)  # float32 -> float64 is exact
        lo = self.indptr[node_ids].copy()
        hi = self.indptr[node_ids + 1].copy()
        start = lo.copy()
        active = lo < hi
        while active.any():
            mid = (lo + hi) >> 1
            probe = np.where(active, mid, 0)
            go_right = active & (self.ts[probe] < t) if len(self.ts) else np.zeros_like(active)
            go_left = active & ~go_right
            lo = np.where(go_right, mid + 1, lo)
            hi = np.where(go_left, mid, hi)
            active = lo < hi
        return start, lo

    def find_neighbors_before(self, node_id, interact_time, return_sampled_probabilities=False):
        s, c = self._cut(np.array([node_id]), np.array([interact_time]))
        s, c = int(s[0]), int(c[0])
        return self.nbr[s:c], self.eid[s:c], self.ts[s:c], None

    def get_historical_neighbors(self, node_ids, node_interact_times, num_neighbors=20):
        assert num_neighbors > 0, "Number of sampled neighbors for each node should be greater than 0!"
        n, k = len(node_ids), int(num_neighbors)
        start, cut = self._cut(node_ids, node_interact_times)
        cnt = np.minimum(cut - start, k)
        col = np.arange(k, dtype=np.int64)[None, :]
        valid = col >= (k - cnt)[:, None]
        pos = cut[:, None] - k + col
        pos = np.where(valid, pos, 0)
        out_nbr = np.zeros((n, k), dtype=np.int64)
        out_eid = np.zeros((n, k), dtype=np.int64)
        out_ts = np.zeros((n, k), dtype=np.float32)
        if len(self.ts):
            out_nbr[valid] = self.nbr[pos[valid]]
            out_eid[valid] = self.eid[pos[valid]]
            out_ts[valid] = self.ts[pos[valid]].astype(np.float32)  # round-to-nearest, as numpy assignment
        return out_nbr, out_eid, out_ts

    def get_historical_neighbors_loop(self, node_ids, node_interact_times, num_neighbors=20):
        """One query at a time, as utils/utils.py:170-209 does (small inputs only)."""
        n, k = len(node_ids), int(num_neighbors)
        out_nbr = np.zeros((n, k), dtype=np.int64)
        out_eid = np.zeros((n, k), dtype=np.int64)
        out_ts = np.zeros((n, k), dtype=np.float32)
        for i, (v, t) in enumerate(zip(node_ids, node_interact_times)):
            a, b = self.indptr[v], self.indptr[v + 1]
            c = a + np.searchsorted(self.ts[a:b], t)
            m = min(c - a, k)
            if m > 0:
                out_nbr[i, k - m:] = self.nbr[c - m:c]
                out_eid[i, k - m:] = self.eid[c - m:c]
                out_ts[i, k - m:] = self.ts[c - m:c]
        return out_nbr, out_eid, out_ts

    def get_multi_hop_neighbors(self, num_hops, node_ids, node_interact_times, num_neighbors=20):
        assert num_hops > 0, "Number of sampled hops should be greater than 0!"
        nbr, eid, ts = self.get_historical_neighbors(node_ids, node_interact_times, num_neighbors)
        nbrs, eids, tss = [nbr], [eid], [ts]
        for _ in range(1, num_hops):
            nbr, eid, ts = self.get_historical_neighbors(nbrs[-1].flatten(), tss[-1].flatten(), num_neighbors)
            nbrs.append(nbr.reshape(len(node_ids), -1))
            eids.append(eid.reshape(len(node_ids), -1))
            tss.append(ts.reshape(len(node_ids), -1))
        return nbrs, eids, tss

    def get_all_first_hop_neighbors(self, node_ids, node_interact_times):
        start, cut = self._cut(node_ids, node_interact_times)
        return ([self.nbr[s:c] for s, c in zip(start, cut)],
                [self.eid[s:c] for s, c in zip(start, cut)],
                [self.ts[s:c] for s, c in zip(start, cut)])

    def reset_random_state(self):  # API parity; 'recent' draws nothing
        pass
