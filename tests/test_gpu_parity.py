"""GPU parity tests (run with -m gpu on the B200 box).  Every check goes through the C ABI
(via the ctypes host layer) and compares against the CPU oracle (oracle/) on the same
seeded inputs and against the golden vectors minted from the real FLiD reference.

Tolerances (BASELINE.json north_star): sampler bit-exact; fp32 embeddings rel 1e-4, stated
here as three checks: every element within 2e-5 of the tensor's scale (max(1, max|want|)),
every significant element (|want| >= 0.1 * scale) within 1e-4 relative, and a mean absolute
error below 2e-6 of the scale; pseudo-label masks identical."""
import ctypes
import os

import numpy as np
import pytest
import torch

import cases
import flid_b200
from flid_b200 import _lib, passes, synth
from oracle import sampler as osamp, tgat as otgat, tgn as otgn, pseudo as opseudo

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def load(name):
    return np.load(os.path.join(G, name))


def assert_fp32_close(got, want, what=""):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.isfinite(got).all(), f"{what}: non-finite values"
    if want.size == 0:
        return
    scale = max(1.0, float(np.abs(want).max()))
    err = np.abs(got - want)
    assert err.max() <= 2e-5 * scale, f"{what}: max abs err {err.max():.3e} (scale {scale:.3g})"
    big = np.abs(want) >= 0.1 * scale
    if big.any():
        rel = (err[big] / np.abs(want[big])).max()
        assert rel <= 1e-4, f"{what}: max relative err {rel:.3e} on elements >= 0.1 * scale"
    assert err.mean() <= 2e-6 * scale, f"{what}: mean abs err {err.mean():.3e}"


def assert_paths_close(a, b, what=""):
    """Two evaluation orders of the same arithmetic (per-slot stream vs projected bulk path): fp32 rounding apart."""
    a, b = a.double(), b.double()
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    assert err <= 2e-5 * scale, f"{what}: paths differ by {err:.3e} (scale {scale:.3g})"


def make_sampler(src, dst, eid, ts, n):
    return flid_b200.NeighborSampler(None, "recent", seed=1, device=DEV, _events=(src, dst, eid, ts, n))


# ===================================================================== sampler
def test_csr_build_matches_oracle_events_and_adj_list():
    src, dst, eid, ts, n = cases.adversarial_events()
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, n)
    s = make_sampler(src, dst, eid, ts, n)
    indptr, nbr, e, t = s._host_csr()
    assert np.array_equal(indptr, o.indptr) and np.array_equal(nbr, o.nbr)
    assert np.array_equal(e, o.eid) and np.array_equal(t, o.ts)
    adj = [[] for _ in range(n + 1)]
    for a, b, c, d in zip(src, dst, eid, ts):
        adj[a].append((b, c, d))
        adj[b].append((a, c, d))
    s2 = flid_b200.NeighborSampler(adj, "recent", seed=1, device=DEV)
    i2, n2, e2, t2 = s2._host_csr()
    assert np.array_equal(i2, o.indptr) and np.array_equal(n2, o.nbr) and np.array_equal(e2, o.eid)
    assert np.array_equal(t2, o.ts)
    assert s.max_degree == int(np.diff(o.indptr).max())


@pytest.mark.parametrize("k", [1, 3, 20])
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_sampler_golden_bit_exact(k, dt):
    g = load("sampler.npz")
    src, dst, eid, ts, n = cases.adversarial_events()
    s = make_sampler(src, dst, eid, ts, n)
    nodes, times = cases.adversarial_queries()
    if dt == "f32":
        times = times.astype(np.float32)
    a, b, c = s.get_historical_neighbors(nodes, times, k)
    assert a.dtype == np.int64 and b.dtype == np.int64 and c.dtype == np.float32
    assert np.array_equal(a, g[f"{dt}_k{k}_nbr"])
    assert np.array_equal(b, g[f"{dt}_k{k}_eid"])
    assert np.array_equal(c, g[f"{dt}_k{k}_ts"])


def test_sampler_multi_hop_and_ragged_apis():
    g = load("sampler.npz")
    src, dst, eid, ts, n = cases.adversarial_events()
    s = make_sampler(src, dst, eid, ts, n)
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, n)
    nodes, times = cases.adversarial_queries()
    nl, el, tl = s.get_multi_hop_neighbors(2, nodes[:300], times[:300], 3)
    for h in range(2):
        assert np.array_equal(nl[h], g[f"hop{h}_nbr"]) and np.array_equal(el[h], g[f"hop{h}_eid"])
        assert np.array_equal(tl[h], g[f"hop{h}_ts"])
    a, b, c = s.get_all_first_hop_neighbors(nodes[:200], times[:200])
    wa, wb, wc = o.get_all_first_hop_neighbors(nodes[:200], times[:200])
    for i in range(200):
        assert np.array_equal(a[i], wa[i]) and np.array_equal(b[i], wb[i]) and np.array_equal(c[i], wc[i])
    x = s.find_neighbors_before(1, 25.0)
    y = o.find_neighbors_before(1, 25.0)
    assert all(np.array_equal(p, q) for p, q in zip(x[:3], y[:3])) and x[3] is None


def test_sampler_errors_and_edges():
    src, dst, eid, ts, n = cases.adversarial_events()
    s = make_sampler(src, dst, eid, ts, n)
    with pytest.raises(AssertionError):
        s.get_historical_neighbors(np.array([1]), np.array([5.0]), 0)
    with pytest.raises(IndexError):
        s.get_historical_neighbors(np.array([n + 1]), np.array([5.0]), 3)
    a, b, c = s.get_historical_neighbors(np.zeros(0, dtype=np.int64), np.zeros(0), 4)
    assert a.shape == (0, 4) and c.dtype == np.float32
    # k far above 32 (GraphMixer asks for 2000 recent neighbours)
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, n)
    nodes = np.array([1, 1, 2, 42, 0, 60], dtype=np.int64)
    times = np.array([1e9, 20.0, 30.0, 1e9, 1e9, 1e9])
    for k in (33, 64, 2000):
        got, want = s.get_historical_neighbors(nodes, times, k), o.get_historical_neighbors(nodes, times, k)
        assert all(np.array_equal(p, q) for p, q in zip(got, want))


@pytest.mark.parametrize("shape", ["wikipedia", "dsub", "fractional"])
def test_sampler_random_graphs_bit_exact(shape):
    if shape == "wikipedia":
        g = synth.wikipedia_shape(seed=3, scale=0.2)
    elif shape == "dsub":
        g = synth.dsub_shape(seed=4, scale=0.2)
    else:
        g = synth.general_graph("frac", 3000, 60000, 1e8, seed=5, exponent=0.9, integral_times=False, dim=4)
    s = make_sampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    o = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    rs = np.random.RandomState(0)
    sel = rs.randint(0, g.num_interactions, 20000)
    nodes = np.concatenate([g.src_node_ids[sel], g.dst_node_ids[sel], rs.randint(0, g.num_nodes + 1, 5000)])
    times = np.concatenate([g.node_interact_times[sel], g.node_interact_times[sel],
                            rs.uniform(0, g.node_interact_times.max() * 1.05, 5000)])
    for k in (20, 30):
        for tt in (times, times.astype(np.float32)):
            got, want = s.get_historical_neighbors(nodes, tt, k), o.get_historical_neighbors(nodes, tt, k)
            for p, q in zip(got, want):
                assert p.dtype == q.dtype and np.array_equal(p, q)
    # sortedness / strict-earlier properties on the device output itself
    a, b, c = s.get_historical_neighbors(nodes, times, 20)
    assert (np.diff(c, axis=1)[a[:, 1:] * a[:, :-1] != 0] >= 0).all()
    assert (c[a != 0].astype(np.float64) <= np.repeat(times[:, None], 20, 1)[a != 0].astype(np.float32)).all()


# ===================================================================== TGAT
def tgat_pair(nf, ef, src, dst, eid, ts, L, heads, p, memo=False):
    s = make_sampler(src, dst, eid, ts, nf.shape[0] - 1)
    m = flid_b200.TGAT(nf, ef, s, 100, L, heads, 0.1, DEV).to(DEV)
    m.load_state_dict({k: v for k, v in p.items() if not k.startswith("_")})
    m.eval()
    m.set_layer_memo(memo)
    return m, s


TGAT_CASES = [("L1_k20", 1, 20, 2, 0.0, False), ("L2_k5", 2, 5, 2, 0.0, False), ("L2_k20_bias", 2, 20, 2, 0.5, False),
              ("L2_k7_zeros", 2, 7, 2, 0.3, True), ("L3_k3", 3, 3, 2, 0.2, False)]


@pytest.mark.parametrize("memo", [False, True])
@pytest.mark.parametrize("name,L,k,heads,bias,zeros", TGAT_CASES)
def test_tgat_golden(name, L, k, heads, bias, zeros, memo):
    g = load("tgat.npz")
    src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
    p = otgat.default_params(172, 172, 100, L, heads, seed=3, time_bias_scale=bias)
    m, _ = tgat_pair(nf, ef, src, dst, eid, ts, L, heads, p, memo)
    sel = g[name + "_sel"]
    with torch.no_grad():
        a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
    assert a.device.type == "cuda" and a.dtype == torch.float32 and a.shape == (len(sel), 172)
    assert_fp32_close(a.cpu().numpy(), g[name + "_src"], name + " src")
    assert_fp32_close(b.cpu().numpy(), g[name + "_dst"], name + " dst")


@pytest.mark.parametrize("shape,L,k,heads", [("wikipedia", 2, 20, 2), ("dsub", 2, 30, 2), ("wikipedia", 1, 20, 4),
                                             ("dsub", 1, 10, 1)])
def test_tgat_vs_oracle_synthetic_shapes(shape, L, k, heads):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = synth.wikipedia_shape(seed=1, scale=0.03) if shape == "wikipedia" else synth.dsub_shape(seed=2, scale=0.03)
    p = otgat.default_params(172, 172, 100, L, heads, seed=9, time_bias_scale=0.25)
    m, _ = tgat_pair(g.node_raw_features, g.edge_raw_features, g.src_node_ids, g.dst_node_ids, g.edge_ids,
                     g.node_interact_times, L, heads, p)
    o = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    e = g.num_interactions
    nb = 60 if L == 2 else 200
    sel = np.concatenate([np.arange(e // 2, e // 2 + nb), np.arange(e - nb, e)])
    with torch.no_grad():
        a, b = m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sel], g.dst_node_ids[sel],
                                                          g.node_interact_times[sel], k)
    wa, wb = otgat.embed_src_dst(p, torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features), o,
                                 g.src_node_ids[sel], g.dst_node_ids[sel], g.node_interact_times[sel], L, k)
    assert_fp32_close(a.cpu().numpy(), wa.numpy(), f"{shape} L{L} src")
    assert_fp32_close(b.cpu().numpy(), wb.numpy(), f"{shape} L{L} dst")


@pytest.mark.parametrize("dn,de,T,heads,L,k", [(64, 128, 36, 2, 2, 7), (32, 16, 32, 4, 1, 32), (256, 172, 128, 1, 2, 3)])
def test_tgat_other_feature_widths(dn, de, T, heads, L, k):
    """Nothing in the kernels is specialised to d = 172 / T = 100: other (multiple-of-4) widths, head counts
    and k up to the 32-slot limit, plain and memoised."""
    rs = np.random.RandomState(dn + de)
    src, dst, eid, ts, _, _ = cases.small_stream()
    n_nodes = int(max(src.max(), dst.max()))
    nf = rs.standard_normal((n_nodes + 1, dn)).astype(np.float32)
    ef = rs.standard_normal((len(src) + 1, de)).astype(np.float32)
    nf[0] = 0
    ef[0] = 0
    p = otgat.default_params(dn, de, T, L, heads, seed=13, time_bias_scale=0.2)
    s = make_sampler(src, dst, eid, ts, n_nodes)
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, n_nodes)
    m = flid_b200.TGAT(nf, ef, s, T, L, heads, 0.1, DEV).to(DEV)
    m.load_state_dict({kk: v for kk, v in p.items() if not kk.startswith("_")})
    m.eval()
    sel = np.arange(420, 480)
    wa, wb = otgat.embed_src_dst(p, torch.from_numpy(nf), torch.from_numpy(ef), o, src[sel], dst[sel], ts[sel], L, k)
    for memo in (False, True):
        m.set_layer_memo(memo)
        with torch.no_grad():
            a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
        assert_fp32_close(a.cpu().numpy(), wa.numpy(), f"widths {dn}/{de}/{T} memo={memo} src")
        assert_fp32_close(b.cpu().numpy(), wb.numpy(), f"widths {dn}/{de}/{T} memo={memo} dst")


def test_tgat_float32_root_times_and_lower_layers():
    """compute_node_temporal_embeddings with float32 times (the recursion's call shape) and
    current_layer_num below num_layers."""
    src, dst, eid, ts, nf, ef = cases.small_stream()
    p = otgat.default_params(172, 172, 100, 2, 2, seed=3, time_bias_scale=0.3)
    m, _ = tgat_pair(nf, ef, src, dst, eid, ts, 2, 2, p)
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    ids, t32 = src[300:340], ts[300:340].astype(np.float32)
    for layer in (1, 2):
        with torch.no_grad():
            got = m.compute_node_temporal_embeddings(ids, t32, layer, 6)
            want = otgat.embed(p, torch.from_numpy(nf), torch.from_numpy(ef), o, ids, t32, layer, 6)
        assert_fp32_close(got.cpu().numpy(), want.numpy(), f"f32 layer {layer}")
    got0 = m.compute_node_temporal_embeddings(ids, t32, 0, 6)
    assert np.array_equal(got0.cpu().numpy(), nf[ids])


def test_tgat_chunking_and_table_do_not_change_bits():
    """Same kernels, different slices: chunked / unchunked and cached-table / per-target query
    folds must be bit-identical (the property the multi-GPU sharding relies on)."""
    g = synth.wikipedia_shape(seed=1, scale=0.03)
    p = otgat.default_params(172, 172, 100, 2, 2, seed=9, time_bias_scale=0.25)
    m, _ = tgat_pair(g.node_raw_features, g.edge_raw_features, g.src_node_ids, g.dst_node_ids, g.edge_ids,
                     g.node_interact_times, 2, 2, p)
    sel = np.arange(g.num_interactions - 500, g.num_interactions)
    args = (g.src_node_ids[sel], g.dst_node_ids[sel], g.node_interact_times[sel], 20)
    with torch.no_grad():
        a0, b0 = m.compute_src_dst_node_temporal_embeddings(*args)
        h = m._engine.handles[2]
        _lib.check(_lib.lib().flid_tgat_set_chunk_targets(h, 21 * 37))
        a1, b1 = m.compute_src_dst_node_temporal_embeddings(*args)
        _lib.check(_lib.lib().flid_tgat_set_chunk_targets(h, 606208))
        a2 = torch.cat([m.compute_src_dst_node_temporal_embeddings(args[0][i:i + 125], args[1][i:i + 125],
                                                                     args[2][i:i + 125], 20)[0]
                        for i in range(0, 500, 125)])
    assert torch.equal(a0, a1) and torch.equal(b0, b1) and torch.equal(a0, a2)
    st = m.last_stats()
    assert st[0] == 125 * 2 * 22 and st[2] == 125 * 2 * 22 and 0 < st[1] <= st[0] * 20


@pytest.mark.parametrize("projected", [False, True])
@pytest.mark.parametrize("shape,L,k", [("wikipedia", 2, 20), ("dsub", 2, 30), ("fractional", 3, 4)])
def test_tgat_layer_memo_is_bit_identical(shape, L, k, projected):
    """The memoised path (one evaluation per adjacency entry and level, flid_tgat_embed_memo) against the
    recursive path, for float64 roots and float32 roots.  With the per-slot stream (``projected=False``) the
    bits are identical; the projected build (per-entry K/V, csrc/bulk_kv.cu) re-associates the arithmetic and
    agrees to fp32 rounding.  In both modes row ranges built separately (the multi-GPU split) give the same
    table bit for bit."""
    same = torch.equal if not projected else (lambda a, b: (assert_paths_close(a, b, "memo vs recursion"), True)[1])
    if shape == "fractional":
        src, dst, eid, ts, nf, ef = cases.small_stream()
        ts = ts + np.random.RandomState(5).rand(len(ts)) * 0.37 + 2.0 ** 24       # float32 rounding matters
        ts.sort()
        n_nodes = nf.shape[0] - 1
    else:
        g = synth.wikipedia_shape(seed=1, scale=0.03) if shape == "wikipedia" else synth.dsub_shape(seed=2, scale=0.03)
        src, dst, eid, ts, nf, ef = (g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times,
                                     g.node_raw_features, g.edge_raw_features)
    p = otgat.default_params(172, 172, 100, L, 2, seed=11, time_bias_scale=0.25)
    m, s = tgat_pair(nf, ef, src, dst, eid, ts, L, 2, p, memo=False)
    m.set_bulk_projection(projected)
    e = len(src)
    sel = np.arange(e - 300, e)
    nodes = np.concatenate([src[sel], dst[sel]])
    times = np.concatenate([ts[sel], ts[sel]])
    with torch.no_grad():
        plain64 = m.compute_node_temporal_embeddings(nodes, times, L, k)
        plain32 = m.compute_node_temporal_embeddings(nodes, times.astype(np.float32), L, k)
        m.set_layer_memo(True)
        memo64 = m.compute_node_temporal_embeddings(nodes, times, L, k)
        memo32 = m.compute_node_temporal_embeddings(nodes, times.astype(np.float32), L, k)
    assert L in m._engine.memo
    assert same(plain64, memo64) and same(plain32, memo32)
    # roots that are events of the graph at float32-exact times take their own lower layers from the
    # memo (one evaluation per root); switched off, every root runs all L levels.  Same bits.
    h, lib = m._engine.handles[L], _lib.lib()
    exact = bool(np.all(times.astype(np.float32).astype(np.float64) == times))
    st = m.last_stats()
    assert st[2] == len(nodes) and st[0] == len(nodes) * (1 if exact else L)
    _lib.check(lib.flid_tgat_set_self_from_memo(h, 0))
    with torch.no_grad():
        full64 = m.compute_node_temporal_embeddings(nodes, times, L, k)
    st = m.last_stats()
    assert st[0] == len(nodes) * L and same(full64, memo64)
    _lib.check(lib.flid_tgat_set_self_from_memo(h, 1))
    # roots that are not events (shifted times, other nodes) must fall back to the full chain
    t_off = times + 0.5
    n_off = np.roll(nodes, 7)
    with torch.no_grad():
        a_memo = m.compute_node_temporal_embeddings(n_off, t_off, L, k)
        m.set_layer_memo(False)
        a_plain = m.compute_node_temporal_embeddings(n_off, t_off, L, k)
        m.set_layer_memo(True)
        m.compute_node_temporal_embeddings(nodes[:4], times[:4], L, k)      # rebuilds the memo for the split-build check
    assert same(a_memo, a_plain)
    # split build == whole build
    tables = m._engine.memo[L][1]
    rows = s.num_entries + 1
    prev = None
    for level, whole in enumerate(tables, start=1):
        part = torch.full_like(whole, float("nan"))
        for lo, hi in ((0, rows // 3), (rows // 3, rows - 1), (rows - 1, rows)):
            _lib.check(lib.flid_tgat_memo_build(h, s.handle, _lib.ptr(m.node_raw_features), _lib.ptr(m.edge_raw_features),
                                                k, level, _lib.ptr(prev), lo, hi, _lib.ptr(part), _lib.stream()))
        assert torch.equal(part[:rows], whole[:rows])
        prev = whole


def test_tgat_layer_memo_auto_and_invalidation():
    src, dst, eid, ts, nf, ef = cases.small_stream()
    p = otgat.default_params(172, 172, 100, 2, 2, seed=3, time_bias_scale=0.1)
    m, s = tgat_pair(nf, ef, src, dst, eid, ts, 2, 2, p, memo="auto")
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    k = 6
    need = (s.num_entries + 1) // ((1 + k) + 1 - 2) + 1          # break-even number of roots
    sel = np.arange(len(src) - 40, len(src))
    with torch.no_grad():
        calls = 0
        while 2 not in m._engine.memo:
            a, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
            calls += 1
            assert calls * 80 <= need + 80, "auto mode never built the memo"
        assert calls > 1, "auto mode built the memo before it paid off"
        a2, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
        assert torch.equal(a, a2)
        # a different k, new weights or a new sampler must not reuse the table
        b, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k + 1)
        wb, _ = otgat.embed_src_dst(p, torch.from_numpy(nf), torch.from_numpy(ef), o, src[sel], dst[sel], ts[sel], 2, k + 1)
        assert_fp32_close(b.cpu().numpy(), wb.numpy(), "memo, other k")
        p2 = otgat.default_params(172, 172, 100, 2, 2, seed=4, time_bias_scale=0.2)
        m.load_state_dict({kk: v for kk, v in p2.items() if not kk.startswith("_")})
        m.set_layer_memo(True)
        c, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
        wc, _ = otgat.embed_src_dst(p2, torch.from_numpy(nf), torch.from_numpy(ef), o, src[sel], dst[sel], ts[sel], 2, k)
        assert_fp32_close(c.cpu().numpy(), wc.numpy(), "memo after load_state_dict")
        half = len(src) // 2
        s2 = make_sampler(src[:half], dst[:half], eid[:half], ts[:half], nf.shape[0] - 1)
        m.set_neighbor_sampler(s2)
        o2 = osamp.OracleSampler.from_events(src[:half], dst[:half], eid[:half], ts[:half], nf.shape[0] - 1)
        d, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
        wd, _ = otgat.embed_src_dst(p2, torch.from_numpy(nf), torch.from_numpy(ef), o2, src[sel], dst[sel], ts[sel], 2, k)
        assert_fp32_close(d.cpu().numpy(), wd.numpy(), "memo after set_neighbor_sampler")


def test_tgat_weight_update_is_picked_up():
    src, dst, eid, ts, nf, ef = cases.small_stream()
    p = otgat.default_params(172, 172, 100, 1, 2, seed=3)
    m, _ = tgat_pair(nf, ef, src, dst, eid, ts, 1, 2, p)
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    sel = np.arange(400, 420)
    with torch.no_grad():
        a0, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 5)
        p2 = otgat.default_params(172, 172, 100, 1, 2, seed=4, time_bias_scale=0.2)
        m.load_state_dict({k: v for k, v in p2.items() if not k.startswith("_")})
        a1, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 5)
    w1, _ = otgat.embed_src_dst(p2, torch.from_numpy(nf), torch.from_numpy(ef), o, src[sel], dst[sel], ts[sel], 1, 5)
    assert not torch.equal(a0, a1)
    assert_fp32_close(a1.cpu().numpy(), w1.numpy(), "after load_state_dict")


def test_tgat_errors():
    src, dst, eid, ts, nf, ef = cases.small_stream()
    p = otgat.default_params(172, 172, 100, 1, 2, seed=3)
    m, _ = tgat_pair(nf, ef, src, dst, eid, ts, 1, 2, p)
    with torch.no_grad():
        with pytest.raises(IndexError):
            m.compute_src_dst_node_temporal_embeddings(np.array([10 ** 6]), np.array([1]), np.array([5.0]), 5)
        with pytest.raises(AssertionError):
            m.compute_src_dst_node_temporal_embeddings(np.array([1]), np.array([1]), np.array([5.0]), 0)


@pytest.mark.parametrize("L,k,tdtype", [(2, 5, "f64"), (1, 20, "f64"), (2, 4, "f32")])
def test_tgat_training_mode_forward_and_gradients(L, k, tdtype):
    """Training-mode calls (M-step batches) go through the differentiable path: device-sampled
    neighbourhoods + torch CUDA ops.  With dropout = 0 the values must match the oracle and the
    eval kernels, and every parameter gradient must match the oracle's autograd."""
    src, dst, eid, ts, nf, ef = cases.small_stream()
    p = otgat.default_params(172, 172, 100, L, 2, seed=5, time_bias_scale=0.3)
    s = make_sampler(src, dst, eid, ts, nf.shape[0] - 1)
    m = flid_b200.TGAT(nf, ef, s, 100, L, 2, 0.0, DEV).to(DEV)
    m.load_state_dict({kk: v for kk, v in p.items() if not kk.startswith("_")})
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    sel = np.arange(440, 470)
    tt = ts[sel] if tdtype == "f64" else ts[sel].astype(np.float32)
    m.train()
    a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], tt, k)
    assert a.requires_grad and a.device.type == "cuda"
    g = torch.Generator().manual_seed(1)
    wa, wb = torch.randn(a.shape, generator=g), torch.randn(b.shape, generator=g)
    ((a * wa.to(DEV)).sum() + (b * wb.to(DEV)).sum()).backward()
    # oracle: same weights as leaf tensors
    po = {kk: (v.clone().requires_grad_(True) if torch.is_tensor(v) else v) for kk, v in p.items()}
    nft, eft = torch.from_numpy(nf), torch.from_numpy(ef)
    oa = otgat.embed(po, nft, eft, o, src[sel], tt, L, k)
    ob = otgat.embed(po, nft, eft, o, dst[sel], tt, L, k)
    ((oa * wa).sum() + (ob * wb).sum()).backward()
    assert_fp32_close(a.detach().cpu().numpy(), oa.detach().numpy(), "train-mode src")
    assert_fp32_close(b.detach().cpu().numpy(), ob.detach().numpy(), "train-mode dst")
    for name, prm in m.named_parameters():
        want = po[name].grad
        assert prm.grad is not None, name
        got = prm.grad.cpu()
        # Frobenius-relative: a ReLU unit whose pre-activation sits within rounding of 0 may flip between the
        # two implementations and change one row of a weight gradient, which a max-norm test would trip on
        rel = float((got - want).norm()) / max(float(want.norm()), 1e-6)
        assert rel <= 2e-3, (name, rel)
    # the eval kernels agree with the differentiable path
    m.eval()
    with torch.no_grad():
        ea, eb = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], tt, k)
    assert_fp32_close(ea.cpu().numpy(), a.detach().cpu().numpy(), "eval kernels vs training path")
    # dropout is live in training mode
    m2 = flid_b200.TGAT(nf, ef, s, 100, L, 2, 0.5, DEV).to(DEV)
    m2.load_state_dict(m.state_dict())
    m2.train()
    d1, _ = m2.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], tt, k)
    d2, _ = m2.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], tt, k)
    assert not torch.equal(d1, d2)


# ===================================================================== TGN
TGN_CASES = [("L1_k5", 1, 5, 25, 12, 0.3), ("L2_k4", 2, 4, 20, 8, 0.0)]


def tgn_model(nf, ef, src, dst, eid, ts, L, p):
    s = make_sampler(src, dst, eid, ts, nf.shape[0] - 1)
    m = flid_b200.MemoryModel(nf, ef, s, 100, "TGN", L, 2, 0.1, device=DEV).to(DEV)
    missing = m.load_state_dict({k: v for k, v in p.items() if not k.startswith("_")}, strict=False)
    assert not missing.unexpected_keys
    m.eval()
    m.memory_bank.__init_memory_bank__()
    return m


@pytest.mark.parametrize("name,L,k,bs,nb,bias", TGN_CASES)
def test_tgn_golden(name, L, k, bs, nb, bias):
    g = load("tgn.npz")
    src, dst, eid, ts, nf, ef = cases.small_stream(num_nodes=30, num_edges=400, seed=11, t_max=2.0e6)
    p = otgn.default_params(172, 172, 100, L, 2, seed=5, time_bias_scale=bias)
    m = tgn_model(nf, ef, src, dst, eid, ts, L, p)
    with torch.no_grad():
        for b in range(nb):
            lo, hi = b * bs, (b + 1) * bs
            a, c = m.compute_src_dst_node_temporal_embeddings(src[lo:hi], dst[lo:hi], ts[lo:hi], eid[lo:hi], True, k)
            assert_fp32_close(torch.cat([a, c]).cpu().numpy(), g[name + "_emb"][b], f"{name} batch {b}")
        a, c = m.compute_src_dst_node_temporal_embeddings(src[hi:hi + bs], dst[hi:hi + bs][::-1].copy(),
                                                          ts[hi:hi + bs], eid[hi:hi + bs], False, k)
        assert_fp32_close(torch.cat([a, c]).cpu().numpy(), g[name + "_neg"], f"{name} negative edges")
    assert_fp32_close(m.memory_bank.node_memories.cpu().numpy(), g[name + "_mem"], "bank memories")
    assert np.array_equal(m.memory_bank.node_last_updated_times.cpu().numpy(), g[name + "_lastupd"])
    msgs = m.memory_bank.node_raw_messages
    pend = sorted(v for v, l in msgs.items() if len(l) > 0)
    assert np.array_equal(np.array(pend), g[name + "_pend_ids"])
    assert_fp32_close(np.stack([msgs[v][-1][0].cpu().numpy() for v in pend]), g[name + "_pend_msg"], "raw messages")
    assert np.array_equal(np.array([msgs[v][-1][1] for v in pend]), g[name + "_pend_ts"])


def test_tgn_state_round_trips_and_pass_driver():
    """backup/reload, state_dict + node_raw_messages checkpoint round trip, and the full-pass
    driver all reproduce the straight run bit for bit; long run vs the oracle."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    src, dst, eid, ts, nf, ef = cases.small_stream(num_nodes=60, num_edges=1200, seed=13, t_max=3.0e6)
    p = otgn.default_params(172, 172, 100, 1, 2, seed=6, time_bias_scale=0.1)
    m = tgn_model(nf, ef, src, dst, eid, ts, 1, p)
    bs, k = 40, 10
    ref_s, ref_d = passes.tgn_pass(m, src, dst, ts, eid, bs, k)
    # oracle over the same 30 batches (state carry across >= 30 consecutive batches)
    o = otgn.OracleTGN(p, torch.from_numpy(nf), torch.from_numpy(ef),
                       osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1), 1, k)
    for lo in range(0, 1200, bs):
        a, c = o.step(src[lo:lo + bs], dst[lo:lo + bs], ts[lo:lo + bs], eid[lo:lo + bs], True)
        assert_fp32_close(ref_s[lo:lo + bs].cpu().numpy(), a.numpy(), f"tgn src batch {lo // bs}")
        assert_fp32_close(ref_d[lo:lo + bs].cpu().numpy(), c.numpy(), f"tgn dst batch {lo // bs}")
    assert_fp32_close(m.memory_bank.node_memories.cpu().numpy(), o.mem.numpy(), "final memories")
    # replay with a backup / detour / reload in the middle, and a checkpoint round trip
    m.memory_bank.__init_memory_bank__()
    outs = []
    with torch.no_grad():
        for lo in range(0, 1200, bs):
            if lo == 400:
                bk = m.memory_bank.backup_memory_bank()
                m.compute_src_dst_node_temporal_embeddings(src[800:840], dst[800:840], ts[800:840], eid[800:840], True, k)
                m.memory_bank.reload_memory_bank(bk)
            if lo == 800:
                sd = {kk: v.clone() for kk, v in m.state_dict().items()}
                raw = m.memory_bank.node_raw_messages
                m.memory_bank.__init_memory_bank__()
                m.load_state_dict(sd)
                m.memory_bank.node_raw_messages = raw
            a, c = m.compute_src_dst_node_temporal_embeddings(src[lo:lo + bs], dst[lo:lo + bs], ts[lo:lo + bs],
                                                              eid[lo:lo + bs], True, k)
            outs.append(a)
    assert torch.equal(torch.cat(outs), ref_s)


def test_tgn_time_travel_is_rejected():
    src, dst, eid, ts, nf, ef = cases.small_stream(num_nodes=30, num_edges=400, seed=11, t_max=2.0e6)
    p = otgn.default_params(172, 172, 100, 1, 2, seed=5)
    m = tgn_model(nf, ef, src, dst, eid, ts, 1, p)
    with torch.no_grad():
        m.compute_src_dst_node_temporal_embeddings(src[200:220], dst[200:220], ts[200:220], eid[200:220], True, 5)
        m.compute_src_dst_node_temporal_embeddings(src[220:240], dst[220:240], ts[220:240], eid[220:240], True, 5)
        with pytest.raises(AssertionError, match="time in the past"):
            for lo in (0, 20, 40, 60):
                m.compute_src_dst_node_temporal_embeddings(src[lo:lo + 20], dst[lo:lo + 20], ts[lo:lo + 20],
                                                           eid[lo:lo + 20], True, 5)


def test_tgn_training_mode_matches_eval_kernels_and_finite_differences():
    """Training-mode TGN batches take the differentiable path (torch GRU + attention on device-sampled
    neighbourhoods, same C state update).  With dropout = 0 it must track the eval kernels batch by
    batch, leave the same memory state, and its gradients must agree with central finite differences
    taken through the *eval kernels* (an independent implementation)."""
    src, dst, eid, ts, nf, ef = cases.small_stream()
    L, k, bs = 1, 5, 20
    p = otgn.default_params(172, 172, 100, L, 2, seed=8, time_bias_scale=0.2)
    s = make_sampler(src, dst, eid, ts, nf.shape[0] - 1)

    def build(dropout):
        m = flid_b200.MemoryModel(nf, ef, s, 100, "TGN", L, 2, dropout, device=DEV).to(DEV)
        missing = m.load_state_dict({kk: v for kk, v in p.items() if not kk.startswith("_")}, strict=False)
        assert not missing.unexpected_keys
        m.memory_bank.__init_memory_bank__()
        return m

    ev, tr = build(0.0), build(0.0)
    ev.eval(), tr.train()
    g = torch.Generator().manual_seed(3)
    for bi in range(6):
        sl = slice(200 + bi * bs, 200 + (bi + 1) * bs)
        if bi == 5:
            backup_ev = ev.memory_bank.backup_memory_bank()
        with torch.no_grad():
            ea, eb = ev.compute_src_dst_node_temporal_embeddings(src[sl], dst[sl], ts[sl], eid[sl], True, k)
        ta, tb = tr.compute_src_dst_node_temporal_embeddings(src[sl], dst[sl], ts[sl], eid[sl], True, k)
        assert ta.requires_grad
        assert_fp32_close(ta.detach().cpu().numpy(), ea.cpu().numpy(), f"batch {bi} src")
        assert_fp32_close(tb.detach().cpu().numpy(), eb.cpu().numpy(), f"batch {bi} dst")
        tr.memory_bank.detach_memory_bank()
    assert torch.equal(ev.memory_bank.node_memories.data, tr.memory_bank.node_memories.data)
    assert torch.equal(ev.memory_bank.node_last_updated_times.data, tr.memory_bank.node_last_updated_times.data)
    assert torch.equal(ev.memory_bank._pending_msg, tr.memory_bank._pending_msg)
    # gradients of the last batch
    wa, wb = torch.randn(ta.shape, generator=g).to(DEV), torch.randn(tb.shape, generator=g).to(DEV)
    ((ta * wa).sum() + (tb * wb).sum()).backward()
    cell = tr.memory_updater.memory_updater
    for prm in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, tr.time_encoder.w.bias,
                tr.embedding_module.merge_layers[0].fc2.bias):
        assert prm.grad is not None and torch.isfinite(prm.grad).all() and float(prm.grad.abs().max()) > 0

    def eval_loss():
        ev.memory_bank.reload_memory_bank(backup_ev)
        with torch.no_grad():
            a, b = ev.compute_src_dst_node_temporal_embeddings(src[sl], dst[sl], ts[sl], eid[sl], False, k)
        return float(((a * wa).sum() + (b * wb).sum()).double())

    checks = [("memory_updater.memory_updater.bias_ih", cell.bias_ih, ev.memory_updater.memory_updater.bias_ih),
              ("memory_updater.memory_updater.bias_hh", cell.bias_hh, ev.memory_updater.memory_updater.bias_hh),
              ("merge fc2 bias", tr.embedding_module.merge_layers[0].fc2.bias, ev.embedding_module.merge_layers[0].fc2.bias)]
    for name, ptr_tr, ptr_ev in checks:
        j = int(ptr_tr.grad.abs().argmax())
        eps = 2e-2
        with torch.no_grad():
            ptr_ev[j] += eps
        lp = eval_loss()
        with torch.no_grad():
            ptr_ev[j] -= 2 * eps
        lm = eval_loss()
        with torch.no_grad():
            ptr_ev[j] += eps
        fd, an = (lp - lm) / (2 * eps), float(ptr_tr.grad[j])
        assert abs(fd - an) <= 3e-2 * max(1.0, abs(an)), (name, fd, an)


@pytest.mark.parametrize("L,k", [(1, 5), (2, 4)])
def test_tgn_training_gradients_match_oracle_autograd(L, k):
    """Training-mode TGN batch after a few state-building batches: embeddings and EVERY parameter gradient
    (GRU updater, time encoder, attention, merge) against the oracle's autograd on the same state."""
    src, dst, eid, ts, nf, ef = cases.small_stream()
    bs = 20
    p = otgn.default_params(172, 172, 100, L, 2, seed=12, time_bias_scale=0.2)
    s = make_sampler(src, dst, eid, ts, nf.shape[0] - 1)
    m = flid_b200.MemoryModel(nf, ef, s, 100, "TGN", L, 2, 0.0, device=DEV).to(DEV)
    missing = m.load_state_dict({kk: v for kk, v in p.items() if not kk.startswith("_")}, strict=False)
    assert not missing.unexpected_keys
    m.memory_bank.__init_memory_bank__()
    m.train()
    po = {kk: (v.clone().requires_grad_(True) if torch.is_tensor(v) and v.dtype.is_floating_point else v)
          for kk, v in p.items()}
    o = otgn.OracleTGN(po, torch.from_numpy(nf), torch.from_numpy(ef),
                       osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1), L, k)
    for bi in range(5):
        sl = slice(300 + bi * bs, 300 + (bi + 1) * bs)
        last = bi == 4
        a, b = m.compute_src_dst_node_temporal_embeddings(src[sl], dst[sl], ts[sl], eid[sl], True, k)
        oa, ob = o.step(src[sl], dst[sl], ts[sl], eid[sl], True, grad=last)
        assert_fp32_close(a.detach().cpu().numpy(), oa.detach().numpy(), f"batch {bi} src")
        assert_fp32_close(b.detach().cpu().numpy(), ob.detach().numpy(), f"batch {bi} dst")
        if not last:
            m.memory_bank.detach_memory_bank()
    g = torch.Generator().manual_seed(4)
    wa, wb = torch.randn(a.shape, generator=g), torch.randn(b.shape, generator=g)
    ((a * wa.to(DEV)).sum() + (b * wb.to(DEV)).sum()).backward()
    ((oa * wa).sum() + (ob * wb).sum()).backward()
    checked = 0
    for name, prm in m.named_parameters():
        if name.startswith("memory_bank.") or name.startswith("memory_updater.memory_bank."):
            continue                      # the bank is state, detached between batches (models/MemoryModel.py:440-445)
        want = po[name].grad if name in po else None
        if want is None:
            continue
        assert prm.grad is not None, name
        rel = float((prm.grad.cpu() - want).norm()) / max(float(want.norm()), 1e-6)
        assert rel <= 2e-3, (name, rel)
        checked += 1
    assert checked >= 2 + 11 * L + 4


# ===================================================================== pseudo labels
@pytest.mark.parametrize("C", [2, 5])
def test_pseudo_label_golden(C):
    g = load("pseudo.npz")
    rs = np.random.RandomState(21)
    emb = None
    for c in (2, 5):
        e = rs.standard_normal((700, 172)).astype(np.float32) * 2.0
        if c == C:
            emb = e
            break
        rs.standard_normal((2, 700, c)), rs.randint(0, c, 700), rs.uniform(0, 1000, 700), rs.rand(700)
    p = opseudo.default_decoder_params(172, C, seed=C)
    dec = flid_b200.MLPClassifier(172, 0.1, C).to(DEV)
    dec.load_state_dict(p)
    dec.eval()
    x = torch.from_numpy(emb).to(DEV)
    with torch.no_grad():
        logits = dec(x)
    assert_fp32_close(logits.cpu().numpy(), g[f"C{C}_logits"], "decoder logits")
    labels, probs = flid_b200.emit_pseudo_labels(dec, x)
    assert_fp32_close(probs.cpu().numpy(), g[f"C{C}_probs"], "probabilities")
    margin = np.sort(g[f"C{C}_probs"], axis=1)
    clear = (margin[:, -1] - margin[:, -2]) > 1e-5
    assert np.array_equal(labels.cpu().numpy()[clear], g[f"C{C}_labels"][clear]) and clear.mean() > 0.99
    # masks from the reference's own fp32 probabilities: must be identical
    store = [torch.from_numpy(s).to(DEV) for s in g[f"C{C}_store"]]
    glab = torch.from_numpy(g[f"C{C}_labels"]).to(torch.float32)
    for thr in (0.3, 0.6, 0.9):
        ps = glab.reshape(1, -1).clone().to(DEV)
        assert np.array_equal(flid_b200.entropy_filter(ps, store, thr).cpu().numpy(), g[f"C{C}_est_{thr}"])
        ps = glab.reshape(1, -1).clone().to(DEV)
        assert np.array_equal(flid_b200.prob_filter(ps, store, thr).cpu().numpy(), g[f"C{C}_cst_{thr}"])
    ps = glab.reshape(2, 350).clone().to(DEV)
    store2 = [s.reshape(2, 350, C) for s in store]
    assert np.array_equal(flid_b200.entropy_filter(ps, store2, 0.6).cpu().numpy(), g[f"C{C}_est2_0.6"])

    class D:
        pass
    full = D()
    full.labels, full.labels_time, full.node_interact_times = g[f"C{C}_true"], g[f"C{C}_lt"], g[f"C{C}_it"]
    for ut in (0, 1):
        ps = glab.reshape(1, -1).clone().to(DEV)
        data = {"full_data": full, "val_offest": np.int64(400), "dataset_name": "wikipedia"}
        r = flid_b200.update_pseudo_labels(data, ps, store, [], "ps", use_transductive=ut, threshold=0.6,
                                           ps_filter="entropy")
        assert np.array_equal(r.cpu().numpy(), g[f"C{C}_upd_ut{ut}"])


@pytest.mark.parametrize("C", [2, 5])
def test_pseudo_label_bulk_path_matches_oracle_and_row_kernel(C):
    """n >= 4096 goes through the tensor-core fc1 + tail kernel; it must agree with the oracle and
    with the one-warp-per-row kernel used for per-batch calls."""
    rs = np.random.RandomState(33)
    emb = (rs.standard_normal((9001, 172)) * 1.5).astype(np.float32)
    p = opseudo.default_decoder_params(172, C, seed=7 + C)
    dec = flid_b200.MLPClassifier(172, 0.1, C).to(DEV)
    dec.load_state_dict(p)
    dec.eval()
    x = torch.from_numpy(emb).to(DEV)
    labels, probs = flid_b200.emit_pseudo_labels(dec, x)
    parts = [flid_b200.emit_pseudo_labels(dec, x[i:i + 1000]) for i in range(0, len(emb), 1000)]
    l_small, p_small = torch.cat([a for a, _ in parts]), torch.cat([b for _, b in parts])
    wl, wp = opseudo.emit(p, torch.from_numpy(emb))
    assert_fp32_close(probs.cpu().numpy(), wp.numpy(), "bulk probabilities vs oracle")
    assert_fp32_close(probs.cpu().numpy(), p_small.cpu().numpy(), "bulk vs row kernel")
    margin = np.sort(wp.numpy(), axis=1)
    clear = (margin[:, -1] - margin[:, -2]) > 1e-5
    assert clear.mean() > 0.99
    assert np.array_equal(labels.cpu().numpy()[clear], wl.numpy()[clear])
    assert np.array_equal(labels.cpu().numpy()[clear], l_small.cpu().numpy()[clear])


def test_e_step_pass_matches_oracle_pipeline():
    """configs[2] in miniature: TGAT L=2 k=20 embeddings -> decoder -> EST filter over 3 stored
    iterations, against the oracle pipeline.  Masks are compared where the oracle's entropy is
    not within 1e-4 of the threshold (embeddings carry fp32 tolerance, not bit equality)."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = synth.reddit_shape(seed=0, scale=0.01)
    o = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    sel = np.arange(g.num_interactions - 150, g.num_interactions)
    src, dst, ts = g.src_node_ids[sel], g.dst_node_ids[sel], g.node_interact_times[sel]
    store_gpu, store_cpu = [], []
    for it in range(3):
        p = otgat.default_params(172, 172, 100, 2, 2, seed=it)
        pd = opseudo.default_decoder_params(172, 2, seed=it)
        m, _ = tgat_pair(g.node_raw_features, g.edge_raw_features, g.src_node_ids, g.dst_node_ids, g.edge_ids,
                         g.node_interact_times, 2, 2, p)
        dec = flid_b200.MLPClassifier(172, 0.1, 2).to(DEV)
        dec.load_state_dict(pd)
        dec.eval()
        pseudo, probs, emb = passes.e_step_pass(m, dec, src, dst, ts, 20, store_gpu, "entropy", 0.9,
                                                return_embeddings=True)
        wa, _ = otgat.embed_src_dst(p, torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features),
                                    o, src, dst, ts, 2, 20)
        assert_fp32_close(emb[0].cpu().numpy(), wa.numpy(), f"iteration {it} embeddings")
        wl, wp = opseudo.emit(pd, wa)
        store_cpu.append(wp)
        assert_fp32_close(probs.cpu().numpy(), wp.numpy(), f"iteration {it} probs")
        want = opseudo.entropy_filter(wl.to(torch.float32).reshape(1, -1).clone(), store_cpu, 0.9)
        acc = torch.softmax(torch.stack(store_cpu).sum(0), dim=1)
        ent = -(acc * torch.log2(acc + 1e-10)).sum(1)
        clear = ((ent - 0.9).abs() > 1e-4) & ((wp[:, 0] - wp[:, 1]).abs() > 1e-4)
        assert clear.float().mean() > 0.95
        assert torch.equal(pseudo.cpu()[0][clear], want[0][clear])
    assert len(store_gpu) == 3


# ===================================================================== full-size properties
def test_full_size_reddit_shape_properties():
    """BASELINE.json configs[2] at full size (10 984 nodes / 672 447 edges), where the oracle would
    take the better part of an hour: size-independent properties instead.
      * sampler: right-aligned zero padding, strictly-earlier and non-decreasing timestamps, every
        returned (neighbour, edge, time) triple is an entry of the queried node, for all 1.34 M roots;
      * embeddings: the memoised bulk pass, the recursion and a permuted / re-chunked evaluation of
        the same roots agree bit for bit; a sample of roots agrees with the oracle within 1e-4."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = synth.reddit_shape(seed=0, scale=1.0)
    e = g.num_interactions
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=DEV)
    nodes = np.concatenate([g.src_node_ids, g.dst_node_ids])
    times = np.concatenate([g.node_interact_times, g.node_interact_times])
    nbr, eid, ts = s.get_historical_neighbors(nodes, times, 20)
    pad = nbr == 0
    assert (pad[:, :-1] >= pad[:, 1:]).all(), "padding must be on the left"
    assert ((eid == 0) == pad).all() and (ts[pad] == 0).all()
    assert (ts.astype(np.float64) < times[:, None])[~pad].all(), "neighbours must be strictly earlier"
    both = ~pad[:, 1:] & ~pad[:, :-1]
    assert (np.diff(ts, axis=1)[both] >= 0).all(), "timestamps must be non-decreasing"
    ev = eid[~pad] - 1                                            # edge ids are 1..E in event order
    owner = np.repeat(nodes, 20).reshape(-1, 20)[~pad]
    other = np.where(g.src_node_ids[ev] == owner, g.dst_node_ids[ev], g.src_node_ids[ev])
    assert ((g.src_node_ids[ev] == owner) | (g.dst_node_ids[ev] == owner)).all()
    assert (other == nbr[~pad]).all() and (g.node_interact_times[ev].astype(np.float32) == ts[~pad]).all()
    # the k most recent: the number of valid slots is min(k, number of earlier interactions of the node)
    order = np.argsort(nodes, kind="stable")
    # embeddings
    p = otgat.default_params(172, 172, 100, 2, 2, seed=2, time_bias_scale=0.1)
    m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 2, 2, 0.1, DEV).to(DEV)
    m.load_state_dict({k: v for k, v in p.items() if not k.startswith("_")})
    m.eval()
    with torch.no_grad():
        a, b = passes.embed_events(m, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 20)   # memoised bulk pass
        assert 2 in m._engine.memo
        rs = np.random.RandomState(0)
        sel = np.sort(rs.choice(e, 3000, replace=False))
        m.set_layer_memo(False)
        pa, pb = m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sel], g.dst_node_ids[sel],
                                                            g.node_interact_times[sel], 20)          # the recursion
        assert_paths_close(a[sel], pa, "bulk pass vs recursion (src)")      # projected bulk pass vs per-slot recursion
        assert_paths_close(b[sel], pb, "bulk pass vs recursion (dst)")
        perm = rs.permutation(len(sel))
        qa, _ = m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sel][perm], g.dst_node_ids[sel][perm],
                                                           g.node_interact_times[sel][perm], 20)
        assert torch.equal(qa, pa[perm])
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    o = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    few = sel[::100]
    wa, wb = otgat.embed_src_dst(p, torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features), o,
                                 g.src_node_ids[few], g.dst_node_ids[few], g.node_interact_times[few], 2, 20)
    assert_fp32_close(a[few].cpu().numpy(), wa.numpy(), "full-size src vs oracle")
    assert_fp32_close(b[few].cpu().numpy(), wb.numpy(), "full-size dst vs oracle")


def test_full_size_dsub_shape_properties():
    """BASELINE.json configs[3] at full size (150 000 nodes / 168 154 edges, k = 30, one year of seconds):
    most neighbourhoods are padded or empty (the all-masked softmax path) and about half of the event times
    are not float32-exact, so roots mix "own lower layers from the memo" and the full chain.  The memoised
    double-way bulk pass must equal the recursion bit for bit and agree with the oracle on a sample."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = synth.dsub_shape(seed=0, scale=1.0)
    e = g.num_interactions
    s = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=DEV)
    nodes = np.concatenate([g.src_node_ids, g.dst_node_ids])
    times = np.concatenate([g.node_interact_times, g.node_interact_times])
    nbr, eid, ts = s.get_historical_neighbors(nodes, times, 30)
    pad = nbr == 0
    assert (pad[:, :-1] >= pad[:, 1:]).all() and (ts.astype(np.float64) < times[:, None])[~pad].all()
    assert pad.all(axis=1).mean() > 0.2, "the shape is meant to exercise empty neighbourhoods"
    exact = times.astype(np.float32).astype(np.float64) == times
    assert 0.2 < exact.mean() < 0.9, "the shape is meant to mix float32-exact and inexact times"
    p = otgat.default_params(172, 172, 100, 2, 2, seed=4, time_bias_scale=0.1)
    m = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, s, 100, 2, 2, 0.1, DEV).to(DEV)
    m.load_state_dict({k: v for k, v in p.items() if not k.startswith("_")})
    m.eval()
    with torch.no_grad():
        a, b = passes.embed_events(m, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 30)
        assert 2 in m._engine.memo
        rs = np.random.RandomState(1)
        sel = np.sort(rs.choice(e, 2000, replace=False))
        m.set_layer_memo(False)
        pa, pb = m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sel], g.dst_node_ids[sel],
                                                            g.node_interact_times[sel], 30)
    assert_paths_close(a[sel], pa, "Dsub bulk pass vs recursion (src)")
    assert_paths_close(b[sel], pb, "Dsub bulk pass vs recursion (dst)")
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    o = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    few = sel[::50]
    wa, wb = otgat.embed_src_dst(p, torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features), o,
                                 g.src_node_ids[few], g.dst_node_ids[few], g.node_interact_times[few], 2, 30)
    assert_fp32_close(a[few].cpu().numpy(), wa.numpy(), "Dsub full-size src vs oracle")
    assert_fp32_close(b[few].cpu().numpy(), wb.numpy(), "Dsub full-size dst vs oracle")


# ===================================================================== round-2 hardening
@pytest.mark.parametrize("memo", [False, True])
def test_tgat_more_than_32_neighbors(memo):
    """num_neighbors above one warp's 32 slots (the reference takes any k, utils/load_configs.py:114):
    hubs fill all 40 slots, most nodes are partly or fully padded."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = synth.wikipedia_shape(seed=3, scale=0.05)
    p = otgat.default_params(172, 172, 100, 2, 2, seed=8, time_bias_scale=0.2)
    m, s = tgat_pair(g.node_raw_features, g.edge_raw_features, g.src_node_ids, g.dst_node_ids, g.edge_ids,
                     g.node_interact_times, 2, 2, p, memo)
    o = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    sel = np.arange(g.num_interactions - 40, g.num_interactions)
    src, dst, ts = g.src_node_ids[sel], g.dst_node_ids[sel], g.node_interact_times[sel]
    got = s.get_historical_neighbors(src, ts, 40)
    want = o.get_historical_neighbors(src, ts, 40)
    assert all(np.array_equal(x, y) for x, y in zip(got, want))
    assert (got[0] != 0).all(axis=1).any() and (got[0] == 0).any(), "case must mix full and padded neighbourhoods"
    with torch.no_grad():
        a, b = m.compute_src_dst_node_temporal_embeddings(src, dst, ts, 40)
    wa, wb = otgat.embed_src_dst(p, torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features), o,
                                 src, dst, ts, 2, 40)
    assert_fp32_close(a.cpu().numpy(), wa.numpy(), "k=40 src")
    assert_fp32_close(b.cpu().numpy(), wb.numpy(), "k=40 dst")


def test_training_path_rejects_more_than_32_neighbors():
    """The training-mode kernels keep one slot per lane; the limit is an explicit error, not a wrong result."""
    src, dst, eid, ts, nf, ef = cases.small_stream()
    p = otgat.default_params(172, 172, 100, 1, 2, seed=3)
    m, _ = tgat_pair(nf, ef, src, dst, eid, ts, 1, 2, p)
    m.train()
    with pytest.raises(ValueError, match="num_neighbors must be <= 32"):
        m.compute_src_dst_node_temporal_embeddings(src[:4], dst[:4], ts[:4], 33)


def test_tgn_wikipedia_shape_60_batches_vs_oracle():
    """configs[1] at the Wikipedia shape (9 227 nodes, hubs with thousands of interactions): 60 consecutive
    batches of 200 events taken from the middle of the stream (rich histories, many pending messages,
    nodes in both roles within a batch), every batch and the final bank against the oracle."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = synth.wikipedia_shape(seed=0, scale=1.0)
    p = otgn.default_params(172, 172, 100, 1, 2, seed=9, time_bias_scale=0.1)
    m = tgn_model(g.node_raw_features, g.edge_raw_features, g.src_node_ids, g.dst_node_ids, g.edge_ids,
                  g.node_interact_times, 1, p)
    o = otgn.OracleTGN(p, torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features),
                       osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids,
                                                       g.node_interact_times, g.num_nodes), 1, 20)
    lo0, bs, nb = 100000, 200, 60
    deg = np.bincount(np.concatenate([g.src_node_ids[:lo0], g.dst_node_ids[:lo0]]))
    assert deg.max() > 5000, "the shape is meant to have hub nodes"
    with torch.no_grad():
        for b in range(nb):
            sl = slice(lo0 + b * bs, lo0 + (b + 1) * bs)
            a, c = m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sl], g.dst_node_ids[sl],
                                                              g.node_interact_times[sl], g.edge_ids[sl], True, 20)
            wa, wc = o.step(g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl], g.edge_ids[sl], True)
            assert_fp32_close(a.cpu().numpy(), wa.numpy(), f"wikipedia-shape TGN src batch {b}")
            assert_fp32_close(c.cpu().numpy(), wc.numpy(), f"wikipedia-shape TGN dst batch {b}")
    assert_fp32_close(m.memory_bank.node_memories.cpu().numpy(), o.mem.numpy(), "bank after 60 batches")
    assert np.array_equal(m.memory_bank.node_last_updated_times.cpu().numpy(), o.last_upd.numpy())
    pend = sorted(v for v, l in m.memory_bank.node_raw_messages.items() if len(l) > 0)
    assert pend == sorted(v for v, l in o.msgs.items() if len(l) > 0) and len(pend) > 1000


def test_tgn_train_prefix_sampler_with_full_size_bank():
    """PTCL/EM_warmup.py:71: the train sampler is built from train_data only, so it may know fewer nodes than
    the bank (one row per node of the full graph) has rows.  Batches inside the prefix work and match the
    oracle; a node the sampler has no list for is the reference's IndexError."""
    src, dst, eid, ts, nf, ef = cases.small_stream(num_nodes=60, num_edges=1200, seed=13, t_max=3.0e6)
    cut = 600
    keep = np.maximum(src[:cut], dst[:cut]) < 50
    ps, pd, pe, pt = src[:cut][keep], dst[:cut][keep], eid[:cut][keep], ts[:cut][keep]
    n_prefix = int(max(ps.max(), pd.max()))
    assert n_prefix < nf.shape[0] - 1
    p = otgn.default_params(172, 172, 100, 1, 2, seed=6, time_bias_scale=0.1)
    s = make_sampler(ps, pd, pe, pt, n_prefix)
    m = flid_b200.MemoryModel(nf, ef, s, 100, "TGN", 1, 2, 0.1, device=DEV).to(DEV)
    m.load_state_dict({k: v for k, v in p.items() if not k.startswith("_")}, strict=False)
    m.eval()
    m.memory_bank.__init_memory_bank__()
    o = otgn.OracleTGN(p, torch.from_numpy(nf), torch.from_numpy(ef),
                       osamp.OracleSampler.from_events(ps, pd, pe, pt, n_prefix), 1, 10)
    with torch.no_grad():
        for lo in range(0, len(ps) - 40, 40):
            sl = slice(lo, lo + 40)
            a, c = m.compute_src_dst_node_temporal_embeddings(ps[sl], pd[sl], pt[sl], pe[sl], True, 10)
            wa, wc = o.step(ps[sl], pd[sl], pt[sl], pe[sl], True)
            assert_fp32_close(torch.cat([a, c]).cpu().numpy(), torch.cat([wa, wc]).numpy(), f"prefix batch {lo // 40}")
        bad = np.array([nf.shape[0] - 1], dtype=np.int64)
        with pytest.raises(IndexError):
            m.compute_src_dst_node_temporal_embeddings(bad, ps[:1], pt[-1:], pe[:1], True, 10)


def test_invalidate_caches_after_data_writes():
    """``param.data`` writes do not bump autograd's version counter: an explicit invalidate picks them up,
    and a rebuilt sampler never inherits the previous sampler's layer memo."""
    src, dst, eid, ts, nf, ef = cases.small_stream()
    p = otgat.default_params(172, 172, 100, 2, 2, seed=3)
    m, s = tgat_pair(nf, ef, src, dst, eid, ts, 2, 2, p, True)
    sel = np.arange(300, 340)
    with torch.no_grad():
        a0, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 5)
        m.merge_layers[1].fc2.bias.data.add_(1.0)
        m.invalidate_caches()
        a1, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 5)
        assert_fp32_close((a1 - a0).cpu().numpy(), np.ones_like(a0.cpu().numpy()), "bias shift after invalidate")
        gens = {s.generation}
        for _ in range(3):
            s2 = make_sampler(src[:200], dst[:200], eid[:200], ts[:200], nf.shape[0] - 1)
            assert s2.generation not in gens
            gens.add(s2.generation)
            m.set_neighbor_sampler(s2)
            b1, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 5)
            m.set_layer_memo(False)
            b2, _ = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 5)
            m.set_layer_memo(True)
            assert_fp32_close(b1.cpu().numpy(), b2.cpu().numpy(), "memo of a rebuilt sampler")
            del s2


def test_decoder_eval_mode_with_grad_keeps_the_graph():
    dec = flid_b200.MLPClassifier(172, 0.1, 2).to(DEV)
    dec.eval()
    x = torch.randn(8, 172, device=DEV, requires_grad=True)
    out = dec(x)
    assert out.requires_grad
    out.sum().backward()
    assert x.grad is not None and float(x.grad.abs().sum()) > 0
    with torch.no_grad():
        fused = dec(x)
    assert_fp32_close(fused.cpu().numpy(), out.detach().cpu().numpy(), "fused decoder vs torch path")


def _tail_decoder(emb, offset_sigmas, seed=0):
    """MLPClassifier whose decision function is one random direction of the embedding with the boundary
    ``offset_sigmas`` standard deviations off the mean (the reference's datasets have a few per cent of
    positive labels, so a trained decoder's boundary sits in the tail of the margin distribution)."""
    dec = flid_b200.MLPClassifier(172, 0.1, 2).to(DEV)
    dec.eval()
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        x = emb.float()
        v = torch.randn(172, generator=gen).to(DEV)
        proj = (x - x.mean(0)) @ v
        v = v / proj.std() * 1.5
        c = (x.mean(0) @ v) + 1.5 * offset_sigmas
        for lin in (dec.fc1, dec.fc2, dec.fc3):
            lin.weight.zero_(), lin.bias.zero_()
        dec.fc1.weight[0], dec.fc1.weight[1] = v, -v
        dec.fc1.bias[0], dec.fc1.bias[1] = -c, c
        dec.fc2.weight[0, 0] = dec.fc2.weight[1, 1] = dec.fc3.weight[0, 0] = dec.fc3.weight[1, 1] = 1.0
    return dec


def test_bf16_projection_mode_tolerance_and_argmax():
    """BASELINE.json north_star, second numeric mode: bf16-rounded operands / fp32 accumulation in the Q/K/V/out
    and MergeLayer projections.  > 100 000 Reddit-shape roots: embeddings and logits within rel 2e-2 of the
    fp32 path (itself pinned to the oracle) and of the oracle on a sample, identical argmax on >= 99.9 % of roots."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = synth.reddit_shape(seed=0, scale=0.08)
    e = g.num_interactions
    assert 2 * e >= 100_000
    p = otgat.default_params(172, 172, 100, 2, 2, seed=2, time_bias_scale=0.1)
    m, s = tgat_pair(g.node_raw_features, g.edge_raw_features, g.src_node_ids, g.dst_node_ids, g.edge_ids,
                     g.node_interact_times, 2, 2, p, True)
    with torch.no_grad():
        a32, b32 = passes.embed_events(m, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 20)
        assert flid_b200.get_numeric_mode() == "f32"
        flid_b200.set_numeric_mode("bf16")
        try:
            a16, b16 = passes.embed_events(m, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 20)
            sel = np.arange(e - 64, e)
            pa, _ = m.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sel], g.dst_node_ids[sel],
                                                               g.node_interact_times[sel], 20)     # per-batch call
        finally:
            flid_b200.set_numeric_mode("f32")
        a32b, _ = passes.embed_events(m, g.src_node_ids, g.dst_node_ids, g.node_interact_times, 20)
    assert torch.equal(a32, a32b), "switching the mode back must restore the fp32 results bit for bit"
    x32, x16 = torch.cat([a32, b32]), torch.cat([a16, b16])
    scale = max(1.0, float(x32.abs().max()))
    err = (x16 - x32).abs()
    assert float(err.max()) <= 2e-2 * scale, f"bf16 mode: max abs err {float(err.max()):.3e} (scale {scale:.3g})"
    assert 1e-5 * scale < float(err.mean()) <= 3e-3 * scale, f"bf16 mode: mean abs err {float(err.mean()):.3e}"
    assert float((pa - a16[sel]).abs().max()) <= 2e-2 * scale
    # oracle (fp32 reference arithmetic) on a sample
    o = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    few = np.sort(np.random.RandomState(0).choice(e, 150, replace=False))
    wa, wb = otgat.embed_src_dst(p, torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features), o,
                                 g.src_node_ids[few], g.dst_node_ids[few], g.node_interact_times[few], 2, 20)
    assert float((a16[few].cpu() - wa).abs().max()) <= 2e-2 * scale
    assert float((b16[few].cpu() - wb).abs().max()) <= 2e-2 * scale
    # labels: boundary two standard deviations off the mean (~2 % positives)
    dec = _tail_decoder(x32, 2.0)
    pr32, lab32, lg32 = dec.score(x32, want_logits=True)
    pr16, lab16, lg16 = dec.score(x16, want_logits=True)
    pos = float(lab32.float().mean())
    assert 0.002 < min(pos, 1 - pos) < 0.2, pos
    agree = float((lab32 == lab16).float().mean())
    assert agree >= 0.999, f"bf16 mode: argmax agreement {agree:.5f} on {x32.shape[0]} roots"
    lscale = max(1.0, float(lg32.abs().max()))
    assert float((lg16 - lg32).abs().max()) <= 2e-2 * lscale


def test_tgn_pass_graph_equals_per_batch_calls():
    """flid_tgn_pass (one C call, CUDA graph per batch, device-side batch counter) against the reference's own loop
    of per-batch calls: embeddings, bank, last-update times and pending messages identical bit for bit; ragged
    last batch; the monotone-time assertion still fires."""
    g = synth.wikipedia_shape(seed=2, scale=0.05)
    p = otgn.default_params(172, 172, 100, 2, 2, seed=4, time_bias_scale=0.1)
    e = g.num_interactions - 37            # not a multiple of the batch size
    src, dst, ts, eid = g.src_node_ids[:e], g.dst_node_ids[:e], g.node_interact_times[:e], g.edge_ids[:e]
    ref = None
    for mode in ("calls", "graph", "nograph"):
        m = tgn_model(g.node_raw_features, g.edge_raw_features, g.src_node_ids, g.dst_node_ids, g.edge_ids,
                      g.node_interact_times, 2, p)
        if mode == "calls":
            a, c = passes.tgn_pass(m, src, dst, ts, eid, 200, 10, per_batch_calls=True)
        else:
            m.memory_bank.__init_memory_bank__()
            a, c = m.embed_pass(src, dst, ts, eid, 200, 10, use_graph=(mode == "graph"))
        state = (a, c, m.memory_bank.node_memories.data.clone(), m.memory_bank.node_last_updated_times.data.clone(),
                 m.memory_bank._pending_msg.clone(), m.memory_bank._pending_ts.clone(), m.memory_bank._has_pending.clone())
        if ref is None:
            ref = state
        else:
            for x, y in zip(state, ref):
                assert torch.equal(x, y), mode
    with pytest.raises(AssertionError, match="time in the past"):
        m.embed_pass(src[:400], dst[:400], ts[:400], eid[:400], 200, 10)      # the bank is already at the end of the stream


def test_bulk_sort_knob_and_wait_event_do_not_change_bits():
    """flid_tgat_set_sort_queries (the owner-partitioned pass hands its roots over already in (node, time) order and
    switches the per-call sort off) and flid_tgat_set_wait_event (the exchange of the last memo level on a side
    stream) change the schedule, not the results: same bits for sorted, unsorted and pre-sorted root lists."""
    g = synth.wikipedia_shape(seed=4, scale=0.08)
    src, dst, eid, ts, nf, ef = (g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times,
                                 g.node_raw_features, g.edge_raw_features)
    L, k = 2, 10
    p = otgat.default_params(172, 172, 100, L, 2, seed=21, time_bias_scale=0.25)
    m, s = tgat_pair(nf, ef, src, dst, eid, ts, L, 2, p, memo=True)
    nodes = np.concatenate([src, dst])
    times = np.concatenate([ts, ts])
    assert len(nodes) >= 8192                                  # the bulk branch that sorts its roots
    with torch.no_grad():
        base = m.compute_node_temporal_embeddings(nodes, times, L, k)
        m._engine.presorted = True                             # what passes._owned_roots sets around its call
        try:
            unsorted = m.compute_node_temporal_embeddings(nodes, times, L, k)
            order = np.lexsort((times, nodes))
            pre = m.compute_node_temporal_embeddings(nodes[order], times[order], L, k)
        finally:
            m._engine.presorted = False
        again = m.compute_node_temporal_embeddings(nodes, times, L, k)
    assert torch.equal(base, unsorted) and torch.equal(base, again)
    assert torch.equal(base[torch.from_numpy(order).to(base.device)], pre)
    # an event recorded on a side stream after some unrelated work: the call waits for it and forgets it
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        junk = torch.randn(1 << 22, device=DEV).sin_()
        ev = torch.cuda.Event()
        ev.record(side)
    h = m._engine.handles[L]
    _lib.check(_lib.lib().flid_tgat_set_wait_event(h, ctypes.c_void_p(ev.cuda_event)))
    with torch.no_grad():
        waited = m.compute_node_temporal_embeddings(nodes, times, L, k)
    assert ev.query() and torch.equal(base, waited)
    del junk
