"""Pin the CPU oracle (oracle/) against the golden vectors minted from the real
FLiD reference (tests/golden/make_golden.py), and against the live reference
when /root/reference is present.  CPU only."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import sampler as osamp, tgat as otgat, tgn as otgn, pseudo as opseudo, graphmixer as omix, tcl as otcl, ref_shim

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name))


# ---------------------------------------------------------------- sampler
def _adv_sampler():
    src, dst, eid, ts, n = cases.adversarial_events()
    return osamp.OracleSampler.from_events(src, dst, eid, ts, n)


def test_sampler_inputs_regenerate():
    g = load("sampler.npz")
    src, dst, eid, ts, n = cases.adversarial_events()
    nodes, times = cases.adversarial_queries()
    assert cases.checksum(src, dst, eid, ts, nodes, times) == g["in_checksum"]


@pytest.mark.parametrize("k", [1, 3, 20])
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_sampler_golden(k, dt):
    g = load("sampler.npz")
    s = _adv_sampler()
    nodes, times = cases.adversarial_queries()
    if dt == "f32":
        times = times.astype(np.float32)
    for fn in (s.get_historical_neighbors, s.get_historical_neighbors_loop):
        a, b, c = fn(nodes, times, k)
        assert a.dtype == np.int64 and b.dtype == np.int64 and c.dtype == np.float32
        assert np.array_equal(a, g[f"{dt}_k{k}_nbr"])
        assert np.array_equal(b, g[f"{dt}_k{k}_eid"])
        assert np.array_equal(c, g[f"{dt}_k{k}_ts"])


def test_sampler_multi_hop_golden():
    g = load("sampler.npz")
    s = _adv_sampler()
    nodes, times = cases.adversarial_queries()
    nl, el, tl = s.get_multi_hop_neighbors(2, nodes[:300], times[:300], 3)
    for h in range(2):
        assert np.array_equal(nl[h], g[f"hop{h}_nbr"])
        assert np.array_equal(el[h], g[f"hop{h}_eid"])
        assert np.array_equal(tl[h], g[f"hop{h}_ts"])


def test_sampler_from_adj_list_and_events_agree():
    src, dst, eid, ts, n = cases.adversarial_events()
    adj = [[] for _ in range(n + 1)]
    for s_, d_, e_, t_ in zip(src, dst, eid, ts):
        adj[s_].append((d_, e_, t_))
        adj[d_].append((s_, e_, t_))
    a = osamp.OracleSampler.from_adj_list(adj)
    b = _adv_sampler()
    for x in ("nbr", "eid", "ts", "indptr"):
        assert np.array_equal(getattr(a, x), getattr(b, x))
    g = load("sampler.npz")
    nodes, times = cases.adversarial_queries()
    ok = nodes <= max(src.max(), dst.max())
    r = b.get_historical_neighbors(nodes[ok], times[ok], 5)
    assert np.array_equal(r[0], g["gns_k5_nbr"]) and np.array_equal(r[1], g["gns_k5_eid"])
    assert np.array_equal(r[2], g["gns_k5_ts"])


def test_sampler_first_hop_and_find_before():
    s = _adv_sampler()
    nodes, times = cases.adversarial_queries()
    nl, el, tl = s.get_all_first_hop_neighbors(nodes[:60], times[:60])
    for i in range(60):
        a, b, c, _ = s.find_neighbors_before(nodes[i], times[i])
        assert np.array_equal(a, nl[i]) and np.array_equal(b, el[i]) and np.array_equal(c, tl[i])
        assert (c < times[i]).all()


# ---------------------------------------------------------------- TGAT
TGAT_CASES = [("L1_k20", 1, 20, 2, 0.0, False), ("L2_k5", 2, 5, 2, 0.0, False), ("L2_k20_bias", 2, 20, 2, 0.5, False),
              ("L2_k7_zeros", 2, 7, 2, 0.3, True), ("L3_k3", 3, 3, 2, 0.2, False)]


@pytest.mark.parametrize("name,L,k,heads,bias,zeros", TGAT_CASES)
def test_tgat_golden(name, L, k, heads, bias, zeros):
    torch.set_num_threads(4)
    g = load("tgat.npz")
    src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
    p = otgat.default_params(172, 172, 100, L, heads, seed=3, time_bias_scale=bias)
    chk = cases.checksum(src, dst, ts, nf, ef, *[v.numpy() for kk, v in p.items() if not kk.startswith("_")])
    assert chk == g[name + "_checksum"], "inputs/weights did not regenerate identically"
    s = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    sel = g[name + "_sel"]
    a, b = otgat.embed_src_dst(p, torch.from_numpy(nf), torch.from_numpy(ef), s, src[sel], dst[sel], ts[sel], L, k)
    # same ops in the same order on the same CPU: expect (near) bit equality
    np.testing.assert_allclose(a.numpy(), g[name + "_src"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(b.numpy(), g[name + "_dst"], rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------- TGN
TGN_CASES = [("L1_k5", 1, 5, 25, 12, 0.3), ("L2_k4", 2, 4, 20, 8, 0.0)]


@pytest.mark.parametrize("name,L,k,bs,nb,bias", TGN_CASES)
def test_tgn_golden(name, L, k, bs, nb, bias):
    torch.set_num_threads(4)
    g = load("tgn.npz")
    src, dst, eid, ts, nf, ef = cases.small_stream(num_nodes=30, num_edges=400, seed=11, t_max=2.0e6)
    p = otgn.default_params(172, 172, 100, L, 2, seed=5, time_bias_scale=bias)
    chk = cases.checksum(src, dst, ts, nf, ef, *[v.numpy() for kk, v in p.items() if not kk.startswith("_")])
    assert chk == g[name + "_checksum"]
    s = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    m = otgn.OracleTGN(p, torch.from_numpy(nf), torch.from_numpy(ef), s, L, k)
    for b in range(nb):
        lo, hi = b * bs, (b + 1) * bs
        a, c = m.step(src[lo:hi], dst[lo:hi], ts[lo:hi], eid[lo:hi], True)
        np.testing.assert_allclose(torch.cat([a, c]).numpy(), g[name + "_emb"][b], rtol=1e-5, atol=2e-6)
    a, c = m.step(src[hi:hi + bs], dst[hi:hi + bs][::-1].copy(), ts[hi:hi + bs], eid[hi:hi + bs], False)
    np.testing.assert_allclose(torch.cat([a, c]).numpy(), g[name + "_neg"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(m.mem.numpy(), g[name + "_mem"], rtol=1e-5, atol=2e-6)
    assert np.array_equal(m.last_upd.numpy(), g[name + "_lastupd"])
    pend = sorted(v for v, l in m.msgs.items() if len(l) > 0)
    assert np.array_equal(np.array(pend), g[name + "_pend_ids"])
    np.testing.assert_allclose(np.stack([m.msgs[v][-1][0].numpy() for v in pend]), g[name + "_pend_msg"],
                               rtol=1e-5, atol=2e-6)
    assert np.array_equal(np.array([m.msgs[v][-1][1] for v in pend]), g[name + "_pend_ts"])


def test_gru_cell_matches_torch():
    torch.manual_seed(0)
    cell = torch.nn.GRUCell(616, 172)
    x, h = torch.randn(33, 616), torch.randn(33, 172)
    with torch.no_grad():
        want = cell(x, h)
        got = otgn.gru_cell(x, h, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------- pseudo labels
@pytest.mark.parametrize("C", [2, 5])
def test_pseudo_golden(C):
    g = load("pseudo.npz")
    rs = np.random.RandomState(21)
    emb = None
    for c in (2, 5):  # regenerate the stream of the generator script
        e = torch.from_numpy(rs.standard_normal((700, 172)).astype(np.float32) * 2.0)
        if c == C:
            emb = e
            break
        rs.standard_normal((2, 700, c))  # the two logit perturbations
        rs.randint(0, c, 700), rs.uniform(0, 1000, 700), rs.rand(700)
    p = opseudo.default_decoder_params(172, C, seed=C)
    with torch.no_grad():
        logits = opseudo.decoder(p, emb)
    np.testing.assert_allclose(logits.numpy(), g[f"C{C}_logits"], rtol=1e-6, atol=1e-6)
    lab, probs = opseudo.emit(p, emb, batch_size=200)
    np.testing.assert_allclose(probs.numpy(), g[f"C{C}_probs"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(lab.numpy(), g[f"C{C}_labels"])
    store = [torch.from_numpy(x) for x in g[f"C{C}_store"]]
    glab = torch.from_numpy(g[f"C{C}_labels"])
    for thr in (0.3, 0.6, 0.9):
        ps = glab.to(torch.float32).reshape(1, -1).clone()
        assert np.array_equal(opseudo.entropy_filter(ps, store, thr).numpy(), g[f"C{C}_est_{thr}"])
        ps = glab.to(torch.float32).reshape(1, -1).clone()
        assert np.array_equal(opseudo.prob_filter(ps, store, thr).numpy(), g[f"C{C}_cst_{thr}"])
    ps = glab.to(torch.float32).reshape(2, 350).clone()
    store2 = [s.reshape(2, 350, C) for s in store]
    assert np.array_equal(opseudo.entropy_filter(ps, store2, 0.6).numpy(), g[f"C{C}_est2_0.6"])
    for ut in (0, 1):
        ps = glab.to(torch.float32).reshape(1, -1).clone()
        r = opseudo.update_pseudo_labels(g[f"C{C}_true"], g[f"C{C}_lt"], g[f"C{C}_it"], 400, ps, store, "ps", ut, 0.6, "entropy")
        assert np.array_equal(r.numpy(), g[f"C{C}_upd_ut{ut}"])


# ---------------------------------------------------------------- GraphMixer
MIXER_CASES = [("L2_k20_g2000", 2, 20, 2000, 0.0, False), ("L2_k5_g7", 2, 5, 7, 0.3, False),
               ("L1_k10_g50_zeros", 1, 10, 50, 0.2, True)]


@pytest.mark.parametrize("name,L,k,gap,bias,zeros", MIXER_CASES)
def test_graphmixer_golden(name, L, k, gap, bias, zeros):
    """GraphMixer restatement (oracle/graphmixer.py) against the live reference's outputs."""
    g = load("graphmixer.npz")
    src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
    p = omix.default_params(172, 100, k, L, seed=5, time_bias_scale=bias)
    assert cases.checksum(src, dst, ts, nf, *[v.numpy() for v in p.values()]) == g[name + "_checksum"]
    s = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    sel = g[name + "_sel"]
    with torch.no_grad():
        a = omix.embed(p, torch.from_numpy(nf), s, src[sel], ts[sel], L, k, gap).numpy()
        b = omix.embed(p, torch.from_numpy(nf), s, dst[sel], ts[sel], L, k, gap).numpy()
    for got, want in ((a, g[name + "_src"]), (b, g[name + "_dst"])):
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


# ---------------------------------------------------------------- TCL
TCL_CASES = [("L2_k20", 2, 20, 0.0, False), ("L1_k6_bias", 1, 6, 0.3, False), ("L2_k4_zeros", 2, 4, 0.2, True)]


@pytest.mark.parametrize("name,L,k,bias,zeros", TCL_CASES)
def test_tcl_golden(name, L, k, bias, zeros):
    """TCL restatement (oracle/tcl.py) against the live reference's outputs."""
    g = load("tcl.npz")
    src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
    p = otcl.default_params(172, 172, 100, L, k + 1, seed=6, time_bias_scale=bias)
    assert cases.checksum(src, dst, ts, nf, ef, *[v.numpy() for v in p.values()]) == g[name + "_checksum"]
    s = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    sel = g[name + "_sel"]
    with torch.no_grad():
        a, b = otcl.embed_src_dst(p, torch.from_numpy(nf), torch.from_numpy(ef), s, src[sel], dst[sel], ts[sel], L, 2, k)
    for got, want in ((a.numpy(), g[name + "_src"]), (b.numpy(), g[name + "_dst"])):
        assert np.abs(got - want).max() <= 5e-6 * max(1.0, np.abs(want).max())


# ---------------------------------------------------------------- live reference (build container only)
@pytest.mark.skipif(not ref_shim.available(), reason="FLiD reference tree not on this machine")
def test_oracle_vs_live_reference_random_graphs():
    ref = ref_shim.load()
    torch.set_num_threads(4)
    for seed in range(3):
        src, dst, eid, ts, nf, ef = cases.small_stream(num_nodes=25 + 10 * seed, num_edges=300, seed=100 + seed,
                                                       t_max=[50.0, 3e6, 4e7][seed])
        n = nf.shape[0] - 1
        adj = [[] for _ in range(n + 1)]
        for s_, d_, e_, t_ in zip(src, dst, eid, ts):
            adj[s_].append((d_, e_, t_))
            adj[d_].append((s_, e_, t_))
        rs_ = ref.NeighborSampler(adj, "recent", seed=0)
        os_ = osamp.OracleSampler.from_events(src, dst, eid, ts, n)
        q = np.random.RandomState(seed)
        nodes = q.randint(0, n + 1, 500)
        times = q.uniform(0, ts.max() * 1.1, 500)
        for k in (2, 10):
            for tt in (times, times.astype(np.float32)):
                want, got = rs_.get_historical_neighbors(nodes, tt, k), os_.get_historical_neighbors(nodes, tt, k)
                for w_, g_ in zip(want, got):
                    assert w_.dtype == g_.dtype and np.array_equal(w_, g_)
        p = otgat.default_params(172, 172, 100, 2, 2, seed=seed, time_bias_scale=0.4)
        m = ref.TGAT(nf, ef, rs_, 100, 2, 2, 0.1, "cpu")
        m.load_state_dict({k_: v for k_, v in p.items() if not k_.startswith("_")})
        m.eval()
        sel = np.arange(200, 230)
        with torch.no_grad():
            wa, wb = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 6)
        ga, gb = otgat.embed_src_dst(p, torch.from_numpy(nf), torch.from_numpy(ef), os_, src[sel], dst[sel], ts[sel], 2, 6)
        np.testing.assert_allclose(ga.numpy(), wa.numpy(), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(gb.numpy(), wb.numpy(), rtol=1e-6, atol=1e-6)
