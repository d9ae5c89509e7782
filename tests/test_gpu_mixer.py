"""GPU parity of the GraphMixer drop-in (flid_b200/graphmixer.py, csrc/mixer.cu) against the golden vectors
minted from the live reference and against the CPU oracle (oracle/graphmixer.py)."""
import os

import numpy as np
import pytest
import torch

import cases
import flid_b200
from flid_b200 import _lib
from oracle import sampler as osamp, graphmixer as omix, tcl as otcl

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
MIXER_CASES = [("L2_k20_g2000", 2, 20, 2000, 0.0, False), ("L2_k5_g7", 2, 5, 7, 0.3, False),
               ("L1_k10_g50_zeros", 1, 10, 50, 0.2, True)]


def close(got, want, what):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= 1e-4 * scale, (what, err, scale)


def build(nf, ef, src, dst, eid, ts, k, L, p, dropout=0.1):
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=DEV, _events=(src, dst, eid, ts, nf.shape[0] - 1))
    m = flid_b200.GraphMixer(nf, ef, s, 100, k, L, dropout=dropout, device=DEV).to(DEV)
    m.load_state_dict(p)
    return m, s


@pytest.mark.parametrize("name,L,k,gap,bias,zeros", MIXER_CASES)
def test_graphmixer_golden_and_oracle(name, L, k, gap, bias, zeros):
    g = np.load(os.path.join(G, "graphmixer.npz"))
    src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
    p = omix.default_params(172, 100, k, L, seed=5, time_bias_scale=bias)
    m, _ = build(nf, ef, src, dst, eid, ts, k, L, p)
    m.eval()
    sel = g[name + "_sel"]
    with torch.no_grad():
        a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], num_neighbors=k, time_gap=gap)
    close(a.cpu().numpy(), g[name + "_src"], name + " src vs reference")
    close(b.cpu().numpy(), g[name + "_dst"], name + " dst vs reference")
    # all events, float32 times, chunked: against the oracle
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    t32 = ts.astype(np.float32)
    m.chunk_queries = 300
    with torch.no_grad():
        got = m.compute_node_temporal_embeddings(src, t32, k, gap)
        want = omix.embed(p, torch.from_numpy(nf), o, src, t32, L, k, gap)
    close(got.cpu().numpy(), want.numpy(), name + " float32 times vs oracle")


@pytest.mark.parametrize("gap", [1, 5, 2000])
def test_neighbor_mean_kernel_vs_numpy(gap):
    src, dst, eid, ts, n = cases.adversarial_events()
    rs = np.random.RandomState(3)
    nf = rs.standard_normal((n + 1, 64)).astype(np.float32)
    nf[0] = rs.standard_normal(64).astype(np.float32) * 0.1          # a non-zero padding row must be followed too
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=DEV, _events=(src, dst, eid, ts, n))
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, n)
    nodes, times = cases.adversarial_queries()
    d_nf = torch.from_numpy(nf).to(DEV)
    d_ids = torch.from_numpy(nodes).to(DEV)
    d_t = torch.from_numpy(times).to(DEV)
    out = torch.empty((len(nodes), 64), dtype=torch.float32, device=DEV)
    _lib.check(_lib.lib().flid_neighbor_mean(s.handle, _lib.ptr(d_nf), 64, _lib.ptr(d_ids), _lib.ptr(d_t), 0, len(nodes),
                                             gap, 1, _lib.ptr(out), _lib.stream()))
    nbr, _, _ = o.get_historical_neighbors(nodes, times, gap)
    feats = torch.from_numpy(nf)[torch.from_numpy(nbr)]
    mask = torch.from_numpy((nbr > 0).astype(np.float32))
    mask[mask == 0] = -1e10
    want = torch.mean(feats * torch.softmax(mask, dim=1).unsqueeze(-1), dim=1) + torch.from_numpy(nf)[torch.from_numpy(nodes)]
    close(out.cpu().numpy(), want.numpy(), f"neighbor mean gap={gap}")
    assert (nbr > 0).sum(1).min() == 0 and (nbr > 0).sum(1).max() >= min(gap, 5)      # empty and full windows covered


def test_graphmixer_training_gradients_and_api():
    src, dst, eid, ts, nf, ef = cases.small_stream()
    L, k, gap = 2, 6, 30
    p = omix.default_params(172, 100, k, L, seed=9, time_bias_scale=0.1)
    m, s = build(nf, ef, src, dst, eid, ts, k, L, p, dropout=0.0)
    m.train()
    sel = np.arange(100, 160)
    a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k, gap)
    assert a.requires_grad
    w = torch.randn(a.shape, generator=torch.Generator().manual_seed(1))
    ((a * w.to(DEV)).sum() + (b * w.to(DEV)).sum()).backward()
    po = {kk: v.clone().requires_grad_(not kk.startswith("time_encoder")) for kk, v in p.items()}
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    oa = omix.embed(po, torch.from_numpy(nf), o, src[sel], ts[sel], L, k, gap)
    ob = omix.embed(po, torch.from_numpy(nf), o, dst[sel], ts[sel], L, k, gap)
    ((oa * w).sum() + (ob * w).sum()).backward()
    close(a.detach().cpu().numpy(), oa.detach().numpy(), "train-mode values")
    for name, prm in m.named_parameters():
        if name.startswith("time_encoder"):
            assert not prm.requires_grad          # frozen in GraphMixer (models/GraphMixer.py:43-45)
            continue
        rel = float((prm.grad.cpu() - po[name].grad).norm()) / max(float(po[name].grad.norm()), 1e-6)
        assert rel <= 2e-3, (name, rel)
    assert set(m.state_dict().keys()) == set(p.keys())
    with pytest.raises(AssertionError):
        m.compute_node_temporal_embeddings(src[:4], ts[:4], 0)
    with pytest.raises(IndexError):
        m.compute_node_temporal_embeddings(np.array([10 ** 6]), np.array([5.0]), k)
    m.set_neighbor_sampler(s)


# ===================================================================== TCL
TCL_CASES = [("L2_k20", 2, 20, 0.0, False), ("L1_k6_bias", 1, 6, 0.3, False), ("L2_k4_zeros", 2, 4, 0.2, True)]


@pytest.mark.parametrize("name,L,k,bias,zeros", TCL_CASES)
def test_tcl_golden_and_oracle(name, L, k, bias, zeros):
    g = np.load(os.path.join(G, "tcl.npz"))
    src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
    p = otcl.default_params(172, 172, 100, L, k + 1, seed=6, time_bias_scale=bias)
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=DEV, _events=(src, dst, eid, ts, nf.shape[0] - 1))
    m = flid_b200.TCL(nf, ef, s, 100, L, 2, k + 1, 0.1, DEV).to(DEV)
    m.load_state_dict(p)
    assert set(m.state_dict().keys()) == set(p.keys())
    m.eval()
    sel = g[name + "_sel"]
    with torch.no_grad():
        a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
    close(a.cpu().numpy(), g[name + "_src"], name + " src vs reference")
    close(b.cpu().numpy(), g[name + "_dst"], name + " dst vs reference")
    # all events, float32 times, chunked: against the oracle
    o = osamp.OracleSampler.from_events(src, dst, eid, ts, nf.shape[0] - 1)
    t32 = ts.astype(np.float32)
    m.chunk_events = 128
    with torch.no_grad():
        ga, gb = m.compute_src_dst_node_temporal_embeddings(src, dst, t32, k)
        wa, wb = otcl.embed_src_dst(p, torch.from_numpy(nf), torch.from_numpy(ef), o, src, dst, t32, L, 2, k)
    close(ga.cpu().numpy(), wa.numpy(), name + " float32 times src vs oracle")
    close(gb.cpu().numpy(), wb.numpy(), name + " float32 times dst vs oracle")
    with pytest.raises(AssertionError):
        m.compute_src_dst_node_temporal_embeddings(src[:4], dst[:4], ts[:4], k + 1)     # num_depths mismatch (TCL.py:184)
    with pytest.raises(IndexError):
        m.compute_src_dst_node_temporal_embeddings(np.array([10 ** 6]), np.array([1]), np.array([5.0]), k)


def test_empty_batches():
    """Zero-length batches return empty tensors through every drop-in (the reference's loops can end on one)."""
    src, dst, eid, ts, nf, ef = cases.small_stream()
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=DEV, _events=(src, dst, eid, ts, nf.shape[0] - 1))
    e = np.zeros(0, dtype=np.int64)
    et = np.zeros(0, dtype=np.float64)
    gm = flid_b200.GraphMixer(nf, ef, s, 100, 5, 1, device=DEV).to(DEV)
    tcl = flid_b200.TCL(nf, ef, s, 100, 1, 2, 6, 0.1, DEV).to(DEV)
    tg = flid_b200.TGAT(nf, ef, s, 100, 2, 2, 0.1, DEV).to(DEV)
    for m, args in ((gm, (e, e, et, 5)), (tcl, (e, e, et, 5)), (tg, (e, e, et, 5))):
        for train in (False, True):
            m.train(train)
            with torch.set_grad_enabled(train):
                a, b = m.compute_src_dst_node_temporal_embeddings(*args)
            assert a.shape == (0, 172) and b.shape == (0, 172), type(m).__name__


# ===================================================================== csrc/dense.cu, kernel by kernel
def _rel(got, want):
    return float((got.double() - want.double()).abs().max()) / max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("rows", [1, 300, 5000])
@pytest.mark.parametrize("n_out,n_in", [(100, 100), (400, 100), (100, 400), (516, 172), (688, 172), (172, 688), (172, 272)])
def test_dense_linear_vs_torch(rows, n_out, n_in):
    """flid_dense (bias, ReLU / GELU, residual, two gathered segments) against float64 torch; both weight images
    (32-column tiles up to 2048 rows, wide tiles above) and both GEMM kernels are reached by these shapes."""
    from flid_b200 import dense
    gen = torch.Generator().manual_seed(rows * 1000 + n_out + n_in)
    w = (torch.randn(n_out, n_in, generator=gen) / n_in ** 0.5).to(DEV)
    b = torch.randn(n_out, generator=gen).to(DEV)
    x = torch.randn(rows, n_in, generator=gen).to(DEV)
    r = torch.randn(rows, n_out, generator=gen).to(DEV)
    cache = dense.DenseWeights()
    ref = x.double() @ w.double().t() + b.double()
    assert _rel(dense.linear(cache, x, w, b), ref) <= 1e-5
    assert _rel(dense.linear(cache, x, w, None), ref - b.double()) <= 1e-5
    assert _rel(dense.linear(cache, x, w, b, act=1), torch.relu(ref)) <= 1e-5
    assert _rel(dense.linear(cache, x, w, b, act=2), torch.nn.functional.gelu(ref)) <= 1e-5
    assert _rel(dense.linear(cache, x, w, b, resid=r), ref + r.double()) <= 1e-5
    assert _rel(dense.linear(cache, x, w, b, act=1, resid=r), torch.relu(ref + r.double())) <= 1e-5
    # two row-gathered segments: torch's cat([t0[i0], t1[i1]], 1)
    w0 = (n_in // 8) * 4
    t0 = torch.randn(37, w0, generator=gen).to(DEV)
    t1 = torch.randn(53, n_in - w0, generator=gen).to(DEV)
    i0 = torch.randint(0, 37, (rows,), generator=gen).to(DEV)
    i1 = torch.randint(0, 53, (rows,), generator=gen).to(DEV)
    ref2 = torch.cat([t0[i0], t1[i1]], 1).double() @ w.double().t() + b.double()
    got2 = dense.linear(cache, t0, w, b, idx=i0.to(torch.int32), x2=t1, idx2=i1.to(torch.int32))
    assert _rel(got2, ref2) <= 1e-5
    # the cached image follows an in-place parameter update
    w.mul_(0.5)
    assert _rel(dense.linear(cache, x, w, b), (ref - b.double()) * 0.5 + b.double()) <= 1e-5
    # strided rows: every 3rd row of x
    if rows >= 3:
        got3 = dense.linear(cache, x, w, b, rows=rows // 3, ldx=3 * n_in)
        assert _rel(got3, (x[::3][:rows // 3].double() @ w.double().t() + b.double())) <= 1e-5
    cache.clear()


@pytest.mark.parametrize("k,hid", [(20, 10), (5, 2), (33, 16), (64, 32)])
def test_token_mix_layernorm_mean_vs_torch(k, hid):
    from flid_b200 import dense
    from flid_b200.graphmixer import MLPMixer
    torch.manual_seed(k)
    m, c = 77, 100
    mixer = MLPMixer(k, c, hid / k + 1e-9, 4.0, 0.0).to(DEV).eval()
    with torch.no_grad():
        mixer.token_norm.weight.uniform_(0.5, 1.5), mixer.token_norm.bias.normal_(0, 0.2)
        mixer.channel_norm.weight.uniform_(0.5, 1.5), mixer.channel_norm.bias.normal_(0, 0.2)
        x = torch.randn(m, k, c, device=DEV)
        tf = mixer.token_feedforward.ffn
        assert tf[0].weight.shape == (hid, k)
        want = mixer.token_feedforward(mixer.token_norm(x.permute(0, 2, 1))).permute(0, 2, 1) + x
        got = torch.empty_like(x)
        _lib.check(_lib.lib().flid_token_mix(_lib.ptr(x), k, c, _lib.ptr(mixer.token_norm.weight), _lib.ptr(mixer.token_norm.bias),
                                             1e-5, _lib.ptr(tf[0].weight), _lib.ptr(tf[0].bias), _lib.ptr(tf[3].weight),
                                             _lib.ptr(tf[3].bias), hid, _lib.ptr(got), m, _lib.stream()))
        assert _rel(got, want) <= 1e-5
        flat = x.reshape(m * k, c)
        assert _rel(dense.layernorm(flat, mixer.channel_norm), mixer.channel_norm(flat)) <= 1e-5
        mean = torch.empty(m, c, device=DEV)
        _lib.check(_lib.lib().flid_token_mean(_lib.ptr(x), k, c, _lib.ptr(mean), c, m, _lib.stream()))
        assert _rel(mean, x.mean(1)) <= 1e-5


@pytest.mark.parametrize("s,heads,d", [(21, 2, 172), (5, 2, 172), (7, 4, 64), (33, 1, 100)])
def test_seq_attention_vs_multihead_attention(s, heads, d):
    """flid_seq_attention + flid_dense in / out projections against nn.MultiheadAttention with a key padding mask
    (self and cross attention), the way TransformerEncoder calls it (models/modules.py:287-300)."""
    from flid_b200 import dense
    torch.manual_seed(s * d)
    m = 41
    mha = torch.nn.MultiheadAttention(d, heads, dropout=0.0).to(DEV).eval()
    cache = dense.DenseWeights()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.3), mha.out_proj.bias.normal_(0, 0.3)
        xq, xkv = torch.randn(m, s, d, device=DEV), torch.randn(m, s, d, device=DEV)
        ids = torch.randint(0, 3, (m, s), device=DEV)
        ids[:, 0] = 1                                  # the node itself is never padding
        ids[3, 1:] = 0                                 # a sequence with no neighbours
        for kv in (xq, xkv):
            want = mha(xq.transpose(0, 1), kv.transpose(0, 1), kv.transpose(0, 1), key_padding_mask=ids == 0)[0].transpose(0, 1)
            w_in, b_in = mha.in_proj_weight, mha.in_proj_bias
            fq, fkv = xq.reshape(m * s, d), kv.reshape(m * s, d)
            if kv is xq:
                qkv = dense.linear(cache, fq, w_in, b_in)
                q, kk, vv = qkv, qkv[:, d:], qkv[:, 2 * d:]
            else:
                q = dense.linear(cache, fq, w_in[:d], b_in[:d])
                kvp = dense.linear(cache, fkv, w_in[d:], b_in[d:])
                kk, vv = kvp, kvp[:, d:]
            ctx = torch.zeros(m * s, d, device=DEV)
            for q_rows in (s, 1):
                _lib.check(_lib.lib().flid_seq_attention(_lib.ptr(q), q.stride(0), _lib.ptr(kk), kk.stride(0), _lib.ptr(vv),
                                                         vv.stride(0), _lib.ptr(ids), s, heads, d // heads, _lib.ptr(ctx), d,
                                                         q_rows, m, _lib.stream()))
                got = dense.linear(cache, ctx, mha.out_proj.weight, mha.out_proj.bias).reshape(m, s, d)
                assert _rel(got[:, :q_rows], want[:, :q_rows]) <= 1e-5, (q_rows, kv is xq)
    cache.clear()


def test_eval_path_equals_torch_modules():
    """The forward-only kernels (model.eval() under no_grad) and the torch modules (grad enabled) agree on both models."""
    src, dst, eid, ts, nf, ef = cases.small_stream()
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=DEV, _events=(src, dst, eid, ts, nf.shape[0] - 1))
    torch.manual_seed(3)
    gm = flid_b200.GraphMixer(nf, ef, s, 100, 20, 2, device=DEV).to(DEV).eval()
    tcl = flid_b200.TCL(nf, ef, s, 100, 2, 2, 21, 0.1, DEV).to(DEV).eval()
    sel = np.arange(50, 450)
    for m in (gm, tcl):
        with torch.no_grad():
            a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 20)
        with torch.enable_grad():
            ra, rb = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], 20)
        assert ra.requires_grad and not a.requires_grad
        close(a.cpu().numpy(), ra.detach().cpu().numpy(), type(m).__name__ + " src")
        close(b.cpu().numpy(), rb.detach().cpu().numpy(), type(m).__name__ + " dst")
