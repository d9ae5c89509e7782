"""GPU tests of the training-mode attention stream (csrc/attn_train.cu, flid_b200/train.py).

The kernels' forward and hand-written backward are compared with a plain torch restatement of
models/modules.py:183-231 (float64 for the kernel-level test, float32 in the reference's literal
op order for the model-level test), with the kernel's own dropout bits fed to the restatement.
Parity with the CPU oracle's autograd at dropout 0 is covered in test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import cases
import flid_b200
from flid_b200 import train
from oracle import tgat as otgat

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def stream_reference(u, table, time_w, time_b, hrow, nbr, eid, dt, edge_feat, keep, p):
    """z of AttnStream with ordinary torch ops in float64 (models/modules.py:197-231 without the projections).
    The time-encoder argument is the reference's float32 fma(dt, w, b) (time_w / time_b are float32 leaves);
    everything after it is float64, so the comparison isolates the kernel's own arithmetic."""
    arg = torch.addcmul(time_b, dt.unsqueeze(-1), time_w)
    x = torch.cat([table[hrow], edge_feat[eid], torch.cos(arg.double())], dim=2)   # [n,k,kd]
    s = torch.einsum('nhd,nkd->nhk', u, x)
    s = s.masked_fill((nbr == 0)[:, None, :], -1e10)
    a = torch.softmax(s, dim=-1)
    a = a * keep.to(a.dtype) / (1.0 - p)
    return torch.einsum('nhk,nkd->nhd', a, x)


@pytest.mark.parametrize("table_grad", [True, False])      # False: the one-pass backward kernel (no row gradients)
@pytest.mark.parametrize("H,k,dn,de,T,p", [(2, 20, 172, 172, 100, 0.0), (2, 20, 172, 172, 100, 0.3),
                                           (1, 5, 64, 32, 20, 0.25), (4, 32, 172, 172, 100, 0.1),
                                           (2, 7, 400, 300, 128, 0.5)])
def test_attn_stream_forward_backward_vs_torch_float64(H, k, dn, de, T, p, table_grad):
    g = torch.Generator().manual_seed(H * 100 + k)
    n, R, E = 301, 40, 500           # few table rows: many duplicate hrow entries exercise the atomic accumulation
    kd = dn + de + T
    table = torch.randn(R, dn, generator=g)
    table[0] = 0
    edge = torch.randn(E, de, generator=g)
    edge[0] = 0
    nbr = torch.randint(1, R, (n, k), generator=g)
    eid = torch.randint(1, E, (n, k), generator=g)
    # left-padded rows as the sampler produces them; rows 0..9 have no neighbour at all
    valid = torch.randint(0, k + 1, (n,), generator=g)
    valid[:10] = 0
    valid[10:20] = k
    pad = torch.arange(k)[None, :] < (k - valid)[:, None]
    nbr[pad], eid[pad] = 0, 0
    dt = torch.rand(n, k, generator=g) * 1e5
    dt[pad] = (torch.rand(n, k, generator=g) * 2e6)[pad]      # padded slots: t - 0
    time_w = torch.from_numpy(1 / 10 ** np.linspace(0, 9, T, dtype=np.float32)) * (1 + 0.1 * torch.randn(T, generator=g))
    time_b = 0.3 * torch.randn(T, generator=g)
    u = torch.randn(n, H, kd, generator=g) * 0.08
    dz = torch.randn(n, H, kd, generator=g)
    hrow = nbr.clone()

    dev = lambda t: t.to(DEV)
    leaf = lambda t: t.to(DEV).requires_grad_(True)
    u1, tab1, w1, b1 = leaf(u), (leaf(table) if table_grad else dev(table)), leaf(time_w), leaf(time_b)
    seed = 1234567 + k
    z = train.AttnStream.apply(u1, tab1, w1, b1, dev(hrow), dev(nbr), dev(eid), dev(dt), dev(edge), p, seed)
    z.backward(dev(dz))
    keep = train.score_keep_mask(seed, n, H, k, p, DEV)
    if p == 0.0:
        assert bool(keep.all())
    else:
        frac = float(keep.float().mean())
        assert abs(frac - (1 - p)) < 0.03, frac
        assert not torch.equal(keep, train.score_keep_mask(seed + 1, n, H, k, p, DEV))

    d64 = lambda t: t.to(DEV, torch.float64).requires_grad_(True)
    u2, tab2, w2, b2 = d64(u), d64(table), leaf(time_w), leaf(time_b)
    z2 = stream_reference(u2, tab2, w2, b2, dev(hrow), dev(nbr), dev(eid), dev(dt), dev(edge).double(), keep, p)
    z2.backward(dev(dz).double())

    def close(got, want, what, tol):
        got, want = got.detach(), want.detach().double()
        scale = max(1.0, float(want.abs().max()))
        err = float((got.double() - want).abs().max())
        assert err <= tol * scale, (what, err, scale)

    close(z, z2, "z", 2e-5)
    close(u1.grad, u2.grad, "du", 5e-5)
    if table_grad:
        close(tab1.grad, tab2.grad, "dtable", 5e-5)
    # time-encoder gradients: sums over n*k slots of O(dt) terms; compare relative to the gradient's own scale
    for got, want, what in ((w1.grad, w2.grad, "dw"), (b1.grad, b2.grad, "db")):
        rel = float((got.double() - want.double()).norm() / want.double().norm())
        assert rel <= 2e-4, (what, rel)
    # rows of all-masked targets take the uniform 1/k over their padded slots (modules.py:217-224)
    assert torch.isfinite(z).all() and float(z.detach()[:10].abs().max()) > 0


def reference_layers(time_encoder, conv_layers, merge_layers, node_feat, edge_feat, levels, depth, k, keeps, p,
                     out_keeps=None):
    """The reference's literal op order (models/modules.py:167-245, models/TGAT.py:68-144) in torch, level-batched,
    with explicit dropout bits (scores; residual_fc output when out_keeps is given, else the layer's nn.Dropout)."""
    w_t, b_t = time_encoder.w.weight.reshape(-1), time_encoder.w.bias
    encode = lambda dt: torch.cos(torch.addcmul(b_t, dt.unsqueeze(-1), w_t))
    h_prev = node_feat[levels[1][0]]
    for l in range(1, depth + 1):
        t_ids, nbr, eid, dt = levels[l]
        n = t_ids.shape[0]
        attn, merge = conv_layers[l - 1], merge_layers[l - 1]
        h_nbr = node_feat[nbr] if l == 1 else h_prev[n:].reshape(n, k, -1)
        te0 = encode(torch.zeros((n, 1), dtype=torch.float32, device=node_feat.device))
        query = residual = torch.cat([h_prev[:n].unsqueeze(1), te0], dim=2)
        kv = torch.cat([h_nbr, edge_feat[eid], encode(dt)], dim=2)
        H, hd = attn.num_heads, attn.head_dim
        q = attn.query_projection(query).reshape(n, 1, H, hd).permute(0, 2, 1, 3)
        kk = attn.key_projection(kv).reshape(n, k, H, hd).permute(0, 2, 1, 3)
        vv = attn.value_projection(kv).reshape(n, k, H, hd).permute(0, 2, 1, 3)
        scores = torch.einsum('bhld,bhnd->bhln', q, kk) * attn.scaling_factor
        scores = scores.masked_fill((nbr == 0)[:, None, None, :], -1e10)
        scores = torch.softmax(scores, dim=-1) * keeps[l - 1][:, :, None, :].float() / (1.0 - p)
        ctx = torch.einsum('bhln,bhnd->bhld', scores, vv).permute(0, 2, 1, 3).flatten(start_dim=2).squeeze(1)
        res = attn.residual_fc(ctx)
        res = attn.dropout(res) if out_keeps is None else res * out_keeps[l - 1].float() / (1.0 - p)
        o = attn.layer_norm(res + residual.squeeze(1))
        h_prev = merge.fc2(merge.act(merge.fc1(torch.cat([o, node_feat[t_ids]], dim=1))))
    return h_prev


@pytest.mark.parametrize("L,k,p", [(2, 5, 0.3), (1, 20, 0.1), (2, 20, 0.0)])
def test_training_forward_backward_with_dropout_vs_reference_op_order(L, k, p):
    src, dst, eid, ts, nf, ef = cases.small_stream()
    prm = otgat.default_params(172, 172, 100, L, 2, seed=11, time_bias_scale=0.3)
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=DEV, _events=(src, dst, eid, ts, nf.shape[0] - 1))

    def model():
        m = flid_b200.TGAT(nf, ef, s, 100, L, 2, p, DEV).to(DEV)
        m.load_state_dict({kk: v for kk, v in prm.items() if not kk.startswith("_")})
        m.train()
        return m

    ma, mb = model(), model()
    sel = np.arange(400, 460)
    nodes, times = np.concatenate([src[sel], dst[sel]]), np.concatenate([ts[sel], ts[sel]])
    seeds = [77, 78]
    a = train.autograd_forward(ma.time_encoder, ma.temporal_conv_layers, ma.merge_layers, s, ma.node_raw_features,
                               ma.edge_raw_features, nodes, times, L, k, True, seeds=seeds)
    levels = train.sample_levels(s, nodes, times, L, k, torch.device(DEV))
    keeps = [train.score_keep_mask(seeds[l - 1], levels[l][0].shape[0], 2, k, p, DEV) for l in range(1, L + 1)]
    out_keeps = [train.output_keep_mask(seeds[l - 1], levels[l][0].shape[0], 272, p, DEV) for l in range(1, L + 1)]
    if p > 0:
        assert abs(float(out_keeps[0].float().mean()) - (1 - p)) < 0.02
    b = reference_layers(mb.time_encoder, mb.temporal_conv_layers, mb.merge_layers, mb.node_raw_features,
                         mb.edge_raw_features, levels, L, k, keeps, p, out_keeps)
    w = torch.randn(a.shape, generator=torch.Generator().manual_seed(2)).to(DEV)
    (a * w).sum().backward()
    (b * w).sum().backward()
    scale = max(1.0, float(b.detach().abs().max()))
    assert float((a - b).detach().abs().max()) <= 1e-4 * scale
    for (name, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert pa.grad is not None and pb.grad is not None, name
        rel = float((pa.grad - pb.grad).norm()) / max(float(pb.grad.norm()), 1e-6)
        assert rel <= 2e-3, (name, rel)
    if p > 0:   # the model-level call draws fresh score seeds from torch's generator: repeatable under manual_seed
        torch.manual_seed(9)
        x1, _ = ma.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
        torch.manual_seed(9)
        x2, _ = ma.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
        x3, _ = ma.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
        assert torch.equal(x1, x2) and not torch.equal(x1, x3)


def test_attn_stream_argument_checks():
    z = torch.zeros
    u = z(4, 2, 444, device=DEV)
    tab, edge = z(10, 172, device=DEV), z(10, 172, device=DEV)
    w, b = z(100, device=DEV), z(100, device=DEV)
    idx = z(4, 33, dtype=torch.int64, device=DEV)
    with pytest.raises(ValueError):
        train.AttnStream.apply(u, tab, w, b, idx, idx, idx, z(4, 33, device=DEV), edge, 0.0, 0)
    idx = z(4, 5, dtype=torch.int64, device=DEV)
    with pytest.raises(ValueError):
        train.AttnStream.apply(u, tab, w, b, idx, idx, idx, z(4, 5, device=DEV), edge, 1.0, 0)
    with pytest.raises(TypeError):
        train.AttnStream.apply(u.double(), tab, w, b, idx, idx, idx, z(4, 5, device=DEV), edge, 0.0, 0)


def test_m_step_then_e_step_round_trip():
    """A miniature PTCL round the way the reference drives it (PTCL/M_step.py:196-325 then PTCL/E_step.py:305-352):
    a few M-step epochs of nn.Sequential(backbone, decoder) on pseudo labels with Adam -- training-mode kernels,
    dropout on -- must reduce the loss; the E-step pass that follows must see the new weights (memo rebuilt) and
    agree with the per-batch eval calls."""
    from flid_b200 import passes
    src, dst, eid, ts, nf, ef = cases.small_stream()
    rs = np.random.RandomState(0)
    labels = (nf[src, 0] + 0.5 * ef[eid, 1] > 0).astype(np.int64)          # learnable from node + edge features
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=DEV, _events=(src, dst, eid, ts, nf.shape[0] - 1))
    torch.manual_seed(0)
    backbone = flid_b200.TGAT(nf, ef, s, 100, 2, 2, 0.1, DEV).to(DEV)
    decoder = flid_b200.MLPClassifier(172, 0.1, 2).to(DEV)
    model = torch.nn.Sequential(backbone, decoder)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    loss_fn = torch.nn.CrossEntropyLoss()
    bs, losses = 50, []
    for epoch in range(4):
        model.train()
        tot = 0.0
        for lo in range(100, 500, bs):
            sl = slice(lo, lo + bs)
            emb, _ = model[0].compute_src_dst_node_temporal_embeddings(src[sl], dst[sl], ts[sl], 10)
            loss = loss_fn(model[1](emb), torch.from_numpy(labels[sl]).to(DEV))
            opt.zero_grad()
            loss.backward()
            opt.step()
            tot += float(loss.detach())
        losses.append(tot)
    assert all(np.isfinite(losses)) and losses[-1] < 0.8 * losses[0], losses
    model.eval()
    pseudo, probs, emb = passes.e_step_pass(model[0], model[1], src, dst, ts, 10, [], "entropy", 0.9, return_embeddings=True)
    with torch.no_grad():
        a, _ = model[0].compute_src_dst_node_temporal_embeddings(src[200:260], dst[200:260], ts[200:260], 10)
    scale = max(1.0, float(a.abs().max()))
    assert float((emb[0][200:260] - a).abs().max()) <= 1e-5 * scale
    assert pseudo.shape == (1, len(src)) and probs.shape == (len(src), 2)
    acc = float((probs.argmax(1).cpu().numpy()[100:500] == labels[100:500]).mean())
    assert acc > 0.6, acc
