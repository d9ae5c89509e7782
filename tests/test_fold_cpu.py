"""CPU check of the algebra the kernels rely on: the folded projections built by
``flid_b200.train.folded_weights`` reproduce MultiHeadAttention's scores and residual_fc(attention output)
(models/modules.py:183-235) for random inputs, and their autograd reaches every unfolded parameter."""
import torch

from flid_b200.tgat import MultiHeadAttention
from flid_b200.train import folded_weights


def test_folded_projections_match_the_unfolded_attention():
    torch.manual_seed(0)
    dn, de, T, H, n, k = 172, 172, 100, 2, 7, 5
    qd, kd = dn + T, dn + de + T
    attn = MultiHeadAttention(dn, de, T, H, dropout=0.0).double()
    q = torch.randn(n, qd, dtype=torch.float64)
    x = torch.randn(n, k, kd, dtype=torch.float64)
    mask = torch.rand(n, k) < 0.3
    mask[0] = True                                    # a target without neighbours
    hd = attn.head_dim
    # reference order (models/modules.py:188-231)
    Q = attn.query_projection(q).view(n, H, hd)
    K = attn.key_projection(x).view(n, k, H, hd)
    V = attn.value_projection(x).view(n, k, H, hd)
    s = torch.einsum('nhd,nkhd->nhk', Q, K) * attn.scaling_factor
    s = s.masked_fill(mask[:, None, :], -1e10)
    a = torch.softmax(s, dim=-1)
    ctx = torch.einsum('nhk,nkhd->nhd', a, V).reshape(n, H * hd)
    want = attn.residual_fc(ctx)
    # folded order (what the kernels evaluate)
    fold_q, fold_o = folded_weights(attn, kd, qd)
    assert fold_q.shape == (H * kd, qd) and fold_o.shape == (qd, H * kd)
    u = (q @ fold_q.t()).view(n, H, kd)
    s2 = torch.einsum('nhd,nkd->nhk', u, x).masked_fill(mask[:, None, :], -1e10)
    assert torch.allclose(s2, s, rtol=1e-10, atol=1e-10)
    z = torch.einsum('nhk,nkd->nhd', torch.softmax(s2, dim=-1), x).reshape(n, H * kd)
    got = z @ fold_o.t() + attn.residual_fc.bias
    assert torch.allclose(got, want, rtol=1e-10, atol=1e-10)
    # uniform 1/k over the padded rows of the empty target (modules.py:217-224)
    assert torch.allclose(torch.softmax(s2, dim=-1)[0], torch.full((H, k), 1.0 / k, dtype=torch.float64))
    # autograd through the folds reaches the unfolded parameters
    got.square().sum().backward()
    for name in ("query_projection", "key_projection", "value_projection", "residual_fc"):
        g = getattr(attn, name).weight.grad
        assert g is not None and float(g.abs().max()) > 0, name


def _kv_perm(c, qd, H):
    """bulk_kv.cu kv_perm: position c of a projected row -> index of the reference's projection output."""
    hd, full = qd // H, (qd // 128) * 128
    if c < full:
        f = c >> 2
        return (f % H) * hd + (f // H) * 4 + (c & 3)
    j = c - full
    return (j % H) * hd + full // H + j // H


def test_projected_bulk_path_algebra():
    """The projected ("per-entry K/V") formulation of flid_b200/csrc/bulk_kv.cu, restated in float64 with the same
    weight folds and head-interleaved row layout, reproduces MultiHeadAttention (models/modules.py:183-235):
    level >= 2 form (qs . K + ut . te, sum a V) and level-1 form (per-entry score from the folded query,
    V = Vn + Ve)."""
    torch.manual_seed(1)
    dn, de, T, H, n, k = 172, 172, 100, 2, 6, 5
    qd, kd, he = dn + T, dn + de + T, dn + de
    hd = qd // H
    attn = MultiHeadAttention(dn, de, T, H, dropout=0.0).double()
    Wq, Wk, Wv = attn.query_projection.weight, attn.key_projection.weight, attn.value_projection.weight
    Wr, br = attn.residual_fc.weight, attn.residual_fc.bias
    perm = torch.tensor([_kv_perm(c, qd, H) for c in range(qd)])
    assert sorted(perm.tolist()) == list(range(qd))
    head_of = torch.tensor([(c >> 2) % H if c < (qd // 128) * 128 else (c - (qd // 128) * 128) % H for c in range(qd)])
    assert torch.equal(head_of, perm // hd)
    h_self = torch.randn(n, dn, dtype=torch.float64)
    te0 = torch.randn(T, dtype=torch.float64)
    hn = torch.randn(n, k, dn, dtype=torch.float64)
    e = torch.randn(n, k, de, dtype=torch.float64)
    te = torch.randn(n, k, T, dtype=torch.float64)
    mask = torch.rand(n, k) < 0.3
    mask[0] = True
    # reference order
    q_in = torch.cat([h_self, te0.expand(n, T)], dim=1)
    x = torch.cat([hn, e, te], dim=2)
    Q = (q_in @ Wq.t()).view(n, H, hd)
    K = (x @ Wk.t()).view(n, k, H, hd)
    Vv = (x @ Wv.t()).view(n, k, H, hd)
    s = torch.einsum('nhd,nkhd->nhk', Q, K) * attn.scaling_factor
    s = s.masked_fill(mask[:, None, :], -1e10)
    a = torch.softmax(s, dim=-1)
    want = torch.einsum('nhk,nkhd->nhd', a, Vv).reshape(n, H * hd) @ Wr.t() + br
    # folds of kv_fold_kernel (scale without the log2(e) factor here: plain softmax below)
    sc = attn.scaling_factor
    fold_q, fold_o = folded_weights(attn, kd, qd)          # [H*kd, qd] (already scaled), [qd, H*kd]
    wqs, cqs = sc * Wq[perm, :dn], sc * (Wq[perm, dn:] @ te0)
    rows_t = torch.tensor([h * kd + he + t for h in range(H) for t in range(T)])
    wut, cut = fold_q[rows_t, :dn], fold_q[rows_t, dn:] @ te0
    wk2, wv2 = Wk[perm, :he], Wv[perm, :he]
    wvn, wve = Wv[perm, :dn], Wv[perm, dn:he]
    wo2 = torch.cat([Wr[:, perm], fold_o[:, rows_t]], dim=1)
    # level >= 2 form
    qs = h_self @ wqs.t() + cqs
    ut = (h_self @ wut.t() + cut).view(n, H, T)
    he_rows = torch.cat([hn, e], dim=2)
    Kp, Vp = he_rows @ wk2.t(), he_rows @ wv2.t()                      # [n, k, qd] interleaved
    onehot = torch.stack([(head_of == h).double() for h in range(H)])  # [H, qd]
    s2 = torch.einsum('nc,nkc,hc->nhk', qs, Kp, onehot) + torch.einsum('nht,nkt->nhk', ut, te)
    assert torch.allclose(s2.masked_fill(mask[:, None, :], -1e10), s, rtol=1e-9, atol=1e-9)
    a2 = torch.softmax(s2.masked_fill(mask[:, None, :], -1e10), dim=-1)
    own = a2[:, head_of, :]                                            # weight of a column's own head, [n, qd, k]
    A = torch.einsum('nck,nkc->nc', own, Vp)
    zt = torch.einsum('nhk,nkt->nht', a2, te).reshape(n, H * T)
    got = torch.cat([A, zt], dim=1) @ wo2.t() + br
    assert torch.allclose(got, want, rtol=1e-9, atol=1e-9)
    # level-1 form: the [h | e] score from the folded query's [h | e] columns, V = Vn + Ve
    u = (q_in @ fold_q.t()).view(n, H, kd)
    s1 = torch.einsum('nhc,nkc->nhk', u[:, :, :he], he_rows)
    s3 = s1 + torch.einsum('nht,nkt->nhk', u[:, :, he:], te)
    assert torch.allclose(s3.masked_fill(mask[:, None, :], -1e10), s, rtol=1e-9, atol=1e-9)
    assert torch.allclose(u[:, :, he:], ut, rtol=1e-10, atol=1e-10)
    V1 = hn @ wvn.t() + e @ wve.t()
    assert torch.allclose(V1, Vp, rtol=1e-10, atol=1e-10)
