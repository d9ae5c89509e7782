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
