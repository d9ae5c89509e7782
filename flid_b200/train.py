"""Training-mode (autograd + dropout) forward of the TGAT / TGN attention layers.

The M-step batches of the reference run ``compute_src_dst_node_temporal_embeddings`` with grad
enabled and ``model.train()`` (PTCL/M_step.py:196-325, NPL/NPL.py:185-314,
PTCL/EM_warmup.py:113-238).  Here a layer is split where its cost is:

* the irregular part -- gather the k neighbour rows, time-encode, score, masked softmax, score
  dropout, weighted sum (models/modules.py:183-231) -- is one sm_100a kernel with a hand-written
  backward kernel (``flid_attn_train_fwd`` / ``flid_attn_train_bwd``, csrc/attn_train.cu), wrapped
  in ``AttnStream`` (a ``torch.autograd.Function``).  No ``[n, k, 444]`` / ``[n, k, 272]`` tensor is
  ever materialised; the backward pass re-gathers the rows instead of saving them;
* the dense algebra -- folding the query through the key projection, the value and residual
  projections, LayerNorm, MergeLayer -- is written with torch matmuls on the *unfolded* parameters,
  so autograd differentiates the folds and every ``state_dict`` parameter receives its gradient.

The neighbourhoods come from the device sampler kernel (bit-exact with the reference).  Level-
batched instead of recursive: level l holds [targets of level l+1 ; their k neighbours], exactly
the multiset of (node, time) pairs the recursion of models/TGAT.py:68-144 visits.

Dropout: the score dropout is drawn by the kernel (Philox4x32-10, seed taken from torch's CPU
generator, so ``torch.manual_seed`` makes runs repeatable); the dropout on the residual_fc output
(models/modules.py:235) is the layer's own ``nn.Dropout``.  The masks are distributed as the
reference's are, not bit-identical to torch's stream (nothing in the reference depends on that).
"""
import numpy as np
import torch

from . import _lib


def _score_seed():
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


class AttnStream(torch.autograd.Function):
    """z[n, H, kd] = sum_j dropout(softmax_j(u_h . x_j))_hj x_j with x_j = [table[hrow_j] | edge[eid_j] | te(dt_j)]."""

    @staticmethod
    def forward(ctx, u, table, time_w, time_b, hrow, nbr, eid, dt, edge_feat, p_drop, seed):
        u, table = u.contiguous(), table.contiguous()
        time_w, time_b = time_w.contiguous(), time_b.contiguous()
        hrow, nbr, eid, dt = hrow.contiguous(), nbr.contiguous(), eid.contiguous(), dt.contiguous()
        n, H, kd = u.shape
        k, dn, de, T = nbr.shape[1], table.shape[1], edge_feat.shape[1], time_w.shape[0]
        if kd != dn + de + T:
            raise ValueError(f"AttnStream: folded query width {kd} != {dn} + {de} + {T}")
        for t in (u, table, time_w, time_b, dt, edge_feat):
            if t.dtype != torch.float32 or not t.is_cuda:
                raise TypeError("AttnStream: float32 CUDA tensors required (flid_b200 has no CPU path)")
        for t in (hrow, nbr, eid):
            if t.dtype != torch.int64 or t.shape != (n, k):
                raise TypeError("AttnStream: hrow / nbr / eid must be int64 [n, k]")
        z = torch.empty_like(u)
        probs = torch.empty((n, H, k), dtype=torch.float32, device=u.device)
        with torch.cuda.device(u.device):
            _lib.check(_lib.lib().flid_attn_train_fwd(
                _lib.ptr(u), _lib.ptr(table), _lib.ptr(hrow), _lib.ptr(nbr), _lib.ptr(eid), _lib.ptr(dt),
                _lib.ptr(edge_feat), _lib.ptr(time_w), _lib.ptr(time_b), n, k, H, dn, de, T, float(p_drop), int(seed),
                _lib.ptr(z), _lib.ptr(probs), _lib.stream()))
        ctx.save_for_backward(u, table, time_w, time_b, hrow, nbr, eid, dt, edge_feat, probs)
        ctx.p_drop, ctx.seed = float(p_drop), int(seed)
        return z

    @staticmethod
    def backward(ctx, dz):
        u, table, time_w, time_b, hrow, nbr, eid, dt, edge_feat, probs = ctx.saved_tensors
        n, H, kd = u.shape
        k, dn, de, T = nbr.shape[1], table.shape[1], edge_feat.shape[1], time_w.shape[0]
        dz = dz.contiguous()
        du = torch.empty_like(u)
        dtable = torch.zeros_like(table) if ctx.needs_input_grad[1] else None
        partial = None
        with torch.cuda.device(u.device):
            if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
                blocks = int(_lib.lib().flid_attn_train_partials(n))
                partial = torch.empty((blocks, 2, T), dtype=torch.float32, device=u.device)
            _lib.check(_lib.lib().flid_attn_train_bwd(
                _lib.ptr(u), _lib.ptr(table), _lib.ptr(hrow), _lib.ptr(nbr), _lib.ptr(eid), _lib.ptr(dt),
                _lib.ptr(edge_feat), _lib.ptr(time_w), _lib.ptr(time_b), n, k, H, dn, de, T, ctx.p_drop, ctx.seed,
                _lib.ptr(probs), _lib.ptr(dz), _lib.ptr(du), _lib.ptr(dtable), _lib.ptr(partial), _lib.stream()))
        dw = db = None
        if partial is not None:
            sums = partial.sum(dim=0)
            dw = sums[0] if ctx.needs_input_grad[2] else None
            db = sums[1] if ctx.needs_input_grad[3] else None
        return du, dtable, dw, db, None, None, None, None, None, None, None


def score_keep_mask(seed, n, num_heads, k, p_drop, device):
    """The kernel's score-dropout bits as a bool tensor [n, H, k] (tests)."""
    keep = torch.empty((n, num_heads, k), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().flid_attn_train_keep_mask(int(seed), n, num_heads, k, float(p_drop), _lib.ptr(keep),
                                                        _lib.stream()))
    return keep.bool()


def attention_layer(attn, merge, time_w, time_b, h_self, merge_self, table, hrow, nbr, eid, dt, edge_feat, training,
                    seed=None):
    """One MultiHeadAttention + MergeLayer evaluation (models/modules.py:167-245, :58-69) for n targets.

    h_self [n, dn]: layer input of the targets; merge_self [n, dn]: MergeLayer's second input;
    table / hrow: where the neighbour slots' layer inputs live.  ``seed`` fixes the score dropout."""
    n, dn = h_self.shape
    H, hd = attn.num_heads, attn.head_dim
    T = time_w.shape[0]
    kd, qd = dn + edge_feat.shape[1] + T, dn + T
    p = float(attn.dropout.p) if training else 0.0
    te0 = torch.cos(time_b)                                    # cos(fma(0, w, b)), models/TGAT.py:90
    query = torch.cat([h_self, te0.expand(n, T)], dim=1)      # also the residual (modules.py:186)
    wq = attn.query_projection.weight.view(H, hd, qd)
    wk = attn.key_projection.weight.view(H, hd, kd)
    wv = attn.value_projection.weight.view(H, hd, kd)
    # score_hj = scale * (Wq_h q) . (Wk_h x_j) = (scale * Wk_h^T Wq_h q) . x_j
    fold = torch.matmul(wk.transpose(1, 2), wq) * attn.scaling_factor          # [H, kd, qd]
    u = torch.matmul(query, fold.reshape(H * kd, qd).t()).view(n, H, kd)
    if p > 0.0 and seed is None:
        seed = _score_seed()
    z = AttnStream.apply(u, table, time_w, time_b, hrow, nbr, eid, dt, edge_feat, p, seed or 0)
    # sum_j a_hj (Wv_h x_j) = Wv_h z_h
    ctx = torch.einsum('nhk,hdk->nhd', z, wv).reshape(n, H * hd)
    out = attn.residual_fc(ctx)
    if training:
        out = attn.dropout(out)
    out = attn.layer_norm(out + query)
    return merge.fc2(merge.act(merge.fc1(torch.cat([out, merge_self], dim=1))))


def sample_levels(sampler, node_ids, node_interact_times, depth, k, device):
    """Top-down sampling: levels[l] = (ids, nbr, eid, dt) for the targets evaluated at layer l."""
    ids = torch.as_tensor(np.asarray(node_ids), dtype=torch.int64, device=device)
    t_np = np.asarray(node_interact_times)
    root_f64 = t_np.dtype != np.float32
    times = torch.as_tensor(t_np, device=device).to(torch.float64 if root_f64 else torch.float32)
    levels = {}
    cur_ids, cur_t64, n_f64 = ids, times.to(torch.float64), (ids.shape[0] if root_f64 else 0)
    for l in range(depth, 0, -1):
        # query times as float64: root-chain targets keep their float64 times, neighbour targets carry the
        # sampler's float32 values, which widen exactly (the comparison the reference's searchsorted makes)
        nbr, eid, ts = sampler.get_historical_neighbors_device(cur_ids, cur_t64, k)
        dt32 = cur_t64.to(torch.float32)[:, None] - ts
        if n_f64:   # models/TGAT.py:120-125: float64 root time minus float32 neighbour time, rounded once
            dt32[:n_f64] = (cur_t64[:n_f64, None] - ts[:n_f64].to(torch.float64)).to(torch.float32)
        levels[l] = (cur_ids, nbr, eid, dt32)
        if l > 1:
            cur_ids = torch.cat([cur_ids, nbr.reshape(-1)])
            cur_t64 = torch.cat([cur_t64, ts.reshape(-1).to(torch.float64)])
    return levels


def autograd_forward(time_encoder, conv_layers, merge_layers, sampler, node_feat, edge_feat, node_ids,
                     node_interact_times, depth, num_neighbors, training, seeds=None):
    """Differentiable ``compute_node_temporal_embeddings`` (models/TGAT.py:68-144;
    models/MemoryModel.py:632-715 when ``node_feat`` is the memory-augmented layer-0 table)."""
    device = node_feat.device
    k = int(num_neighbors)
    assert k > 0, 'Number of sampled neighbors for each node should be greater than 0!'
    if k > 32:
        raise ValueError("flid_b200 training path: num_neighbors must be <= 32")
    levels = sample_levels(sampler, node_ids, node_interact_times, depth, k, device)
    w_t, b_t = time_encoder.w.weight.reshape(-1), time_encoder.w.bias
    h_prev = node_feat[levels[1][0]]
    out = h_prev
    for l in range(1, depth + 1):
        t_ids, nbr, eid, dt = levels[l]
        n = t_ids.shape[0]
        if l == 1:
            table, hrow = node_feat, nbr          # rows by neighbour id
        else:                                     # rows n.. of the previous level are this level's neighbours
            table = h_prev
            hrow = n + torch.arange(n * k, dtype=torch.int64, device=device).view(n, k)
        out = attention_layer(conv_layers[l - 1], merge_layers[l - 1], w_t, b_t, h_prev[:n], node_feat[t_ids], table,
                              hrow, nbr, eid, dt, edge_feat, training, None if seeds is None else seeds[l - 1])
        h_prev = out
    return out
