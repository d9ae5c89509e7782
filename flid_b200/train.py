"""Training-mode (autograd + dropout) forward of the TGAT / TGN attention layers.

The M-step batches of the reference run ``compute_src_dst_node_temporal_embeddings`` with grad
enabled and ``model.train()`` (PTCL/M_step.py:196-325, NPL/NPL.py:185-314,
PTCL/EM_warmup.py:113-238).  Here a layer is split where its cost is:

* one attention + MergeLayer evaluation is ONE autograd node, ``TrainLayer``: its forward and its
  backward are each one C-ABI call (``flid_train_layer_fwd`` / ``_bwd``, csrc/train_layer.cu) that
  launches the tcgen05 GEMMs, the attention-stream kernels (gather, time encoding, masked softmax,
  score dropout, weighted sum and their hand-written backward, csrc/attn_train.cu), LayerNorm,
  dropout and the weight-gradient reductions back to back.  No ``[n, k, 444]`` / ``[n, k, 272]``
  tensor is ever materialised; the backward pass re-gathers the rows instead of saving them;
* the projections enter folded (``folded_weights``: a few tiny differentiable torch matmuls on the
  *unfolded* parameters), so autograd carries the folded gradients back to every ``state_dict``
  parameter;
* ``AttnStream`` exposes the attention stream alone as an autograd Function (custom layers, tests).

The neighbourhoods come from the device sampler kernel (bit-exact with the reference).  Level-
batched instead of recursive: level l holds [targets of level l+1 ; their k neighbours], exactly
the multiset of (node, time) pairs the recursion of models/TGAT.py:68-144 visits.

Dropout (scores, models/modules.py:224, and residual_fc output, :235) is drawn by the kernels
(Philox4x32-10; one seed per layer call taken from torch's CPU generator, so ``torch.manual_seed``
makes runs repeatable) and regenerated in the backward pass.  The masks are distributed as the
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


_LAYER_TENSORS = ("fold_q", "fold_o", "res_b", "ln_w", "ln_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "time_w", "time_b")


class TrainLayer(torch.autograd.Function):
    """One attention + merge layer through ``flid_train_layer_fwd`` / ``flid_train_layer_bwd`` (csrc/train_layer.cu).

    Differentiable inputs: q [n, qd], merge_self [n, dn], table [R, dn] and the eleven weight tensors of
    ``_LAYER_TENSORS`` (folded projections included); index tensors and edge features carry no gradient."""

    @staticmethod
    def forward(ctx, q, merge_self, table, hrow, nbr, eid, dt, edge_feat, p_drop, seed, num_heads, *weights):
        # hrow: int64 [n, k] rows of `table`, or a Python int r0: slot (i, j) reads table row r0 + i * k + j
        q, merge_self, table = q.contiguous(), merge_self.contiguous(), table.contiguous()
        weights = tuple(w.contiguous() for w in weights)
        nbr, eid, dt = nbr.contiguous(), eid.contiguous(), dt.contiguous()
        hrow_t, hrow_off = (None, int(hrow)) if isinstance(hrow, int) else (hrow.contiguous(), 0)
        n, qd = q.shape
        k, dn, de = nbr.shape[1], table.shape[1], edge_feat.shape[1]
        T, H = qd - dn, int(num_heads)
        zw = H * (dn + de + T)
        for t in (q, merge_self, table, dt, edge_feat) + weights:
            if t.dtype != torch.float32 or not t.is_cuda:
                raise TypeError("TrainLayer: float32 CUDA tensors required (flid_b200 has no CPU path)")
        if weights[0].shape != (zw, qd) or weights[1].shape != (qd, zw) or merge_self.shape != (n, dn):
            raise ValueError("TrainLayer: inconsistent shapes")
        dev = q.device
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        saved = (new(n, zw), new(n, H, k), new(n, zw), new(n, qd), new(n, qd), new(n, dn))   # u probs z y ln hid
        out, pre = new(n, dn), new(n, qd)
        wst = _lib.TrainWeights(*[w.data_ptr() for w in weights])
        sst = _lib.TrainSaved(*[t.data_ptr() for t in saved])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().flid_train_layer_fwd(
                wst, _lib.ptr(q), _lib.ptr(merge_self), _lib.ptr(table), _lib.ptr(hrow_t), hrow_off, _lib.ptr(nbr),
                _lib.ptr(eid), _lib.ptr(dt), _lib.ptr(edge_feat), n, k, H, dn, de, T, float(p_drop), int(seed), sst,
                _lib.ptr(pre), _lib.ptr(out), _lib.stream()))
        ctx.save_for_backward(q, merge_self, table, nbr, eid, dt, edge_feat, *weights, *saved)
        ctx.hrow = hrow_t      # an index tensor without autograd history
        ctx.meta = (float(p_drop), int(seed), H, len(weights), hrow_off)
        return out

    @staticmethod
    def backward(ctx, d_out):
        p_drop, seed, H, nw, hrow_off = ctx.meta
        q, merge_self, table, nbr, eid, dt, edge_feat = ctx.saved_tensors[:7]
        weights, saved = ctx.saved_tensors[7:7 + nw], ctx.saved_tensors[7 + nw:]
        hrow = ctx.hrow
        n, qd = q.shape
        k, dn, de = nbr.shape[1], table.shape[1], edge_feat.shape[1]
        T = qd - dn
        dev = q.device
        d_out = d_out.contiguous()
        flat = torch.zeros((sum(w.numel() for w in weights),), dtype=torch.float32, device=dev)   # one memset
        grads, off = [], 0
        for w in weights:
            grads.append(flat[off:off + w.numel()].view(w.shape))
            off += w.numel()
        grads = tuple(grads)
        d_q = torch.empty_like(q)
        d_cat = torch.empty((n, qd + dn), dtype=torch.float32, device=dev)
        d_table = torch.zeros_like(table) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(dev):
            floats = int(_lib.lib().flid_train_layer_scratch_floats(n, k, H, dn, de, T))
            scratch = torch.empty((floats,), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().flid_train_layer_bwd(
                _lib.TrainWeights(*[w.data_ptr() for w in weights]), _lib.ptr(q), _lib.ptr(merge_self), _lib.ptr(table),
                _lib.ptr(hrow), hrow_off, _lib.ptr(nbr), _lib.ptr(eid), _lib.ptr(dt), _lib.ptr(edge_feat), n, k, H, dn, de, T,
                p_drop, seed, _lib.TrainSaved(*[t.data_ptr() for t in saved]), _lib.ptr(d_out), _lib.ptr(d_q),
                _lib.ptr(d_cat), _lib.ptr(d_table), _lib.TrainWeights(*[g.data_ptr() for g in grads]), _lib.ptr(scratch),
                _lib.stream()))
        return (d_q, d_cat[:, qd:], d_table, None, None, None, None, None, None, None, None) + grads


def folded_weights(attn, kd, qd):
    """(fold_q [H*kd, qd], fold_o [qd, H*kd]) of one MultiHeadAttention, built with differentiable torch ops:
    score_hj = scale (Wq_h q).(Wk_h x_j) = (scale Wk_h^T Wq_h q).x_j;  residual_fc(sum_j a_hj Wv_h x_j) = (Wr_h Wv_h) z_h."""
    H, hd = attn.num_heads, attn.head_dim
    wq = attn.query_projection.weight.view(H, hd, qd)
    wk = attn.key_projection.weight.view(H, hd, kd)
    wv = attn.value_projection.weight.view(H, hd, kd)
    fold_q = (torch.matmul(wk.transpose(1, 2), wq) * attn.scaling_factor).reshape(H * kd, qd)
    fold_o = torch.matmul(attn.residual_fc.weight, torch.block_diag(*wv.unbind(0)))        # [qd, H*hd] . [H*hd, H*kd]
    return fold_q, fold_o


def attention_layer(attn, merge, time_w, time_b, h_self, merge_self, table, hrow, nbr, eid, dt, edge_feat, training,
                    seed=None):
    """One MultiHeadAttention + MergeLayer evaluation (models/modules.py:167-245, :58-69) for n targets.

    h_self [n, dn]: layer input of the targets; merge_self [n, dn]: MergeLayer's second input;
    table / hrow: where the neighbour slots' layer inputs live.  ``seed`` fixes both dropouts."""
    n, dn = h_self.shape
    T = time_w.shape[0]
    kd, qd = dn + edge_feat.shape[1] + T, dn + T
    p = float(attn.dropout.p) if training else 0.0
    te0 = torch.cos(time_b)                                    # cos(fma(0, w, b)), models/TGAT.py:90
    query = torch.cat([h_self, te0.expand(n, T)], dim=1)      # also the residual (modules.py:186)
    fold_q, fold_o = folded_weights(attn, kd, qd)
    if p > 0.0 and seed is None:
        seed = _score_seed()
    return TrainLayer.apply(query, merge_self, table, hrow, nbr, eid, dt, edge_feat, p, seed or 0, attn.num_heads,
                            fold_q, fold_o, attn.residual_fc.bias, attn.layer_norm.weight, attn.layer_norm.bias,
                            merge.fc1.weight, merge.fc1.bias, merge.fc2.weight, merge.fc2.bias, time_w, time_b)


def output_keep_mask(seed, n, qd, p_drop, device):
    """The kernel's residual_fc-output dropout bits of a layer call with this seed, bool [n, qd] (tests)."""
    keep = torch.empty((n, qd), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().flid_train_layer_out_keep_mask(int(seed), n, qd, float(p_drop), _lib.ptr(keep),
                                                             _lib.stream()))
    return keep.bool()


def sample_levels(sampler, node_ids, node_interact_times, depth, k, device):
    """Top-down sampling in one C call: levels[l] = (ids, nbr, eid, dt) for the targets evaluated at layer l."""
    ids = np.ascontiguousarray(node_ids, dtype=np.int64)
    t_np = np.asarray(node_interact_times)
    n = ids.shape[0]
    if n and (int(ids.min()) < 0 or int(ids.max()) > sampler.num_nodes):
        raise IndexError("flid_b200 training path: node id outside the graph")
    with torch.cuda.device(device):
        d_ids = _lib.to_device(ids, np.int64, device, "tr_ids")
        d_t = _lib.to_device(t_np, np.float64, device, "tr_times")        # float32 -> float64 is exact
        levels, ptrs, nl = {}, [[], [], [], [], []], n
        for l in range(depth, 0, -1):
            tens = (torch.empty((nl,), dtype=torch.int64, device=device),
                    torch.empty((nl,), dtype=torch.float64, device=device),
                    torch.empty((nl, k), dtype=torch.int64, device=device),
                    torch.empty((nl, k), dtype=torch.int64, device=device),
                    torch.empty((nl, k), dtype=torch.float32, device=device))
            for lst, t in zip(ptrs, tens):
                lst.insert(0, t.data_ptr())
            levels[l] = (tens[0], tens[2], tens[3], tens[4])
            nl *= 1 + k
        arrs = [(_lib.c_void * depth)(*lst) for lst in ptrs]
        _lib.check(_lib.lib().flid_train_sample_levels(sampler.handle, _lib.ptr(d_ids), _lib.ptr(d_t),
                                                       1 if t_np.dtype == np.float32 else 0, n, int(k), depth, *arrs,
                                                       _lib.stream()))
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
            table, hrow = h_prev, n
        out = attention_layer(conv_layers[l - 1], merge_layers[l - 1], w_t, b_t, h_prev[:n], node_feat[t_ids], table,
                              hrow, nbr, eid, dt, edge_feat, training, None if seeds is None else seeds[l - 1])
        h_prev = out
    return out
