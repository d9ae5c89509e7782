"""Training-mode (autograd + dropout) forward of the TGAT / TGN attention layers.

The M-step batches of the reference run ``compute_src_dst_node_temporal_embeddings`` with grad
enabled and ``model.train()`` (PTCL/M_step.py:196-325, NPL/NPL.py:185-314,
PTCL/EM_warmup.py:113-238).  Here a layer is split where its cost is:

* the whole L-layer stack is ONE autograd node, ``TrainStack``: its forward and its backward are
  each one C-ABI call (``flid_train_model_fwd`` / ``_bwd``, csrc/train_layer.cu) that launches, per
  layer, the tcgen05 GEMMs, the attention-stream kernels (gather, time encoding, masked softmax,
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
_NW = len(_LAYER_TENSORS)


def _saved_layout(n, k, H, dn, de, T):
    """float counts of one layer's flid_train_saved members, in struct order"""
    qd, zw = dn + T, H * (dn + de + T)
    return (n * qd, n * dn, n * zw, n * H * k, n * zw, n * qd, n * qd, n * dn, n * dn)   # q merge_self u probs z y ln hid out


class TrainStack(torch.autograd.Function):
    """The whole L-layer attention stack as ONE autograd node: forward = ``flid_train_model_fwd``, backward =
    ``flid_train_model_bwd`` (csrc/train_layer.cu), one C call each.

    Differentiable inputs: node_feat [R, dn] (the layer-0 table), te0 [T] = cos(time_b), and 11 weight tensors
    per layer in ``_LAYER_TENSORS`` order (folded projections included).  ``levels`` / ``seeds`` are plain Python
    lists (index l-1 for layer l)."""

    @staticmethod
    def forward(ctx, node_feat, edge_feat, te0, levels, seeds, p_drop, num_heads, k, *weights):
        node_feat, edge_feat, te0 = node_feat.contiguous(), edge_feat.contiguous(), te0.contiguous()
        weights = tuple(w.contiguous() for w in weights)
        L, H = len(levels), int(num_heads)
        dn, de, T = node_feat.shape[1], edge_feat.shape[1], te0.shape[0]
        qd, zw = dn + T, H * (dn + de + T)
        dev = node_feat.device
        for t in (node_feat, edge_feat, te0) + weights:
            if t.dtype != torch.float32 or not t.is_cuda:
                raise TypeError("TrainStack: float32 CUDA tensors required (flid_b200 has no CPU path)")
        if len(weights) != _NW * L or weights[0].shape != (zw, qd) or weights[1].shape != (qd, zw):
            raise ValueError("TrainStack: inconsistent weight shapes")
        n_top = levels[L - 1][0].shape[0]
        out = torch.empty((n_top, dn), dtype=torch.float32, device=dev)
        layouts = [_saved_layout(levels[l][0].shape[0], k, H, dn, de, T) for l in range(L)]
        total = sum(sum(lay) for lay in layouts) - layouts[L - 1][-1] + levels[0][0].shape[0] * qd
        buf = torch.empty((total,), dtype=torch.float32, device=dev)      # every saved activation + pre_scratch
        base, off, saved_ptrs = buf.data_ptr(), 0, []
        for l, lay in enumerate(layouts):
            ptrs = []
            for j, cnt in enumerate(lay):
                if l == L - 1 and j == len(lay) - 1:
                    ptrs.append(out.data_ptr())                           # the top layer's output is the result
                else:
                    ptrs.append(base + 4 * off)
                    off += cnt
            saved_ptrs.append(ptrs)
        pre_ptr = base + 4 * off
        wst = (_lib.TrainWeights * L)(*[_lib.TrainWeights(*[w.data_ptr() for w in weights[_NW * l:_NW * (l + 1)]])
                                        for l in range(L)])
        lst = (_lib.TrainLevel * L)(*[_lib.TrainLevel(lv[0].data_ptr(), lv[1].data_ptr(), lv[2].data_ptr(),
                                                      lv[3].data_ptr(), lv[0].shape[0]) for lv in levels])
        sst = (_lib.TrainSaved * L)(*[_lib.TrainSaved(*p) for p in saved_ptrs])
        sds = (_lib.C.c_uint64 * L)(*[int(s) for s in list(seeds)[:L]])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().flid_train_model_fwd(wst, lst, sst, _lib.ptr(node_feat), _lib.ptr(edge_feat),
                                                       _lib.ptr(te0), L, int(k), H, dn, de, T, float(p_drop), sds,
                                                       _lib.c_void(pre_ptr), _lib.stream()))
        ctx.save_for_backward(node_feat, edge_feat, te0, *weights)
        ctx.keep = (levels, buf)                # index tensors and activations (no autograd history)
        ctx.structs = (wst, lst, sst, sds)
        ctx.meta = (L, H, int(k), float(p_drop))
        return out

    @staticmethod
    def backward(ctx, d_out):
        L, H, k, p_drop = ctx.meta
        node_feat, edge_feat, te0 = ctx.saved_tensors[:3]
        weights = ctx.saved_tensors[3:]
        wst, lst, sst, sds = ctx.structs
        levels = ctx.keep[0]
        dn, de, T = node_feat.shape[1], edge_feat.shape[1], te0.shape[0]
        dev = node_feat.device
        d_out = d_out.contiguous()
        sizes = [w.numel() for w in weights]
        flat = torch.zeros((sum(sizes) + T,), dtype=torch.float32, device=dev)      # one memset for every gradient
        parts = flat.split(sizes + [T])
        grads = tuple(g.view(w.shape) for g, w in zip(parts, weights))
        d_te0 = parts[-1]
        d_node = torch.zeros_like(node_feat) if ctx.needs_input_grad[0] else None
        base = flat.data_ptr()
        offs, o = [], 0
        for sz in sizes:
            offs.append(base + 4 * o)
            o += sz
        gst = (_lib.TrainWeights * L)(*[_lib.TrainWeights(*offs[_NW * l:_NW * (l + 1)]) for l in range(L)])
        with torch.cuda.device(dev):
            floats = int(_lib.lib().flid_train_model_scratch_floats(levels[L - 1][0].shape[0], k, L, H, dn, de, T))
            scratch = torch.empty((floats,), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().flid_train_model_bwd(wst, lst, sst, _lib.ptr(node_feat), _lib.ptr(edge_feat), L, k, H,
                                                       dn, de, T, p_drop, sds, _lib.ptr(d_out), gst, _lib.ptr(d_te0),
                                                       _lib.ptr(d_node), _lib.ptr(scratch), _lib.stream()))
        return (d_node, None, d_te0, None, None, None, None, None) + grads


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
        levels, ptrs, nl, keep_alive = {}, [[], [], [], [], []], n, []
        for l in range(depth, 0, -1):
            tens = (torch.empty((nl,), dtype=torch.int64, device=device),
                    torch.empty((nl,), dtype=torch.float64, device=device),
                    torch.empty((nl, k), dtype=torch.int64, device=device),
                    torch.empty((nl, k), dtype=torch.int64, device=device),
                    torch.empty((nl, k), dtype=torch.float32, device=device))
            keep_alive.append(tens)      # the C call below reads and writes every level's tensors through raw pointers
            for lst, t in zip(ptrs, tens):
                lst.insert(0, t.data_ptr())
            levels[l] = (tens[0], tens[2], tens[3], tens[4])
            nl *= 1 + k
        arrs = [(_lib.c_void * depth)(*lst) for lst in ptrs]
        _lib.check(_lib.lib().flid_train_sample_levels(sampler.handle, _lib.ptr(d_ids), _lib.ptr(d_t),
                                                       1 if t_np.dtype == np.float32 else 0, n, int(k), depth, *arrs,
                                                       _lib.stream()))
        del keep_alive               # stream-ordered free: safe after the launches were queued
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
    if len(node_ids) == 0:                     # an empty batch: nothing to sample, nothing to differentiate
        return node_feat.new_zeros((0, node_feat.shape[1]))
    levels = sample_levels(sampler, node_ids, node_interact_times, depth, k, device)
    w_t, b_t = time_encoder.w.weight.reshape(-1), time_encoder.w.bias
    dn, T = node_feat.shape[1], b_t.shape[0]
    kd, qd = dn + edge_feat.shape[1] + T, dn + T
    p = float(conv_layers[0].dropout.p) if training else 0.0
    weights = []
    for l in range(depth):
        attn, merge = conv_layers[l], merge_layers[l]
        weights += [*folded_weights(attn, kd, qd), attn.residual_fc.bias, attn.layer_norm.weight, attn.layer_norm.bias,
                    merge.fc1.weight, merge.fc1.bias, merge.fc2.weight, merge.fc2.bias, w_t, b_t]
    if seeds is None:
        seeds = [_score_seed() if p > 0.0 else 0 for _ in range(depth)]
    te0 = torch.cos(b_t)                                       # cos(fma(0, w, b)), models/TGAT.py:90
    return TrainStack.apply(node_feat, edge_feat, te0, [levels[l] for l in range(1, depth + 1)], list(seeds), p,
                            conv_layers[0].num_heads, k, *weights)
