"""TEST INFRASTRUCTURE ONLY -- torch-fp32 CPU restatement of FLiD's pseudo-label scoring.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this; the product never does.

Parity status: PINNED for the decoder and the two filters against the reference
(``models/modules.py:72-97``, ``PTCL/utils.py:38-123`` imported by
``tests/golden/make_golden.py``).  The emission loop (``PTCL/E_step.py:305-352``)
lives in a module that cannot be imported (it pulls in the missing
``models.EdgeBank``), so ``emit`` restates those lines and is pinned only through
its parts (decoder + softmax + max).

* ``decoder``          -- MLPClassifier 172->80->10->C, ReLU, dropout = id in eval
* ``emit``             -- per batch: logits -> softmax(dim=1) -> max -> (label, probs)
* ``entropy_filter``   -- EST: p = softmax(sum_iters probs); H = -sum p*log2(p+1e-10); H > thr => -1
* ``prob_filter``      -- CST: max_c probs[-1] < thr => -1
* ``update_pseudo_labels`` -- filter, then ground-truth overwrite where t == labels_time
  (single-way datasets, modes 'ps'/'gt', optional train-only mask)
"""
import numpy as np
import torch
import torch.nn.functional as F


def default_decoder_params(input_dim=172, num_classes=2, seed=0):
    g = torch.Generator().manual_seed(seed + 77)

    def lin(o, i):
        b = 1.0 / np.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b, (torch.rand(o, generator=g) * 2 - 1) * b

    p = {}
    p["fc1.weight"], p["fc1.bias"] = lin(80, input_dim)
    p["fc2.weight"], p["fc2.bias"] = lin(10, 80)
    p["fc3.weight"], p["fc3.bias"] = lin(num_classes, 10)
    return p


def decoder(p, x):
    x = F.relu(F.linear(x, p["fc1.weight"], p["fc1.bias"]))
    x = F.relu(F.linear(x, p["fc2.weight"], p["fc2.bias"]))
    return F.linear(x, p["fc3.weight"], p["fc3.bias"])


def emit(p, embeddings, batch_size=200):
    """PTCL/E_step.py:310-352 for a single-way dataset: labels int64[E], probs f32[E,C]."""
    labels, probs = [], []
    with torch.no_grad():
        for lo in range(0, embeddings.shape[0], batch_size):
            pr = torch.softmax(decoder(p, embeddings[lo:lo + batch_size]), dim=1)
            _, lab = torch.max(pr, dim=1)
            labels.append(lab.to(torch.long)), probs.append(pr)
    return torch.cat(labels, dim=0), torch.cat(probs, dim=0)


def entropy_filter(ps_labels, ps_labels_store, threshold=0.6):
    """ps_labels float32 [1,E] or [2,E]; store = list of [E,C] (or [2,E,C]) probs."""
    acc = torch.sum(torch.stack(ps_labels_store), dim=0)
    double_way = ps_labels.shape[0] == 2
    if double_way:
        ps_labels = torch.cat([ps_labels[0], ps_labels[1]], dim=0).reshape(1, -1)
        acc = torch.cat([acc[0], acc[1]], dim=0)
    pr = F.softmax(acc, dim=1)
    ent = -torch.sum(pr * torch.log2(pr + 1e-10), dim=1)
    ps_labels[:, ent > threshold] = -1
    if double_way:
        half = ps_labels.shape[1] // 2
        ps_labels = torch.cat([ps_labels[:, :half], ps_labels[:, half:]], dim=0)
    return ps_labels


def prob_filter(ps_labels, ps_labels_store, threshold=0.6):
    pr = ps_labels_store[-1]
    double_way = ps_labels.shape[0] == 2
    if double_way:
        ps_labels = torch.cat([ps_labels[0], ps_labels[1]], dim=0).reshape(1, -1)
        pr = torch.cat([pr[0], pr[1]], dim=0)
    ps_labels[:, torch.max(pr, dim=1)[0] < threshold] = -1
    if double_way:
        half = ps_labels.shape[1] // 2
        ps_labels = torch.cat([ps_labels[:, :half], ps_labels[:, half:]], dim=0)
    return ps_labels


def update_pseudo_labels(true_labels, labels_time, interact_times, val_offset, pseudo_labels, store,
                         mode="ps", use_transductive=0, threshold=0.6, ps_filter="none"):
    """PTCL/utils.py:69-123, single-way branch, no file saving."""
    if ps_filter == "entropy":
        pseudo_labels = entropy_filter(pseudo_labels, store, threshold)
    elif ps_filter == "probability":
        pseudo_labels = prob_filter(pseudo_labels, store, threshold)
    if mode == "gt":
        pseudo_labels[0, :] = torch.from_numpy(true_labels.astype("float32"))
        return pseudo_labels
    mask = interact_times == labels_time
    if use_transductive:
        mask = mask & (np.arange(pseudo_labels.shape[1]) < val_offset)
    mask = torch.from_numpy(mask).to(torch.bool)
    pseudo_labels[0, mask] = torch.from_numpy(true_labels[mask.numpy()].astype("float32"))
    return pseudo_labels
