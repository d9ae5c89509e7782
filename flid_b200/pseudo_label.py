"""Pseudo-label scoring on the device: decoder MLP + softmax/argmax emission and the EST /
CST filters, mirroring

* ``models.modules.MLPClassifier`` (``models/modules.py:72-97``) -- same constructor and
  ``state_dict`` keys (fc1/fc2/fc3); eval-mode forward runs ``flid_pseudo_label``;
* the emission loop of ``PTCL/E_step.py:305-352`` -> ``emit_pseudo_labels``;
* ``entropy_filter`` / ``prob_filter`` / ``update_pseudo_labels`` of ``PTCL/utils.py:38-123``
  (same names, argument order and in-place semantics on the ``[1, E]`` / ``[2, E]`` label tensor).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class MLPClassifier(nn.Module):
    def __init__(self, input_dim: int, dropout: float = 0.1, num_classes: int = 2):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, 80)
        self.fc2 = nn.Linear(80, 10)
        self.fc3 = nn.Linear(10, num_classes)
        self.act = nn.ReLU()
        self.dropout = nn.Dropout(dropout)

    def _weights(self):
        w = _lib.MlpWeights()
        ps = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.fc3.weight, self.fc3.bias]
        for p in ps:
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("flid_b200.MLPClassifier: parameters must be contiguous float32 CUDA tensors")
        (w.fc1_w, w.fc1_b, w.fc2_w, w.fc2_b, w.fc3_w, w.fc3_b) = [p.data_ptr() for p in ps]
        w.input_dim, w.hidden1 = self.fc1.in_features, self.fc1.out_features
        w.hidden2, w.num_classes = self.fc2.out_features, self.fc3.out_features
        return w

    def score(self, x: torch.Tensor, want_logits: bool = False):
        """x float32 [n, in] on the device -> (probs [n, C], labels int64 [n], logits or None)."""
        dev = _lib.require_cuda(x.device)
        x = x.detach().to(torch.float32).contiguous()
        n, c = x.shape[0], self.fc3.out_features
        with torch.cuda.device(dev):
            probs = torch.empty((n, c), dtype=torch.float32, device=dev)
            labels = torch.empty(n, dtype=torch.int64, device=dev)
            logits = torch.empty((n, c), dtype=torch.float32, device=dev) if want_logits else None
            w = self._weights()
            _lib.check(_lib.lib().flid_pseudo_label(C.byref(w), _lib.ptr(x), n, _lib.ptr(probs), _lib.ptr(labels),
                                                    _lib.ptr(logits), _lib.stream()))
        return probs, labels, logits

    def forward(self, x: torch.Tensor):
        """models/modules.py:86-97.  The fused kernel (forward-only, no dropout) serves eval-mode calls that
        need no graph; training mode (dropout, also under no_grad as in the reference) and any call whose
        input or parameters require grad take the torch path."""
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.training or needs_graph or x.device.type != "cuda":
            # decoder training on cached embeddings (PTCL/E_step.py) is plain torch autograd, as in the reference
            x = self.dropout(self.act(self.fc1(x)))
            x = self.dropout(self.act(self.fc2(x)))
            return self.fc3(x)
        return self.score(x, want_logits=True)[2]


def emit_pseudo_labels(decoder: MLPClassifier, embeddings: torch.Tensor):
    """PTCL/E_step.py:305-352 for a single-way dataset, all events in one launch:
    returns (labels int64 [E], probabilities float32 [E, C])."""
    probs, labels, _ = decoder.score(embeddings)
    return labels, probs


def _flatten(ps_labels, store_item):
    if ps_labels.shape[0] == 2:   # double-way layout: [2, E] labels, [2, E, C] probabilities
        return store_item.reshape(-1, store_item.shape[-1])
    return store_item


def entropy_filter(ps_labels, ps_labels_store, threshold=0.6):
    """EST (PTCL/utils.py:38-54): rows whose entropy of softmax(sum of stored probs) exceeds
    ``threshold`` get label -1 (in place on ``ps_labels``, which is also returned)."""
    dev = _lib.require_cuda(ps_labels.device)
    assert ps_labels.dtype == torch.float32 and ps_labels.is_contiguous()
    flat = [_flatten(ps_labels, s).to(dev, torch.float32).contiguous() for s in ps_labels_store]
    n, c = flat[0].shape
    assert ps_labels.numel() == n
    ptrs = (C.c_void_p * len(flat))(*[t.data_ptr() for t in flat])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().flid_entropy_filter(ptrs, len(flat), n, c, float(threshold), _lib.ptr(ps_labels),
                                                  _lib.stream()))
    return ps_labels


def prob_filter(ps_labels, ps_labels_store, threshold=0.6):
    """CST (PTCL/utils.py:56-67): rows whose max probability in the LAST stored iteration is
    below ``threshold`` get label -1."""
    dev = _lib.require_cuda(ps_labels.device)
    assert ps_labels.dtype == torch.float32 and ps_labels.is_contiguous()
    last = _flatten(ps_labels, ps_labels_store[-1]).to(dev, torch.float32).contiguous()
    n, c = last.shape
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().flid_prob_filter(_lib.ptr(last), n, c, float(threshold), _lib.ptr(ps_labels),
                                               _lib.stream()))
    return ps_labels


def update_pseudo_labels(data, pseudo_labels, pseudo_labels_store, double_way_dataset, mode,
                         use_transductive=0, save=False, save_path=0, threshold=0.6, iter_num=-1, ps_filter='none'):
    """PTCL/utils.py:69-123: filter on the device, then overwrite with ground truth where the
    event time equals the label time (host masks, exactly as the reference builds them)."""
    import os
    if save:
        os.makedirs(save_path, exist_ok=True)
        torch.save(pseudo_labels, os.path.join(save_path, f'raw_{iter_num}.pt'))
    if ps_filter == 'entropy':
        pseudo_labels = entropy_filter(pseudo_labels, pseudo_labels_store, threshold=threshold)
    elif ps_filter == 'probability':
        pseudo_labels = prob_filter(pseudo_labels, pseudo_labels_store, threshold=threshold)
    full = data['full_data']
    true_labels, labels_times, interact_times = full.labels, full.labels_time, full.node_interact_times
    train_mask = np.arange(pseudo_labels.shape[1]) < data['val_offest']
    dev = pseudo_labels.device

    def put(row, mask, values):
        m = torch.from_numpy(mask).to(torch.bool).to(dev)
        pseudo_labels[row, m] = torch.from_numpy(values[mask].astype('float32')).to(dev)

    if data['dataset_name'] in double_way_dataset:
        for row in (0, 1):
            mask = interact_times == labels_times[row]
            put(row, mask & train_mask if use_transductive else mask, true_labels[row])
    elif mode == 'ps':
        mask = interact_times == labels_times
        put(0, mask & train_mask if use_transductive else mask, true_labels)
    elif mode == 'gt':
        pseudo_labels[0, :] = torch.from_numpy(true_labels.astype('float32')).to(dev)
    if save:
        torch.save(pseudo_labels, os.path.join(save_path, f'updated_{iter_num}.pt'))
    return pseudo_labels
