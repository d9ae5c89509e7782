"""ctypes binding of include/flid_b200.h (the C-ABI shared library).

There is deliberately no fallback: if ``libflid_b200.so`` cannot be built/loaded, or no
CUDA device is present when a compute entry point is called, an exception is raised.
PyTorch is used only for device memory, pinned staging buffers and the current stream.
"""
import ctypes as C
import os
import threading

import numpy as np
import torch

from . import build as _build

c_i64p = C.POINTER(C.c_int64)
c_void = C.c_void_p


class LayerWeights(C.Structure):
    _fields_ = [(n, c_void) for n in ("query_w", "key_w", "value_w", "ln_w", "ln_b", "res_w", "res_b",
                                      "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class TgnState(C.Structure):
    _fields_ = [("num_rows", C.c_int64)] + [(n, c_void) for n in (
        "memories", "last_updated", "pending_msg", "pending_ts", "has_pending", "next_memories", "layer0", "scratch")]


class GruWeights(C.Structure):
    _fields_ = [(n, c_void) for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]


class MlpWeights(C.Structure):
    _fields_ = [(n, c_void) for n in ("fc1_w", "fc1_b", "fc2_w", "fc2_b", "fc3_w", "fc3_b")] + \
               [(n, C.c_int) for n in ("input_dim", "hidden1", "hidden2", "num_classes")]


class TrainWeights(C.Structure):   # flid_train_weights and flid_train_grads (same members)
    _fields_ = [(n, c_void) for n in ("fold_q", "fold_o", "res_b", "ln_w", "ln_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
                                      "time_w", "time_b")]


class TrainSaved(C.Structure):
    _fields_ = [(n, c_void) for n in ("q", "merge_self", "u", "probs", "z", "y", "ln", "hid", "out")]


class TrainLevel(C.Structure):
    _fields_ = [(n, c_void) for n in ("ids", "nbr", "eid", "dt")] + [("n", C.c_int64)]


_SIGNATURES = {
    "flid_last_error": (C.c_char_p, []),
    "flid_abi_version": (C.c_int, []),
    "flid_launch_count": (C.c_int64, []),
    "flid_debug_gemm": (C.c_int, [C.c_int, c_void, C.c_int64, c_void, C.c_int, c_void, C.c_int64, C.c_int, c_void,
                                  C.c_int64, c_void, c_void, C.c_int64, C.c_int64, C.c_int, C.c_int, c_void]),
    "flid_debug_gemm_time": (C.c_int, [c_void, C.c_int64, c_void, C.c_int, c_void, C.c_int64, C.c_int, c_void,
                                       C.c_int64, c_void, c_void, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_float), c_void]),
    "flid_graph_build_events": (C.c_int, [c_void, c_void, c_void, c_void, C.c_int64, C.c_int64, C.c_int,
                                          C.POINTER(c_void), c_void]),
    "flid_graph_build_entries": (C.c_int, [c_void, c_void, c_void, c_void, C.c_int64, C.c_int64, C.c_int,
                                           C.POINTER(c_void), c_void]),
    "flid_graph_free": (None, [c_void]),
    "flid_graph_info": (C.c_int, [c_void, c_i64p, c_i64p, c_i64p]),
    "flid_graph_export_host": (C.c_int, [c_void, c_void, c_void, c_void, c_void]),
    "flid_sample_recent": (C.c_int, [c_void, c_void, c_void, C.c_int, C.c_int64, C.c_int, c_void, c_void, c_void,
                                     c_void]),
    "flid_sample_cut": (C.c_int, [c_void, c_void, c_void, C.c_int, C.c_int64, c_void, c_void, c_void]),
    "flid_tgat_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(c_void)]),
    "flid_tgat_free": (None, [c_void]),
    "flid_tgat_set_weights": (C.c_int, [c_void, c_void, c_void, C.POINTER(LayerWeights), c_void]),
    "flid_tgat_cache_node_table": (C.c_int, [c_void, c_void, C.c_int64, c_void]),
    "flid_tgat_refresh_node_rows": (C.c_int, [c_void, c_void, c_void, C.c_int64, c_void]),
    "flid_tgat_embed": (C.c_int, [c_void, c_void, c_void, c_void, c_void, c_void, C.c_int, C.c_int64, C.c_int,
                                  c_void, c_void]),
    "flid_tgat_memo_build": (C.c_int, [c_void, c_void, c_void, c_void, C.c_int, C.c_int, c_void, C.c_int64, C.c_int64,
                                       c_void, c_void]),
    "flid_tgat_embed_memo": (C.c_int, [c_void, c_void, c_void, c_void, C.POINTER(c_void), c_void, c_void, C.c_int,
                                       C.c_int64, C.c_int, c_void, c_void]),
    "flid_tgat_set_self_from_memo": (C.c_int, [c_void, C.c_int]),
    "flid_tgat_set_chunk_targets": (C.c_int, [c_void, C.c_int64]),
    "flid_tgat_set_numeric_mode": (C.c_int, [c_void, C.c_int]),
    "flid_tgat_set_ln_fold": (C.c_int, [c_void, C.c_int]),
    "flid_tgat_set_sort_queries": (C.c_int, [c_void, C.c_int]),
    "flid_tgat_set_wait_event": (C.c_int, [c_void, c_void]),
    "flid_tgat_bulk_invalidate": (C.c_int, [c_void]),
    "flid_tgat_memo_build_owner_range": (C.c_int, [c_void, c_void, c_void, c_void, C.c_int, C.c_int, c_void, C.c_int64,
                                                   C.c_int64, C.c_int, c_void, c_void]),
    "flid_tgat_set_bulk_range": (C.c_int, [c_void, C.c_int64, C.c_int64]),
    "flid_graph_export_mirror": (C.c_int, [c_void, c_void, c_void]),
    "flid_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(c_void), c_void]),
    "flid_peer_open": (C.c_int, [c_void, C.POINTER(c_void)]),
    "flid_peer_close": (C.c_int, [c_void]),
    "flid_peer_free": (C.c_int, [c_void]),
    "flid_memo_exchange_p2p": (C.c_int, [c_void, c_void, C.POINTER(c_void), c_i64p, C.c_int, C.c_int, C.c_int, c_void]),
    "flid_tgat_set_bulk_projection": (C.c_int, [c_void, C.c_int]),
    "flid_tgat_profile": (C.c_int, [c_void, C.c_int]),
    "flid_tgat_profile_read": (C.c_int, [c_void, C.POINTER(C.c_double), c_i64p]),
    "flid_tgat_last_stats": (C.c_int, [c_void, c_i64p]),
    "flid_tgn_reset": (C.c_int, [C.POINTER(TgnState), c_void, C.c_int, C.c_int, c_void]),
    "flid_tgn_rebuild": (C.c_int, [c_void, C.POINTER(TgnState), C.POINTER(GruWeights), c_void, c_void]),
    "flid_tgn_step": (C.c_int, [c_void, c_void, C.POINTER(TgnState), C.POINTER(GruWeights), c_void, c_void, c_void,
                                c_void, c_void, c_void, C.c_int64, C.c_int, C.c_int, c_void, c_void, c_void]),
    "flid_tgn_pass": (C.c_int, [c_void, c_void, C.POINTER(TgnState), C.POINTER(GruWeights), c_void, c_void, c_void,
                                c_void, c_void, c_void, C.c_int64, C.c_int64, C.c_int, c_void, c_void, c_void, C.c_int,
                                c_void]),
    "flid_pseudo_label": (C.c_int, [C.POINTER(MlpWeights), c_void, C.c_int64, c_void, c_void, c_void, c_void]),
    "flid_entropy_filter": (C.c_int, [C.POINTER(c_void), C.c_int, C.c_int64, C.c_int, C.c_float, c_void, c_void]),
    "flid_prob_filter": (C.c_int, [c_void, C.c_int64, C.c_int, C.c_float, c_void, c_void]),
    "flid_neighbor_mean": (C.c_int, [c_void, c_void, C.c_int, c_void, c_void, C.c_int, C.c_int64, C.c_int, C.c_int, c_void,
                                     c_void]),
    "flid_dense_weight_create": (c_void, [c_void, C.c_int64, C.c_int, C.c_int, c_void]),
    "flid_dense_weight_update": (C.c_int, [c_void, c_void, C.c_int64, c_void]),
    "flid_dense_weight_free": (None, [c_void]),
    "flid_dense": (C.c_int, [c_void, c_void, c_void, C.c_int64, C.c_int, c_void, c_void, C.c_int64, C.c_int, c_void, c_void,
                             C.c_int64, C.c_int, c_void, C.c_int64, C.c_int64, c_void]),
    "flid_time_rows": (C.c_int, [c_void, c_void, c_void, c_void, C.c_int, c_void, C.c_int64, c_void]),
    "flid_token_mix": (C.c_int, [c_void, C.c_int, C.c_int, c_void, c_void, C.c_float, c_void, c_void, c_void, c_void, C.c_int,
                                 c_void, C.c_int64, c_void]),
    "flid_token_mean": (C.c_int, [c_void, C.c_int, C.c_int, c_void, C.c_int64, C.c_int64, c_void]),
    "flid_row_layernorm": (C.c_int, [c_void, C.c_int64, c_void, c_void, C.c_float, c_void, C.c_int64, C.c_int64, C.c_int, c_void]),
    "flid_add_periodic_rows": (C.c_int, [c_void, c_void, C.c_int, C.c_int, C.c_int64, c_void]),
    "flid_seq_attention": (C.c_int, [c_void, C.c_int64, c_void, C.c_int64, c_void, C.c_int64, c_void, C.c_int, C.c_int, C.c_int,
                                     c_void, C.c_int64, C.c_int, C.c_int64, c_void]),
    "flid_attn_train_partials": (C.c_int64, [C.c_int64]),
    "flid_attn_train_fwd": (C.c_int, [c_void] * 9 + [C.c_int64] + [C.c_int] * 5 + [C.c_float, C.c_uint64, c_void,
                                                                                  c_void, c_void]),
    "flid_attn_train_bwd": (C.c_int, [c_void] * 9 + [C.c_int64] + [C.c_int] * 5 + [C.c_float, C.c_uint64] +
                            [c_void] * 6),
    "flid_train_layer_out_keep_mask": (C.c_int, [C.c_uint64, C.c_int64, C.c_int, C.c_float, c_void, c_void]),
    "flid_train_sample_levels": (C.c_int, [c_void, c_void, c_void, C.c_int, C.c_int64, C.c_int, C.c_int] +
                                 [C.POINTER(c_void)] * 5 + [c_void]),
    "flid_train_model_scratch_floats": (C.c_int64, [C.c_int64] + [C.c_int] * 6),
    "flid_train_model_fwd": (C.c_int, [C.POINTER(TrainWeights), C.POINTER(TrainLevel), C.POINTER(TrainSaved), c_void,
                                       c_void, c_void] + [C.c_int] * 6 + [C.c_float, C.POINTER(C.c_uint64), c_void,
                                                                          c_void]),
    "flid_train_model_bwd": (C.c_int, [C.POINTER(TrainWeights), C.POINTER(TrainLevel), C.POINTER(TrainSaved), c_void,
                                       c_void] + [C.c_int] * 6 + [C.c_float, C.POINTER(C.c_uint64), c_void,
                                                                  C.POINTER(TrainWeights), c_void, c_void, c_void,
                                                                  c_void]),
    "flid_attn_train_keep_mask": (C.c_int, [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_float, c_void, c_void]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


def library_path():
    return _build.LIB


def lib():
    """Load (building first if sources changed and nvcc is present) the CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = _build.build(verbose=bool(os.environ.get("FLID_VERBOSE_BUILD")))
            handle = C.CDLL(path)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the symbol is missing: fail loudly
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


def check(status):
    if status == 0:
        return
    msg = lib().flid_last_error().decode("utf-8", "replace")
    if status == 3:
        raise IndexError(msg)
    if status == 4 or msg.startswith(("Number of sampled", "The sum of node_feat_dim")):
        raise AssertionError(msg)
    if status == 1:
        raise ValueError(msg)
    raise RuntimeError(f"flid_b200 CUDA error: {msg}")


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("flid_b200 has no CPU fallback: a CUDA device (B200, sm_100a) is required")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError(f"flid_b200 has no CPU fallback: device must be CUDA, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def ptr(t):
    return c_void(t.data_ptr()) if t is not None else c_void(None)


def stream():
    return c_void(torch.cuda.current_stream().cuda_stream)


class _PinnedPool:
    """Reusable pinned staging buffers keyed by (tag, dtype); grow-only.  A buffer is not
    rewritten by the host before the async copy that last read it has finished."""

    def __init__(self):
        self.bufs = {}
        self.events = {}

    def get(self, tag, numel, dtype):
        key = (tag, dtype)
        ev = self.events.pop(key, None)
        if ev is not None:
            ev.synchronize()
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < numel:
            buf = torch.empty(max(numel, 1024), dtype=dtype, pin_memory=True)
            self.bufs[key] = buf
        return buf[:numel]

    def mark(self, tag, dtype):
        ev = torch.cuda.Event()
        ev.record()
        self.events[(tag, dtype)] = ev


_pool = _PinnedPool()


def to_device(arr, dtype, device, tag, sync_follows=False):
    """numpy -> device tensor through a pinned staging buffer (async on the current stream).
    ``sync_follows``: the caller synchronises the stream before anyone can stage under this tag
    again (every compute entry point that reports errors does), so no guard event is recorded."""
    a = np.ascontiguousarray(arr, dtype=dtype)
    t = torch.from_numpy(a)
    stage = _pool.get(tag, t.numel(), t.dtype)
    stage.copy_(t.reshape(-1))
    out = stage.to(device, non_blocking=True).reshape(a.shape)
    if not sync_follows:
        _pool.mark(tag, t.dtype)
    return out


def to_device_concat(arrays, dtype, device, tag, sync_follows=False):
    """concatenate(arrays) on the device without materialising the concatenation on the host: each
    piece is copied into its slice of one pinned staging buffer, one async H2D copy follows."""
    parts = [np.asarray(a) for a in arrays]
    total = sum(p.shape[0] for p in parts)
    tdtype = torch.from_numpy(np.empty(0, dtype=dtype)).dtype
    stage = _pool.get(tag, total, tdtype)
    view = stage.numpy()
    off = 0
    for p in parts:
        n = p.shape[0]
        np.copyto(view[off:off + n], p, casting="same_kind")
        off += n
    out = stage.to(device, non_blocking=True)
    if not sync_follows:
        _pool.mark(tag, tdtype)
    return out


def to_host(t, tag, copy=True):
    """device tensor -> numpy array via pinned staging (synchronises the current stream).  ``copy=False`` returns
    a view of the staging buffer, valid until the next ``to_host`` with the same tag."""
    stage = _pool.get(tag, t.numel(), t.dtype)
    stage.copy_(t.reshape(-1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    out = stage.numpy().reshape(t.shape)
    return out.copy() if copy else out
