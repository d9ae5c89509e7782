"""Host side of csrc/dense.cu: the forward-only dense layers GraphMixer / TCL use in evaluation
(``model.eval()`` under ``torch.no_grad()``).  Every ``nn.Linear`` becomes one ``flid_dense`` call
on the tcgen05 3xTF32 GEMM; the tiled weight images are cached per parameter and re-tiled when the
parameter's version counter moves (optimizer step, ``load_state_dict``).  Training keeps the torch
modules, whose autograd the reference relies on."""
import os

import torch

from . import _lib


class DenseWeights:
    """Cache of ``flid_dense_weight`` handles keyed by (storage pointer, shape, row stride)."""

    def __init__(self):
        self._h = {}

    def handle(self, weight: torch.Tensor):
        """``weight``: [n_out, n_in] float32 CUDA view with unit column stride (a row slice of a parameter is fine)."""
        assert weight.dim() == 2 and weight.stride(1) == 1 and weight.dtype == torch.float32 and weight.is_cuda
        key = (weight.data_ptr(), weight.shape[0], weight.shape[1], weight.stride(0))
        ent = self._h.get(key)
        if ent is None:
            h = _lib.lib().flid_dense_weight_create(_lib.ptr(weight), weight.stride(0), weight.shape[0], weight.shape[1],
                                                    _lib.stream())
            if not h:
                _lib.check(1)
            ent = [h, weight._version]
            self._h[key] = ent
        elif ent[1] != weight._version:
            _lib.check(_lib.lib().flid_dense_weight_update(ent[0], _lib.ptr(weight), weight.stride(0), _lib.stream()))
            ent[1] = weight._version
        return ent[0]

    def __deepcopy__(self, memo):
        return DenseWeights()         # handles belong to one module instance; the copy re-tiles on first use

    def clear(self):
        if self._h and torch.cuda.is_available():
            torch.cuda.synchronize()
        for h, _ in self._h.values():
            _lib.lib().flid_dense_weight_free(h)
        self._h = {}

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass


def linear(cache: DenseWeights, x, weight, bias=None, *, act=0, resid=None, idx=None, x2=None, idx2=None, rows=None, ldx=None,
           out=None):
    """act([x[idx] | x2[idx2]] @ weight.T + bias + resid) as one GEMM launch.  ``x`` / ``x2``: 2-D float32 with unit
    column stride; ``idx`` / ``idx2``: int32 row gathers; ``rows``: number of output rows when no plain segment tells
    it; ``ldx``: row stride override for ``x`` (e.g. every S-th row of a [m * S, d] matrix); ``out``: destination view
    (unit column stride, any row stride)."""
    m = int(rows if rows is not None else (idx.shape[0] if idx is not None else x.shape[0]))
    n_out = weight.shape[0]
    if out is None:
        out = torch.empty((m, n_out), dtype=torch.float32, device=x.device)
    assert out.shape == (m, n_out) and out.stride(1) == 1
    w0 = x.shape[1]
    w1 = x2.shape[1] if x2 is not None else 0
    assert w0 + w1 == weight.shape[1], (w0, w1, tuple(weight.shape))
    _lib.check(_lib.lib().flid_dense(cache.handle(weight), _lib.ptr(x), _lib.ptr(idx), int(ldx if ldx is not None else x.stride(0)),
                                     w0, _lib.ptr(x2), _lib.ptr(idx2), x2.stride(0) if x2 is not None else 0, w1,
                                     _lib.ptr(bias), _lib.ptr(resid), resid.stride(0) if resid is not None else 0, int(act),
                                     _lib.ptr(out), out.stride(0), m, _lib.stream()))
    return out


def layernorm(x, norm: torch.nn.LayerNorm, out=None):
    """``norm(x)`` over the last dimension of a 2-D tensor (in place when ``out is x``)."""
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.lib().flid_row_layernorm(_lib.ptr(x), x.stride(0), _lib.ptr(norm.weight), _lib.ptr(norm.bias), float(norm.eps),
                                             _lib.ptr(out), out.stride(0), x.shape[0], x.shape[1], _lib.stream()))
    return out


def time_rows(dt, ids, time_encoder):
    """cos(dt * w + b) rows [n, T]; rows whose id is 0 are zero when ``ids`` is given."""
    w, b = time_encoder.w.weight.reshape(-1), time_encoder.w.bias
    dt = dt.reshape(-1).contiguous()
    flat_ids = ids.reshape(-1).contiguous() if ids is not None else None
    out = torch.empty((dt.shape[0], w.shape[0]), dtype=torch.float32, device=dt.device)
    _lib.check(_lib.lib().flid_time_rows(_lib.ptr(dt), _lib.ptr(flat_ids), _lib.ptr(w), _lib.ptr(b), w.shape[0], _lib.ptr(out),
                                         dt.shape[0], _lib.stream()))
    return out


def fast_path(module: torch.nn.Module) -> bool:
    """The forward-only kernels apply when nothing needs a gradient and dropout is off (FLID_DENSE=0: measurement knob,
    keeps the torch modules)."""
    return (not module.training) and (not torch.is_grad_enabled()) and os.environ.get("FLID_DENSE", "1") != "0"
