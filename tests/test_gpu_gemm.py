"""Projection-GEMM back ends through the C ABI: tcgen05 3xTF32 vs fp32 SIMT vs float64 torch."""
import numpy as np
import pytest
import torch

from flid_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(backend, a0, idx0, a1, w, bias, m, relu):
    lib = _lib.lib()
    n = w.shape[0]
    w0, w1 = a0.shape[1], (a1.shape[1] if a1 is not None else 0)
    c = torch.full((m, n + 3), -7.0, dtype=torch.float32, device=DEV)      # ldc > n: untouched columns must survive
    with torch.cuda.device(DEV):
        _lib.check(lib.flid_debug_gemm(backend, _lib.ptr(a0), a0.stride(0), _lib.ptr(idx0), w0, _lib.ptr(a1),
                                       a1.stride(0) if a1 is not None else 0, w1, _lib.ptr(w), w.stride(0),
                                       _lib.ptr(bias), _lib.ptr(c), c.stride(0), m, n, int(relu), _lib.stream()))
        torch.cuda.synchronize()
    assert (c[:, n:] == -7.0).all()
    return c[:, :n]


CASES = [  # m, n, w0, w1, gather, bias, relu
    (300, 272, 888, 0, False, True, False),      # out-projection (+ residual_fc) shape
    (1000, 172, 272, 172, True, True, True),     # MergeLayer fc1 on [attention | raw] with gathered raw rows
    (129, 172, 172, 0, False, True, False),      # MergeLayer fc2
    (128, 888, 172, 0, True, True, False),       # query fold
    (77, 516, 616, 0, True, True, False),        # GRU input projection
    (5, 16, 4, 0, False, False, False),          # tiny / ragged
    (4096, 272, 888, 0, False, True, False),
    (300, 172, 172, 172, True, True, True),      # two segments whose boundary falls inside a 16-float K chunk
    (2500, 80, 172, 0, False, True, True),       # decoder fc1 (double-buffered accumulators, deep A ring)
    (700, 444, 172, 0, True, True, False),       # single-head query fold: one 448-column tile (two MMA column groups)
    (40000, 272, 444, 0, False, True, False),    # more work items than SMs, ragged last tile
]


@pytest.mark.parametrize("m,n,w0,w1,gather,use_bias,relu", CASES)
def test_tc_gemm_matches_simt_and_fp64(m, n, w0, w1, gather, use_bias, relu):
    g = torch.Generator(device="cpu").manual_seed(m * 1000 + n)
    rows0 = 2 * m + 3 if gather else m
    a0 = torch.randn(rows0, w0, generator=g).to(DEV)
    idx = torch.randint(0, rows0, (m,), generator=g).to(torch.int32).to(DEV) if gather else None
    a1 = torch.randn(m, w1, generator=g).to(DEV) if w1 else None
    w = (torch.randn(n, w0 + w1, generator=g) / np.sqrt(w0 + w1)).to(DEV)
    bias = torch.randn(n, generator=g).to(DEV) if use_bias else None
    a_full = a0[idx.long()] if gather else a0
    if w1:
        a_full = torch.cat([a_full, a1], dim=1)
    want = a_full.double() @ w.double().t()
    if use_bias:
        want = want + bias.double()
    if relu:
        want = want.clamp_min(0)
    simt = run(0, a0, idx, a1, w, bias, m, relu)
    tc = run(1, a0, idx, a1, w, bias, m, relu)
    scale = float(want.abs().max())
    e_simt = float((simt.double() - want).abs().max()) / scale
    e_tc = float((tc.double() - want).abs().max()) / scale
    assert e_simt < 2e-6, f"simt rel err {e_simt:.2e}"
    # the tensor core aligns/truncates its fp32 accumulation, so the error grows mildly with K
    assert e_tc < 1.5e-5, f"tcgen05 3xTF32 rel err {e_tc:.2e}"
