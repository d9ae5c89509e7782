#!/usr/bin/env python
"""Time the tcgen05 3xTF32 GEMM on the four shapes of the TGAT projection chain.

    python tools/gemm_probe.py [--m 65536] [--reps 20]

Prints per shape: device time per launch, tensor-pipe utilisation against the tf32 MMA floor
(3 MMAs x N/2 cycles per 8 K-elements per 128-row tile at the measured SM clock) and the A-operand
bandwidth.  Development tool (not part of the product path or of bench.py).
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flid_b200 import _lib  # noqa: E402

SHAPES = [  # name, N, w0, w1, gather1, relu
    ("out_proj  K=888 N=272", 272, 888, 0, False, False),
    ("merge_fc1 K=444 N=172", 172, 272, 172, True, True),
    ("merge_fc2 K=172 N=172", 172, 172, 0, False, False),
    ("qfold     K=172 N=888", 888, 172, 0, False, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.lib()
    m = args.m
    sm_clock = 1.92e9
    for i, (name, n, w0, w1, gather, relu) in enumerate(SHAPES):
        if args.only >= 0 and i != args.only:
            continue
        a0 = torch.randn(m, w0, device=dev)
        a1 = torch.randn(20000, w1, device=dev) if w1 else None
        idx1 = torch.randint(0, 20000, (m,), device=dev, dtype=torch.int32) if w1 else None
        # the hook gathers segment 0 only; emulate the fc1 layout with a pre-gathered second segment
        a1g = a1[idx1.long()].contiguous() if w1 else None
        w = torch.randn(n, w0 + w1, device=dev) / (w0 + w1) ** 0.5
        bias = torch.randn(n, device=dev)
        c = torch.empty(m, n, device=dev)
        ms = C.c_float(0)
        with torch.cuda.device(dev):
            _lib.check(lib.flid_debug_gemm_time(_lib.ptr(a0), a0.stride(0), None, w0, _lib.ptr(a1g),
                                                a1g.stride(0) if w1 else 0, w1, _lib.ptr(w), w.stride(0),
                                                _lib.ptr(bias), _lib.ptr(c), c.stride(0), m, n, int(relu), args.reps,
                                                C.byref(ms), _lib.stream()))
        # correctness on a sample of rows (the timing hook leaves the last launch's output in c)
        torch.cuda.synchronize()
        rows = torch.randint(0, m, (2048,), device=dev)
        a_full = a0[rows] if not w1 else torch.cat([a0[rows], a1g[rows]], dim=1)
        want = a_full.double() @ w.double().t() + bias.double()
        if relu:
            want = want.clamp_min(0)
        err = float((c[rows].double() - want).abs().max()) / float(want.abs().max())
        k = w0 + w1
        n_pad = -(-n // 16) * 16
        floor_cycles = (-(-m // 128)) * (-(-k // 8)) * 3 * (n_pad / 2) / 148
        floor_ms = floor_cycles / sm_clock * 1e3
        gbs = m * k * 4 / (ms.value * 1e-3) / 1e9
        print(f"{name}  M={m}: {ms.value * 1e3:8.1f} us   mma-floor {floor_ms * 1e3:6.1f} us ({100 * floor_ms / ms.value:4.1f}% of tensor peak)"
              f"   A-read {gbs:7.0f} GB/s   rel.err {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
