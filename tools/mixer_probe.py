"""Evaluation-path timing of the GraphMixer / TCL drop-ins on a Wikipedia-shaped stream: the library's own dense
kernels (csrc/dense.cu) against the same modules composed from torch CUDA ops (FLID_DENSE=0).
    python tools/mixer_probe.py [events] [batch]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flid_b200  # noqa: E402


def main():
    n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 157474
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rs = np.random.RandomState(0)
    n_users, n_items = 8227, 1000
    src = rs.randint(1, n_users + 1, n_events).astype(np.int64)
    dst = (n_users + rs.randint(1, n_items + 1, n_events)).astype(np.int64)
    ts = np.sort(rs.uniform(0, 2.6e6, n_events))
    eid = np.arange(1, n_events + 1, dtype=np.int64)
    n_nodes = n_users + n_items
    nf = np.zeros((n_nodes + 1, 172), np.float32)
    ef = rs.standard_normal((n_events + 1, 172)).astype(np.float32)
    ef[0] = 0
    dev = "cuda:0"
    s = flid_b200.NeighborSampler(None, "recent", seed=1, device=dev, _events=(src, dst, eid, ts, n_nodes))
    torch.manual_seed(0)
    models = {"GraphMixer": flid_b200.GraphMixer(nf, ef, s, 100, 20, 2, device=dev).to(dev).eval(),
              "TCL": flid_b200.TCL(nf, ef, s, 100, 2, 2, 21, 0.1, dev).to(dev).eval()}
    out = {"events": n_events, "batch": batch or n_events}
    for name, m in models.items():
        res = {}
        for mode in ("1", "0"):
            os.environ["FLID_DENSE"] = mode
            vals = []
            for it in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                with torch.no_grad():
                    if batch:
                        for lo in range(0, n_events, batch):
                            a, b = m.compute_src_dst_node_temporal_embeddings(src[lo:lo + batch], dst[lo:lo + batch],
                                                                              ts[lo:lo + batch], 20)
                    else:
                        a, b = m.compute_src_dst_node_temporal_embeddings(src, dst, ts, 20)
                torch.cuda.synchronize()
                vals.append(time.perf_counter() - t0)
                chk = float(a.double().abs().sum() + b.double().abs().sum())
            res["kernels" if mode == "1" else "torch_modules"] = {"s_per_pass": round(min(vals), 4), "checksum": chk}
        res["speedup"] = round(res["torch_modules"]["s_per_pass"] / res["kernels"]["s_per_pass"], 2)
        out[name] = res
    if os.environ.get("MIXER_PROFILE"):
        os.environ["FLID_DENSE"] = "1"
        from torch.profiler import profile, ProfilerActivity
        for name, m in models.items():
            with profile(activities=[ProfilerActivity.CUDA]) as prof, torch.no_grad():
                m.compute_src_dst_node_temporal_embeddings(src, dst, ts, 20)
                torch.cuda.synchronize()
            rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]
            out[name]["top_kernels_ms"] = [[e.key[:60], e.count, round(e.device_time_total / 1e3, 2)] for e in rows]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
