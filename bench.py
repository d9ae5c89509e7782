#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

Workload (configs[2], the configuration the metric is quoted on): one PTCL E-step
pseudo-label pass -- TGAT (2 layers, 20 recent neighbours, d=172, T=100, 2 heads)
embeddings of both endpoints of every event of a Reddit-shape synthetic graph
(10 984 nodes / 672 447 edges), decoder MLP -> softmax/argmax, EST entropy filter over a
3-iteration probability store.  One "step" = one full pass (1 344 894 root queries).
metric = temporal embeddings (root queries) per second, whole job.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

N > 1: launched under torchrun, one rank per GPU; the graph / features / weights are
replicated, contiguous event ranges are sharded (strong scaling of one pass) and the
(label, probabilities) rows are all-gathered with NCCL.
--impl reference: the reference's CPU path (the parity-pinned oracle port; the reference
is pure Python and cannot travel to the GPU box) on the host cores, bounded sample per step.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is a CPU measurement on all host cores
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_NBR, LAYERS, HEADS, DN, DE, TD = 20, 2, 2, 172, 172, 100
METRIC = "temporal embeddings/sec (TGAT 2-layer, 20 nbrs), E-step pseudo-label pass"
UNIT = "root queries/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(scale):
    from flid_b200 import synth
    return synth.reddit_shape(seed=0, scale=scale)


def measured_tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained bf16)"


def measured_traffic_per_eval():
    """DRAM bytes per attention evaluation from the committed ncu --set full capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r1_attention_traffic.json")
    try:
        d = json.load(open(path))
        return float(d["dram_bytes_per_eval"]), d["capture"]
    except Exception:
        return None, None


def algorithmic_bytes(evals_l1, evals_l2, valid_slots, k):
    """SURVEY.md 8(d) A(k, l), with the measured number of valid (non-padded) neighbour slots
    instead of k for the gathered rows: per slot 4*dn + 4*de gathered, per evaluation 20*k index
    bytes + 16 (id, time) + self row + output row (+ raw row for the merge at layer >= 2)."""
    per_eval = 20 * k + 16 + 4 * DN + 4 * DN
    return valid_slots * (4 * DN + 4 * DE) + (evals_l1 + evals_l2) * per_eval + evals_l2 * 4 * DN


# --------------------------------------------------------------------------- CPU reference arm
def oracle_pass_rate(g, num_batches, threads, seed=2):
    """Root queries/s of the oracle port (literal reference algorithm, torch CPU fp32) on a
    bounded sample: ``num_batches`` calls of 200 events spread over the stream, + decoder/EST.
    This is the only place bench.py touches oracle/ (the cpu_baseline / --impl reference legs)."""
    from oracle import sampler as osamp, tgat as otgat, pseudo as opseudo
    torch.set_num_threads(threads)
    s = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    p = otgat.default_params(DN, DE, TD, LAYERS, HEADS, seed=seed)
    pd = opseudo.default_decoder_params(DN, 2, seed=seed)
    nf, ef = torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features)
    e = g.num_interactions
    starts = np.linspace(0, max(e - 200, 0), num_batches).astype(np.int64)
    t0 = time.perf_counter()
    roots = 0
    for lo in starts:
        sl = slice(int(lo), int(lo) + 200)
        a, b = otgat.embed_src_dst(p, nf, ef, s, g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl],
                                   LAYERS, K_NBR)
        lab, pr = opseudo.emit(pd, a)
        opseudo.entropy_filter(lab.to(torch.float32).reshape(1, -1), [pr, pr, pr], 0.9)
        roots += 2 * len(a)
    dt = time.perf_counter() - t0
    return roots / dt, dt, roots


def run_reference(args, rank):
    if rank != 0:
        return
    g = build_problem(args.scale)
    cores = os.cpu_count() or 1
    nb = args.ref_batches
    rates = []
    for _ in range(max(args.warmup, 0)):
        oracle_pass_rate(g, 1, cores)
    t_all = 0.0
    for _ in range(args.steps):
        r, dt, roots = oracle_pass_rate(g, nb, cores)
        rates.append(r)
        t_all += dt
    value = float(np.mean(rates))
    sample = f"{nb} calls of 200 events (400 root queries each) spread over the stream, per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t_all / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(g), "layers": LAYERS, "num_neighbors": K_NBR, "batch": 200,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(g):
    return (f"configs[2]: TGAT E-step pseudo-label pass + EST filter, Reddit-shape synthetic graph "
            f"({g.num_nodes} nodes / {g.num_interactions} edges, d=172), L=2, k=20")


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import flid_b200
    from flid_b200 import _lib, passes

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    g = build_problem(args.scale)
    e = g.num_interactions
    sampler = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    model = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, sampler, TD, LAYERS, HEADS, 0.1, dev).to(dev)
    model.eval()
    dec = flid_b200.MLPClassifier(DN, 0.1, 2).to(dev)
    dec.eval()

    def load_weights(seed):
        """random-init weights of the reference architecture: torch.manual_seed(seed) + default init
        (TimeEncoder keeps its fixed 1/10^linspace(0,9,T) frequencies, models/modules.py:19-21)."""
        torch.manual_seed(seed)
        with torch.no_grad():
            for mod in list(model.modules()) + list(dec.modules()):
                if isinstance(mod, (torch.nn.Linear, torch.nn.LayerNorm)) and mod is not model.time_encoder.w:
                    mod.reset_parameters()

    lo, hi, per = passes.shard_bounds(e, rank, world)
    src_h, dst_h, t_h = g.src_node_ids[lo:hi], g.dst_node_ids[lo:hi], g.node_interact_times[lo:hi]
    n_loc = hi - lo
    nodes_d = torch.from_numpy(np.concatenate([src_h, dst_h])).to(dev)
    times_d = torch.from_numpy(np.concatenate([t_h, t_h])).to(dev)

    def finish(emb_src, store_prev):
        labels, probs = flid_b200.emit_pseudo_labels(dec, emb_src)
        if world > 1:
            packed = torch.cat([labels.to(torch.float32).unsqueeze(1), probs], dim=1)
            full = passes.all_gather_rows(packed, e, per, dist)
            labels, probs = full[:, 0].to(torch.int64), full[:, 1:].contiguous()
        pseudo = labels.to(torch.float32).reshape(1, -1).contiguous()
        flid_b200.entropy_filter(pseudo, store_prev + [probs], 0.9)
        return pseudo, probs

    use_memo = args.memo == "on"
    model.set_layer_memo(False)          # the memo is driven explicitly below so that every pass pays for its build
    pass_stats = {}

    def memo_build(collect=False):
        """An E-step pass follows an M-step (new weights), so the weights are re-uploaded and the layer
        memo is rebuilt inside every timed pass: rows sharded over the ranks, all-gathered in place
        over NCCL."""
        # new weights every pass: the weight upload (float64 folds, hi/lo tiling, per-node query-fold table) is
        # redone inside the timed region as well
        model._engine.versions.clear()
        if not use_memo:
            return
        model._engine.memo.clear()
        model._engine.build_stats = [0, 0, 0] if collect else None     # summed over the pieces of a sharded build
        model.build_layer_memo(K_NBR, sharded=world > 1)
        if collect:
            pass_stats["build"] = tuple(model._engine.build_stats)
            model._engine.build_stats = None

    def step_device(store_prev, collect=False):
        with torch.no_grad():
            memo_build(collect)
            model._engine.memo_mode = use_memo
            both = model.compute_node_temporal_embeddings(nodes_d, times_d, LAYERS, K_NBR)
            if collect:
                pass_stats["embed"] = model.last_stats()
        return finish(both[:n_loc], store_prev)

    def step_e2e(store_prev):
        with torch.no_grad():
            memo_build()
            a, _ = model.compute_src_dst_node_temporal_embeddings(src_h, dst_h, t_h, K_NBR)   # host numpy in
        pseudo, probs = finish(a, store_prev)
        return _lib.to_host(pseudo, "b_pseudo"), _lib.to_host(probs, "b_probs")               # host numpy out

    # probability store of the two earlier EM iterations (weights re-seeded 0, 1), then seed 2
    store = []
    for seed in (0, 1):
        load_weights(seed)
        store.append(step_device([])[1])
    load_weights(2)
    handle = None

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device(store, collect=True)
    handle = model._engine.handles[LAYERS]
    lib = _lib.lib()

    # ---- timed region: device-resident inputs
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    _lib.check(lib.flid_tgat_profile(handle, 1))
    sync_all()
    launches0 = lib.flid_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device(store)
    ev1.record()
    sync_all()
    launches = lib.flid_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    prof_ms = (ctypes.c_double * 4)()
    prof_n = (ctypes.c_int64 * 4)()
    _lib.check(lib.flid_tgat_profile_read(handle, prof_ms, prof_n))
    _lib.check(lib.flid_tgat_profile(handle, 0))

    # ---- end-to-end: host buffers in, host labels/probs out, through the drop-in API
    step_e2e(store)
    sync_all()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e(store)
    e1.record()
    sync_all()
    wall_e2e = time.perf_counter() - t0
    ms_e2e = max(e0.elapsed_time(e1), 1000.0 * wall_e2e)     # host-side staging counts too
    clock_info = clocks.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        roots_per_step = 2 * e
        value = roots_per_step * args.steps / (ms_total / 1000.0)
        e2e_value = roots_per_step * args.steps / (ms_e2e / 1000.0)
        # roofline of the dominant kernel (attention stream) on rank 0's shard
        roots_loc = 2 * n_loc
        if use_memo:
            # memo rows (layer 1) + per root: layer 2, and layer 1 only for roots that are not graph events
            evals_l2 = roots_loc
            evals_l1 = pass_stats["build"][0] + pass_stats["embed"][0] - roots_loc
            valid = pass_stats["build"][1] + pass_stats["embed"][1]
        else:
            evals_l1, evals_l2 = roots_loc * (1 + K_NBR), roots_loc
            valid = pass_stats["embed"][1]
            assert evals_l1 + evals_l2 == pass_stats["embed"][0]
        alg = algorithmic_bytes(evals_l1, evals_l2, valid, K_NBR)                # per step, rank 0
        attn_ms, attn_n = prof_ms[2], prof_n[2]
        peak, peak_src = measured_peaks()
        achieved = (alg * args.steps / 1e9) / (attn_ms / 1000.0) if attn_ms > 0 else 0.0
        per_eval, traffic_src = measured_traffic_per_eval()
        traffic = per_eval * (evals_l1 + evals_l2) * args.steps / max(attn_n, 1) if per_eval else None
        # second-largest kernel class: the projection chain (3xTF32 tcgen05 GEMMs + LayerNorm)
        flops_eval = 2.0 * (888 * 272 + 444 * 172 + 172 * 172)          # out-projection + MergeLayer, fp32-equivalent
        chain_ms = prof_ms[3] + prof_ms[1]
        qfold_flops = 2.0 * 172 * 888 * (roots_loc if use_memo else (evals_l1 + evals_l2))
        useful_tf = ((evals_l1 + evals_l2) * flops_eval + qfold_flops) * args.steps / (chain_ms / 1000.0) / 1e12 if chain_ms > 0 else 0.0
        tpeak, tpeak_src = measured_tensor_peak()
        secondary = {"bound": "tensor", "kernel": "gemm_tc_kernel<1> chain (tcgen05 kind::tf32, 3 MMAs per product for fp32-grade accuracy) "
                     "+ ln_kernel", "achieved": useful_tf, "executed_tf32": 3.0 * useful_tf, "peak": tpeak, "unit": "TFLOP/s",
                     "frac": useful_tf / tpeak, "peak_source": tpeak_src,
                     "note": "fp32-equivalent useful FLOPs against the measured bf16 peak; the kernel is L2->SM bandwidth "
                             "bound (DESIGN.md section 4), tensor pipe 29-65 % busy by shape (profiles/r1_final_kernels_chunk606k.txt)"}
        cores = os.cpu_count() or 1
        cpu_base = None
        if world == 1:     # reported on rank 0 at N = 1 only (torchrun pins OMP to one thread per rank)
            cpu_rate, cpu_dt, cpu_roots = oracle_pass_rate(g, args.ref_batches, cores)
            cpu_base = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{args.ref_batches} calls of 200 events (400 root queries each) spread over "
                                  f"the stream, {cpu_dt:.1f} s of CPU work"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(g), "layers": LAYERS, "num_neighbors": K_NBR, "heads": HEADS,
                       "roots_per_step": roots_per_step, "parallelism": f"query-sharded x{world}, graph replicated",
                       "layer_memo": ("on: h1 of every adjacency entry evaluated once per pass (rebuilt inside each "
                                      "timed step), %d attention evaluations per step on rank 0 instead of %d"
                                      % (evals_l1 + evals_l2, roots_loc * (2 + K_NBR))) if use_memo else "off",
                       "l2_policy": "inputs larger than L2 (edge feature table %.0f MB, %.0f MB of embeddings written "
                                    "per step; L2 is 126 MB)" % (g.edge_raw_features.nbytes / 1e6,
                                                                   roots_per_step * DN * 4 / 1e6)},
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(24 * n_loc), "d2h_bytes_per_step": int(12 * e)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "attn_pk_kernel<2,3,2> (gather + time-encode + masked softmax + "
                         "weighted sum, packed fp32 pairs)", "attention_evals_per_step": int(evals_l1 + evals_l2), "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg * args.steps / max(attn_n, 1),
                         "launches": int(attn_n), "avg_launch_ms": attn_ms / max(attn_n, 1),
                         "kernel_ms_per_step": {"level_sample": prof_ms[0] / args.steps,
                                                "query_fold_gemm": prof_ms[1] / args.steps,
                                                "attention_stream": prof_ms[2] / args.steps,
                                                "out_ln_merge_chain": prof_ms[3] / args.steps},
                         "secondary": secondary},
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debugging only; 1.0 = BASELINE config)")
    ap.add_argument("--memo", default="on", choices=["on", "off"],
                    help="layer memo of the bulk pass (off = the reference's recursion, 22 evaluations per root)")
    ap.add_argument("--ref-batches", type=int, default=120, help="oracle calls of 200 events in the CPU sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
