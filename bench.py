#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 0..4]

Default (--config 2, the configuration the metric is quoted on): one PTCL E-step pseudo-label
pass -- TGAT (2 layers, 20 recent neighbours, d=172, T=100, 2 heads) embeddings of both
endpoints of every event of a Reddit-shape synthetic graph (10 984 nodes / 672 447 edges),
decoder MLP -> softmax/argmax, EST entropy filter over a 3-iteration probability store.
One "step" = one full pass (1 344 894 root queries).  metric = temporal embeddings (root
queries) per second, whole job.

The other BASELINE.json configs emit the same line shape:
  --config 0  TGAT E-step inference pass on the Wikipedia-shape graph (the reference's own CPU-runnable case);
              the reference's per-batch loop (788 calls of B=200) is timed beside it as a secondary field
  --config 1  TGN (MemoryModel, GRU updater, last-message aggregator), Wikipedia shape, chronological
              batches of 200, single GPU (N > 1: independent replicas, "scaling": "weak")
  --config 3  TGAT L=2 k=30 double-way pass on the Dsub-shape graph (150 000 nodes / 168 154 edges)
  --config 4  1 M nodes / 50 M edges: recent-neighbour sampling + attention aggregation (L=1) for 4 M root
              queries in chunks of 64 k (L=2 as a secondary field at N=1)

N > 1: launched under torchrun, one rank per GPU; the graph / features / weights are replicated,
root queries are sharded (strong scaling of one pass) and the (label, probabilities) rows are
gathered with NCCL.  After the timed region rank 0 repeats the pass unsharded and checks that the
gathered results are identical bit for bit ("sharded_equals_single").
--impl reference: the reference's CPU path on the host cores -- the real reference modules when a copy of
the tree is present (oracle/ref_shim.py: /root/reference, baseline/_ref/, oracle/_ref/), else the
parity-pinned oracle port -- on a bounded sample per step.
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

HEADS, DN, DE, TD = 2, 172, 172, 100
UNIT = "root queries/s"

CONFIGS = {
    0: dict(kind="tgat", shape="wikipedia", layers=2, k=20, double_way=False, batch=200,
            metric="temporal embeddings/sec (TGAT 2-layer, 20 nbrs), E-step inference pass, Wikipedia shape"),
    1: dict(kind="tgn", shape="wikipedia", layers=1, k=20, double_way=False, batch=200,
            metric="temporal embeddings/sec (TGN, GRU updater, last-message aggregator, 20 nbrs), chronological "
                   "pass in batches of 200, Wikipedia shape"),
    2: dict(kind="tgat", shape="reddit", layers=2, k=20, double_way=False, batch=200,
            metric="temporal embeddings/sec (TGAT 2-layer, 20 nbrs), E-step pseudo-label pass"),
    3: dict(kind="tgat", shape="dsub", layers=2, k=30, double_way=True, batch=200,
            metric="temporal embeddings/sec (TGAT 2-layer, 30 nbrs), double-way inference pass, Dsub shape"),
    4: dict(kind="scaling", shape="scaling", layers=1, k=20, double_way=False, batch=65536,
            metric="temporal embeddings/sec (recent-neighbour sampling + 1-layer attention aggregation, 20 nbrs), "
                   "1 M nodes / 50 M edges"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained bf16)"


def measured_traffic_per_eval():
    """DRAM bytes per attention evaluation from the committed ncu --set full capture (profiles/); only valid
    for the configuration it was captured on (configs[2], one GPU)."""
    for name in ("r2_attention_traffic.json", "r1_attention_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            return float(d["dram_bytes_per_eval"]), d["capture"]
        except Exception:
            continue
    return None, None


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


# --------------------------------------------------------------------------- workloads
def build_graph(cfg, scale):
    from flid_b200 import synth
    shape = cfg["shape"]
    if shape == "wikipedia":
        return synth.wikipedia_shape(seed=0, scale=scale)
    if shape == "reddit":
        return synth.reddit_shape(seed=0, scale=scale)
    if shape == "dsub":
        return synth.dsub_shape(seed=0, scale=scale)
    return synth.scaling_shape(seed=0, num_nodes=max(1000, int(1_000_000 * scale)),
                               num_edges=max(20000, int(50_000_000 * scale)))


def workload_name(ci, cfg, g):
    e, n = g.num_interactions, g.num_nodes
    if ci == 0:
        return (f"configs[0]: TGAT E-step inference pass (decoder + EST filter), Wikipedia-shape synthetic graph "
                f"({n} nodes / {e} edges, d=172), L=2, k=20")
    if ci == 1:
        return (f"configs[1]: TGN (MemoryModel, GRU updater, last-message aggregator) chronological pass, "
                f"Wikipedia-shape synthetic graph ({n} nodes / {e} edges, d=172), L=1, k=20, batches of 200")
    if ci == 2:
        return (f"configs[2]: TGAT E-step pseudo-label pass + EST filter, Reddit-shape synthetic graph "
                f"({n} nodes / {e} edges, d=172), L=2, k=20")
    if ci == 3:
        return (f"configs[3]: TGAT 2-layer 30-neighbor double-way inference pass (decoder + EST filter on both "
                f"endpoints), Dsub-shape synthetic graph ({n} nodes / {e} edges, 2 classes)")
    return (f"configs[4]: synthetic temporal graph {n} nodes / {e} edges, recent-neighbour sampling + 1-layer "
            f"attention aggregation, root queries from the last 10 % of time in chunks of 65 536")


def algorithmic_bytes(evals_l1, evals_up, valid_slots, k):
    """SURVEY.md 8(d) A(k, l), with the measured number of valid (non-padded) neighbour slots
    instead of k for the gathered rows: per slot 4*dn + 4*de gathered, per evaluation 20*k index
    bytes + 16 (id, time) + self row + output row (+ raw row for the merge at layer >= 2)."""
    per_eval = 20 * k + 16 + 4 * DN + 4 * DN
    return valid_slots * (4 * DN + 4 * DE) + (evals_l1 + evals_up) * per_eval + evals_up * 4 * DN


def scaling_roots(g, total, seed=5):
    """configs[4] queries: events drawn uniformly from the last 10 % of time, one endpoint each."""
    rs = np.random.RandomState(seed)
    e = g.num_interactions
    lo = int(np.searchsorted(g.node_interact_times, 0.9 * g.node_interact_times[-1]))
    ev = rs.randint(lo, e, total)
    side = rs.randint(0, 2, total).astype(bool)
    nodes = np.where(side, g.src_node_ids[ev], g.dst_node_ids[ev]).astype(np.int64)
    return nodes, g.node_interact_times[ev].astype(np.float64)


# --------------------------------------------------------------------------- CPU reference arm
def cpu_tgat_rate(cfg, g, num_batches, threads, seed=2):
    """Root queries/s of the reference's CPU path on a bounded sample: ``num_batches`` calls of 200 events
    spread over the stream, + decoder / EST filter.  The real reference modules are used when a copy of the
    tree is present on this machine (kind "reference"), else the parity-pinned oracle port (kind "port").
    This is the only place bench.py touches oracle/ (the cpu_baseline / --impl reference legs)."""
    from oracle import ref_shim
    from oracle import sampler as osamp, tgat as otgat, pseudo as opseudo
    torch.set_num_threads(threads)
    L, k, two = cfg["layers"], cfg["k"], cfg["double_way"]
    p = otgat.default_params(DN, DE, TD, L, HEADS, seed=seed)
    pd = opseudo.default_decoder_params(DN, 2, seed=seed)
    e = g.num_interactions
    starts = np.linspace(0, max(e - 200, 0), num_batches).astype(np.int64)
    kind = "port"
    if ref_shim.available():
        ref = ref_shim.load()
        data = ref.Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, np.zeros(e))
        rs = ref.get_neighbor_sampler(data, "recent", seed=1)
        model = ref.TGAT(g.node_raw_features, g.edge_raw_features, rs, TD, L, HEADS, 0.1, "cpu")
        model.load_state_dict({kk: v for kk, v in p.items() if not kk.startswith("_")})
        dec = ref.MLPClassifier(DN, 0.1, 2)
        dec.load_state_dict(pd)
        model.eval(), dec.eval()
        kind = "reference"

        def one(sl):
            with torch.no_grad():
                a, b = model.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sl], g.dst_node_ids[sl],
                                                                      g.node_interact_times[sl], k)
                x = torch.cat([a, b]) if two else a
                pr = torch.softmax(dec(x), dim=1)
                lab = pr.argmax(dim=1).to(torch.float32).reshape(2 if two else 1, -1)
                st = pr.reshape(2, -1, 2) if two else pr
                ref.entropy_filter(lab, [st, st, st], 0.9)
            return 2 * len(a)
    else:
        s = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times,
                                            g.num_nodes)
        nf, ef = torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features)

        def one(sl):
            a, b = otgat.embed_src_dst(p, nf, ef, s, g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl],
                                       L, k)
            lab, pr = opseudo.emit(pd, torch.cat([a, b]) if two else a)
            opseudo.entropy_filter(lab.to(torch.float32).reshape(1, -1), [pr, pr, pr], 0.9)
            return 2 * len(a)
    t0 = time.perf_counter()
    roots = 0
    for lo in starts:
        roots += one(slice(int(lo), int(lo) + 200))
    dt = time.perf_counter() - t0
    sample = f"{num_batches} calls of 200 events (400 root queries each) spread over the stream"
    return roots / dt, dt, sample, kind


def cpu_tgn_rate(cfg, g, num_batches, threads, seed=2):
    """configs[1] on the CPU: the first ``num_batches`` chronological batches of 200 from a reset bank."""
    from oracle import ref_shim
    from oracle import sampler as osamp, tgn as otgn
    torch.set_num_threads(threads)
    L, k = cfg["layers"], cfg["k"]
    p = otgn.default_params(DN, DE, TD, L, HEADS, seed=seed)
    e = g.num_interactions
    nb = min(num_batches, e // 200)
    kind = "port"
    if ref_shim.available():
        ref = ref_shim.load()
        data = ref.Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, np.zeros(e))
        rs = ref.get_neighbor_sampler(data, "recent", seed=1)
        model = ref.MemoryModel(g.node_raw_features, g.edge_raw_features, rs, TD, "TGN", L, HEADS, 0.1, device="cpu")
        model.load_state_dict({kk: v for kk, v in p.items() if not kk.startswith("_")}, strict=False)
        model.eval()
        model.memory_bank.__init_memory_bank__()
        kind = "reference"

        def one(sl):
            with torch.no_grad():
                model.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sl], g.dst_node_ids[sl],
                                                               g.node_interact_times[sl], g.edge_ids[sl], True, k)
    else:
        o = otgn.OracleTGN(p, torch.from_numpy(g.node_raw_features), torch.from_numpy(g.edge_raw_features),
                           osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids,
                                                           g.node_interact_times, g.num_nodes), L, k)

        def one(sl):
            o.step(g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl], g.edge_ids[sl], True)
    t0 = time.perf_counter()
    for b in range(nb):
        one(slice(b * 200, (b + 1) * 200))
    dt = time.perf_counter() - t0
    return 400 * nb / dt, dt, f"the first {nb} chronological batches of 200 events from a reset memory bank", kind


def cpu_scaling_rate(cfg, num_queries, threads, seed=2):
    """configs[4] on the CPU: the same generator law at 1/10 size (100 k nodes / 5 M edges; the numpy CSR of the
    full graph alone takes minutes), L=1, k=20, ``num_queries`` root queries in calls of 4 096."""
    from flid_b200 import synth
    from oracle import sampler as osamp, tgat as otgat
    torch.set_num_threads(threads)
    g = synth.scaling_shape(seed=0, num_nodes=100_000, num_edges=5_000_000)
    rs = np.random.RandomState(1)
    nf = torch.from_numpy(rs.standard_normal((g.num_nodes + 1, DN)).astype(np.float32))
    ef = torch.from_numpy(rs.standard_normal((g.num_interactions + 1, DE)).astype(np.float32))
    s = osamp.OracleSampler.from_events(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes)
    p = otgat.default_params(DN, DE, TD, 1, HEADS, seed=seed)
    nodes, times = scaling_roots(g, num_queries)
    t0 = time.perf_counter()
    for lo in range(0, num_queries, 4096):
        otgat.embed(p, nf, ef, s, nodes[lo:lo + 4096], times[lo:lo + 4096], 1, cfg["k"])
    dt = time.perf_counter() - t0
    return num_queries / dt, dt, (f"{num_queries} root queries in calls of 4 096 on the same generator law at 1/10 "
                                  f"size (100 000 nodes / 5 000 000 edges), L=1, k=20"), "port"


def cpu_rate(ci, cfg, g, args, threads):
    if cfg["kind"] == "tgn":
        return cpu_tgn_rate(cfg, g, args.ref_batches, threads)
    if cfg["kind"] == "scaling":
        return cpu_scaling_rate(cfg, 4096 * max(1, args.ref_batches // 8), threads)
    return cpu_tgat_rate(cfg, g, args.ref_batches, threads)


def cpu_baseline_block(ci, cfg, g, args):
    cores = os.cpu_count() or 1
    rate, dt, sample, kind = cpu_rate(ci, cfg, g, args, cores)
    out = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{sample}, {dt:.1f} s of CPU work"}
    if cores > 2:       # the reference's own thread setting (train.py:31-34: 2 intra-op threads), a smaller sample
        small = argparse.Namespace(**vars(args))
        small.ref_batches = max(8, args.ref_batches // 4)
        r2, dt2, s2, _ = cpu_rate(ci, cfg, g, small, 2)
        out["two_threads"] = {"value": r2, "cores": 2, "sample": f"{s2}, {dt2:.1f} s of CPU work"}
    torch.set_num_threads(cores)
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    ci, cfg = args.config, CONFIGS[args.config]
    g = build_graph(cfg, args.scale if cfg["kind"] != "scaling" else min(args.scale, 0.02))
    cores = os.cpu_count() or 1
    one = argparse.Namespace(**vars(args))
    one.ref_batches = 1 if cfg["kind"] != "scaling" else 8
    for _ in range(max(args.warmup, 0)):
        cpu_rate(ci, cfg, g, one, cores)
    rates, t_all, sample, kind = [], 0.0, "", "port"
    for _ in range(args.steps):
        r, dt, sample, kind = cpu_rate(ci, cfg, g, args, cores)
        rates.append(r)
        t_all += dt
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_all / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak" if cfg["kind"] == "tgn" else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(ci, cfg, g) if cfg["kind"] != "scaling"
                   else "configs[4]: synthetic temporal graph 1000000 nodes / 50000000 edges, recent-neighbour sampling + "
                        "1-layer attention aggregation, root queries from the last 10 % of time in chunks of 65 536",
                   "layers": cfg["layers"], "num_neighbors": cfg["k"], "batch": 200, "sample": sample + ", per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + ", per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
class Bench:
    def __init__(self, args, rank, world, local_rank):
        import torch.distributed as dist
        self.args, self.rank, self.world = args, rank, world
        self.dist = dist
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def sync_all(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K calls of fn bracketed by barrier + synchronize, CUDA events on the current stream; also the wall clock
        (host-side staging counts for the end-to-end number)."""
        self.sync_all()
        t0 = time.perf_counter()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record()
        for i in range(steps):
            fn()
            marks[i + 1].record()
        self.sync_all()
        self.last_step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        return marks[0].elapsed_time(marks[-1]), 1000.0 * (time.perf_counter() - t0)

    def max_over_ranks(self, *vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def load_weights(modules, skip, seed):
    """random-init weights of the reference architecture: torch.manual_seed(seed) + default init
    (TimeEncoder keeps its fixed 1/10^linspace(0,9,T) frequencies, models/modules.py:19-21)."""
    torch.manual_seed(seed)
    with torch.no_grad():
        for root in modules:
            for mod in root.modules():
                if isinstance(mod, (torch.nn.Linear, torch.nn.LayerNorm, torch.nn.GRUCell)) and mod is not skip:
                    mod.reset_parameters()


def separable_decoder(dec, emb_sample, seed=0):
    """A decoder whose EST mask is mixed on this model's embeddings: the first layer reads one random
    direction of the (standardised) embedding, the logit gap is spread so that roughly half of the events
    fall below the 0.9-bit entropy threshold.  Default-initialised decoders give entropy ~ 1 everywhere (all
    labels filtered), which exercises nothing of the filter."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        x = emb_sample.float()
        mu, sd = x.mean(0), x.std(0).clamp_min(1e-6)
        v = torch.randn(x.shape[1], generator=g).to(x.device)
        v = v / sd / v.norm()
        proj = (x - mu) @ v
        s = 1.5 / proj.std().clamp_min(1e-6)
        for lin in (dec.fc1, dec.fc2, dec.fc3):
            lin.weight.zero_(), lin.bias.zero_()
        dec.fc1.weight[0], dec.fc1.weight[1] = v * s, -v * s
        dec.fc1.bias[0], dec.fc1.bias[1] = -(mu @ v) * s, (mu @ v) * s
        dec.fc2.weight[0, 0], dec.fc2.weight[1, 1] = 1.0, 1.0
        dec.fc3.weight[0, 0], dec.fc3.weight[1, 1] = 1.0, 1.0


def run_tgat(b, ci, cfg):
    """configs[0], [2], [3]: bulk E-step pass through flid_b200.passes.e_step_pass."""
    import flid_b200
    from flid_b200 import _lib, passes
    args, rank, world, dev, dist = b.args, b.rank, b.world, b.dev, b.dist
    L, K_NBR, two = cfg["layers"], cfg["k"], cfg["double_way"]
    g = build_graph(cfg, args.scale)
    e = g.num_interactions
    sampler = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    model = flid_b200.TGAT(g.node_raw_features, g.edge_raw_features, sampler, TD, L, HEADS, 0.1, dev).to(dev)
    model.eval()
    dec = flid_b200.MLPClassifier(DN, 0.1, 2).to(dev)
    dec.eval()
    use_memo = args.memo == "on"
    model.set_layer_memo(True if use_memo else False)
    src_d = torch.from_numpy(g.src_node_ids).to(dev)
    dst_d = torch.from_numpy(g.dst_node_ids).to(dev)
    t_d = torch.from_numpy(g.node_interact_times).to(dev)
    ways = 2 if two else 1
    pass_stats = {}

    def new_weights(collect=False):
        """An E-step pass follows an M-step (new weights), so inside every timed pass the weights are
        re-uploaded (float64 folds, hi/lo tiling, per-node query-fold table) and the layer memo is rebuilt
        (rows sharded over the ranks when N > 1)."""
        model.invalidate_caches()
        model._engine.build_stats = [0, 0, 0] if collect else None

    def one_pass(src, dst, t, store_prev, sharded, collect=False):
        new_weights(collect)
        out = passes.e_step_pass(model, dec, src, dst, t, K_NBR, list(store_prev), "entropy", 0.9, sharded=sharded,
                                 double_way=two)
        if collect:
            pass_stats["build"] = tuple(model._engine.build_stats or (0, 0, 0))
            pass_stats["embed"] = model.last_stats()
            model._engine.build_stats = None
        return out

    def step_device(store_prev, collect=False):
        return one_pass(src_d, dst_d, t_d, store_prev, world > 1, collect)

    def step_e2e(store_prev):
        pseudo, probs, _ = one_pass(g.src_node_ids, g.dst_node_ids, g.node_interact_times, store_prev, world > 1)
        # host numpy out (views of the pinned staging buffers: no second host copy)
        t0 = time.perf_counter()
        out = _lib.to_host(pseudo, "b_pseudo", copy=False), _lib.to_host(probs, "b_probs", copy=False)
        passes._mark("to_host", t0)
        return out

    # probability store of the two earlier EM iterations (weights re-seeded 0, 1), then seed 2; the decoder is
    # made separable on the seed-0 embeddings so that the EST mask is mixed (asserted below)
    def seed_iteration(seed):
        load_weights([model], model.time_encoder.w, seed)
        with torch.no_grad():
            model.invalidate_caches()
            a0, _ = model.compute_src_dst_node_temporal_embeddings(g.src_node_ids[-4096:], g.dst_node_ids[-4096:],
                                                                   g.node_interact_times[-4096:], K_NBR)
        separable_decoder(dec, a0, seed)

    store = []
    for seed in (0, 1):
        seed_iteration(seed)
        store.append(step_device([])[1])
    seed_iteration(2)

    for _ in range(max(args.warmup, 3)):
        pseudo, probs, _ = step_device(store, collect=True)
    kept = float((pseudo >= 0).float().mean())
    handle = model._engine.handles[L]
    lib = _lib.lib()

    # ---- timed region: device-resident inputs
    clocks = ClockSampler(b.local_rank)
    if rank == 0:
        clocks.start()
    _lib.check(lib.flid_tgat_profile(handle, 1))
    step_device(store)         # untimed: the event timer creates its CUDA events on first use
    _lib.check(lib.flid_tgat_profile(handle, 1))
    b.sync_all()
    launches0 = lib.flid_launch_count()
    passes.trace_report()      # (FLID_PASS_TRACE) drop the warm-up phases
    ms_total, wall_dev = b.timed(lambda: step_device(store), args.steps)
    pass_trace = passes.trace_report()
    step_ms = list(b.last_step_ms)
    launches = lib.flid_launch_count() - launches0
    prof_ms = (ctypes.c_double * 4)()
    prof_n = (ctypes.c_int64 * 4)()
    _lib.check(lib.flid_tgat_profile_read(handle, prof_ms, prof_n))
    _lib.check(lib.flid_tgat_profile(handle, 0))

    # ---- end-to-end: host buffers in, host labels/probs out, through the public pass API
    step_e2e(store)
    passes.trace_report()
    ev_ms, wall_ms = b.timed(lambda: step_e2e(store), args.steps)
    e2e_trace = passes.trace_report()
    e2e_trace["event_ms_total"], e2e_trace["wall_ms_total"] = ev_ms, wall_ms
    ms_e2e = max(ev_ms, wall_ms)
    clock_info = clocks.stop() if rank == 0 else None
    ms_total, ms_e2e = b.max_over_ranks(ms_total, ms_e2e)

    if os.environ.get("FLID_BENCH_TIMELINE") and rank == 0:
        dump_timeline(lambda: step_device(store), os.environ["FLID_BENCH_TIMELINE"], world)
    elif os.environ.get("FLID_BENCH_TIMELINE"):
        step_device(store)      # the other ranks take part in the pass's collectives
        step_device(store)

    # ---- outside the timed region: sharded == single, bit for bit
    same = None
    if world > 1:
        p_sh, pr_sh, _ = step_device(store)
        if rank == 0:
            p_1, pr_1, _ = one_pass(src_d, dst_d, t_d, store, False)
            same = bool(torch.equal(p_sh, p_1) and torch.equal(pr_sh, pr_1))
        b.sync_all()

    # ---- secondary numbers (N = 1 only, outside the timed region)
    secondary_modes = {}
    if world == 1 and not args.no_secondary:
        secondary_modes = tgat_secondary(b, ci, cfg, g, model, dec, store, step_device, pseudo, probs)

    if rank != 0:
        return
    roots_per_step = 2 * e
    value = roots_per_step * args.steps / (ms_total / 1000.0)
    e2e_value = roots_per_step * args.steps / (ms_e2e / 1000.0)
    lo, hi, _ = passes.shard_bounds(e, rank, world)
    roots_loc = int(pass_stats["embed"][2])        # root queries answered by rank 0 (the roots of the nodes it owns)
    top_evals = pass_stats["embed"][0]
    build_evals = pass_stats["build"][0]
    if use_memo:
        evals_up = roots_loc                                     # layer L of every local root
        evals_l1 = build_evals + top_evals - roots_loc           # memo rows + lower layers of roots not in the memo
        valid = pass_stats["build"][1] + pass_stats["embed"][1]
    else:
        evals_up, evals_l1 = roots_loc, top_evals - roots_loc
        valid = pass_stats["embed"][1]
    alg = algorithmic_bytes(evals_l1, evals_up, valid, K_NBR)                # per step, rank 0
    attn_ms, attn_n = prof_ms[2], prof_n[2]
    peak, peak_src = measured_peaks()
    achieved = (alg * args.steps / 1e9) / (attn_ms / 1000.0) if attn_ms > 0 else 0.0
    per_eval, traffic_src = measured_traffic_per_eval()
    traffic = None
    if per_eval and ci == 2 and world == 1:
        traffic = per_eval * (evals_l1 + evals_up) * args.steps / max(attn_n, 1)
    chain_ms = prof_ms[3] + prof_ms[1]
    gemm_flops = pass_gemm_flops(model, evals_l1 + evals_up, roots_loc)
    useful_tf = gemm_flops * args.steps / (chain_ms / 1000.0) / 1e12 if chain_ms > 0 else 0.0
    tpeak, tpeak_src = measured_tensor_peak()
    secondary = {"bound": "tensor", "kernel": "projection chain: tcgen05 kind::tf32 GEMMs (3 MMAs per product for "
                 "fp32-grade accuracy) + LayerNorm", "achieved": useful_tf, "executed_tf32": 3.0 * useful_tf,
                 "peak": tpeak, "unit": "TFLOP/s", "frac": useful_tf / tpeak, "peak_source": tpeak_src,
                 "frac_of_tf32_peak": 3.0 * useful_tf / (tpeak / 2.0),
                 "note": "fp32-equivalent useful FLOPs against the measured bf16 peak; frac_of_tf32_peak relates the "
                         "executed tf32 MMAs to half the bf16 peak (the dense tf32 rate)"}
    cpu_base = cpu_baseline_block(ci, cfg, g, args) if world == 1 else None
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(ci, cfg, g), "layers": L, "num_neighbors": K_NBR, "heads": HEADS,
                   "roots_per_step": roots_per_step, "parallelism": (f"owner-partitioned x{world} (queries routed to the rank that owns their node, memo rows "
                                   f"exchanged once per level), graph replicated") if world > 1 else "one GPU",
                   "layer_memo": ("on: lower layers of every adjacency entry evaluated once per pass (rebuilt inside "
                                  "each timed step), %d attention evaluations per step on rank 0 instead of %d"
                                  % (evals_l1 + evals_up, roots_loc * sum((1 + K_NBR) ** i for i in range(L))))
                   if use_memo else "off",
                   "est_kept_fraction": kept, "step_ms": step_ms, "wall_ms_per_step": wall_dev / args.steps,
                   "l2_policy": "inputs larger than L2 (edge feature table %.0f MB, %.0f MB of embeddings written "
                                "per step; L2 is 126 MB)" % (g.edge_raw_features.nbytes / 1e6,
                                                               roots_per_step * DN * 4 / 1e6)},
        "clocks": clock_info,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(24 * (hi - lo)), "d2h_bytes_per_step": int(12 * e * ways)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "attention stream (gather + time-encode + masked softmax + weighted "
                     "sum, packed fp32 pairs)", "attention_evals_per_step": int(evals_l1 + evals_up),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src if traffic else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg * args.steps / max(attn_n, 1),
                     "launches": int(attn_n), "avg_launch_ms": attn_ms / max(attn_n, 1),
                     "kernel_ms_per_step": {"level_sample": prof_ms[0] / args.steps,
                                            "query_side_gemm": prof_ms[1] / args.steps,
                                            "attention_stream": prof_ms[2] / args.steps,
                                            "out_ln_merge_chain": prof_ms[3] / args.steps},
                     "secondary": secondary},
        "cpu_baseline": cpu_base,
    }
    if same is not None:
        line["sharded_equals_single"] = same
        plans = list(model._engine.shard_plans.values())
        line["config"]["memo_row_exchange"] = ("one kernel storing into the peers' tables over NVLink (CUDA IPC)"
                                               if plans and plans[0].p2p else "all_to_all_single (NCCL)")
    if os.environ.get("FLID_PASS_TRACE") == "1":
        line["pass_trace_ms_per_step"] = {k: v / args.steps for k, v in pass_trace.items()}
        line["e2e_trace_ms_per_step"] = {k: v / args.steps for k, v in e2e_trace.items()}
    line.update(secondary_modes)
    print(json.dumps(line), flush=True)
    assert 0.05 < kept < 0.95, f"EST mask is degenerate: kept fraction {kept}"
    assert same is not False, "sharded pass differs from the single-GPU pass"


def pass_gemm_flops(model, evals, roots):
    """fp32-equivalent FLOPs of the projection GEMMs of one pass on this rank (folded formulation)."""
    per_eval = 2.0 * (888 * 272 + 444 * 172 + 172 * 172)      # out-projection + MergeLayer
    return evals * per_eval + 2.0 * 172 * 888 * roots         # + query fold of the top-layer targets


def tgat_secondary(b, ci, cfg, g, model, dec, store, step_device, pseudo_f32, probs_f32):
    """Secondary fields of the TGAT configs at N = 1: the bf16-projection numeric mode (north_star: rel 2e-2,
    identical argmax on >= 99.9 % of nodes) and the reference's own per-batch loop (B = 200 calls)."""
    import flid_b200
    out = {}
    args = b.args
    K_NBR = cfg["k"]
    e = g.num_interactions
    if hasattr(flid_b200, "set_numeric_mode"):
        try:
            flid_b200.set_numeric_mode("bf16")
            from flid_b200 import _lib
            lib, handle = _lib.lib(), model._engine.handles[model.num_layers]
            for _ in range(3):
                ps_b, pr_b, _ = step_device(store)
            _lib.check(lib.flid_tgat_profile(handle, 1))
            step_device(store)
            _lib.check(lib.flid_tgat_profile(handle, 1))
            ms, _ = b.timed(lambda: step_device(store), args.steps)
            pm, pn = (ctypes.c_double * 4)(), (ctypes.c_int64 * 4)()
            _lib.check(lib.flid_tgat_profile_read(handle, pm, pn))
            _lib.check(lib.flid_tgat_profile(handle, 0))
            agree = float((pr_b.argmax(-1) == probs_f32.argmax(-1)).float().mean())
            out["bf16_projections"] = {"value": 2 * e * args.steps / (ms / 1000.0), "unit": UNIT,
                                       "ms_per_step": ms / args.steps, "step_ms": list(b.last_step_ms),
                                       "kernel_ms_per_step": {"level_sample": pm[0] / args.steps,
                                                              "query_side_gemm": pm[1] / args.steps,
                                                              "attention_stream": pm[2] / args.steps,
                                                              "out_ln_merge_chain": pm[3] / args.steps},
                                       "argmax_agreement_with_f32": agree,
                                       "max_abs_prob_diff": float((pr_b - probs_f32).abs().max()),
                                       "note": "projection GEMMs with bf16-rounded operands and fp32 accumulation "
                                               "(one MMA per product); the headline stays f32"}
        finally:
            flid_b200.set_numeric_mode("f32")
    # the unchanged callers' loop: E/200 calls of compute_src_dst_node_temporal_embeddings with host batches
    nb = min(-(-e // 200), 1000)
    model.set_layer_memo("auto")
    model.invalidate_caches()
    with torch.no_grad():
        model.build_layer_memo(K_NBR)
        for i in range(20):
            sl = slice(i * 200, (i + 1) * 200)
            model.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sl], g.dst_node_ids[sl],
                                                           g.node_interact_times[sl], K_NBR)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(nb):
            sl = slice(i * 200, (i + 1) * 200)
            model.compute_src_dst_node_temporal_embeddings(g.src_node_ids[sl], g.dst_node_ids[sl],
                                                           g.node_interact_times[sl], K_NBR)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out["per_batch_loop"] = {"calls": nb, "batch": 200, "us_per_call": 1e6 * dt / nb,
                             "value": 400 * nb / dt, "unit": UNIT,
                             "note": "compute_src_dst_node_temporal_embeddings per batch of 200 events with host "
                                     "numpy inputs (the reference's own loop shape), layer memo built once"}
    model.set_layer_memo(True)
    return out


def run_tgn(b, ci, cfg):
    """configs[1]: chronological TGN pass in batches of 200 (sequential memory updates: one GPU per replica)."""
    import flid_b200
    from flid_b200 import _lib, passes
    args, rank, world, dev = b.args, b.rank, b.world, b.dev
    L, K_NBR = cfg["layers"], cfg["k"]
    g = build_graph(cfg, args.scale)
    e = g.num_interactions
    sampler = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    model = flid_b200.MemoryModel(g.node_raw_features, g.edge_raw_features, sampler, TD, "TGN", L, HEADS, 0.1,
                                  device=dev).to(dev)
    load_weights([model], model.time_encoder.w, 2)
    model.eval()
    lib = _lib.lib()

    def step(per_batch_calls=False):
        return passes.tgn_pass(model, g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, cfg["batch"],
                               K_NBR, per_batch_calls=per_batch_calls)

    def step_e2e():
        a, c = step()
        return _lib.to_host(a, "b_tgn_src", copy=False), _lib.to_host(c, "b_tgn_dst", copy=False)

    for _ in range(max(args.warmup, 3)):
        step()
    handle = model._engine.handles[L]
    clocks = ClockSampler(b.local_rank)
    if rank == 0:
        clocks.start()
    launches0 = lib.flid_launch_count()
    ms_total, _ = b.timed(step, args.steps)
    launches = lib.flid_launch_count() - launches0
    # per-kernel-class times from one extra pass with the event timer on (direct launches: the timer's events
    # cannot be read back from a captured graph), outside the timed region
    _lib.check(lib.flid_tgat_profile(handle, 1))
    step()
    prof_ms = (ctypes.c_double * 4)()
    prof_n = (ctypes.c_int64 * 4)()
    _lib.check(lib.flid_tgat_profile_read(handle, prof_ms, prof_n))
    _lib.check(lib.flid_tgat_profile(handle, 0))
    for i in range(4):
        prof_ms[i] *= args.steps      # the report below divides by the number of timed steps
        prof_n[i] *= args.steps
    step_e2e()
    ev_ms, wall_ms = b.timed(step_e2e, args.steps)
    ms_e2e = max(ev_ms, wall_ms)
    clock_info = clocks.stop() if rank == 0 else None
    ms_total, ms_e2e = b.max_over_ranks(ms_total, ms_e2e)
    loop_ms = None
    if world == 1 and not args.no_secondary:      # the unchanged callers' loop: one Python call per batch of 200
        step(True)
        loop_ms, _ = b.timed(lambda: step(True), 1)
    if rank != 0:
        return
    nb = -(-e // cfg["batch"])
    roots = 2 * e * world                      # replicas only: every rank runs its own full pass
    value = roots * args.steps / (ms_total / 1000.0)
    # valid slots per pass are not tracked per batch; the stream reads at most k rows per evaluation
    evals = 2 * e * sum((1 + K_NBR) ** i for i in range(L))
    alg = algorithmic_bytes(evals - 2 * e if L > 1 else evals, 2 * e if L > 1 else 0, evals * K_NBR, K_NBR)
    attn_ms, attn_n = prof_ms[2], prof_n[2]
    peak, peak_src = measured_peaks()
    achieved = (alg * args.steps / 1e9) / (attn_ms / 1000.0) if attn_ms > 0 else 0.0
    cpu_base = cpu_baseline_block(ci, cfg, g, args) if world == 1 else None
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(ci, cfg, g), "layers": L, "num_neighbors": K_NBR, "heads": HEADS,
                   "batch": cfg["batch"], "batches_per_step": nb, "us_per_batch": 1000.0 * ms_total / args.steps / nb,
                   "driver": "flid_tgn_pass: one C call per pass, the launch sequence of a batch captured as a CUDA graph "
                             "and replayed with a device-side batch counter",
                   "roots_per_step": roots,
                   "parallelism": "replicas only (memory updates are a chain over batches)" if world > 1 else "one GPU",
                   "l2_policy": "inputs larger than L2 are not possible at this shape (edge table %.0f MB); every batch "
                                "rewrites the bank rows and the 2 x B x 172 outputs" % (g.edge_raw_features.nbytes / 1e6)},
        "clocks": clock_info,
        "e2e": {"value": roots * args.steps / (ms_e2e / 1000.0), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(32 * e), "d2h_bytes_per_step": int(2 * e * DN * 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "attention stream of the per-batch embedding (400 targets per launch: "
                     "latency-bound, not bandwidth-bound)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg * args.steps / max(attn_n, 1), "launches": int(attn_n),
                     "avg_launch_ms": attn_ms / max(attn_n, 1),
                     "note": "upper bound on bytes: k rows per evaluation assumed valid",
                     "kernel_ms_per_step": {"level_sample": prof_ms[0] / args.steps,
                                            "query_side_gemm": prof_ms[1] / args.steps,
                                            "attention_stream": prof_ms[2] / args.steps,
                                            "out_ln_merge_chain": prof_ms[3] / args.steps}},
        "cpu_baseline": cpu_base,
    }
    if loop_ms is not None:
        line["per_batch_loop"] = {"ms_per_pass": loop_ms, "us_per_batch": 1000.0 * loop_ms / nb,
                                  "value": 2 * e / (loop_ms / 1000.0), "unit": UNIT,
                                  "note": "compute_src_dst_node_temporal_embeddings called per batch of 200 from Python "
                                          "(the reference's own loop shape): same kernels, ~25 launches and one "
                                          "synchronisation per batch"}
    print(json.dumps(line), flush=True)


def dump_timeline(step, out_dir, world):
    """Development aid (FLID_BENCH_TIMELINE=<dir>): one extra pass under torch.profiler after the timed region;
    writes the device timeline summary of rank 0 -- kernel busy time, span, the idle gaps and what surrounds them."""
    from torch.profiler import profile, ProfilerActivity
    step()                      # re-primes the caches keyed by the inputs (the e2e steps used host arrays)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ev = [x for x in prof.events() if str(x.device_type).endswith("CUDA") and x.time_range.end > x.time_range.start]
    ev.sort(key=lambda x: x.time_range.start)
    if not ev:
        return
    t0, t1 = ev[0].time_range.start, max(x.time_range.end for x in ev)
    busy, cur_end, gaps = 0.0, t0, []
    for i, x in enumerate(ev):
        a, z = x.time_range.start, x.time_range.end
        if a > cur_end:
            if a - cur_end >= 15.0:
                gaps.append([round(a - cur_end, 1), round(cur_end - t0, 1), ev[i - 1].name[:48], x.name[:48]])
            busy += z - a
            cur_end = z
        elif z > cur_end:
            busy += z - cur_end
            cur_end = z
    by = {}
    for x in ev:
        k = x.name[:60]
        by[k] = [by.get(k, [0, 0])[0] + 1, by.get(k, [0, 0])[1] + (x.time_range.end - x.time_range.start)]
    top = sorted(([k, v[0], round(v[1], 1)] for k, v in by.items()), key=lambda r: -r[2])
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"timeline_N{world}.json"), "w") as f:
        json.dump({"span_us": round(t1 - t0, 1), "busy_us": round(busy, 1), "kernels": len(ev),
                   "gaps_ge_15us": sorted(gaps, key=lambda r: -r[0])[:60], "gap_total_us": round(sum(r[0] for r in gaps), 1),
                   "by_kernel_us": top[:70],
                   "sequence": [[round(x.time_range.start - t0, 1), round(x.time_range.end - x.time_range.start, 1), x.name[:44]]
                                for x in ev]}, f, indent=0)


def run_scaling(b, ci, cfg):
    """configs[4]: 1 M nodes / 50 M edges, feature tables generated on the device, 4 M root queries."""
    import flid_b200
    from flid_b200 import _lib, passes
    args, rank, world, dev, dist = b.args, b.rank, b.world, b.dev, b.dist
    K_NBR, chunk = cfg["k"], cfg["batch"]
    g = build_graph(cfg, args.scale)
    e, n_nodes = g.num_interactions, g.num_nodes
    sampler = flid_b200.get_neighbor_sampler(g, "recent", seed=1, device=dev)
    models = {}
    gen = torch.Generator(device=dev).manual_seed(1)
    node_feat = torch.randn((n_nodes + 1, DN), device=dev, generator=gen)
    edge_feat = torch.randn((e + 1, DE), device=dev, generator=gen)
    node_feat[0] = 0
    edge_feat[0] = 0

    def make_model(L):
        m = flid_b200.TGAT(np.zeros((2, DN), np.float32), np.zeros((2, DE), np.float32), sampler, TD, L, HEADS, 0.0,
                           dev).to(dev)
        m.node_raw_features, m.edge_raw_features = node_feat, edge_feat
        load_weights([m], m.time_encoder.w, 2)
        m.eval()
        return m

    models[1] = make_model(1)
    total = max(chunk, int(4_000_000 * args.scale))
    nodes_h, times_h = scaling_roots(g, total)
    lo, hi, per = passes.shard_bounds(total, rank, world)
    nodes_d = torch.from_numpy(nodes_h[lo:hi]).to(dev)
    times_d = torch.from_numpy(times_h[lo:hi]).to(dev)
    n_loc = hi - lo
    out = torch.empty((n_loc, DN), dtype=torch.float32, device=dev)
    stats = {}

    def step_device(L=1, collect=False):
        m = models[L]
        ev = vs = 0
        with torch.no_grad():
            for c0 in range(0, n_loc, chunk):
                out[c0:c0 + chunk] = m.compute_node_temporal_embeddings(nodes_d[c0:c0 + chunk], times_d[c0:c0 + chunk],
                                                                        L, K_NBR)
                if collect:
                    st = m.last_stats()
                    ev, vs = ev + st[0], vs + st[1]
        if collect:
            stats[L] = (ev, vs)

    host_out = torch.empty((n_loc, DN), dtype=torch.float32, pin_memory=True)

    def step_e2e():
        m = models[1]
        with torch.no_grad():
            for c0 in range(0, n_loc, chunk):
                sl = slice(lo + c0, min(lo + c0 + chunk, hi))
                r = m.compute_node_temporal_embeddings(nodes_h[sl], times_h[sl], 1, K_NBR)     # host numpy in
                host_out[c0:c0 + r.shape[0]].copy_(r, non_blocking=True)                       # host out
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device(1, collect=True)
    handle = models[1]._engine.handles[1]
    lib = _lib.lib()
    clocks = ClockSampler(b.local_rank)
    if rank == 0:
        clocks.start()
    _lib.check(lib.flid_tgat_profile(handle, 1))
    launches0 = lib.flid_launch_count()
    ms_total, _ = b.timed(lambda: step_device(1), args.steps)
    launches = lib.flid_launch_count() - launches0
    prof_ms = (ctypes.c_double * 4)()
    prof_n = (ctypes.c_int64 * 4)()
    _lib.check(lib.flid_tgat_profile_read(handle, prof_ms, prof_n))
    _lib.check(lib.flid_tgat_profile(handle, 0))
    step_e2e()
    ev_ms, wall_ms = b.timed(step_e2e, args.steps)
    ms_e2e = max(ev_ms, wall_ms)
    clock_info = clocks.stop() if rank == 0 else None
    ms_total, ms_e2e = b.max_over_ranks(ms_total, ms_e2e)

    same = None
    if world > 1:        # sharded == single on the first chunk of every rank, gathered on rank 0
        head = out[:min(chunk, n_loc)].contiguous()
        pieces = [torch.empty_like(head) for _ in range(world)] if rank == 0 else None
        dist.gather(head, pieces, dst=0)
        if rank == 0:
            same = True
            with torch.no_grad():
                for r in range(world):
                    rlo = passes.shard_bounds(total, r, world)[0]
                    sl = slice(rlo, rlo + head.shape[0])
                    ref = models[1].compute_node_temporal_embeddings(nodes_h[sl], times_h[sl], 1, K_NBR)
                    same = same and bool(torch.equal(ref, pieces[r]))
        b.sync_all()

    two_layer = None
    if world == 1 and not args.no_secondary:
        free = torch.cuda.mem_get_info(dev)[0]
        need = (sampler.num_entries + 1) * DN * 4 * 1.15 + 12e9
        if free > need:
            models[2] = make_model(2)
            models[2].set_layer_memo(True)
            with torch.no_grad():
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                models[2].build_layer_memo(K_NBR)
                torch.cuda.synchronize()
                t_build = time.perf_counter() - t0
            step_device(2)
            ms2, _ = b.timed(lambda: step_device(2), 1)
            two_layer = {"memo_build_ms": 1000.0 * t_build, "memo_rows": sampler.num_entries + 1,
                         "memo_evals_per_s": (sampler.num_entries + 1) / t_build,
                         "roots_ms": ms2, "value": total / (ms2 / 1000.0), "unit": UNIT,
                         "note": "L=2 with the layer memo: one-off build over every adjacency entry, then the same "
                                 "4 M root queries"}
    if rank != 0:
        return
    value = total * args.steps / (ms_total / 1000.0)
    ev, vs = stats[1]
    alg = algorithmic_bytes(ev, 0, vs, K_NBR)
    attn_ms, attn_n = prof_ms[2], prof_n[2]
    peak, peak_src = measured_peaks()
    achieved = (alg * args.steps / 1e9) / (attn_ms / 1000.0) if attn_ms > 0 else 0.0
    cpu_base = cpu_baseline_block(ci, cfg, g, args) if world == 1 else None
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(ci, cfg, g), "layers": 1, "num_neighbors": K_NBR, "heads": HEADS,
                   "roots_per_step": total, "chunk": chunk,
                   "parallelism": f"query-sharded x{world}, graph replicated ({(edge_feat.numel() + node_feat.numel()) * 4 / 2**30:.1f} GiB of features per GPU)",
                   "l2_policy": "inputs larger than L2 (edge feature table %.1f GB; L2 is 126 MB)" % (edge_feat.numel() * 4 / 1e9)},
        "clocks": clock_info,
        "e2e": {"value": total * args.steps / (ms_e2e / 1000.0), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(16 * n_loc), "d2h_bytes_per_step": int(n_loc * DN * 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "attention stream (gather + time-encode + masked softmax + weighted sum)",
                     "attention_evals_per_step": int(ev), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg * args.steps / max(attn_n, 1), "launches": int(attn_n),
                     "avg_launch_ms": attn_ms / max(attn_n, 1),
                     "kernel_ms_per_step": {"level_sample": prof_ms[0] / args.steps,
                                            "query_side_gemm": prof_ms[1] / args.steps,
                                            "attention_stream": prof_ms[2] / args.steps,
                                            "out_ln_merge_chain": prof_ms[3] / args.steps},
                     "sampler": {"bound": "hbm", "kernel": "level_sample_kernel", "bytes_per_query": 800,
                                 "achieved": (total / world * 800 * args.steps / 1e9) / (prof_ms[0] / 1000.0) if prof_ms[0] > 0 else 0.0,
                                 "unit": "GB/s"}},
        "cpu_baseline": cpu_base,
    }
    if two_layer is not None:
        line["two_layer"] = two_layer
    if same is not None:
        line["sharded_equals_single"] = same
    print(json.dumps(line), flush=True)
    assert same is not False, "sharded pass differs from the single-GPU pass"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json configs[i]; 2 (default) is the configuration the metric is quoted on")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debugging only; 1.0 = BASELINE config)")
    ap.add_argument("--memo", default="on", choices=["on", "off"],
                    help="layer memo of the bulk pass (off = the reference's recursion, 22 evaluations per root)")
    ap.add_argument("--ref-batches", type=int, default=120, help="oracle calls of 200 events in the CPU sample")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary fields (bf16 mode, per-batch loop)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    cfg = CONFIGS[args.config]
    b = Bench(args, rank, world, local_rank)
    {"tgat": run_tgat, "tgn": run_tgn, "scaling": run_scaling}[cfg["kind"]](b, args.config, cfg)
    b.finish()


if __name__ == "__main__":
    main()
