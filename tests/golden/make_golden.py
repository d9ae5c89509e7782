"""Mint the golden vectors by running the REAL FLiD reference (imported from
/root/reference through ``oracle/ref_shim.py``) on the deterministic inputs of
``cases.py``.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/{sampler,tgat,tgn,pseudo,graphmixer,tcl}.npz
(``python tests/golden/make_golden.py graphmixer`` regenerates one file)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import ref_shim, tgat as otgat, tgn as otgn, pseudo as opseudo, graphmixer as omix, tcl as otcl  # noqa: E402

torch.set_num_threads(4)
ref = ref_shim.load()


def ref_sampler(src, dst, eid, ts, num_nodes):
    """get_neighbor_sampler builds lists up to max node id; pad to num_nodes by hand."""
    adj = [[] for _ in range(num_nodes + 1)]
    for s, d, e, t in zip(src, dst, eid, ts):
        adj[s].append((d, e, t))
        adj[d].append((s, e, t))
    return ref.NeighborSampler(adj_list=adj, sample_neighbor_strategy="recent", seed=1)


def strip(p):
    return {k: v for k, v in p.items() if not k.startswith("_")}


def golden_sampler():
    src, dst, eid, ts, n = cases.adversarial_events()
    s = ref_sampler(src, dst, eid, ts, n)
    nodes, times = cases.adversarial_queries()
    out = {"in_checksum": cases.checksum(src, dst, eid, ts, nodes, times)}
    for k in (1, 3, 20):
        a, b, c = s.get_historical_neighbors(nodes, times, k)
        out[f"f64_k{k}_nbr"], out[f"f64_k{k}_eid"], out[f"f64_k{k}_ts"] = a, b, c
        a, b, c = s.get_historical_neighbors(nodes, times.astype(np.float32), k)
        out[f"f32_k{k}_nbr"], out[f"f32_k{k}_eid"], out[f"f32_k{k}_ts"] = a, b, c
    nl, el, tl = s.get_multi_hop_neighbors(2, nodes[:300], times[:300], 3)
    for h in range(2):
        out[f"hop{h}_nbr"], out[f"hop{h}_eid"], out[f"hop{h}_ts"] = nl[h], el[h], tl[h]
    # get_neighbor_sampler itself (Data record) on the same stream
    data = ref.Data(src, dst, ts, eid, np.zeros(len(src)))
    s2 = ref.get_neighbor_sampler(data, "recent", seed=1)
    ok = nodes <= max(src.max(), dst.max())
    a, b, c = s2.get_historical_neighbors(nodes[ok], times[ok], 5)
    out["gns_k5_nbr"], out["gns_k5_eid"], out["gns_k5_ts"] = a, b, c
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("sampler.npz", len(nodes), "queries")


TGAT_CASES = [
    # name, L, k, heads, bias_scale, node_zeros, n_events
    ("L1_k20", 1, 20, 2, 0.0, False, 40),
    ("L2_k5", 2, 5, 2, 0.0, False, 30),
    ("L2_k20_bias", 2, 20, 2, 0.5, False, 12),
    ("L2_k7_zeros", 2, 7, 2, 0.3, True, 20),
    ("L3_k3", 3, 3, 2, 0.2, False, 10),
]


def golden_tgat():
    out = {}
    for name, L, k, heads, bias, zeros, nev in TGAT_CASES:
        src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
        n = nf.shape[0] - 1
        s = ref_sampler(src, dst, eid, ts, n)
        p = otgat.default_params(172, 172, 100, L, heads, seed=3, time_bias_scale=bias)
        m = ref.TGAT(nf, ef, s, time_feat_dim=100, num_layers=L, num_heads=heads, dropout=0.1, device="cpu")
        m.load_state_dict(strip(p))
        m.eval()
        sel = np.linspace(len(src) // 3, len(src) - 1, nev).astype(np.int64)
        with torch.no_grad():
            a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], k)
        out[name + "_src"], out[name + "_dst"], out[name + "_sel"] = a.numpy(), b.numpy(), sel
        out[name + "_checksum"] = cases.checksum(src, dst, ts, nf, ef, *[v.numpy() for v in strip(p).values()])
        print("tgat", name, a.shape, float(a.abs().mean()))
    np.savez_compressed(os.path.join(HERE, "tgat.npz"), **out)


TGN_CASES = [
    # name, L, k, batch, n_batches, bias
    ("L1_k5", 1, 5, 25, 12, 0.3),
    ("L2_k4", 2, 4, 20, 8, 0.0),
]


def golden_tgn():
    out = {}
    for name, L, k, bs, nb, bias in TGN_CASES:
        src, dst, eid, ts, nf, ef = cases.small_stream(num_nodes=30, num_edges=400, seed=11, t_max=2.0e6)
        n = nf.shape[0] - 1
        s = ref_sampler(src, dst, eid, ts, n)
        p = otgn.default_params(172, 172, 100, L, 2, seed=5, time_bias_scale=bias)
        m = ref.MemoryModel(nf, ef, s, time_feat_dim=100, model_name="TGN", num_layers=L, num_heads=2,
                            dropout=0.1, device="cpu")
        sd = strip(p)
        missing = m.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, missing
        m.eval()
        m.memory_bank.__init_memory_bank__()
        embs = []
        with torch.no_grad():
            for b in range(nb):
                lo, hi = b * bs, (b + 1) * bs
                a, c = m.compute_src_dst_node_temporal_embeddings(src[lo:hi], dst[lo:hi], ts[lo:hi], eid[lo:hi], True, k)
                embs.append(torch.cat([a, c], 0).numpy())
            # one negative-edge call (no state change) at the end
            a, c = m.compute_src_dst_node_temporal_embeddings(src[hi:hi + bs], dst[hi:hi + bs][::-1].copy(),
                                                              ts[hi:hi + bs], eid[hi:hi + bs], False, k)
            out[name + "_neg"] = torch.cat([a, c], 0).numpy()
        out[name + "_emb"] = np.stack(embs)
        out[name + "_mem"] = m.memory_bank.node_memories.data.numpy().copy()
        out[name + "_lastupd"] = m.memory_bank.node_last_updated_times.data.numpy().copy()
        # pending last message per node (what the aggregator would pick next)
        pend_ids = sorted(v for v, l in m.memory_bank.node_raw_messages.items() if len(l) > 0)
        out[name + "_pend_ids"] = np.array(pend_ids, dtype=np.int64)
        out[name + "_pend_msg"] = np.stack([m.memory_bank.node_raw_messages[v][-1][0].numpy() for v in pend_ids])
        out[name + "_pend_ts"] = np.array([m.memory_bank.node_raw_messages[v][-1][1] for v in pend_ids])
        out[name + "_checksum"] = cases.checksum(src, dst, ts, nf, ef, *[v.numpy() for v in sd.values()])
        print("tgn", name, out[name + "_emb"].shape, float(np.abs(out[name + "_emb"]).mean()))
    np.savez_compressed(os.path.join(HERE, "tgn.npz"), **out)


MIXER_CASES = [
    # name, layers, k (= num_tokens), time_gap, bias, node_zeros, n_events
    ("L2_k20_g2000", 2, 20, 2000, 0.0, False, 40),
    ("L2_k5_g7", 2, 5, 7, 0.3, False, 40),
    ("L1_k10_g50_zeros", 1, 10, 50, 0.2, True, 25),
]


def golden_graphmixer():
    """GraphMixer (models/GraphMixer.py): SURVEY 8(f) rank 4, another consumer of the sampler."""
    out = {}
    for name, L, k, gap, bias, zeros, nev in MIXER_CASES:
        src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
        s = ref_sampler(src, dst, eid, ts, nf.shape[0] - 1)
        p = omix.default_params(172, 100, k, L, seed=5, time_bias_scale=bias)
        m = ref.GraphMixer(nf, ef, s, time_feat_dim=100, num_tokens=k, num_layers=L, dropout=0.1, device="cpu")
        m.load_state_dict(p)
        m.eval()
        sel = np.linspace(0, len(src) - 1, nev).astype(np.int64)       # includes nodes without any history
        with torch.no_grad():
            a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], num_neighbors=k, time_gap=gap)
        out[name + "_src"], out[name + "_dst"], out[name + "_sel"] = a.numpy(), b.numpy(), sel
        out[name + "_checksum"] = cases.checksum(src, dst, ts, nf, *[v.numpy() for v in p.values()])
        print("graphmixer", name, a.shape, float(a.abs().mean()))
    np.savez_compressed(os.path.join(HERE, "graphmixer.npz"), **out)


TCL_CASES = [
    # name, layers, k (num_depths = k + 1), bias, node_zeros, n_events
    ("L2_k20", 2, 20, 0.0, False, 30),
    ("L1_k6_bias", 1, 6, 0.3, False, 40),
    ("L2_k4_zeros", 2, 4, 0.2, True, 25),
]


def golden_tcl():
    """TCL (models/TCL.py): SURVEY 8(f) rank 4, another consumer of the sampler."""
    out = {}
    for name, L, k, bias, zeros, nev in TCL_CASES:
        src, dst, eid, ts, nf, ef = cases.small_stream(node_zeros=zeros)
        s = ref_sampler(src, dst, eid, ts, nf.shape[0] - 1)
        p = otcl.default_params(172, 172, 100, L, k + 1, seed=6, time_bias_scale=bias)
        m = ref.TCL(nf, ef, s, time_feat_dim=100, num_layers=L, num_heads=2, num_depths=k + 1, dropout=0.1, device="cpu")
        m.load_state_dict(p)
        m.eval()
        sel = np.linspace(0, len(src) - 1, nev).astype(np.int64)
        with torch.no_grad():
            a, b = m.compute_src_dst_node_temporal_embeddings(src[sel], dst[sel], ts[sel], num_neighbors=k)
        out[name + "_src"], out[name + "_dst"], out[name + "_sel"] = a.numpy(), b.numpy(), sel
        out[name + "_checksum"] = cases.checksum(src, dst, ts, nf, ef, *[v.numpy() for v in p.values()])
        print("tcl", name, a.shape, float(a.abs().mean()))
    np.savez_compressed(os.path.join(HERE, "tcl.npz"), **out)


def golden_pseudo():
    out = {}
    rs = np.random.RandomState(21)
    for C in (2, 5):
        emb = torch.from_numpy(rs.standard_normal((700, 172)).astype(np.float32) * 2.0)
        p = opseudo.default_decoder_params(172, C, seed=C)
        dec = ref.MLPClassifier(input_dim=172, dropout=0.1, num_classes=C)
        dec.load_state_dict(p)
        dec.eval()
        with torch.no_grad():
            logits = dec(emb)
            probs = torch.softmax(logits, dim=1)
            lab = torch.max(probs, dim=1)[1]
        store = [probs]
        for it in range(2):  # two more "EM iterations" with perturbed logits
            store.append(torch.softmax(logits + torch.from_numpy(rs.standard_normal(logits.shape).astype(np.float32)), dim=1))
        out[f"C{C}_logits"], out[f"C{C}_probs"], out[f"C{C}_labels"] = logits.numpy(), probs.numpy(), lab.numpy()
        out[f"C{C}_store"] = torch.stack(store).numpy()
        for thr in (0.3, 0.6, 0.9):
            ps = lab.to(torch.float32).reshape(1, -1).clone()
            out[f"C{C}_est_{thr}"] = ref.entropy_filter(ps, store, threshold=thr).numpy().copy()
            ps = lab.to(torch.float32).reshape(1, -1).clone()
            out[f"C{C}_cst_{thr}"] = ref.prob_filter(ps, store, threshold=thr).numpy().copy()
        # double-way layout [2, E/2]
        half = 350
        ps = lab.to(torch.float32).reshape(2, half).clone()
        store2 = [s.reshape(2, half, C) for s in store]
        out[f"C{C}_est2_0.6"] = ref.entropy_filter(ps, store2, threshold=0.6).numpy().copy()
        # update_pseudo_labels (single-way, 'ps', with and without transductive mask)
        true = rs.randint(0, C, 700)
        it = np.sort(rs.uniform(0, 1000, 700)).round()
        lt = it.copy()
        lt[rs.rand(700) < 0.7] += 5.0
        data = {"full_data": ref.Data(np.ones(700, int), np.ones(700, int), it, np.arange(1, 701), true, lt),
                "val_offest": np.int64(400), "dataset_name": "wikipedia"}
        for ut in (0, 1):
            ps = lab.to(torch.float32).reshape(1, -1).clone()
            r = ref.update_pseudo_labels(data, ps, store, [], "ps", use_transductive=ut, threshold=0.6,
                                         ps_filter="entropy")
            out[f"C{C}_upd_ut{ut}"] = r.numpy().copy()
        out[f"C{C}_true"], out[f"C{C}_it"], out[f"C{C}_lt"] = true, it, lt
    np.savez_compressed(os.path.join(HERE, "pseudo.npz"), **out)
    print("pseudo.npz done")


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, fn in (("sampler", golden_sampler), ("tgat", golden_tgat), ("tgn", golden_tgn), ("pseudo", golden_pseudo),
                     ("graphmixer", golden_graphmixer), ("tcl", golden_tcl)):
        if not only or name in only:
            fn()
