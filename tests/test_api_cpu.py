"""Host-side behaviour of the drop-in classes that needs no GPU."""
import numpy as np
import torch


def test_models_survive_deepcopy_without_sharing_native_state():
    """copy.deepcopy(model) (EMA / best-model snapshots in user code): parameters are copied, the native engine and
    the dense-layer weight cache of the copy start empty, the (immutable) sampler object is shared."""
    import copy

    import flid_b200
    from flid_b200.dense import DenseWeights
    rs = np.random.RandomState(0)
    nf = rs.standard_normal((6, 172)).astype(np.float32)
    ef = rs.standard_normal((9, 172)).astype(np.float32)
    sampler = flid_b200.NeighborSampler.__new__(flid_b200.NeighborSampler)      # no device: an empty shell is enough here
    for make in (lambda: flid_b200.TGAT(nf, ef, sampler, 100, 2, 2, 0.1, "cpu"),
                 lambda: flid_b200.GraphMixer(nf, ef, sampler, 100, 5, 1, device="cpu"),
                 lambda: flid_b200.TCL(nf, ef, sampler, 100, 1, 2, 6, 0.1, "cpu")):
        m = make()
        c = copy.deepcopy(m)
        assert c.neighbor_sampler is m.neighbor_sampler
        for (ka, a), (kb, b) in zip(m.state_dict().items(), c.state_dict().items()):
            assert ka == kb and torch.equal(a, b) and (a.numel() == 0 or a.data_ptr() != b.data_ptr())
        if hasattr(m, "_engine"):
            assert c._engine is not m._engine and c._engine.handles == {} and c._engine.dims == m._engine.dims
        if hasattr(m, "_dense"):
            assert isinstance(c._dense, DenseWeights) and c._dense is not m._dense and c._dense._h == {}


def test_bench_reads_the_round2_traffic_capture():
    """bench.py's roofline.traffic comes from the committed ncu --set full capture of the final kernels."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    per_eval, source = bench.measured_traffic_per_eval()
    assert 4000.0 < per_eval < 12000.0 and "round-2" in source and "r2_final_attention_pass" in source
