"""CPU-only checks of the C-ABI boundary: the library builds/loads without a GPU, exports
every symbol that include/flid_b200.h declares, and the Python host layer refuses to run
without CUDA instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import flid_b200
from flid_b200 import _lib, passes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "flid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flid_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/flid_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes table and header disagree"
    assert lib.flid_abi_version() == 1
    assert lib.flid_launch_count() >= 0


def test_library_is_in_tree_and_sm100a():
    path = _lib.library_path()
    assert os.path.dirname(path) == os.path.join(ROOT, "flid_b200")
    assert os.path.isfile(path)


def test_argument_errors_without_gpu():
    lib = _lib.lib()
    h = ctypes.c_void_p(None)
    assert lib.flid_tgat_create(172, 172, 100, 2, 3, ctypes.byref(h)) != 0      # 272 % 3 != 0
    assert b"divided by num_heads" in lib.flid_last_error()
    assert lib.flid_tgat_create(170, 172, 100, 2, 2, ctypes.byref(h)) != 0      # rows not 16-byte multiples
    assert lib.flid_tgat_create(172, 172, 100, 2, 2, ctypes.byref(h)) == 0
    lib.flid_tgat_free(h)
    with pytest.raises(AssertionError):
        _lib.check(lib.flid_sample_recent(ctypes.c_void_p(1), None, None, 0, 4, 0, None, None, None, None))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        flid_b200.NeighborSampler([[], [(1, 1, 1.0)]], "recent")
    with pytest.raises(NotImplementedError):
        flid_b200.NeighborSampler([[]], "uniform", seed=0)
    with pytest.raises(ValueError):
        flid_b200.NeighborSampler([[]], "bogus")
    m = flid_b200.TGAT(np.zeros((4, 172), np.float32), np.zeros((4, 172), np.float32), None, 100, 2, 2, 0.1, "cpu")
    for mode in (m.eval, m.train):
        mode()
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.compute_src_dst_node_temporal_embeddings(np.array([1]), np.array([2]), np.array([1.0]), 20)


def test_sampler_consumers_have_no_cpu_fallback_either():
    nf, ef = np.zeros((4, 172), np.float32), np.zeros((4, 172), np.float32)
    for m in (flid_b200.GraphMixer(nf, ef, None, 100, 5, 1), flid_b200.TCL(nf, ef, None, 100, 1, 2, 6)):
        assert not m.time_encoder.w.weight.is_cuda
        with pytest.raises((RuntimeError, TypeError)):
            m.compute_src_dst_node_temporal_embeddings(np.array([1]), np.array([2]), np.array([1.0]), 5)


def test_state_dict_keys_match_reference_layout():
    m = flid_b200.TGAT(np.zeros((4, 172), np.float32), np.zeros((4, 172), np.float32), None, 100, 2, 2, 0.1, "cpu")
    sd = m.state_dict()
    assert sd["time_encoder.w.weight"].shape == (100, 1) and sd["time_encoder.w.bias"].shape == (100,)
    for l in range(2):
        a = f"temporal_conv_layers.{l}."
        assert sd[a + "query_projection.weight"].shape == (272, 272)
        assert sd[a + "key_projection.weight"].shape == (272, 444)
        assert sd[a + "value_projection.weight"].shape == (272, 444)
        assert sd[a + "residual_fc.weight"].shape == (272, 272) and sd[a + "residual_fc.bias"].shape == (272,)
        assert sd[a + "layer_norm.weight"].shape == (272,)
        assert sd[f"merge_layers.{l}.fc1.weight"].shape == (172, 444)
        assert sd[f"merge_layers.{l}.fc2.weight"].shape == (172, 172)
    assert len(sd) == 2 + 2 * 11
    g = flid_b200.MemoryModel(np.zeros((5, 172), np.float32), np.zeros((9, 172), np.float32), None, 100, "TGN", 1)
    sd = g.state_dict()
    assert sd["memory_bank.node_memories"].shape == (5, 172)
    assert sd["memory_updater.memory_bank.node_last_updated_times"].shape == (5,)
    assert sd["memory_updater.memory_updater.weight_ih"].shape == (516, 616)
    assert sd["memory_updater.memory_updater.weight_hh"].shape == (516, 172)
    assert "embedding_module.time_encoder.w.weight" in sd
    assert "embedding_module.temporal_conv_layers.0.key_projection.weight" in sd
    d = flid_b200.MLPClassifier(172, 0.1, 2).state_dict()
    assert d["fc1.weight"].shape == (80, 172) and d["fc2.weight"].shape == (10, 80) and d["fc3.weight"].shape == (2, 10)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present")
def test_state_dict_keys_equal_live_reference():
    from oracle import ref_shim
    ref = ref_shim.load()
    nf, ef = np.zeros((5, 172), np.float32), np.zeros((9, 172), np.float32)
    ours = flid_b200.TGAT(nf, ef, None, 100, 2, 2, 0.1, "cpu").state_dict()
    theirs = ref.TGAT(nf, ef, None, 100, 2, 2, 0.1, "cpu").state_dict()
    assert {k: tuple(v.shape) for k, v in ours.items()} == {k: tuple(v.shape) for k, v in theirs.items()}
    assert torch.equal(ours["time_encoder.w.weight"], theirs["time_encoder.w.weight"])
    ours = flid_b200.MemoryModel(nf, ef, None, 100, "TGN", 2).state_dict()
    theirs = ref.MemoryModel(nf, ef, None, 100, "TGN", 2).state_dict()
    assert {k: tuple(v.shape) for k, v in ours.items()} == {k: tuple(v.shape) for k, v in theirs.items()}
    assert {k: tuple(v.shape) for k, v in flid_b200.MLPClassifier(172).state_dict().items()} == \
           {k: tuple(v.shape) for k, v in ref.MLPClassifier(172).state_dict().items()}
    ours = flid_b200.GraphMixer(nf, ef, None, 100, 20, 2).state_dict()
    theirs = ref.GraphMixer(nf, ef, None, 100, 20, 2).state_dict()
    assert {k: tuple(v.shape) for k, v in ours.items()} == {k: tuple(v.shape) for k, v in theirs.items()}
    ours = flid_b200.TCL(nf, ef, None, 100, 2, 2, 21).state_dict()
    theirs = ref.TCL(nf, ef, None, 100, 2, 2, 21).state_dict()
    assert {k: tuple(v.shape) for k, v in ours.items()} == {k: tuple(v.shape) for k, v in theirs.items()}


def test_shard_bounds_cover_everything_in_order():
    for n in (0, 1, 7, 200, 672447):
        for w in (1, 2, 3, 4, 8):
            got, per = [], None
            for r in range(w):
                lo, hi, per = passes.shard_bounds(n, r, w)
                assert 0 <= hi - lo <= per
                got += list(range(lo, hi)) if n < 1000 else []
                if r == w - 1:
                    assert hi == n
            if n < 1000:
                assert got == list(range(n))
