"""TEST INFRASTRUCTURE ONLY -- import shim for the *real* FLiD reference tree.

The reference cannot be imported as shipped: ``utils/DataLoader.py:239`` is a
SyntaxError and ``evaluate_models_utils.py:12`` imports a module that is not in
the tree.  The hot-path modules only need
``utils.DataLoader.Data`` (``utils/utils.py:6``), so we register a stub package
``utils`` whose ``__path__`` points at the reference and a stub
``utils.DataLoader`` that holds a 7-field record with the field names of
``utils/DataLoader.py:46-65``.

The reference tree lives in the build container (``/root/reference``); on another
machine a copy is looked for under ``baseline/_ref/`` and ``oracle/_ref/`` (never committed).  This shim is used by ``tests/golden/make_golden.py``
(to mint golden vectors) and by the ``not gpu`` tests that pin the oracle against
the reference when the tree is present.  Nothing in ``flid_b200/`` imports it.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the build container's read-only tree, then the places a copy may have been put for the GPU box
# (both git-ignored, neither gpurun-ignored); nested one level deep as a pip --target / unpacked archive would be
REFERENCE_ROOTS = ("/root/reference", os.path.join(_REPO, "baseline", "_ref"), os.path.join(_REPO, "oracle", "_ref"))


def reference_root():
    for root in REFERENCE_ROOTS:
        cands = [root]
        if os.path.isdir(root):
            cands += [os.path.join(root, d) for d in sorted(os.listdir(root)) if os.path.isdir(os.path.join(root, d))]
        for c in cands:
            if os.path.isfile(os.path.join(c, "models", "TGAT.py")) and os.path.isfile(os.path.join(c, "utils", "utils.py")):
                return c
    return None


def available() -> bool:
    return reference_root() is not None


class _Data:
    """Field-for-field stand-in for utils/DataLoader.py:46-65."""

    def __init__(self, src_node_ids, dst_node_ids, node_interact_times, edge_ids, labels, labels_time=None):
        self.src_node_ids = src_node_ids
        self.dst_node_ids = dst_node_ids
        self.node_interact_times = node_interact_times
        self.edge_ids = edge_ids
        self.labels = labels
        self.labels_time = labels_time
        self.num_interactions = len(src_node_ids)
        self.unique_node_ids = set(src_node_ids) | set(dst_node_ids)
        self.num_unique_nodes = len(self.unique_node_ids)


_loaded = None


def load():
    """Return a namespace with the reference's hot-path classes (or raise)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("FLiD reference tree not present on this machine")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if root not in sys.path:
        sys.path.insert(0, root)
    if "utils" in sys.modules and not getattr(sys.modules["utils"], "_flid_shim", False):
        raise RuntimeError("a different top-level 'utils' module is already imported")
    pkg = types.ModuleType("utils")
    pkg.__path__ = [os.path.join(root, "utils")]
    pkg._flid_shim = True
    sys.modules["utils"] = pkg
    dl = types.ModuleType("utils.DataLoader")
    dl.Data = _Data
    sys.modules["utils.DataLoader"] = dl

    ns = types.SimpleNamespace()
    from utils.utils import NeighborSampler, get_neighbor_sampler  # noqa: E402
    from models.modules import TimeEncoder, MultiHeadAttention, MergeLayer, MLPClassifier  # noqa: E402
    from models.TGAT import TGAT  # noqa: E402
    from models.MemoryModel import MemoryModel  # noqa: E402
    from models.GraphMixer import GraphMixer  # noqa: E402
    from models.TCL import TCL  # noqa: E402
    from PTCL.utils import entropy_filter, prob_filter, update_pseudo_labels  # noqa: E402

    ns.Data = _Data
    ns.NeighborSampler = NeighborSampler
    ns.get_neighbor_sampler = get_neighbor_sampler
    ns.TimeEncoder = TimeEncoder
    ns.MultiHeadAttention = MultiHeadAttention
    ns.MergeLayer = MergeLayer
    ns.MLPClassifier = MLPClassifier
    ns.TGAT = TGAT
    ns.MemoryModel = MemoryModel
    ns.GraphMixer = GraphMixer
    ns.TCL = TCL
    ns.entropy_filter = entropy_filter
    ns.prob_filter = prob_filter
    ns.update_pseudo_labels = update_pseudo_labels
    _loaded = ns
    return ns
