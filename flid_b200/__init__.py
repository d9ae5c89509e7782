"""flid_b200 -- B200-native (sm_100a) temporal-embedding hot path of FLiD behind FLiD's own
Python API: ``NeighborSampler`` / ``get_neighbor_sampler``, ``TGAT``, ``MemoryModel('TGN')``,
``MLPClassifier`` + pseudo-label filters, and ``GraphMixer`` / ``TCL`` as further consumers of the sampler.  All compute goes through the C-ABI CUDA library
``libflid_b200.so`` (include/flid_b200.h); there is no CPU fallback.
"""
from .sampler import NeighborSampler, get_neighbor_sampler  # noqa: F401
from .tgat import TGAT, TimeEncoder, MultiHeadAttention, MergeLayer, set_numeric_mode, get_numeric_mode  # noqa: F401
from .memory_model import MemoryModel, MemoryBank  # noqa: F401
from .graphmixer import GraphMixer  # noqa: F401
from .tcl import TCL  # noqa: F401
from .pseudo_label import (MLPClassifier, emit_pseudo_labels, entropy_filter, prob_filter,  # noqa: F401
                           update_pseudo_labels)

__all__ = ["NeighborSampler", "get_neighbor_sampler", "TGAT", "MemoryModel", "MemoryBank", "GraphMixer", "TCL", "MLPClassifier",
           "emit_pseudo_labels", "entropy_filter", "prob_filter", "update_pseudo_labels", "set_numeric_mode",
           "get_numeric_mode"]
