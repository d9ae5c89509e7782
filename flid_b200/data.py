"""Processed-dataset files of the reference -> the arrays the hot path consumes.

Mirrors the loading half of ``get_PTCL_data`` / ``get_link_prediction_data``
(utils/DataLoader.py:229-272): ``ml_{name}.csv`` (columns ``u, i, ts, idx`` plus ``label, last_ts``
or, for the double-way datasets, ``label_u, label_i, last_u_ts, last_i_ts``), ``ml_{name}.npy``
(edge features ``[E+1, de]``, row 0 = padding) and ``ml_{name}_node.npy`` (``[N+1, dn]``), with the
reference's zero padding of both tables to the model width (172; 384 for ``oag``).  The
train / val / test splitting and the EM bookkeeping of the reference stay in the reference's
drivers; this module only produces what ``get_neighbor_sampler`` and the models take.
Host-side only (pandas parses the csv, as in the reference).
"""
import os

import numpy as np
import pandas as pd

DOUBLE_WAY_DATASETS = ('arxiv', 'oag')           # utils/DataLoader.py:262


class Data:
    """The reference's interaction record (utils/DataLoader.py:46-65), same attribute names."""

    def __init__(self, src_node_ids, dst_node_ids, node_interact_times, edge_ids, labels, labels_time=None):
        self.src_node_ids = src_node_ids
        self.dst_node_ids = dst_node_ids
        self.node_interact_times = node_interact_times
        self.edge_ids = edge_ids
        self.labels = labels
        self.labels_time = labels_time
        self.num_interactions = len(src_node_ids)
        self.unique_node_ids = set(src_node_ids.tolist()) | set(dst_node_ids.tolist())
        self.num_unique_nodes = len(self.unique_node_ids)


def pad_features(table, width, what, dataset_name):
    """Zero-pad a feature table on the right to ``width`` columns (utils/DataLoader.py:245-256)."""
    assert width >= table.shape[1], f'{what} feature dimension in dataset {dataset_name} is bigger than {width}!'
    if table.shape[1] < width:
        table = np.concatenate([table, np.zeros((table.shape[0], width - table.shape[1]))], axis=1)
    return table


def load_processed(dataset_name, root='./processed_data', feat_dim=None):
    """-> (node_raw_features [N+1, d], edge_raw_features [E+1, d], full_data: Data)."""
    base = os.path.join(root, dataset_name)
    csv_path = os.path.join(base, f'ml_{dataset_name}.csv')
    frame = pd.read_csv(csv_path)
    names = set(frame.columns)
    table = {c: frame[c].values for c in frame.columns}
    edge_raw_features = np.load(os.path.join(base, f'ml_{dataset_name}.npy'))
    node_raw_features = np.load(os.path.join(base, f'ml_{dataset_name}_node.npy'))
    if feat_dim is None:
        feat_dim = 384 if dataset_name == 'oag' else 172          # utils/DataLoader.py:236-244
    node_raw_features = pad_features(node_raw_features, feat_dim, 'Node', dataset_name)
    edge_raw_features = pad_features(edge_raw_features, feat_dim, 'Edge', dataset_name)
    src = table['u'].astype(np.int64)
    dst = table['i'].astype(np.int64)
    ts = table['ts'].astype(np.float64)
    eid = table['idx'].astype(np.int64)
    if dataset_name in DOUBLE_WAY_DATASETS and 'label_u' in names:
        labels = [table['label_u'], table['label_i']]
        labels_time = [table['last_u_ts'], table['last_i_ts']]
    else:
        labels = table['label']
        labels_time = table['last_ts'] if 'last_ts' in names else None
    return node_raw_features, edge_raw_features, Data(src, dst, ts, eid, labels, labels_time)
