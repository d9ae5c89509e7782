"""Host-side loader of the reference's processed-dataset files (flid_b200/data.py)."""

import numpy as np
import pytest

from flid_b200 import data


def write_dataset(tmp, name, double_way):
    d = tmp / name
    d.mkdir()
    rs = np.random.RandomState(0)
    E, N = 50, 12
    u, i = rs.randint(1, 7, E), rs.randint(7, N + 1, E)
    ts = np.sort(rs.uniform(0, 1e6, E)).round(3)
    idx = np.arange(1, E + 1)
    if double_way:
        cols = "u,i,ts,label_u,label_i,idx,last_u_ts,last_i_ts"
        rows = np.stack([u, i, ts, rs.randint(0, 3, E), rs.randint(0, 3, E), idx, ts, ts], 1)
    else:
        cols = ",u,i,ts,label,idx,last_ts"
        rows = np.stack([np.arange(E), u, i, ts, rs.randint(0, 2, E), idx, ts], 1)
    np.savetxt(d / f"ml_{name}.csv", rows, delimiter=",", header=cols, comments="", fmt="%.3f")
    np.save(d / f"ml_{name}.npy", rs.standard_normal((E + 1, 20)))
    np.save(d / f"ml_{name}_node.npy", np.zeros((N + 1, 172)))
    return u, i, ts, idx


@pytest.mark.parametrize("name,double_way", [("toy", False), ("arxiv", True)])
def test_load_processed(tmp_path, name, double_way):
    u, i, ts, idx = write_dataset(tmp_path, name, double_way)
    nf, ef, full = data.load_processed(name, root=str(tmp_path))
    assert nf.shape == (13, 172) and ef.shape == (51, 172)
    assert np.all(ef[:, 20:] == 0)                      # right zero padding to the model width
    assert full.src_node_ids.dtype == np.int64 and full.node_interact_times.dtype == np.float64
    assert np.array_equal(full.src_node_ids, u) and np.array_equal(full.dst_node_ids, i)
    assert np.allclose(full.node_interact_times, ts) and np.array_equal(full.edge_ids, idx)
    assert full.num_interactions == 50 and full.num_unique_nodes == len(set(u) | set(i))
    if double_way:
        assert isinstance(full.labels, list) and len(full.labels) == 2 and len(full.labels_time) == 2
    else:
        assert full.labels.shape == (50,) and full.labels_time.shape == (50,)


def test_feature_width_check(tmp_path):
    write_dataset(tmp_path, "toy", False)
    with pytest.raises(AssertionError):
        data.load_processed("toy", root=str(tmp_path), feat_dim=100)
