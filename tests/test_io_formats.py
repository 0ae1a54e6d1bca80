"""The on-disk formats either side of the layer (pygcn_b200/io.py): the `.cites` edge list of the
(commented) Cora loader and the three file levels of `utils.load_adj`.  Host logic only -- the device product is
replaced by a numpy stand-in here; `tests/test_gpu_optin.py` runs the same calls on the GPU."""
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import gcn_oracle as O
from pygcn_b200 import io as IO

CORA = "/root/reference/data/cora/cora.cites"


@pytest.mark.skipif(not os.path.exists(CORA), reason="the reference tree is not mounted here (GPU box)")
def test_read_cites_gives_the_golden_cora_edges(golden):
    """cora.cites through read_cites == the edge array the reference-generated fixture was built from."""
    g = golden("cora_pipeline.npz")
    edges, ids = IO.read_cites(CORA)
    assert edges.dtype == np.int32 and ids.shape[0] == int(g["n"]) == 2708
    assert np.array_equal(edges, g["edges"])
    idx, val = O.build_normalized_adjacency(edges[:, 0], edges[:, 1], ids.shape[0])
    assert np.array_equal(idx, g["indices"]) and np.array_equal(val, g["values"])


def test_read_cites_maps_ids_like_the_reference_loader(tmp_path):
    """pygcn/utils.py:354-359: idx_map = {id: position in the .content order}; edges = map(idx_map.get, ...)."""
    p = tmp_path / "toy.cites"
    p.write_text("35\t1033\n35\t103482\n1033 35\n40 40\n")
    edges, ids = IO.read_cites(str(p))
    assert ids.tolist() == [35, 40, 1033, 103482]
    assert edges.tolist() == [[0, 2], [0, 3], [2, 0], [1, 1]]
    content_order = np.array([103482, 7, 35, 1033, 40])          # a node without links, another order
    idx_map = {j: i for i, j in enumerate(content_order.tolist())}  # the reference's line
    raw = np.genfromtxt(str(p), dtype=np.int32)
    want = np.array(list(map(idx_map.get, raw.flatten())), dtype=np.int32).reshape(raw.shape)
    edges, ids = IO.read_cites(str(p), content_order)
    assert np.array_equal(edges, want) and np.array_equal(ids, content_order)
    repeated = np.array([35, 40, 1033, 103482, 35])                # enumerate keeps the last position of a repeated id
    idx_map = {j: i for i, j in enumerate(repeated.tolist())}
    want = np.array(list(map(idx_map.get, raw.flatten())), dtype=np.int32).reshape(raw.shape)
    assert np.array_equal(IO.read_cites(str(p), repeated)[0], want)
    with pytest.raises(ValueError):
        IO.read_cites(str(p), np.array([35, 40, 1033]))
    bad = tmp_path / "bad.cites"
    bad.write_text("1 2 3\n")
    with pytest.raises(ValueError):
        IO.read_cites(str(bad))
    one = tmp_path / "one.cites"
    one.write_text("5 9\n")
    edges, ids = IO.read_cites(str(one))
    assert edges.tolist() == [[0, 1]] and ids.tolist() == [5, 9]


def test_load_adj_files_walks_the_three_levels_of_utils_load_adj(tmp_path, golden):
    """pygcn/utils.py:93-132 on the reference-generated fixture: pickle -> avg_array npy -> adj npy, each level found on
    the next call; the result is the reference's own adjacency (<= 1e-5) and its float32 cast."""
    c = golden("load_adj.npz")
    hours = [sp.csr_matrix(h) for h in c["hours"]]
    root, out = tmp_path / "mob", tmp_path / "out"
    (root / "SanFrancisco").mkdir(parents=True)
    out.mkdir()
    full = "San_Francisco_Oakland_Hayward_CA"
    with open(root / "SanFrancisco" / (full + IO.CBG_PICKLE_SUFFIX), "wb") as f:
        pickle.dump(hours, f)
    calls = []

    def product(avg):  # stand-in for the device product (functional.load_adj), float64 like the reference's loop
        calls.append(tuple(avg.shape))
        a = avg.double().numpy()
        return torch.from_numpy((a.T @ a).astype(np.float32))

    with pytest.raises(ValueError):
        IO.load_adj_files("SanFrancisco", str(root), str(out), "cpu", product=product)  # no full name for the pickle
    adj0, _ = IO.load_adj_files("SanFrancisco", str(root), str(out), "cpu", msa_name_full=full, product=product)
    assert sorted(os.listdir(out)) == []  # the default leaves the reference's cache directory untouched
    adj, n = IO.load_adj_files("SanFrancisco", str(root), str(out), "cpu", msa_name_full=full, product=product, save=True)
    assert torch.equal(adj, adj0)
    assert n == c["adj"].shape[0] and adj.dtype == torch.float32 and calls == [tuple(c["hours"].shape[1:])] * 2
    assert O.normwise_err(adj.numpy(), c["adj"]) < 1e-5
    avg_o, _ = O.cbg_adjacency(c["hours"])
    assert np.array_equal(np.load(out / "avg_array_SanFrancisco.npy"), avg_o)       # utils.py:116-121, float64
    # the device product is saved under its own name: adj_<msa>.npy stays the reference's fp64 double-loop cache
    assert np.load(out / "adj_SanFrancisco.gcnb200.npy").dtype == np.float64 and not os.path.exists(out / "adj_SanFrancisco.npy")
    np.save(out / "adj_SanFrancisco.npy", c["adj"])                                 # (what the reference itself writes, utils.py:129)
    adj1, n1 = IO.load_adj_files("SanFrancisco", "/nonexistent", str(out), "cpu", product=product)  # level 1: cache hit
    assert n1 == n and O.normwise_err(adj1.numpy(), c["adj"]) < 1e-6 and len(calls) == 2
    os.remove(out / "adj_SanFrancisco.npy")
    adj2, _ = IO.load_adj_files("SanFrancisco", "/nonexistent", str(out), "cpu", product=product, save=False)  # level 2
    assert torch.equal(adj2, adj) and len(calls) == 3 and not os.path.exists(out / "adj_SanFrancisco.npy")
    assert np.array_equal(IO.average_visits(hours), avg_o)
    with pytest.raises(ValueError):
        IO.average_visits([])


def test_read_cites_random_files_against_the_loader_lines(tmp_path):
    """Seeded sweep: random paper ids, link lists and `.content` orders through read_cites == the loader's own
    statements (pygcn/utils.py:354-359) executed on the same file."""
    rs = np.random.default_rng(11)
    for case in range(25):
        n = int(rs.integers(1, 40))
        ids = rs.choice(np.arange(1, 10 ** 6), size=n, replace=False)
        e = int(rs.integers(1, 120))
        raw = ids[rs.integers(0, n, size=(e, 2))]
        p = tmp_path / ("c%d.cites" % case)
        np.savetxt(p, raw, fmt="%d", delimiter="\t" if case % 2 else " ")
        order = rs.permutation(ids)
        if case % 3 == 0:  # nodes that appear in no link
            order = np.concatenate([order, np.arange(10 ** 6 + 1, 10 ** 6 + 4)])
        idx_map = {j: i for i, j in enumerate(order.tolist())}
        edges_unordered = np.genfromtxt(str(p), dtype=np.int32).reshape(-1, 2)
        want = np.array(list(map(idx_map.get, edges_unordered.flatten())), dtype=np.int32).reshape(edges_unordered.shape)
        got, back = IO.read_cites(str(p), order)
        assert np.array_equal(got, want) and np.array_equal(back, order)
        got_default, ids_default = IO.read_cites(str(p))
        assert np.array_equal(ids_default[got_default], raw)  # the default order is the sorted ids that occur
        assert np.array_equal(ids_default, np.unique(raw))
