"""The on-disk formats on either side of the layer: the citation edge list the (commented)
Cora loader reads and the cached dense CBG adjacency of `utils.load_adj`.  Host-side parsing only; the graphs and
products themselves are built by the device paths (`Graph.from_edges`, `functional.load_adj`).

  * `<dataset>.cites` -- one link per line, `<ID of cited paper> <ID of citing paper>` (data/cora/README:25-29), read by
    `np.genfromtxt(..., dtype=np.int32)` and mapped to node indices through the order of the `.content` file
    (pygcn/utils.py:354-359); entry (edges[:, 0], edges[:, 1]) = 1 of the raw adjacency (utils.py:360-362).
  * `adj_<msa>.npy` / `avg_array_<msa>.npy` / `<msa>/<full name>_2020-03-01_to_2020-05-02.pkl` -- the three levels of
    `utils.load_adj` (pygcn/utils.py:93-132): the cached dense adjacency, the cached hourly average of the POI x CBG
    visit matrices, the pickled list of scipy sparse matrices.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

CBG_PICKLE_SUFFIX = "_2020-03-01_to_2020-05-02.pkl"  # pygcn/utils.py:100


def read_cites(path, ids=None):
    """Parse a `.cites` edge list -> (edges int32 [E, 2] of node indices, ids int64 [N]).

    `ids` is the node order: the first column of the `.content` file in the reference's loader (pygcn/utils.py:354-355
    builds `idx_map = {paper id: position}` from it).  None -> the sorted unique paper ids of the edge list itself
    (the fork ships `cora.cites` without `cora.content`).  A paper id missing from `ids` makes the
    reference's `idx_map.get` return None and its int32 conversion fail (utils.py:358-359); here it is a ValueError."""
    raw = np.loadtxt(path, dtype=np.int64, ndmin=2)
    if raw.size == 0:
        raw = raw.reshape(0, 2)
    if raw.shape[1] != 2:
        raise ValueError("%s: expected two paper ids per line, found %d columns" % (path, raw.shape[1]))
    if ids is None:
        ids = np.unique(raw)
        edges = np.searchsorted(ids, raw)
    else:
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        order = np.argsort(ids, kind="stable")
        sorted_ids = ids[order]
        if sorted_ids.size > 1 and (sorted_ids[1:] == sorted_ids[:-1]).any():
            # a dict built by enumerate keeps the LAST position of a repeated id (utils.py:355)
            keep = np.ones(sorted_ids.size, dtype=bool)
            keep[:-1] = sorted_ids[1:] != sorted_ids[:-1]
            sorted_ids, order = sorted_ids[keep], order[keep]
        pos = np.searchsorted(sorted_ids, raw)
        pos_c = np.minimum(pos, max(sorted_ids.size - 1, 0))
        if sorted_ids.size == 0 or (sorted_ids[pos_c] != raw).any():
            missing = raw[(sorted_ids[pos_c] != raw)] if sorted_ids.size else raw
            raise ValueError("%s: paper id %d is not in the node list" % (path, int(missing.flat[0])))
        edges = order[pos_c]
    return edges.astype(np.int32).reshape(-1, 2), ids


def graph_from_cites(path, device, ids=None):
    """`.cites` file -> the normalised adjacency the Cora loader builds (pygcn/utils.py:356-376: coo of ones at
    (edges[:, 0], edges[:, 1]), max-symmetrised, + I, row-normalised) as a device `Graph`; returns (graph, ids)."""
    from .graph import Graph

    edges, ids = read_cites(path, ids)
    dev = torch.device(device)
    src = torch.from_numpy(np.ascontiguousarray(edges[:, 0])).to(dev)
    dst = torch.from_numpy(np.ascontiguousarray(edges[:, 1])).to(dev)
    return Graph.from_edges(src, dst, int(ids.shape[0])), ids


def average_visits(poi_cbg_visits_list):
    """pygcn/utils.py:116-120: the mean of the hourly POI x CBG visit matrices (scipy sparse or dense), accumulated in
    float64 in list order."""
    if len(poi_cbg_visits_list) == 0:
        raise ValueError("empty visit list")
    avg = np.zeros(poi_cbg_visits_list[0].shape, dtype=np.float64)
    for m in poi_cbg_visits_list:
        avg += m.toarray() if hasattr(m, "toarray") else np.asarray(m)
    avg /= len(poi_cbg_visits_list)
    return avg


def load_adj_files(msa_name, mob_data_root, output_root, device, msa_name_full=None, product=None, save=False):
    """`utils.load_adj(msa_name, mob_data_root, output_root)` (pygcn/utils.py:93-132) with its three file levels, the
    O(N^2 M) double loop replaced by the device product: returns (adj fp32 [n_cbg, n_cbg] on `device`, n_cbg).

      1. `<output_root>/adj_<msa>.npy` exists          -> load it (utils.py:94-97), `torch.FloatTensor(adj)` cast;
      2. else `<output_root>/avg_array_<msa>.npy`      -> adj = avg^T avg on the device (functional.load_adj);
      3. else `<mob_data_root>/<msa>/<full name>_2020-03-01_to_2020-05-02.pkl` -> average (utils.py:116-120), saved as
         avg_array_<msa>.npy like utils.py:121, then 2.
    save=True (off by default) also writes the results back: avg_array_<msa>.npy like utils.py:121, and the adjacency under
    its OWN name adj_<msa>.gcnb200.npy -- never the reference's cache adj_<msa>.npy, which the reference trusts as its
    fp64 double-loop result (this product is an fp32 / 3xTF32 device product, ~1e-6 from it and not bit-symmetric).  `msa_name_full` replaces the reference's constants.MSA_NAME_FULL_DICT lookup (utils.py:99; that
    table lives outside this path).  `product(avg fp32 device tensor) -> adj` defaults to functional.load_adj."""
    dev = torch.device(device)
    adj_path = os.path.join(output_root, "adj_%s.npy" % msa_name)
    if os.path.exists(adj_path):
        adj = np.load(adj_path)
        if adj.ndim != 2 or adj.shape[0] != adj.shape[1]:
            raise ValueError("%s: expected a square matrix, found shape %s" % (adj_path, adj.shape))
        return torch.from_numpy(np.ascontiguousarray(adj, dtype=np.float32)).to(dev), int(adj.shape[0])
    avg_path = os.path.join(output_root, "avg_array_%s.npy" % msa_name)
    if os.path.exists(avg_path):
        avg = np.asarray(np.load(avg_path), dtype=np.float64)
    else:
        if msa_name_full is None:
            raise ValueError("msa_name_full is needed to locate the visit pickle (constants.MSA_NAME_FULL_DICT[%r] in the "
                             "reference, utils.py:99)" % msa_name)
        with open(os.path.join(mob_data_root, msa_name, msa_name_full + CBG_PICKLE_SUFFIX), "rb") as f:
            visits = pickle.load(f)
        avg = average_visits(visits)
        if save:
            np.save(avg_path, avg)
    if avg.ndim != 2:
        raise ValueError("%s: expected a [n_poi, n_cbg] matrix, found shape %s" % (avg_path, avg.shape))
    if product is None:
        from .functional import load_adj as product
    adj = product(torch.from_numpy(np.ascontiguousarray(avg, dtype=np.float32)).to(dev))
    if save:
        np.save(os.path.join(output_root, "adj_%s.gcnb200.npy" % msa_name), adj.detach().cpu().numpy().astype(np.float64))
    return adj, int(avg.shape[1])
