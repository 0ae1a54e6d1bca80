#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own code --
    /root/reference/pygcn/layers.py   (GraphConvolution, the hot path)
    /root/reference/pygcn/utils.py    (normalize, sparse_mx_to_torch_sparse_tensor)
-- runs it on seeded inputs with the installed torch (CPU) / scipy, and stores
inputs (or their seeds) and the reference's outputs as small .npz files.  The
Cora loader body of utils.py:348-382 is inside a string literal upstream and
cannot be imported; its three graph-building statements (utils.py:360-368) are
executed here with scipy exactly as written there, feeding the imported
``normalize`` / ``sparse_mx_to_torch_sparse_tensor``.

Nothing here is used at test time except the .npz files it writes.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load_reference():
    # utils.py imports matplotlib (absent here) and `constants` from ../gt-generator
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, os.path.join(REF, "gt-generator"))
    sys.path.insert(0, os.path.join(REF, "pygcn"))

    def by_path(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    layers = by_path("layers", os.path.join(REF, "pygcn", "layers.py"))
    utils = by_path("utils", os.path.join(REF, "pygcn", "utils.py"))
    models = by_path("models", os.path.join(REF, "pygcn", "models.py"))
    return layers, utils, models


def rng_inputs(seed, shape):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def ref_pipeline(utils, edges, n):
    """utils.py:360-368 as written, then the imported normalize / to-torch functions."""
    adj = sp.coo_matrix((np.ones(edges.shape[0]), (edges[:, 0], edges[:, 1])), shape=(n, n), dtype=np.float32)
    adj = adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)
    adj = utils.normalize(adj + sp.eye(adj.shape[0]))
    return utils.sparse_mx_to_torch_sparse_tensor(adj)


def run_layer(layers, fin, fout, bias, x, adj, g, seed=42, x_requires_grad=True):
    torch.manual_seed(seed)
    gc = layers.GraphConvolution(fin, fout, bias=bias)
    xt = torch.from_numpy(x).clone().requires_grad_(x_requires_grad)
    out = gc(xt, adj)
    out.backward(torch.from_numpy(g))
    res = {
        "weight": gc.weight.detach().numpy().copy(),
        "out": out.detach().numpy().copy(),
        "dW": gc.weight.grad.numpy().copy(),
    }
    if bias:
        res["bias"] = gc.bias.detach().numpy().copy()
        res["db"] = gc.bias.grad.numpy().copy()
    if x_requires_grad:
        res["dX"] = xt.grad.numpy().copy()
    return res


def make_apply_bn(models):
    """apply_bn.npz: the reference's own `GCN.apply_bn` (pygcn/models.py:41-45) as the models call it,
    `apply_bn(F.relu(y))` (models.py:49,53), forward and autograd backward, on CPU.  The method moves its fresh
    BatchNorm1d to the GPU with `.cuda()`; there is none here, so `nn.Module.cuda` is a no-op while it runs (a shim in
    this harness, the reference's lines execute unchanged)."""
    import torch.nn.functional as F

    cases = {}
    real_cuda = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, device=None: self
    try:
        for name, (n, f, seed, relu, offset) in {
            "cbg32": (1100, 32, 71, True, 0.0),        # hidden width of the CBG models, float4 path
            "odd7": (257, 7, 72, True, 0.0),           # 7 classes: scalar path, ragged row count
            "plain16": (400, 16, 73, False, 2.5),      # apply_bn alone on a shifted input (mean >> 0)
            "deadcol": (300, 8, 74, True, 0.0),        # a column that ReLU zeroes entirely (var = 0 -> rstd = eps^-1/2)
            "two_rows": (2, 4, 75, False, 0.0),        # the smallest batch torch accepts
        }.items():
            y = rng_inputs(seed, (n, f)) + np.float32(offset)
            if name == "deadcol":
                y[:, 3] = -np.abs(y[:, 3]) - 1.0
            g = rng_inputs(seed + 100, (n, f))
            yt = torch.from_numpy(y.copy()).requires_grad_(True)
            out = models.GCN.apply_bn(None, F.relu(yt) if relu else yt)
            out.backward(torch.from_numpy(g))
            cases[name + "/y"] = y
            cases[name + "/g"] = g
            cases[name + "/relu"] = np.int64(relu)
            cases[name + "/out"] = out.detach().numpy().copy()
            cases[name + "/dy"] = yt.grad.numpy().copy()
    finally:
        torch.nn.Module.cuda = real_cuda
    np.savez_compressed(os.path.join(OUT, "apply_bn.npz"), **cases)


def main():
    torch.set_num_threads(1)
    layers, utils, models = _load_reference()
    if sys.argv[1:] == ["apply_bn"]:  # only this fixture (the others are unchanged)
        make_apply_bn(models)
        print("wrote apply_bn.npz", os.path.getsize(os.path.join(OUT, "apply_bn.npz")) // 1024, "KiB")
        return
    make_apply_bn(models)

    # ---------------------------------------------------------------- Cora graph pipeline
    raw = np.loadtxt(os.path.join(REF, "data", "cora", "cora.cites"), dtype=np.int64)
    ids = np.unique(raw)  # cora.content is absent: node order = sorted paper id
    edges = np.searchsorted(ids, raw).astype(np.int32)
    n = ids.shape[0]
    adj = ref_pipeline(utils, edges, n)
    idx = adj._indices().numpy().copy()
    val = adj._values().numpy().copy()
    assert idx.dtype == np.int64 and val.dtype == np.float32
    np.savez_compressed(
        os.path.join(OUT, "cora_pipeline.npz"),
        edges=edges, n=np.int64(n), indices=idx, values=val,
        is_coalesced=np.bool_(adj.is_coalesced()),
    )
    print("cora: n", n, "nnz", val.shape[0])

    # small synthetic edge lists through the same pipeline (duplicates, self edges, isolated nodes)
    pipe = {}
    for k, (nn, ne, seed) in enumerate([(37, 120, 3), (257, 900, 4), (64, 10, 5)]):
        r = np.random.default_rng(seed)
        e = r.integers(0, nn - 3, size=(ne, 2)).astype(np.int32)  # last 3 nodes isolated
        e = np.concatenate([e, e[: ne // 5]], axis=0)  # duplicate edges
        a = ref_pipeline(utils, e, nn)
        pipe[f"edges{k}"] = e
        pipe[f"n{k}"] = np.int64(nn)
        pipe[f"indices{k}"] = a._indices().numpy().copy()
        pipe[f"values{k}"] = a._values().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "pipeline_small.npz"), **pipe)

    # ---------------------------------------------------------------- parameter init (layers.py:23-29)
    init = {}
    torch.manual_seed(42)
    gc = layers.GraphConvolution(64, 32)
    init["w_64_32"] = gc.weight.detach().numpy().copy()
    init["b_64_32"] = gc.bias.detach().numpy().copy()
    torch.manual_seed(42)
    g1 = layers.GraphConvolution(8, 32)
    g2 = layers.GraphConvolution(32, 32)
    g3 = layers.GraphConvolution(32, 32, bias=False)
    for i, gg in enumerate((g1, g2, g3)):
        init[f"stack_w{i}"] = gg.weight.detach().numpy().copy()
        if gg.bias is not None:
            init[f"stack_b{i}"] = gg.bias.detach().numpy().copy()
    init["repr"] = np.array(repr(gc))
    init["state_keys"] = np.array(sorted(gc.state_dict().keys()))
    np.savez_compressed(os.path.join(OUT, "init.npz"), **init)

    # ---------------------------------------------------------------- layer fwd/bwd cases
    cases = {}

    def put(name, d):
        for k, v in d.items():
            cases[f"{name}/{k}"] = v

    # (1) Cora L1 (true shape 1433->16) and L2 (16->7); X regenerated from its seed at test time
    x = rng_inputs(1, (n, 1433))
    g = rng_inputs(2, (n, 16))
    r = run_layer(layers, 1433, 16, True, x, adj, g)
    r["x_checksum"] = np.float64(x.astype(np.float64).sum())
    r["dX_head"] = r.pop("dX")[:32].copy()  # keep the fixture small
    put("cora_l1", r)
    x = rng_inputs(3, (n, 16))
    g = rng_inputs(4, (n, 7))
    put("cora_l2", run_layer(layers, 16, 7, True, x, adj, g))

    # (2) ragged random COO with duplicates and empty rows, odd widths, non-contiguous X
    rs = np.random.default_rng(11)
    nn, nnz = 300, 3000
    rows = rs.integers(0, nn, nnz)
    rows[rows % 17 == 0] = 1  # empty rows + one long row
    cols = rs.integers(0, nn, nnz)
    vals = rs.standard_normal(nnz).astype(np.float32)
    a2 = torch.sparse_coo_tensor(np.vstack([rows, cols]), vals, (nn, nn))  # unsorted, uncoalesced, duplicates
    xw = rng_inputs(12, (nn, 40))
    g = rng_inputs(13, (nn, 7))
    torch.manual_seed(42)
    gc = layers.GraphConvolution(33, 7)
    xt = torch.from_numpy(xw).clone().requires_grad_(True)
    out = gc(xt[:, :33], a2)  # column-slice view, as models.py:345 does
    out.backward(torch.from_numpy(g))
    put("ragged", dict(rows=rows, cols=cols, vals=vals, n=np.int64(nn), weight=gc.weight.detach().numpy().copy(),
                       bias=gc.bias.detach().numpy().copy(), out=out.detach().numpy().copy(),
                       dW=gc.weight.grad.numpy().copy(), db=gc.bias.grad.numpy().copy(), dXfull=xt.grad.numpy().copy()))

    # (3) dense strided adjacency (the fork's live scripts, utils.py:124-132), fork shapes 8->32
    nd = 96
    ad = np.abs(rng_inputs(21, (nd, nd)))
    ad[ad < 0.3] = 0.0
    x = rng_inputs(22, (nd, 8))
    g = rng_inputs(23, (nd, 32))
    r = run_layer(layers, 8, 32, True, x, torch.Tensor(ad), g)
    r["adj"] = ad
    put("dense", r)

    # (4) no bias, CSR adjacency, input without grad
    a4 = a2.coalesce().to_sparse_csr()
    x = rng_inputs(31, (nn, 16))
    g = rng_inputs(32, (nn, 16))
    r = run_layer(layers, 16, 16, False, x, a4, g, x_requires_grad=False)
    put("nobias_csr", r)

    # (5) torch.spmm functional, rectangular
    rs = np.random.default_rng(41)
    m_, k_, f_ = 50, 80, 24
    rr = rs.integers(0, m_, 400)
    cc = rs.integers(0, k_, 400)
    vv = rs.standard_normal(400).astype(np.float32)
    a5 = torch.sparse_coo_tensor(np.vstack([rr, cc]), vv, (m_, k_))
    d5 = torch.from_numpy(rng_inputs(42, (k_, f_))).requires_grad_(True)
    o5 = torch.spmm(a5, d5)
    g5 = rng_inputs(43, (m_, f_))
    o5.backward(torch.from_numpy(g5))
    put("spmm_rect", dict(rows=rr, cols=cc, vals=vv, shape=np.array([m_, k_, f_]), out=o5.detach().numpy().copy(),
                          dB=d5.grad.numpy().copy()))

    # (6) the unchanged 3-layer caller: models.GeneratorGCN (relu after every layer, models.py:103-111)
    e6 = pipe["edges1"]
    a6 = ref_pipeline(utils, e6, 257)
    torch.manual_seed(42)
    m = models.GeneratorGCN(8, 32, 32, 0.5, 10)
    x = torch.from_numpy(rng_inputs(51, (257, 8))).requires_grad_(True)
    y = m(x, a6)
    g = rng_inputs(52, (257, 32))
    y.backward(torch.from_numpy(g))
    d6 = {"out": y.detach().numpy().copy(), "dX": x.grad.numpy().copy()}
    for name, p in m.named_parameters():
        d6["param:" + name] = p.detach().numpy().copy()
        d6["grad:" + name] = p.grad.numpy().copy()
    put("stack3", d6)

    np.savez_compressed(os.path.join(OUT, "layer_cases.npz"), **cases)

    # ---------------------------------------------------------------- load_adj (utils.py:93-132): CBG adjacency
    # The reference's own function, fed a synthetic hourly POI x CBG visit list through the file layout it
    # expects (<mob_data_root>/<msa>/<full name>_2020-03-01_to_2020-05-02.pkl): adj = avg^T avg, dense fp32.
    import pickle
    import tempfile

    constants = sys.modules.get("constants") or __import__("constants")
    rs = np.random.default_rng(61)
    n_poi, n_cbg, n_hours = 53, 31, 6
    hours = []
    for _ in range(n_hours):
        dense = rs.poisson(0.4, size=(n_poi, n_cbg)).astype(np.float64) * rs.random((n_poi, n_cbg))
        hours.append(sp.csr_matrix(dense))
    with tempfile.TemporaryDirectory() as root, tempfile.TemporaryDirectory() as outdir:
        msa = "SanFrancisco"
        os.makedirs(os.path.join(root, msa))
        with open(os.path.join(root, msa, "%s_2020-03-01_to_2020-05-02.pkl" % constants.MSA_NAME_FULL_DICT[msa]), "wb") as f:
            pickle.dump(hours, f)
        # First call: `np.zeros(..) += <scipy sparse>` turns avg_array into an np.matrix, and the matrix product
        # in the double loop (utils.py:128) raises -- after avg_array_<msa>.npy has been saved.  A second call
        # loads that file as a plain ndarray and completes: that is how the function runs at all upstream.
        try:
            adj_ref, num_cbgs = utils.load_adj(msa, root, outdir)
        except ValueError:
            adj_ref, num_cbgs = utils.load_adj(msa, root, outdir)
    assert num_cbgs == n_cbg and adj_ref.dtype == torch.float32
    np.savez_compressed(os.path.join(OUT, "load_adj.npz"), hours=np.stack([h.toarray() for h in hours]),
                        adj=adj_ref.numpy().copy())
    print("wrote", sorted(os.listdir(OUT)))
    for f in os.listdir(OUT):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
