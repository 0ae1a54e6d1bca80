"""GPU tests of the (ReLU ->) fresh BatchNorm op, the programmatic-dependent-launch chain and the device side of the
file loaders.  Written at the end of round 1 (then gated, never run); first run on hardware in round 2
(profiles/r02_pytest_optin_first_hardware_run.txt: 19 passed) and part of the default `pytest -m gpu` suite since.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


def dev():
    return torch.device("cuda:0")


@pytest.fixture()
def pdl():
    """Switches programmatic dependent launch on for the test body (csrc/common.cuh), restores plain stream order."""
    from pygcn_b200 import _lib

    lib = _lib.load()

    def switch(on):
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_PDL, 1 if on else 0), "set_tuning")
    yield switch
    lib.gcnb_set_tuning(_lib.TUNE_PDL, 0)


def _graph(P, n, n_edges, seed):
    rs = np.random.default_rng(seed)
    src = torch.from_numpy(rs.integers(0, n, n_edges).astype(np.int32)).to(dev())
    dst = torch.from_numpy(rs.integers(0, n, n_edges).astype(np.int32)).to(dev())
    return P.Graph.from_edges(src, dst, n)


@pytest.mark.parametrize("n,fin,fout,relu", [(20000, 64, 32, False), (20000, 64, 32, True), (20000, 16, 7, True),
                                             (30000, 100, 256, False), (6000, 8, 32, True)])
def test_pdl_chain_is_bit_identical_to_plain_launches(pdl, n, fin, fout, relu):
    """The same kernels in the same order: with the PDL instantiations (launch_dependents first, wait before the first
    dependent access) out, dX, dW and db must equal the plain-launch results bit for bit, eagerly and as a CUDA-graph
    replay, over several back-to-back steps (the overlap only exists between consecutive kernels)."""
    import pygcn_b200 as P

    gr = _graph(P, n, 20 * n, seed=fout)
    gen = torch.Generator(device=dev()).manual_seed(fin)
    x = torch.randn(n, fin, generator=gen, device=dev())
    g = torch.randn(n, fout, generator=gen, device=dev())
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout, fuse_relu=relu).to(dev())

    def step():
        layer.weight.grad = None
        layer.bias.grad = None
        xt = x.clone().requires_grad_(True)
        out = layer(xt, gr)
        out.backward(g)
        return out.detach().clone(), xt.grad.clone(), layer.weight.grad.clone(), layer.bias.grad.clone()

    pdl(False)
    want = step()
    pdl(True)
    for _ in range(4):
        got = step()
        torch.cuda.synchronize()
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    # graph replay with the programmatic edges captured
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    xs = x.clone()
    layer.weight.grad = None
    layer.bias.grad = None
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        o_static = layer(xs, gr)
        o_static.backward(g)
    for _ in range(3):
        cg.replay()
    torch.cuda.synchronize()
    assert torch.equal(o_static.detach(), want[0])
    assert torch.equal(layer.weight.grad, want[2]) and torch.equal(layer.bias.grad, want[3])


# ---------------------------------------------------------------- (ReLU ->) fresh BatchNorm
def _torch_apply_bn(y, relu):
    """The reference's lines (pygcn/models.py:41-45, 49): a fresh nn.BatchNorm1d on F.relu(y)."""
    a = torch.relu(y) if relu else y
    return torch.nn.BatchNorm1d(y.shape[1]).to(y.device)(a)


def _normwise(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", ["cbg32", "odd7", "plain16", "deadcol", "two_rows"])
def test_apply_bn_matches_reference_fixture(golden, name):
    """apply_bn through the C ABI against the reference's own apply_bn(F.relu(y)) outputs and gradients
    (tests/golden/apply_bn.npz): <= 1e-5 norm-wise (the 2-row batch's gradient is a cancellation: 1e-4)."""
    import pygcn_b200 as P

    c = golden("apply_bn.npz")
    relu = bool(c[name + "/relu"])
    y = torch.from_numpy(c[name + "/y"]).to(dev()).requires_grad_(True)
    out = P.apply_bn(y, relu=relu)
    out.backward(torch.from_numpy(c[name + "/g"]).to(dev()))
    assert _normwise(out.detach().cpu(), torch.from_numpy(c[name + "/out"])) < 1e-5
    assert _normwise(y.grad.cpu(), torch.from_numpy(c[name + "/dy"])) < (1e-4 if name == "two_rows" else 1e-5)


@pytest.mark.parametrize("n,f,relu", [(100000, 32, True), (100000, 32, False), (50001, 47, True), (232965, 256, True),
                                      (3, 1, True), (70000, 1040, False)])
def test_apply_bn_matches_torch_batchnorm(n, f, relu):
    """Against torch's own CUDA BatchNorm1d + ReLU + autograd on the same input (what the reference executes on a GPU),
    an fp64 evaluation arbitrating; float4 and scalar paths, a panel wider than one column tile, bit-identical reruns."""
    import pygcn_b200 as P

    gen = torch.Generator(device=dev()).manual_seed(n + f)
    y = (torch.randn(n, f, generator=gen, device=dev()) + 0.25).requires_grad_(True)
    g = torch.randn(n, f, generator=gen, device=dev())
    out = P.apply_bn(y, relu=relu)
    out.backward(g)
    dy = y.grad.clone()
    y.grad = None
    ref = _torch_apply_bn(y, relu)
    ref.backward(g)
    y64 = y.detach().double().requires_grad_(True)
    a64 = torch.relu(y64) if relu else y64
    ex = (a64 - a64.mean(0)) / torch.sqrt(a64.var(0, unbiased=False) + 1e-5)
    ex.backward(g.double())
    tol = 1e-5 if n >= 16 else 1e-3  # (a 3-row batch differentiates through a near-cancellation, torch's result too)
    for ours, theirs, exact in ((out.detach(), ref.detach(), ex.detach()), (dy, y.grad, y64.grad)):
        assert _normwise(ours, exact) <= max(tol, 2 * _normwise(theirs, exact))
    y.grad = None
    out2 = P.apply_bn(y, relu=relu)
    out2.backward(g)
    assert torch.equal(out2, out) and torch.equal(y.grad, dy)


def test_apply_bn_after_the_fused_relu_layer_and_on_views():
    """As the models use it: apply_bn(F.relu(gc(x, adj))) == apply_bn(gc_fused_relu(x, adj)) == apply_bn(gc(x, adj),
    relu=True), gradients reaching the layer's weights; a column-slice view as input; torch's error for a 1-row batch."""
    import pygcn_b200 as P

    n, fin, fout = 20000, 64, 32
    gr = _graph(P, n, 20 * n, seed=3)
    gen = torch.Generator(device=dev()).manual_seed(9)
    x = torch.randn(n, fin, generator=gen, device=dev())
    g = torch.randn(n, fout, generator=gen, device=dev())
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout).to(dev())
    fused = P.GraphConvolution(fin, fout, fuse_relu=True).to(dev())
    fused.load_state_dict(layer.state_dict())
    res = []
    for fn in (lambda: _torch_apply_bn(torch.relu(layer(x, gr)), False), lambda: P.apply_bn(layer(x, gr), relu=True),
               lambda: P.apply_bn(fused(x, gr))):
        for m in (layer, fused):
            m.zero_grad()
        o = fn()
        o.backward(g)
        used = fused if fused.weight.grad is not None else layer
        res.append((o.detach(), used.weight.grad.clone(), used.bias.grad.clone()))
    for o, dw, db in res[1:]:
        assert _normwise(o, res[0][0]) < 1e-5 and _normwise(dw, res[0][1]) < 1e-5
        assert float((db - res[0][2]).abs().max()) < 1e-5 * float(res[0][1].abs().max())  # (db of a layer under BN is ~0)
    wide = torch.randn(5000, 40, generator=gen, device=dev())
    view = wide[:, 3:35]  # row stride 40 > width 32, start not 16-byte aligned: scalar path
    assert _normwise(P.apply_bn(view, relu=True), _torch_apply_bn(view, True)) < 1e-5
    with pytest.raises(ValueError):
        P.apply_bn(torch.zeros(1, 8, device=dev()))
    with pytest.raises(RuntimeError):
        P.apply_bn(torch.zeros(4, 8))


# ---------------------------------------------------------------- on-disk formats (pygcn_b200/io.py)
def test_graph_from_cites_and_load_adj_files_on_the_device(tmp_path, golden):
    """The file loaders composed with the measured device paths: a `.cites` file written from the golden Cora edges ->
    the golden adjacency bit for bit; the three file levels of utils.load_adj with the tcgen05 product -> the
    reference's adjacency (<= 1e-5), cached like the reference caches it."""
    import pickle

    import scipy.sparse as sp

    from pygcn_b200 import io as IO

    g = golden("cora_pipeline.npz")
    ids = np.arange(1000, 1000 + int(g["n"]), dtype=np.int64) * 7            # any increasing paper ids
    p = tmp_path / "cora.cites"
    np.savetxt(p, ids[g["edges"]], fmt="%d", delimiter="\t")
    gr, ids_back = IO.graph_from_cites(str(p), dev())
    coo = gr.to_sparse_coo()
    assert np.array_equal(ids_back, ids)
    assert np.array_equal(coo._indices().cpu().numpy(), g["indices"])
    assert np.array_equal(coo._values().cpu().numpy(), g["values"])

    c = golden("load_adj.npz")
    root, out = tmp_path / "mob", tmp_path / "out"
    (root / "SanFrancisco").mkdir(parents=True)
    out.mkdir()
    with open(root / "SanFrancisco" / ("X" + IO.CBG_PICKLE_SUFFIX), "wb") as f:
        pickle.dump([sp.csr_matrix(h) for h in c["hours"]], f)
    adj, n = IO.load_adj_files("SanFrancisco", str(root), str(out), dev(), msa_name_full="X", save=True)
    assert n == c["adj"].shape[0] and adj.is_cuda and adj.dtype == torch.float32
    assert _normwise(adj.cpu(), torch.from_numpy(c["adj"])) < 1e-5
    # saved under its own name, never as the reference's fp64 cache adj_<msa>.npy; the averaged visits are cached like utils.py:121
    assert os.path.exists(out / "adj_SanFrancisco.gcnb200.npy") and not os.path.exists(out / "adj_SanFrancisco.npy")
    adj1, _ = IO.load_adj_files("SanFrancisco", "/nonexistent", str(out), dev())  # level 2: from avg_array_<msa>.npy
    assert torch.equal(adj1, adj)


def test_apply_bn_on_batched_samples_equals_the_per_sample_loop():
    """x[B, N, F] (the batched layer's output layout): one pass over the node-major [N, B*F] panel == apply_bn on every
    sample, the loop of GCN_OVER_MLP.forward (pygcn/models.py:343-349); gradients too; composed with the batched layer."""
    import pygcn_b200 as P

    b, n, fin, f = 5, 20000, 8, 32
    gr = _graph(P, n, 10 * n, seed=4)
    gen = torch.Generator(device=dev()).manual_seed(12)
    x = torch.randn(b, n, fin, generator=gen, device=dev())
    g = torch.randn(b, n, f, generator=gen, device=dev())
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, f).to(dev())
    layer.zero_grad()
    out = P.apply_bn(layer(x, gr), relu=True)
    out.backward(g)
    dw = layer.weight.grad.clone()
    layer.zero_grad()
    outs = torch.stack([_torch_apply_bn(layer(x[i], gr), True) for i in range(b)])
    outs.backward(g)
    assert out.shape == (b, n, f)
    assert _normwise(out.detach(), outs.detach()) < 1e-5 and _normwise(dw, layer.weight.grad) < 1e-5
    plain = torch.randn(3, 1000, 12, generator=gen, device=dev())  # batch-major storage: one copy into the node-major layout
    want = torch.stack([_torch_apply_bn(plain[i], False) for i in range(3)])
    assert _normwise(P.apply_bn(plain), want) < 1e-5
