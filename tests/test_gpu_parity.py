"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's goldens.

Bars: index / byte work bit-exact; float work max|a-b|/max|b| <= 1e-5.
Everything here needs a B200: run with `pytest -m gpu`.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rng_inputs
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5


def dev():
    return torch.device("cuda:0")


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev())


def err(a, b):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return O.normwise_err(a, b)


@pytest.fixture(scope="module")
def P():
    import pygcn_b200

    assert pygcn_b200._lib.load().gcnb_check_device() == 0, pygcn_b200._lib.last_error()
    return pygcn_b200


# ------------------------------------------------------------------ graph build: bit-exact
def _edges_graph(P, e, n, **kw):
    return P.Graph.from_edges(cu(e[:, 0]), cu(e[:, 1]), n, **kw)


def test_cora_pipeline_bit_exact(P, golden):
    g = golden("cora_pipeline.npz")
    n = int(g["n"])
    gr = _edges_graph(P, g["edges"], n)
    coo = gr.to_sparse_coo()
    idx = coo._indices().cpu().numpy()
    val = coo._values().cpu().numpy()
    assert idx.dtype == np.int64 and val.dtype == np.float32
    assert np.array_equal(idx, g["indices"])
    assert np.array_equal(val.view(np.uint32), g["values"].view(np.uint32))
    assert not coo.is_coalesced()  # same flag the reference's constructor leaves
    # CSR and CSR^T against the oracle's restatement
    rowptr, col, v = (t.cpu().numpy() for t in gr.csr())
    assert np.array_equal(rowptr, O.coo_to_csr(g["indices"], n))
    assert np.array_equal(col, g["indices"][1]) and np.array_equal(v, g["values"])
    t_rowptr, t_col, t_val = O.transpose_csr(g["indices"], g["values"], n, n)
    rp, c, vv = (t.cpu().numpy() for t in gr.csr(transpose=True))
    assert np.array_equal(rp, t_rowptr) and np.array_equal(c, t_col)
    assert np.array_equal(vv.view(np.uint32), t_val.view(np.uint32))
    assert gr.pattern_symmetric
    _, counts = O.degree_bins(O.coo_to_csr(g["indices"], n))
    assert gr.bin_rows == list(counts) and gr.max_degree == 169


@pytest.mark.parametrize("k", [0, 1, 2])
def test_small_pipelines_bit_exact(P, golden, k):
    g = golden("pipeline_small.npz")
    gr = _edges_graph(P, g[f"edges{k}"], int(g[f"n{k}"]))
    coo = gr.to_sparse_coo()
    assert np.array_equal(coo._indices().cpu().numpy(), g[f"indices{k}"])
    assert np.array_equal(coo._values().cpu().numpy().view(np.uint32), g[f"values{k}"].view(np.uint32))


def test_pipeline_flag_subsets_and_empty(P):
    rs = np.random.default_rng(5)
    n = 50
    e = rs.integers(0, n, size=(200, 2)).astype(np.int32)
    # no symmetrise / no self loops / no normalise: plain multiplicity matrix
    gr = _edges_graph(P, e, n, symmetrize=False, self_loops=False, row_normalize=False)
    r, c, v = O.edges_to_counts(e[:, 0], e[:, 1], n)
    coo = gr.to_sparse_coo()
    assert np.array_equal(coo._indices().cpu().numpy(), np.vstack([r, c]))
    assert np.array_equal(coo._values().cpu().numpy(), v)
    # symmetrise only
    gr = _edges_graph(P, e, n, symmetrize=True, self_loops=False, row_normalize=False)
    r2, c2, v2 = O.symmetrize_max(r, c, v, n)
    coo = gr.to_sparse_coo()
    assert np.array_equal(coo._indices().cpu().numpy(), np.vstack([r2, c2]))
    assert np.array_equal(coo._values().cpu().numpy(), v2)
    # empty edge list: identity after + I and normalisation
    gr = P.Graph.from_edges(torch.empty(0, dtype=torch.int32, device=dev()),
                            torch.empty(0, dtype=torch.int32, device=dev()), 7)
    coo = gr.to_sparse_coo()
    assert np.array_equal(coo._indices().cpu().numpy(), np.vstack([np.arange(7), np.arange(7)]))
    assert np.array_equal(coo._values().cpu().numpy(), np.ones(7, np.float32))
    # n = 0
    gr = P.Graph.from_edges(torch.empty(0, dtype=torch.int32, device=dev()),
                            torch.empty(0, dtype=torch.int32, device=dev()), 0)
    assert gr.nnz == 0 and gr.n_rows == 0
    with pytest.raises(RuntimeError):
        _edges_graph(P, np.array([[0, 99]], np.int32), 10)


def test_coo_unsorted_duplicates_layout(P, golden):
    c = golden("layer_cases.npz")
    n = int(c["ragged/n"])
    rows, cols, vals = c["ragged/rows"], c["ragged/cols"], c["ragged/vals"]
    adj = torch.sparse_coo_tensor(cu(np.vstack([rows, cols])), cu(vals), (n, n))
    gr = P.Graph.from_torch(adj)
    order = np.lexsort((np.arange(rows.size), cols, rows))  # (row, col) stable
    rowptr, col, v = (t.cpu().numpy() for t in gr.csr())
    assert np.array_equal(col, cols[order]) and np.array_equal(v, vals[order])
    assert np.array_equal(rowptr, O.coo_to_csr(np.vstack([rows[order], cols[order]]), n))
    assert gr.nnz == rows.size and gr.bin_rows[0] > 0  # empty rows exist
    # CSR input gives the same handle contents as the coalesced COO
    csr = adj.coalesce().to_sparse_csr()
    g2 = P.Graph.from_torch(csr)
    assert np.array_equal(g2.csr()[0].cpu().numpy(), csr.crow_indices().cpu().numpy())
    assert np.array_equal(g2.csr()[1].cpu().numpy(), csr.col_indices().cpu().numpy())
    # out-of-range index is an error, not a crash
    bad = torch.sparse_coo_tensor(cu(np.array([[0], [5]])), cu(np.ones(1, np.float32)), (4, 4),
                                  check_invariants=False)
    with pytest.raises(RuntimeError):
        P.Graph.from_torch(bad)


# ------------------------------------------------------------------ layer vs the reference's goldens
def _run_layer(P, w, b, x, adj, g, relu=False, x_grad=True):
    layer = P.GraphConvolution(w.shape[0], w.shape[1], bias=b is not None, fuse_relu=relu).to(dev())
    with torch.no_grad():
        layer.weight.copy_(cu(w))
        if b is not None:
            layer.bias.copy_(cu(b))
    xt = x.clone().requires_grad_(x_grad)
    out = layer(xt, adj)
    out.backward(g)
    return layer, xt, out


def test_layer_cora_golden(P, golden):
    p = golden("cora_pipeline.npz")
    c = golden("layer_cases.npz")
    n = int(p["n"])
    gr = _edges_graph(P, p["edges"], n)
    layer, xt, out = _run_layer(P, c["cora_l1/weight"], c["cora_l1/bias"], cu(rng_inputs(1, (n, 1433))), gr,
                                cu(rng_inputs(2, (n, 16))))
    assert err(out, c["cora_l1/out"]) < TOL
    assert err(layer.weight.grad, c["cora_l1/dW"]) < TOL
    assert err(layer.bias.grad, c["cora_l1/db"]) < TOL
    assert err(xt.grad[:32], c["cora_l1/dX_head"]) < TOL
    # the same adjacency passed as the torch sparse tensor the reference builds
    adj = torch.sparse_coo_tensor(cu(p["indices"]), cu(p["values"]), (n, n))
    layer, xt, out = _run_layer(P, c["cora_l2/weight"], c["cora_l2/bias"], cu(rng_inputs(3, (n, 16))), adj,
                                cu(rng_inputs(4, (n, 7))))
    assert err(out, c["cora_l2/out"]) < TOL
    assert err(layer.weight.grad, c["cora_l2/dW"]) < TOL
    assert err(layer.bias.grad, c["cora_l2/db"]) < TOL
    assert err(xt.grad, c["cora_l2/dX"]) < TOL


def test_layer_ragged_noncontiguous_golden(P, golden):
    c = golden("layer_cases.npz")
    n = int(c["ragged/n"])
    adj = torch.sparse_coo_tensor(cu(np.vstack([c["ragged/rows"], c["ragged/cols"]])), cu(c["ragged/vals"]), (n, n))
    xw = cu(rng_inputs(12, (n, 40))).requires_grad_(True)
    layer = P.GraphConvolution(33, 7).to(dev())
    with torch.no_grad():
        layer.weight.copy_(cu(c["ragged/weight"]))
        layer.bias.copy_(cu(c["ragged/bias"]))
    out = layer(xw[:, :33], adj)  # column-slice view, as pygcn/models.py:345 passes it
    out.backward(cu(rng_inputs(13, (n, 7))))
    assert err(out, c["ragged/out"]) < TOL
    assert err(layer.weight.grad, c["ragged/dW"]) < TOL
    assert err(layer.bias.grad, c["ragged/db"]) < TOL
    assert err(xw.grad, c["ragged/dXfull"]) < TOL


def test_layer_dense_adj_golden(P, golden):
    c = golden("layer_cases.npz")
    layer, xt, out = _run_layer(P, c["dense/weight"], c["dense/bias"], cu(rng_inputs(22, (96, 8))),
                                cu(c["dense/adj"]), cu(rng_inputs(23, (96, 32))))
    assert err(out, c["dense/out"]) < TOL
    assert err(layer.weight.grad, c["dense/dW"]) < TOL
    assert err(layer.bias.grad, c["dense/db"]) < TOL
    assert err(xt.grad, c["dense/dX"]) < TOL


@pytest.mark.parametrize("n,fin,fout", [(2943, 8, 32), (300, 16, 7), (257, 33, 48)])
def test_dense_route_fork_shape(P, n, fin, fout):
    """The fork's live scripts pass a dense `torch.Tensor(adj)` (pygcn/utils.py:124-132, N = 2943 for
    San Francisco): dense handles run as tensor-core GEMMs, results must match the reference's
    `torch.spmm(dense, .)` == mm semantics (fp64 arbiter) for forward and all gradients."""
    gen = torch.Generator(device=dev()).manual_seed(n)
    v = torch.rand(40, n, generator=gen, device=dev())
    adj = (v.t() @ v) / 40.0  # adj[i][j] = sum_p avg[p,i] * avg[p,j], fully dense, un-normalised
    if n == 300:
        adj = adj * (torch.rand(n, n, generator=gen, device=dev()) < 0.4)  # 40 % dense
    gr = P.Graph.from_torch(adj)
    assert gr.dense_route and gr.nnz == int((adj != 0).sum())
    x = torch.randn(n, fin, generator=gen, device=dev())
    g = torch.randn(n, fout, generator=gen, device=dev())
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout).to(dev())
    xt = x.clone().requires_grad_(True)
    out = layer(xt, adj)  # the tensor itself: converted once, cached
    out.backward(g)
    assert P.as_graph(adj) is P.as_graph(adj)
    w64 = layer.weight.detach().double().requires_grad_(True)
    b64 = layer.bias.detach().double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    ref = adj.double() @ (x64 @ w64) + b64
    ref.backward(g.double())
    for mine, r in ((out, ref), (layer.weight.grad, w64.grad), (layer.bias.grad, b64.grad), (xt.grad, x64.grad)):
        assert ((mine.double() - r).abs().max() / r.abs().max()).item() < TOL
    # fused ReLU epilogue on the dense route
    l2 = P.GraphConvolution(fin, fout, fuse_relu=True).to(dev())
    l2.load_state_dict(layer.state_dict())
    assert ((l2(x, adj) - torch.relu(ref.detach()).float()).abs().max() / ref.abs().max()).item() < TOL


def test_fused_dropout_relu_epilogue(P):
    """bias + ReLU + dropout mask in the SpMM epilogue and the matching masked backward, against the
    same ops composed in torch (F.relu then mask * 1/(1-p)), CSR and dense routes."""
    n = 2000
    src, dst = _powerlaw_graph(n, 30000, seed=9, hub_deg=1500)
    gr = P.Graph.from_edges(cu(src), cu(dst), n)
    dense = gr.to_sparse_coo().to_dense()
    gen = torch.Generator(device=dev()).manual_seed(4)
    x = torch.randn(n, 20, generator=gen, device=dev())
    g = torch.randn(n, 12, generator=gen, device=dev())
    mask = torch.rand(n, 12, generator=gen, device=dev()) >= 0.3
    torch.manual_seed(42)
    layer = P.GraphConvolution(20, 12).to(dev())
    for adj in (gr, dense):
        for relu in (False, True):
            xt = x.clone().requires_grad_(True)
            layer.zero_grad()
            out = P.gcn_layer(xt, adj, layer.weight, layer.bias, relu=relu, dropout_mask=mask, dropout_p=0.3)
            out.backward(g)
            w = layer.weight.detach().double().requires_grad_(True)
            b = layer.bias.detach().double().requires_grad_(True)
            x64 = x.double().requires_grad_(True)
            z = dense.double() @ (x64 @ w) + b
            ref = (torch.relu(z) if relu else z) * mask.double() / 0.7
            ref.backward(g.double())
            for mine, r in ((out, ref), (layer.weight.grad, w.grad), (layer.bias.grad, b.grad), (xt.grad, x64.grad)):
                assert ((mine.double() - r).abs().max() / r.abs().max()).item() < TOL
    # module option: active in training mode only
    drop = P.GraphConvolution(20, 12, dropout=0.5).to(dev())
    drop.eval()
    assert torch.equal(drop(x, gr), P.gcn_layer(x, gr, drop.weight, drop.bias))
    drop.train()
    frac = (drop(x, gr) == 0).float().mean().item()
    assert 0.4 < frac < 0.6


@pytest.mark.parametrize("kind", ["csr", "dense"])
@pytest.mark.parametrize("relu", [False, True])
def test_batched_samples_match_per_sample_loop(P, kind, relu):
    """x [B, N, Fin] in one pass (one GEMM + one SpMM of width B*Fout) == the reference's per-sample loop
    (GCN_OVER_MLP.forward, pygcn/models.py:343-349) over the unbatched layer, forward and all gradients."""
    n, bsz, fin, fout = 1500, 5, 8, 32
    src, dst = _powerlaw_graph(n, 20000, seed=21, hub_deg=1200)
    gr = P.Graph.from_edges(cu(src), cu(dst), n)
    adj = gr if kind == "csr" else gr.to_sparse_coo().to_dense()
    gen = torch.Generator(device=dev()).manual_seed(8)
    x = torch.randn(bsz, n, fin + 3, generator=gen, device=dev())[:, :, :fin]  # column-slice view (models.py:345)
    g = torch.randn(bsz, n, fout, generator=gen, device=dev())
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout, fuse_relu=relu).to(dev())
    xb = x.clone().requires_grad_(True)
    out = layer(xb, adj)
    assert out.shape == (bsz, n, fout)
    out.backward(g)
    gw, gb = layer.weight.grad.clone(), layer.bias.grad.clone()
    layer.zero_grad()
    xs = x.clone().requires_grad_(True)
    outs = torch.stack([layer(xs[i], adj) for i in range(bsz)])
    outs.backward(g)
    for mine, ref in ((out, outs), (xb.grad, xs.grad), (gw, layer.weight.grad), (gb, layer.bias.grad)):
        assert ((mine - ref).abs().max() / ref.abs().max()).item() < TOL
    # fp64 arbiter for the batched path itself
    a64 = gr.to_sparse_coo().to_dense().double()
    ref64 = a64 @ (x.double() @ layer.weight.detach().double()) + layer.bias.detach().double()
    if relu:
        ref64 = torch.relu(ref64)
    assert ((out.double() - ref64).abs().max() / ref64.abs().max()).item() < TOL


def test_load_adj_matches_reference_golden_and_oracle(P, golden):
    """the CBG adjacency avg^T avg of utils.load_adj on the device, against the
    reference-generated fixture and, at a fork-like size (2943 CBGs), against the oracle."""
    c = golden("load_adj.npz")
    avg, adj_o = O.cbg_adjacency(c["hours"])
    adj = P.load_adj(cu(avg.astype(np.float32)))
    assert adj.shape == c["adj"].shape and adj.dtype == torch.float32
    assert err(adj, c["adj"]) < TOL and err(adj, adj_o) < TOL
    rs = np.random.default_rng(5)
    hours = rs.poisson(0.05, size=(3, 4000, 2943)) * rs.random((3, 4000, 2943))
    avg, adj_o = O.cbg_adjacency(hours)
    adj = P.load_adj(cu(avg.astype(np.float32)))
    assert err(adj, adj_o) < TOL
    # and it feeds the layer like the scripts' dense `adj` does (policy-generator.py:339)
    gr = P.Graph.from_torch(adj)
    assert gr.n_rows == 2943


def test_cora_two_layer_training_follows_torch_reference(P, golden):
    """the upstream 2-layer GCN training loop (dropout after layer 1, log_softmax +
    NLL, Adam -- the semantics of the commented-out Cora loader, pygcn/utils.py:348-382 / upstream train.py)
    on the real cora.cites graph with synthetic features and planted labels.  Our layers (fused ReLU +
    fused dropout mask) and the reference's three lines on torch CUDA ops start from the same weights
    and see the same dropout masks: the loss curves must agree step by step, and training must work."""
    c = golden("cora_pipeline.npz")
    n = int(c["n"])
    e = c["edges"]
    gr = P.Graph.from_edges(cu(e[:, 0]), cu(e[:, 1]), n)
    coo = gr.to_sparse_coo().coalesce()
    rs = np.random.default_rng(77)
    labels = torch.from_numpy(rs.integers(0, 7, n)).to(dev())
    feats = torch.from_numpy((rs.standard_normal((n, 64)) + 2.0 * np.eye(7)[labels.cpu().numpy()].repeat(10, 1)[:, :64]
                              ).astype(np.float32)).to(dev())
    idx_train = torch.arange(0, 140, device=dev())
    torch.manual_seed(42)
    g1, g2 = P.GraphConvolution(64, 16, fuse_relu=True).to(dev()), P.GraphConvolution(16, 7).to(dev())
    w1, b1 = (t.detach().clone().requires_grad_(True) for t in (g1.weight, g1.bias))
    w2, b2 = (t.detach().clone().requires_grad_(True) for t in (g2.weight, g2.bias))
    opt_a = torch.optim.Adam(list(g1.parameters()) + list(g2.parameters()), lr=0.01, weight_decay=5e-4)
    opt_b = torch.optim.Adam([w1, b1, w2, b2], lr=0.01, weight_decay=5e-4)
    gen = torch.Generator(device=dev()).manual_seed(3)
    la, lb = [], []
    for step in range(30):
        mask = torch.rand(n, 16, generator=gen, device=dev()) >= 0.5
        opt_a.zero_grad()
        h = P.gcn_layer(feats, gr, g1.weight, g1.bias, relu=True, dropout_mask=mask, dropout_p=0.5)
        out = F.log_softmax(g2(h, gr), dim=1)
        loss = F.nll_loss(out[idx_train], labels[idx_train])
        loss.backward()
        opt_a.step()
        opt_b.zero_grad()
        hr = torch.relu(torch.spmm(coo, torch.mm(feats, w1)) + b1) * mask / 0.5
        outr = F.log_softmax(torch.spmm(coo, torch.mm(hr, w2)) + b2, dim=1)
        lossr = F.nll_loss(outr[idx_train], labels[idx_train])
        lossr.backward()
        opt_b.step()
        la.append(loss.item())
        lb.append(lossr.item())
    assert la[-1] < 0.5 * la[0]                                  # it trains
    assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-4 * max(la)  # and follows the reference trajectory
    assert ((g1.weight - w1).abs().max() / w1.abs().max()).item() < 1e-3


def test_layer_nobias_csr_golden(P, golden):
    c = golden("layer_cases.npz")
    n = int(c["ragged/n"])
    adj = torch.sparse_coo_tensor(cu(np.vstack([c["ragged/rows"], c["ragged/cols"]])), cu(c["ragged/vals"]),
                                  (n, n)).coalesce().to_sparse_csr()
    layer, xt, out = _run_layer(P, c["nobias_csr/weight"], None, cu(rng_inputs(31, (n, 16))), adj,
                                cu(rng_inputs(32, (n, 16))), x_grad=False)
    assert layer.bias is None and xt.grad is None
    assert err(out, c["nobias_csr/out"]) < TOL
    assert err(layer.weight.grad, c["nobias_csr/dW"]) < TOL


def test_spmm_functional_golden(P, golden):
    c = golden("layer_cases.npz")
    m, k, f = (int(v) for v in c["spmm_rect/shape"])
    adj = torch.sparse_coo_tensor(cu(np.vstack([c["spmm_rect/rows"], c["spmm_rect/cols"]])), cu(c["spmm_rect/vals"]),
                                  (m, k))
    d = cu(rng_inputs(42, (k, f))).requires_grad_(True)
    out = P.spmm(adj, d)
    out.backward(cu(rng_inputs(43, (m, f))))
    assert err(out, c["spmm_rect/out"]) < TOL
    assert err(d.grad, c["spmm_rect/dB"]) < TOL


@pytest.mark.parametrize("fused", [False, True])
def test_stack3_unchanged_caller_golden(P, golden, fused):
    """relu(gc(x, adj)) x3 exactly as models.GeneratorGCN.forward (pygcn/models.py:103-111)."""
    p = golden("pipeline_small.npz")
    c = golden("layer_cases.npz")
    gr = _edges_graph(P, p["edges1"], 257)
    torch.manual_seed(42)
    gcs = [P.GraphConvolution(8, 32, fuse_relu=fused), P.GraphConvolution(32, 32, fuse_relu=fused),
           P.GraphConvolution(32, 32, fuse_relu=fused)]
    for k, gc in enumerate(gcs, 1):  # same seed, same draw order as the reference model
        assert np.array_equal(gc.weight.detach().numpy(), c[f"stack3/param:gc{k}.weight"])
        assert np.array_equal(gc.bias.detach().numpy(), c[f"stack3/param:gc{k}.bias"])
        gc.to(dev())
    x = cu(rng_inputs(51, (257, 8))).requires_grad_(True)
    h = x
    with torch.autograd.set_detect_anomaly(True):  # policy-generator.py:419
        for gc in gcs:
            h = F.relu(gc(h, gr))
        h.backward(cu(rng_inputs(52, (257, 32))), retain_graph=True)  # policy-generator.py:420
    assert err(h, c["stack3/out"]) < TOL
    assert err(x.grad, c["stack3/dX"]) < TOL
    for k, gc in enumerate(gcs, 1):
        assert err(gc.weight.grad, c[f"stack3/grad:gc{k}.weight"]) < TOL
        assert err(gc.bias.grad, c[f"stack3/grad:gc{k}.bias"]) < TOL
    assert sorted(gcs[0].state_dict().keys()) == ["bias", "weight"]


# ------------------------------------------------------------------ CUDA path vs the oracle on seeded inputs
def _powerlaw_graph(n, n_edges, seed, hub_deg=0):
    rs = np.random.default_rng(seed)
    src = (n * rs.random(n_edges) ** 3).astype(np.int64)  # skewed: low ids are hubs
    dst = rs.integers(0, n, n_edges)
    if hub_deg:
        src = np.concatenate([src, np.full(hub_deg, 3)])
        dst = np.concatenate([dst, rs.choice(n, hub_deg, replace=False)])
    return src.astype(np.int32), dst.astype(np.int32)


_KERNELS = {"auto": 0, "rows": 1, "group": 2, "tma": 3}


@pytest.fixture
def spmm_kernel(P, request):
    """Forces one of the SpMM kernels for the short-row bins (gcnb_set_tuning), restores auto after."""
    from pygcn_b200 import _lib

    lib = _lib.load()
    name, variant = request.param if isinstance(request.param, tuple) else (request.param, -1)
    if name == "stream":  # the streaming kernel wherever the view is eligible (auto keeps it for big graphs);
        # variant = gathered rows in flight per lane, paired with one of the L2 hint modes (2 -> 2, 4 -> 1, 8 -> 0)
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_KERNEL, 0), "set_tuning")
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_STREAM, 2), "set_tuning")
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_STREAM_BATCH, max(variant, 0)), "set_tuning")
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_STREAM_HINT, {2: 2, 4: 1, 8: 0}.get(variant, 0)), "set_tuning")
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_STREAM_HOT_MB, 1), "set_tuning")  # small graphs: some hot, some cold columns
    else:
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_STREAM, 0 if name != "auto" else 1), "set_tuning")
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_KERNEL, _KERNELS[name]), "set_tuning")
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_GROUP_VARIANT, variant), "set_tuning")
    yield name
    lib.gcnb_set_tuning(_lib.TUNE_SPMM_KERNEL, 0)
    lib.gcnb_set_tuning(_lib.TUNE_SPMM_GROUP_VARIANT, -1)
    lib.gcnb_set_tuning(_lib.TUNE_SPMM_STREAM, 1)
    lib.gcnb_set_tuning(_lib.TUNE_STREAM_BATCH, 0)
    lib.gcnb_set_tuning(_lib.TUNE_STREAM_HINT, 0)
    lib.gcnb_set_tuning(_lib.TUNE_STREAM_HOT_MB, 48)


@pytest.mark.parametrize("spmm_kernel", ["auto", "rows", "group", "tma", ("group", 0), ("group", 1), ("group", 2), ("group", 3),
                                         ("group", 4), ("group", 7), ("group", 13), ("group", 14), "stream", ("stream", 2),
                                         ("stream", 4), ("stream", 8)], indirect=True)
@pytest.mark.parametrize("fin,fout", [(64, 32), (5, 1), (9, 3), (16, 7), (33, 47), (100, 256), (20, 600)])
def test_layer_vs_oracle_widths_and_long_rows(P, fin, fout, spmm_kernel):
    n = 6000
    src, dst = _powerlaw_graph(n, 60000, seed=fin * 1000 + fout, hub_deg=3000)
    idx, val = O.build_normalized_adjacency(src, dst, n)
    gr = P.Graph.from_edges(cu(src), cu(dst), n)
    assert gr.nnz == idx.shape[1] and gr.n_long_chunks > 0 and gr.max_degree >= 3000
    x = rng_inputs(1, (n, fin))
    g = rng_inputs(2, (n, fout))
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout).to(dev())
    w = layer.weight.detach().cpu().numpy()
    b = layer.bias.detach().cpu().numpy()
    xt = cu(x).requires_grad_(True)
    out = layer(xt, gr)
    out.backward(cu(g))
    _, o_ref = O.c_layer_forward(x, w, b, idx, val, n)
    dw, db, dx, _ = O.c_layer_backward(x, w, True, idx, val, n, g)
    _, o64 = O.c_layer_forward(x, w, b, idx, val, n, dtype=np.float64)
    assert err(out, o_ref) < TOL and err(out, o64) < TOL
    assert err(layer.weight.grad, dw) < TOL
    assert err(layer.bias.grad, db) < TOL
    assert err(xt.grad, dx) < TOL
    # fused ReLU epilogue + masked backward against relu() composed on the oracle
    layer2 = P.GraphConvolution(fin, fout, fuse_relu=True).to(dev())
    layer2.load_state_dict(layer.state_dict())
    xt2 = cu(x).requires_grad_(True)
    out2 = layer2(xt2, gr)
    out2.backward(cu(g))
    assert err(out2, np.maximum(o_ref, 0)) < TOL
    # the mask comes from the CUDA output itself (entries within rounding of zero may differ in
    # sign from the oracle's); it must agree with the oracle's everywhere else
    o2 = out2.detach().cpu().numpy()
    flips = (o2 > 0) != (o_ref > 0)
    assert flips.sum() <= 4 and (flips.sum() == 0 or np.abs(o_ref[flips]).max() < 1e-5 * np.abs(o_ref).max())
    gm = O.relu_backward(g, o2)
    dw2, db2, dx2, _ = O.c_layer_backward(x, w, True, idx, val, n, gm)
    assert err(layer2.weight.grad, dw2) < TOL and err(layer2.bias.grad, db2) < TOL and err(xt2.grad, dx2) < TOL


@pytest.mark.parametrize("association", ["reference", "aggregate_first", "auto"])
@pytest.mark.parametrize("fin,fout,relu,need_dx", [(16, 64, False, True), (100, 256, True, False), (100, 256, True, True),
                                                   (5, 9, True, True), (33, 47, False, False), (64, 32, True, True)])
def test_layer_association_orders_match_the_oracle(P, association, fin, fout, relu, need_dx):
    """adj @ (X @ W) (the reference's order, pygcn/layers.py:33-34) and (adj @ X) @ W (what "auto" picks when
    in_features < out_features) are the same layer: output, dW, db and dX against the oracle, fused ReLU and a
    dropout mask included; dX only when the input asks for it."""
    n = 5000
    src, dst = _powerlaw_graph(n, 40000, seed=fin + fout, hub_deg=2100)
    idx, val = O.build_normalized_adjacency(src, dst, n)
    gr = P.Graph.from_edges(cu(src), cu(dst), n)
    x = rng_inputs(3, (n, fin))
    g = rng_inputs(4, (n, fout))
    torch.manual_seed(7)
    layer = P.GraphConvolution(fin, fout, fuse_relu=relu, association=association).to(dev())
    if association == "aggregate_first" and fin % 4 != 0:  # rows of X are not 16-byte aligned: only "auto" may decline
        with pytest.raises(RuntimeError, match="aligned"):
            layer(cu(x), gr)
        return
    w = layer.weight.detach().cpu().numpy()
    b = layer.bias.detach().cpu().numpy()
    xt = cu(x).requires_grad_(need_dx)
    out = layer(xt, gr)
    out.backward(cu(g))
    _, o_ref = O.c_layer_forward(x, w, b, idx, val, n)
    if relu:
        o2 = out.detach().cpu().numpy()
        assert err(out, np.maximum(o_ref, 0)) < TOL
        gm = O.relu_backward(g, o2)
    else:
        assert err(out, o_ref) < TOL
        gm = g
    dw, db, dx, _ = O.c_layer_backward(x, w, True, idx, val, n, gm)
    assert err(layer.weight.grad, dw) < TOL and err(layer.bias.grad, db) < TOL
    if need_dx:
        assert err(xt.grad, dx) < TOL
    else:
        assert xt.grad is None
    # dropout mask epilogue in either order
    keep = torch.rand(n, fout, device=dev()) > 0.4
    o_m = P.gcn_layer(cu(x), gr, layer.weight.detach(), layer.bias.detach(), relu=relu, dropout_mask=keep, dropout_p=0.4,
                      association=association)
    want = (np.maximum(o_ref, 0) if relu else o_ref) * keep.cpu().numpy() / 0.6
    assert err(o_m, want) < TOL
    with pytest.raises(ValueError):
        P.gcn_layer(cu(x), gr, layer.weight.detach(), association="sideways")


# ------------------------------------------------------------------ bf16 panel tier (north_star: 2e-2)
TOL_BF16 = 2e-2


def _to_bf16(P, src, ld_dst=None):
    import ctypes

    from pygcn_b200 import _lib

    lib = _lib.load()
    n, f = src.shape
    ld = ld_dst or (f + 7) // 8 * 8
    dst = torch.full((n, ld), float("nan"), dtype=torch.bfloat16, device=src.device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.gcnb_to_bf16(n, f, src.data_ptr(), src.stride(0), dst.data_ptr(), ld, st), "gcnb_to_bf16")
    return dst


def _spmm_raw(P, gr, dense, f, bf16, flags=0, bias=None):
    import ctypes

    from pygcn_b200 import _lib

    lib = _lib.load()
    rows = gr.n_cols if flags & _lib.SPMM_TRANSPOSE else gr.n_rows
    out = torch.empty(rows, f, device=dense.device)
    ws = torch.empty(max(int(lib.gcnb_spmm_workspace_bytes(gr._h, flags, f)), 256), dtype=torch.uint8, device=dense.device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lib.gcnb_spmm_bf16 if bf16 else lib.gcnb_spmm
    _lib.check(fn(gr._h, flags, dense.data_ptr(), dense.stride(0), f, bias.data_ptr() if bias is not None else None,
                  out.data_ptr(), f, ws.data_ptr(), ws.numel(), st), "spmm")
    return out


@pytest.mark.parametrize("f", [1, 3, 7, 8, 9, 32, 47, 100, 256])
def test_to_bf16_is_round_to_nearest_even_with_zero_padding(P, f):
    gen = torch.Generator(device=dev()).manual_seed(f)
    wide = torch.randn(777, f + 5, generator=gen, device=dev()) * 100
    wide[5, 0] = float("inf")
    wide[6, 0] = 1.00390625  # exactly between two bf16 values: ties to even
    src = wide[:, :f]  # row stride > width
    got = _to_bf16(P, src)
    assert torch.equal(got[:, :f].view(torch.int16), src.to(torch.bfloat16).view(torch.int16))
    assert (got[:, f:].float() == 0).all()


@pytest.mark.parametrize("spmm_kernel", ["auto", "rows", "group", ("group", 0), ("group", 1), ("group", 3), ("group", 13),
                                         ("group", 14), "stream", ("stream", 2)], indirect=True)
@pytest.mark.parametrize("f", [1, 3, 7, 8, 24, 32, 47, 64, 100, 256, 600, 1100])
def test_spmm_bf16_panel_equals_fp32_kernel_on_the_rounded_panel(P, f, spmm_kernel):
    """gcnb_spmm_bf16 gathers bf16 rows and accumulates in fp32: on a panel that is already bf16-representable
    it must agree with the fp32 kernel (1e-5: the two may pick different kernels, hence summation orders, for
    a width), forward, transposed, with bias + ReLU, on a power-law graph with split long rows."""
    from pygcn_b200 import _lib

    n = 5000
    src, dst = _powerlaw_graph(n, 50000, seed=f, hub_deg=2500)
    gr = P.Graph.from_edges(cu(src), cu(dst), n)
    assert gr.n_long_chunks > 0
    gen = torch.Generator(device=dev()).manual_seed(f)
    dense = torch.randn(n, f, generator=gen, device=dev())
    bias = torch.randn(f, generator=gen, device=dev())
    panel = _to_bf16(P, dense)
    rounded = panel[:, :f].float().contiguous()
    for flags, b in ((0, None), (_lib.SPMM_TRANSPOSE, None), (_lib.SPMM_RELU, bias)):
        got = _spmm_raw(P, gr, panel, f, True, flags, b)
        want = _spmm_raw(P, gr, rounded, f, False, flags, b)
        assert err(got, want.cpu().numpy()) < TOL, (f, flags)
        full = _spmm_raw(P, gr, dense, f, False, flags, b)  # the fp32 panel: the tier's bound
        assert err(got, full.cpu().numpy()) < TOL_BF16


@pytest.mark.parametrize("spmm_kernel", ["stream", ("stream", 2), ("stream", 4), ("stream", 8)], indirect=True)
@pytest.mark.parametrize("f", [4, 48, 100, 256, 300])
def test_stream_spmm_item_boundaries_and_epilogues(P, f, spmm_kernel):
    """The streaming SpMM (spmm_stream.cu) on a CSR whose rows start and end exactly on the 1024-entry item
    boundaries, span whole items, or hold one entry; every epilogue (bias, ReLU, dropout mask, accumulate), a strided
    output, fp32 and bf16 panels, forward and transposed -- against torch's CUDA CSR product; twice, bit-identical."""
    import ctypes

    from pygcn_b200 import _lib

    lib = _lib.load()
    rs = np.random.default_rng(f)
    lengths = np.concatenate([[1024, 1024, 512, 512, 2048, 1, 1023, 3072, 5, 4096 + 7], rs.integers(1, 40, 3000),
                              np.ones(700, np.int64), [9000], rs.integers(1, 6, 2000)])
    n = len(lengths)
    crow = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    col = rs.integers(0, n, crow[-1])
    val = rs.standard_normal(crow[-1]).astype(np.float32)
    csr = torch.sparse_csr_tensor(cu(crow), cu(col), cu(val), (n, n))
    gr = P.Graph.from_torch(csr)
    assert gr.nnz == crow[-1]
    dense = torch.randn(n, f, generator=torch.Generator(device=dev()).manual_seed(1), device=dev())
    ref = torch.sparse.mm(csr, dense)
    ref_t = torch.sparse.mm(csr.to_sparse_coo().t().coalesce().to_sparse_csr(), dense)
    got = _spmm_raw(P, gr, dense, f, False)
    assert err(got, ref.cpu().numpy()) < TOL
    assert torch.equal(got, _spmm_raw(P, gr, dense, f, False))
    assert err(_spmm_raw(P, gr, dense, f, False, _lib.SPMM_TRANSPOSE), ref_t.cpu().numpy()) < TOL
    bias = torch.randn(f, device=dev())
    assert err(_spmm_raw(P, gr, dense, f, False, _lib.SPMM_RELU, bias), torch.relu(ref + bias).cpu().numpy()) < TOL
    # accumulate into a strided output: out[:, 3:3+f] of a wider buffer += A dense
    wide = torch.randn(n, f + 9, device=dev())
    keep = wide.clone()
    view = wide[:, 4:4 + f]
    ws = torch.empty(max(int(lib.gcnb_spmm_workspace_bytes(gr._h, 0, f)), 256), dtype=torch.uint8, device=dev())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.gcnb_spmm(gr._h, _lib.SPMM_ACCUMULATE, dense.data_ptr(), f, f, None, view.data_ptr(), wide.stride(0),
                             ws.data_ptr(), ws.numel(), st), "spmm")
    assert err(view, (keep[:, 4:4 + f] + ref).cpu().numpy()) < TOL
    assert torch.equal(wide[:, :4], keep[:, :4]) and torch.equal(wide[:, 4 + f:], keep[:, 4 + f:])
    # bf16 panel
    panel = _to_bf16(P, dense)
    want = torch.sparse.mm(csr, panel[:, :f].float())
    assert err(_spmm_raw(P, gr, panel, f, True), want.cpu().numpy()) < TOL
    # fused dropout mask + ReLU through the layer API (mask epilogue of the streaming kernel)
    keepmask = torch.rand(n, f, device=dev()) > 0.3
    w = torch.eye(f, device=dev())
    out = P.gcn_layer(dense, gr, w, bias, relu=True, dropout_mask=keepmask, dropout_p=0.3)
    want = torch.relu(ref + bias) * keepmask / 0.7
    assert err(out, want.cpu().numpy()) < TOL


def test_stream_spmm_falls_back_when_a_row_is_empty(P):
    """A view with an empty row has no entry to carry the end-of-row tag: it takes the row kernels."""
    from pygcn_b200 import _lib

    lib = _lib.load()
    _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_STREAM, 2), "set_tuning")
    try:
        n = 300
        rs = np.random.default_rng(0)
        lengths = rs.integers(0, 30, n)
        lengths[7] = 0
        crow = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        col = rs.integers(0, n, crow[-1])
        val = rs.standard_normal(crow[-1]).astype(np.float32)
        csr = torch.sparse_csr_tensor(cu(crow), cu(col), cu(val), (n, n))
        gr = P.Graph.from_torch(csr)
        dense = torch.randn(n, 64, device=dev())
        assert err(_spmm_raw(P, gr, dense, 64, False), torch.sparse.mm(csr, dense).cpu().numpy()) < TOL
    finally:
        lib.gcnb_set_tuning(_lib.TUNE_SPMM_STREAM, 1)


def test_spmm_bf16_rejects_unaligned_panels_and_dense_route(P):
    from pygcn_b200 import _lib

    n = 300
    src, dst = _powerlaw_graph(n, 3000, seed=1)
    gr = P.Graph.from_edges(cu(src), cu(dst), n)
    panel = torch.zeros(n, 40, dtype=torch.bfloat16, device=dev())
    with pytest.raises(_lib.GcnbError):
        _spmm_raw(P, gr, panel[:, 1:34], 33, True)  # rows not 16-byte aligned
    dense_adj = torch.rand(64, 64, device=dev())
    gd = P.Graph.from_torch(dense_adj)
    assert gd.dense_route
    with pytest.raises(_lib.GcnbError):
        _spmm_raw(P, gd, torch.zeros(64, 8, dtype=torch.bfloat16, device=dev()), 8, True)
    # the layer falls back to the fp32-tier kernels there instead of failing
    layer = P.GraphConvolution(8, 8, precision="bf16").to(dev())
    x = torch.randn(64, 8, device=dev())
    ref = dense_adj @ (x @ layer.weight) + layer.bias
    assert err(layer(x, gd), ref.detach().cpu().numpy()) < TOL


@pytest.mark.parametrize("fin,fout,relu", [(64, 32, False), (64, 32, True), (16, 7, True), (100, 256, False), (33, 47, True)])
def test_layer_bf16_tier_within_2e_2_of_the_reference_lines(P, fin, fout, relu):
    """precision="bf16": output and every gradient within 2e-2 (norm-wise) of the reference's three lines in
    fp32 (pygcn/layers.py:33-36 + autograd), measured ~2e-3; dX only when asked."""
    n = 6000
    src, dst = _powerlaw_graph(n, 60000, seed=fin + fout, hub_deg=3000)
    gr = P.Graph.from_edges(cu(src), cu(dst), n)
    coo = gr.to_sparse_coo()
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout, fuse_relu=relu, precision="bf16").to(dev())
    x = cu(rng_inputs(1, (n, fin))).requires_grad_(True)
    g = cu(rng_inputs(2, (n, fout)))
    out = layer(x, gr)
    out.backward(g)
    w = layer.weight.detach().clone().requires_grad_(True)
    b = layer.bias.detach().clone().requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    o_pre = torch.spmm(coo, torch.mm(x2, w)) + b
    o_ref = F.relu(o_pre) if relu else o_pre
    gm = g
    if relu:
        # the ReLU mask comes from OUR forward output: entries within the tier's error of zero may sit on the
        # other side of it than the fp32 run's (a max-norm comparison of gradients under two different masks
        # measures the mask, not the arithmetic); everywhere else the masks must agree
        flips = (out.detach() > 0) != (o_pre.detach() > 0)
        assert flips.float().mean().item() < 0.01
        assert (not flips.any()) or o_pre.detach()[flips].abs().max().item() < TOL_BF16 * o_pre.detach().abs().max().item()
        gm = g * (out.detach() > 0)
    o_pre.backward(gm)
    errs = {"out": err(out, o_ref.detach().cpu().numpy()), "dW": err(layer.weight.grad, w.grad.cpu().numpy()),
            "db": err(layer.bias.grad, b.grad.cpu().numpy()), "dX": err(x.grad, x2.grad.cpu().numpy())}
    assert all(v < TOL_BF16 for v in errs.values()), errs
    assert errs["out"] > 1e-6  # the panel really is bf16 (an fp32 run would sit at ~1e-7)
    # without input grad no dX is produced
    x3 = x.detach().clone()
    layer.zero_grad()
    layer(x3, gr).backward(g)
    assert x3.grad is None and layer.weight.grad is not None


def test_gemm_generic_strides(P):
    rs = np.random.default_rng(7)
    for (m, n, k) in [(1, 1, 1), (130, 33, 70), (64, 32, 20000), (257, 100, 16), (1000, 7, 1433)]:
        a = rs.standard_normal((m, k)).astype(np.float32)
        b = rs.standard_normal((k, n)).astype(np.float32)
        ref = a.astype(np.float64) @ b.astype(np.float64)
        assert err(P.mm(cu(a), cu(b)), ref) < TOL
        assert err(P.mm(cu(a.T.copy()).t(), cu(b)), ref) < TOL  # A given transposed (dW = X^T dS)
        assert err(P.mm(cu(a), cu(b.T.copy()).t()), ref) < TOL  # B given transposed (dX = dS W^T)
        assert err(P.mm(cu(np.hstack([a, a]))[:, :k], cu(b)), ref) < TOL  # row stride > width


def test_gemm_tcgen05_3xtf32(P):
    """tcgen05 kind::tf32 kernels with the 3-term split hold the fp32 bar (1e-5 norm-wise) on every
    product shape of the layer: X W ("rows"), dS W^T ("rows", B transposed), X^T dS ("tn")."""
    gen = torch.Generator(device=dev()).manual_seed(5)

    def rnd(*shape):
        return torch.randn(*shape, generator=gen, device=dev())

    def nerr(c, ref):
        return ((c.double() - ref).abs().max() / ref.abs().max()).item()

    for (m, n, k, bt) in [(128, 32, 32, False), (100_000, 32, 64, False), (100_000, 64, 32, True), (5000, 47, 100, False),
                          (3000, 600, 16, False), (2708, 16, 1432, False), (20_000, 256, 608, False), (333, 7, 16, True)]:
        a = rnd(m, k)
        b = rnd(n, k).t() if bt else rnd(k, n)
        ref = a.double() @ b.double()
        assert nerr(P.mm(a, b, precision="tf32x3"), ref) < TOL, (m, n, k, bt)
        a2 = rnd(m, k + 8)[:, :k]  # row stride > K (column-slice view, pygcn/models.py:345)
        assert nerr(P.mm(a2, b, precision="tf32x3"), a2.double() @ b.double()) < TOL
    for (r, m, n) in [(32, 128, 32), (100_000, 64, 32), (2708, 1432, 16), (30_000, 256, 256), (4097, 100, 48), (7, 8, 4)]:
        x, y = rnd(r, m), rnd(r, n)
        ref = x.double().t() @ y.double()
        c = P.mm(x.t(), y, precision="tf32x3")
        assert nerr(c, ref) < TOL, (r, m, n)
        assert torch.equal(c, P.mm(x.t(), y, precision="tf32x3"))  # fixed-order split-K reduction
    # the layer on the tensor-core path against the CUDA-core path
    n = 50_000
    gr = P.Graph.from_edges(torch.randint(0, n, (n * 10,), generator=gen, device=dev(), dtype=torch.int32),
                            torch.randint(0, n, (n * 10,), generator=gen, device=dev(), dtype=torch.int32), n)
    x, g = rnd(n, 64), rnd(n, 32)
    res = {}
    for prec in ("fp32", "tf32x3"):
        torch.manual_seed(42)
        layer = P.GraphConvolution(64, 32, precision=prec).to(dev())
        xt = x.clone().requires_grad_(True)
        out = layer(xt, gr)
        out.backward(g)
        res[prec] = (out.detach(), layer.weight.grad, layer.bias.grad, xt.grad)
    for a_, b_ in zip(res["fp32"], res["tf32x3"]):
        assert ((a_ - b_).abs().max() / a_.abs().max()).item() < TOL


# ------------------------------------------------------------------ full-size properties (BASELINE config 1)
@pytest.fixture(scope="module")
def cbg(P):
    n, deg = 100_000, 100
    gen = torch.Generator(device=dev()).manual_seed(0)
    src = torch.randint(0, n, (n * deg // 2,), generator=gen, device=dev(), dtype=torch.int32)
    dst = torch.randint(0, n, (n * deg // 2,), generator=gen, device=dev(), dtype=torch.int32)
    return P.Graph.from_edges(src, dst, n), n


def test_cbg_full_size_properties(P, cbg):
    gr, n = cbg
    assert 9_900_000 < gr.nnz < 10_200_000 and gr.pattern_symmetric
    gen = torch.Generator(device=dev()).manual_seed(1)
    ones = torch.ones(n, 32, device=dev())
    # rows of D^-1(A+I) sum to one
    assert (P.spmm(gr, ones) - 1).abs().max().item() < 1e-5
    # adjoint identity <A x, y> == <x, A^T y>: ties the backward SpMM to the forward one
    x = torch.randn(n, 32, generator=gen, device=dev(), requires_grad=True)
    y = torch.randn(n, 32, generator=gen, device=dev())
    ax = P.spmm(gr, x)
    ax.backward(y)
    lhs = (ax.double() * y.double()).sum().item()
    rhs = (x.detach().double() * x.grad.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0) * 10
    # linearity of the layer without bias
    layer = P.GraphConvolution(64, 32, bias=False).to(dev())
    a = torch.randn(n, 64, generator=gen, device=dev())
    b = torch.randn(n, 64, generator=gen, device=dev())
    with torch.no_grad():
        lin = layer(2.0 * a - 3.0 * b, gr)
        comb = 2.0 * layer(a, gr) - 3.0 * layer(b, gr)
    assert ((lin - comb).abs().max() / comb.abs().max()).item() < TOL
    # run-to-run determinism (atomic-free accumulation)
    with torch.no_grad():
        assert torch.equal(layer(a, gr), layer(a, gr))
    # against torch's own CUDA spmm on the exported tensor (library cross-check, not the oracle)
    coo = gr.to_sparse_coo()
    ref = torch.spmm(coo, x.detach())
    assert ((ax.detach() - ref).abs().max() / ref.abs().max()).item() < TOL


def test_cbg_backward_matches_torch_autograd(P, cbg):
    gr, n = cbg
    coo = gr.to_sparse_coo().coalesce()
    gen = torch.Generator(device=dev()).manual_seed(3)
    x = torch.randn(n, 64, generator=gen, device=dev())
    g = torch.randn(n, 32, generator=gen, device=dev())
    torch.manual_seed(42)
    layer = P.GraphConvolution(64, 32).to(dev())
    xt = x.clone().requires_grad_(True)
    layer(xt, gr).backward(g)
    w = layer.weight.detach().clone().requires_grad_(True)
    b = layer.bias.detach().clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    (torch.spmm(coo, torch.mm(xr, w)) + b).backward(g)  # the reference's three lines on CUDA
    for mine, ref in ((layer.weight.grad, w.grad), (layer.bias.grad, b.grad), (xt.grad, xr.grad)):
        assert ((mine - ref).abs().max() / ref.abs().max()).item() < TOL


# ------------------------------------------------------------------ full-size properties (BASELINE configs 2, 3)
def _rmat_edges(n, n_edges, seed, a=0.57, b=0.19, c=0.19):
    """R-MAT endpoints on the device (a,b,c,d = .57,.19,.19,.05), folded into [0, n)."""
    gen = torch.Generator(device=dev()).manual_seed(seed)
    bits = max(1, (n - 1).bit_length())
    src = torch.zeros(n_edges, dtype=torch.int64, device=dev())
    dst = torch.zeros(n_edges, dtype=torch.int64, device=dev())
    for _ in range(bits):
        r = torch.rand(n_edges, generator=gen, device=dev())
        src = src * 2 + (r >= a + b).to(torch.int64)
        dst = dst * 2 + (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
    return (src % n).to(torch.int32), (dst % n).to(torch.int32)


def _shape_properties(P, gr, n, fin, fout, seed):
    """Size-independent checks of one layer on a big graph: rows of D^-1(A+I) sum to one, adjoint identity
    between the forward and the transposed SpMM, linearity, run-to-run determinism, agreement of the
    forward and of every gradient with torch's CUDA ops on the exported tensor."""
    gen = torch.Generator(device=dev()).manual_seed(seed)
    ones = torch.ones(n, 8, device=dev())
    # fp32 accumulation of deg terms 1/deg: rounding grows like sqrt(deg) (hub rows of the R-MAT graph
    # hold > 1e5 entries); 4 ulp * sqrt(max degree), never tighter than 1e-5
    assert (P.spmm(gr, ones) - 1).abs().max().item() < max(1e-5, 2.4e-7 * gr.max_degree ** 0.5)
    x = torch.randn(n, fout, generator=gen, device=dev(), requires_grad=True)
    y = torch.randn(n, fout, generator=gen, device=dev())
    ax = P.spmm(gr, x)
    ax.backward(y)
    lhs = (ax.double() * y.double()).sum().item()
    rhs = (x.detach().double() * x.grad.double()).sum().item()
    # both sides are sums of ~n*fout products that largely cancel: the scale of their rounding is
    # |Ax| |y| (Cauchy-Schwarz), not the value of the sum
    scale = ax.detach().double().norm().item() * y.double().norm().item()
    assert abs(lhs - rhs) <= 1e-6 * scale
    csr = gr.to_sparse_coo().coalesce().to_sparse_csr()
    ref = torch.sparse.mm(csr, x.detach())
    assert ((ax.detach() - ref).abs().max() / ref.abs().max()).item() < TOL
    del ref, ax
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout).to(dev())
    a = torch.randn(n, fin, generator=gen, device=dev())
    g = torch.randn(n, fout, generator=gen, device=dev())
    at = a.clone().requires_grad_(True)
    out = layer(at, gr)
    out.backward(g)
    with torch.no_grad():
        assert torch.equal(layer(a, gr), out)  # atomic-free: bit-stable
    w = layer.weight.detach().clone().requires_grad_(True)
    b = layer.bias.detach().clone().requires_grad_(True)
    ar = a.clone().requires_grad_(True)
    o_ref = torch.sparse.mm(csr, torch.mm(ar, w)) + b  # the reference's three lines on CUDA
    o_ref.backward(g)
    pairs = ((out, o_ref), (layer.weight.grad, w.grad), (layer.bias.grad, b.grad), (at.grad, ar.grad))
    errs = [((mine.detach() - r.detach()).abs().max() / r.detach().abs().max()).item() for mine, r in pairs]
    if max(errs) >= TOL:
        # two fp32 computations of a 2.4 M-row reduction may differ by more than 1e-5 from each other:
        # the fp64 run of the same formula arbitrates: our error must be below
        # max(1e-5, 2 x torch's own fp32 error)
        w64 = layer.weight.detach().double().requires_grad_(True)
        b64 = layer.bias.detach().double().requires_grad_(True)
        a64 = a.double().requires_grad_(True)
        o64 = torch.sparse.mm(csr.to(torch.float64), torch.mm(a64, w64)) + b64
        o64.backward(g.double())
        for (mine, r), r64 in zip(pairs, (o64, w64.grad, b64.grad, a64.grad)):
            scale = r64.detach().abs().max()
            mine_err = ((mine.detach().double() - r64.detach()).abs().max() / scale).item()
            ref_err = ((r.detach().double() - r64.detach()).abs().max() / scale).item()
            assert mine_err < max(TOL, 2 * ref_err), (mine_err, ref_err)


def test_reddit_shape_full_size_properties(P):
    """BASELINE configs[2]: N=232 965, ~114.6 M stored entries (uniform), 602 -> 256: wide-row SpMM
    (warp-per-row kernel, two float4 chunks per lane) and compute-relevant tcgen05 GEMMs."""
    n = 232_965
    gen = torch.Generator(device=dev()).manual_seed(0)
    src = torch.randint(0, n, (n * 246,), generator=gen, device=dev(), dtype=torch.int32)
    dst = torch.randint(0, n, (n * 246,), generator=gen, device=dev(), dtype=torch.int32)
    gr = P.Graph.from_edges(src, dst, n)
    del src, dst
    assert 113_000_000 < gr.nnz < 116_000_000 and gr.pattern_symmetric
    _shape_properties(P, gr, n, 602, 256, seed=5)


def test_products_shape_full_size_properties(P):
    """BASELINE configs[3]: N=2 449 029, ~62 M stored entries, R-MAT (power-law: every degree bin and
    the long-row split are populated), 100 -> 256 and the 256 -> 47 output layer."""
    n = 2_449_029
    src, dst = _rmat_edges(n, 31_000_000, seed=7)
    gr = P.Graph.from_edges(src, dst, n)
    del src, dst
    assert 40_000_000 < gr.nnz < 66_000_000 and gr.n_long_chunks > 0 and gr.max_degree > 1024
    assert all(b > 0 for b in gr.bin_rows[1:])
    _shape_properties(P, gr, n, 100, 256, seed=6)
    _shape_properties(P, gr, n, 256, 47, seed=7)


@pytest.mark.parametrize("world", [1, 3, 4])
def test_partitioned_graph_build_is_bit_identical_to_cutting_the_full_graph(P, world):
    """dist.build_partitioned (no rank ever holds the whole adjacency) == row blocks cut out of the
    single-GPU Graph.from_edges result: same rowptr / columns / fp32 values bit for bit, for A and A^T.
    Ranks are emulated one after the other on this GPU; the all-gather of the row sums is a concatenation."""
    from pygcn_b200 import dist as D

    n = 4000
    src, dst = _powerlaw_graph(n, 30000, seed=30 + world, hub_deg=1500)
    src = np.concatenate([src, src[:500]])  # duplicate edges: counts > 1
    dst = np.concatenate([dst, dst[:500]])
    s, d = cu(src), cu(dst)
    full = P.Graph.from_edges(s, d, n)
    bounds = D.partition_rows_by_nnz(full.csr()[0].cpu(), world)
    pad = D.DistGraph.padded_rows(bounds)
    stage1 = [D._partition_counts(s, d, n, bounds[r], bounds[r + 1]) for r in range(world)]
    rowsum_global = torch.cat([t[4] for t in stage1])
    for r in range(world):
        ref = D.DistGraph.from_graph(full, r, world, bounds, split=False)
        lrp, lcol, a, rows, _ = stage1[r]
        mine = D._partition_blocks(r, world, bounds, pad, lrp, lcol, a, rows, rowsum_global.clone(), full.nnz)
        for got, want in ((mine.fwd_remote, ref.fwd_remote or ref.fwd_diag), (mine.bwd_remote, ref.bwd_remote or ref.bwd_diag)):
            assert got.shape == want.shape and got.nnz == want.nnz
            for x, y in zip(got.csr(), want.csr()):
                assert torch.equal(x, y)
    # and the single-rank entry point end to end
    dg = D.build_partitioned(s, d, n, 0, 1)
    assert dg.fwd_remote.nnz == full.nnz and torch.equal(dg.fwd_remote.csr()[2], full.csr()[2])


# ------------------------------------------------------------------ multi-GPU building blocks on one GPU
@pytest.mark.parametrize("world", [1, 3, 8])
def test_row_partition_blocks_emulated_on_one_gpu(P, world):
    """All ranks of the 1-D row partition emulated sequentially on one device (pygcn_b200/dist.py):
    sum_q A[p,q] S_q with the ACCUMULATE epilogue must equal the rows of the full SpMM, forward and
    transposed, and the full layer backward must equal the sum of the ranks' contributions."""
    from pygcn_b200 import dist as D

    n = 5000
    src, dst = _powerlaw_graph(n, 40000, seed=world, hub_deg=2000)
    gr = P.Graph.from_edges(cu(src), cu(dst), n)
    bounds = D.partition_rows_by_nnz(gr.csr()[0].cpu(), world)
    gen = torch.Generator(device=dev()).manual_seed(world)
    x = torch.randn(n, 24, generator=gen, device=dev())
    g = torch.randn(n, 12, generator=gen, device=dev())
    torch.manual_seed(42)
    layer = P.GraphConvolution(24, 12, precision="fp32").to(dev())
    xt = x.clone().requires_grad_(True)
    ref_out = layer(xt, gr)
    ref_out.backward(g)
    ops = D.CudaOps("fp32")
    w, b = layer.weight.detach(), layer.bias.detach()
    dgs = [D.DistGraph.from_graph(gr, p, world, bounds, split=True) for p in range(world)]
    assert sum(dg.nnz_local for dg in dgs) == gr.nnz
    pad = dgs[0].pad_rows

    def gathered(parts):  # what the all-gather of the padded slots delivers (padding poisoned with NaN)
        buf = torch.full((world * pad, parts[0].shape[1]), float("nan"), device=dev())
        for q, t in enumerate(parts):
            buf[q * pad: q * pad + t.shape[0]] = t
        return buf

    s_all = gathered([ops.gemm(x[bounds[q]:bounds[q + 1]], w) for q in range(world)])
    g_all = gathered([g[bounds[q]:bounds[q + 1]] for q in range(world)])
    dw_sum = torch.zeros_like(w)
    db_sum = torch.zeros_like(b)
    for p, dg in enumerate(dgs):
        r0, r1 = bounds[p], bounds[p + 1]
        out = torch.empty(r1 - r0, 12, device=dev())
        if world == 1:
            ops.spmm_block(dg.fwd_diag, s_all[: r1 - r0], out, False, b, False)
        else:
            ops.spmm_block(dg.fwd_diag, s_all[p * pad: p * pad + pad], out, False)
            ops.spmm_block(dg.fwd_remote, s_all, out, True, b, False)
        assert ((out - ref_out[r0:r1].detach()).abs().max() / ref_out.abs().max()).item() < TOL
        ds = torch.empty(r1 - r0, 12, device=dev())
        ops.spmm_block(dg.bwd_diag, g_all[p * pad: p * pad + pad], ds, False)
        if world > 1:
            ops.spmm_block(dg.bwd_remote, g_all, ds, True)
        dw_sum += ops.gemm(x[r0:r1].t(), ds)
        staged = torch.empty(pad, 12, device=dev())
        db_p, _ = ops.colsum(g[r0:r1].contiguous(), None, staged)
        assert torch.equal(staged[: r1 - r0], g[r0:r1])
        db_sum += db_p
        dx = ops.gemm(ds, w.t())
        assert ((dx - xt.grad[r0:r1]).abs().max() / xt.grad.abs().max()).item() < TOL
    assert ((dw_sum - layer.weight.grad).abs().max() / layer.weight.grad.abs().max()).item() < TOL
    assert ((db_sum - layer.bias.grad).abs().max() / layer.bias.grad.abs().max()).item() < TOL
    if world > 1:  # unsplit variant: the whole row block over the all-gathered panel
        dg = D.DistGraph.from_graph(gr, 1, world, bounds, split=False)
        r0, r1 = bounds[1], bounds[2]
        out = torch.empty(r1 - r0, 12, device=dev())
        ops.spmm_block(dg.fwd_remote, s_all, out, False, b, False)
        assert ((out - ref_out[r0:r1].detach()).abs().max() / ref_out.abs().max()).item() < TOL
        assert dg.fwd_diag is None and dg.nnz_local == dgs[1].nnz_local
        assert D.DistGraph.from_graph(gr, 1, world, bounds).split == (dgs[1].fwd_diag.nnz >= 0.6 * dgs[1].nnz_local)
    with pytest.raises(RuntimeError):  # blocks carry no transpose
        P.spmm(dgs[0].fwd_diag, torch.zeros(dgs[0].fwd_diag.n_cols, 4, device=dev(), requires_grad=True)).sum().backward()
    # single-process DistGraphConvolution (world 1) equals the plain layer
    if world == 1:
        dl = D.DistGraphConvolution(24, 12, precision="fp32").to(dev())
        dl.inner.load_state_dict(layer.state_dict())
        xt2 = x.clone().requires_grad_(True)
        o2 = dl(xt2, dgs[0])
        o2.backward(g)
        assert torch.equal(o2, ref_out) and torch.allclose(xt2.grad, xt.grad, rtol=0, atol=0)
        assert torch.equal(dl.inner.weight.grad, layer.weight.grad)


def test_halo_pack_and_gemm_ex_through_the_c_abi(P):
    """Two entry points of round 2 called directly: gcnb_halo_pack (rows every peer reads -> one send buffer, vector
    and scalar copies) and gcnb_gemm_ex (product + bias + ReLU; the TMA-fed tcgen05 kernel's fused epilogue on the big
    shape, the in-place pass on the small one)."""
    import ctypes

    from pygcn_b200 import _lib

    lib = _lib.load()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device=dev()).manual_seed(3)
    for f, ld in ((256, 256), (100, 104), (47, 47)):
        panel = torch.randn(5000, ld, generator=gen, device=dev())
        rows = [torch.randint(0, 5000, (k,), generator=gen, device=dev(), dtype=torch.int32) for k in (700, 0, 1234)]
        i64 = ctypes.c_int64 * 4
        h = ctypes.c_void_p()
        send = i64(700, 0, 0, 1234)  # this rank is 1 of 4
        cat = torch.cat(rows).contiguous()
        _lib.check(lib.gcnb_halo_create(1, 4, send, cat.data_ptr(), i64(0, 0, 0, 0), i64(0, 0, 0, 0), st, ctypes.byref(h)), "halo_create")
        assert lib.gcnb_halo_send_rows(h) == 1934
        buf = torch.full((1934, f), float("nan"), device=dev())
        _lib.check(lib.gcnb_halo_pack(h, panel.data_ptr(), ld, f, buf.data_ptr(), st), "halo_pack")
        assert torch.equal(buf, panel[cat.long(), :f])
        lib.gcnb_halo_free(h)
    for m, k, n in ((300000, 100, 256), (500, 9, 7)):
        a = torch.randn(m, k, generator=gen, device=dev())
        w = torch.randn(k, n, generator=gen, device=dev())
        b = torch.randn(n, generator=gen, device=dev())
        out = torch.empty(m, n, device=dev())
        ws = torch.empty(max(int(lib.gcnb_gemm_workspace_bytes(m, n, k, _lib.GEMM_AUTO)), 256), dtype=torch.uint8, device=dev())
        for relu in (0, 1):
            _lib.check(lib.gcnb_gemm_ex(m, n, k, a.data_ptr(), k, 1, w.data_ptr(), n, 1, out.data_ptr(), n, b.data_ptr(), relu,
                                        _lib.GEMM_AUTO, ws.data_ptr(), ws.numel(), st), "gemm_ex")
            want = (a.double() @ w.double() + b.double())
            want = torch.relu(want) if relu else want
            assert err(out, want.float().cpu().numpy()) < TOL


# ------------------------------------------------------------------ error behaviour
def test_errors(P, golden):
    layer = P.GraphConvolution(4, 3)
    with pytest.raises(RuntimeError):  # CPU tensors: explicit error, no fallback
        layer(torch.zeros(5, 4), torch.eye(5))
    layer = layer.to(dev())
    adj = torch.eye(5, device=dev())
    with pytest.raises(RuntimeError):  # inner dimension mismatch
        layer(torch.zeros(5, 6, device=dev()), adj)
    with pytest.raises(RuntimeError):  # rows of x vs adj
        layer(torch.zeros(6, 4, device=dev()), adj)
    with pytest.raises(RuntimeError):  # dtype
        layer(torch.zeros(5, 4, device=dev(), dtype=torch.float64), adj)
    out = layer(torch.zeros(5, 4, device=dev()), adj)
    assert out.shape == (5, 3) and torch.allclose(out, layer.bias.expand(5, 3))
