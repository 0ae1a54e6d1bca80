"""Pin the oracle (oracle/gcn_oracle.py + .c) against the reference's own outputs.

The golden .npz files were produced by tests/golden/make_golden.py, which runs
/root/reference/pygcn/{layers,utils}.py.  Index/byte work must match bit-exactly;
float work within 1e-5 norm-wise and within 1e-6 of the fp64 run.
"""
import numpy as np
import pytest

from conftest import rng_inputs
from oracle import gcn_oracle as O

TOL = 1e-5


def test_cora_pipeline_bit_exact(golden):
    g = golden("cora_pipeline.npz")
    e = g["edges"]
    idx, val = O.build_normalized_adjacency(e[:, 0], e[:, 1], int(g["n"]))
    assert idx.dtype == np.int64 and val.dtype == np.float32
    assert np.array_equal(idx, g["indices"])
    assert np.array_equal(val.view(np.uint32), g["values"].view(np.uint32))
    assert idx.shape[1] == 13264
    # row-major, columns ascending (what the device CSR build must reproduce)
    key = idx[0] * int(g["n"]) + idx[1]
    assert np.all(np.diff(key) > 0)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_small_pipelines_bit_exact(golden, k):
    g = golden("pipeline_small.npz")
    e = g[f"edges{k}"]
    idx, val = O.build_normalized_adjacency(e[:, 0], e[:, 1], int(g[f"n{k}"]))
    assert np.array_equal(idx, g[f"indices{k}"])
    assert np.array_equal(val.view(np.uint32), g[f"values{k}"].view(np.uint32))


def test_init_bounds(golden):
    g = golden("init.npz")
    wb, bb = O.init_bounds(64, 32)
    assert np.abs(g["w_64_32"]).max() <= wb and np.abs(g["w_64_32"]).max() > 0.95 * wb
    assert np.abs(g["b_64_32"]).max() <= bb
    assert str(g["repr"]) == "GraphConvolution (64 -> 32)"
    assert list(g["state_keys"]) == ["bias", "weight"]


def _check_layer(c, prefix, x, idx, val, n, g, has_bias=True, dx_key="dX", use_c=True):
    w = c[f"{prefix}/weight"]
    b = c[f"{prefix}/bias"] if has_bias else None
    fwd = O.c_layer_forward if use_c else O.layer_forward
    bwd = O.c_layer_backward if use_c else O.layer_backward
    _, out = fwd(x, w, b, idx, val, n)
    dw, db, dx, _ = bwd(x, w, has_bias, idx, val, x.shape[0], g)
    _, out64 = fwd(x, w, b, idx, val, n, dtype=np.float64)
    assert O.normwise_err(out, c[f"{prefix}/out"]) < TOL
    assert O.normwise_err(out, out64) < 2e-6
    assert O.normwise_err(c[f"{prefix}/out"], out64) < 2e-6
    assert O.normwise_err(dw, c[f"{prefix}/dW"]) < TOL
    if has_bias:
        assert O.normwise_err(db, c[f"{prefix}/db"]) < TOL
    return dx


def test_layer_cora(golden):
    p = golden("cora_pipeline.npz")
    c = golden("layer_cases.npz")
    n = int(p["n"])
    idx, val = p["indices"], p["values"]
    x = rng_inputs(1, (n, 1433))
    assert abs(float(x.astype(np.float64).sum()) - float(c["cora_l1/x_checksum"])) < 1e-9
    dx = _check_layer(c, "cora_l1", x, idx, val, n, rng_inputs(2, (n, 16)))
    assert O.normwise_err(dx[:32], c["cora_l1/dX_head"]) < TOL
    dx = _check_layer(c, "cora_l2", rng_inputs(3, (n, 16)), idx, val, n, rng_inputs(4, (n, 7)))
    assert O.normwise_err(dx, c["cora_l2/dX"]) < TOL


@pytest.mark.parametrize("use_c", [True, False])
def test_layer_ragged_duplicates(golden, use_c):
    c = golden("layer_cases.npz")
    n = int(c["ragged/n"])
    idx = np.vstack([c["ragged/rows"], c["ragged/cols"]]).astype(np.int64)
    xw = rng_inputs(12, (n, 40))
    dx = _check_layer(c, "ragged", np.ascontiguousarray(xw[:, :33]), idx, c["ragged/vals"], n,
                      rng_inputs(13, (n, 7)), use_c=use_c)
    full = np.zeros((n, 40), dtype=np.float32)
    full[:, :33] = dx
    assert O.normwise_err(full, c["ragged/dXfull"]) < TOL


def test_layer_dense_adj(golden):
    c = golden("layer_cases.npz")
    ad = c["dense/adj"]
    r, cc = np.nonzero(ad)
    idx = np.vstack([r, cc]).astype(np.int64)
    dx = _check_layer(c, "dense", rng_inputs(22, (96, 8)), idx, ad[r, cc], 96, rng_inputs(23, (96, 32)))
    assert O.normwise_err(dx, c["dense/dX"]) < TOL


def test_layer_nobias(golden):
    c = golden("layer_cases.npz")
    n = int(c["ragged/n"])
    idx = np.vstack([c["ragged/rows"], c["ragged/cols"]]).astype(np.int64)
    _check_layer(c, "nobias_csr", rng_inputs(31, (n, 16)), idx, c["ragged/vals"], n, rng_inputs(32, (n, 16)),
                 has_bias=False)


def test_transpose_and_bins(golden):
    p = golden("cora_pipeline.npz")
    n = int(p["n"])
    idx, val = p["indices"], p["values"]
    rowptr = O.coo_to_csr(idx, n)
    assert rowptr[-1] == idx.shape[1]
    t_rowptr, t_col, t_val = O.transpose_csr(idx, val, n, n)
    # pattern of the Cora pipeline is symmetric, values are not
    assert np.array_equal(t_rowptr, rowptr) and np.array_equal(t_col, idx[1])
    assert not np.array_equal(t_val, val)
    b, counts = O.degree_bins(rowptr)
    assert counts.sum() == n and counts[0] == 0
    deg = np.diff(rowptr)
    assert deg.min() == 2 and deg.max() == 169


def test_stack3_relu(golden):
    """models.GeneratorGCN restated on the oracle: relu(gc(x, adj)) x3 (models.py:103-111)."""
    p = golden("pipeline_small.npz")
    c = golden("layer_cases.npz")
    idx, val, n = p["indices1"], p["values1"], 257
    x = rng_inputs(51, (n, 8))
    g = rng_inputs(52, (n, 32))
    acts, h = [], x
    for k in (1, 2, 3):
        _, o = O.c_layer_forward(h, c[f"stack3/param:gc{k}.weight"], c[f"stack3/param:gc{k}.bias"], idx, val, n)
        acts.append((h, o))
        h = np.maximum(o, 0)
    assert O.normwise_err(h, c["stack3/out"]) < TOL
    grad = g
    for k in (3, 2, 1):
        hin, o = acts[k - 1]
        grad = O.relu_backward(grad, o)
        dw, db, dx, _ = O.c_layer_backward(hin, c[f"stack3/param:gc{k}.weight"], True, idx, val, n, grad)
        assert O.normwise_err(dw, c[f"stack3/grad:gc{k}.weight"]) < TOL
        assert O.normwise_err(db, c[f"stack3/grad:gc{k}.bias"]) < TOL
        grad = dx
    assert O.normwise_err(grad, c["stack3/dX"]) < TOL


def test_ref_port_matches_golden(golden):
    """oracle/ref_layer_torch.py (what the reference arm of bench.py times) reproduces the
    reference's own outputs."""
    import torch

    from oracle import ref_layer_torch as R

    c = golden("layer_cases.npz")
    p = golden("cora_pipeline.npz")
    n = int(p["n"])
    adj = R.make_reference_adj(torch.from_numpy(p["indices"]), torch.from_numpy(p["values"]), (n, n))
    out, dw, db = R.reference_layer_fwdbwd(
        torch.from_numpy(rng_inputs(3, (n, 16))), torch.from_numpy(c["cora_l2/weight"]),
        torch.from_numpy(c["cora_l2/bias"]), adj, torch.from_numpy(rng_inputs(4, (n, 7))))
    # same library calls as the reference; only the BLAS thread count may differ from the golden run
    assert O.normwise_err(out.numpy(), c["cora_l2/out"]) < 1e-6
    assert O.normwise_err(dw.numpy(), c["cora_l2/dW"]) < 1e-6 and O.normwise_err(db.numpy(), c["cora_l2/db"]) < 1e-6
    csr = R.make_reference_adj(torch.from_numpy(p["indices"]), torch.from_numpy(p["values"]), (n, n), "csr")
    out2, _, _ = R.reference_layer_fwdbwd(
        torch.from_numpy(rng_inputs(3, (n, 16))), torch.from_numpy(c["cora_l2/weight"]),
        torch.from_numpy(c["cora_l2/bias"]), csr, torch.from_numpy(rng_inputs(4, (n, 7))))
    assert O.normwise_err(out2.numpy(), c["cora_l2/out"]) < TOL


def test_cbg_adjacency_oracle_matches_reference_load_adj(golden):
    """oracle.cbg_adjacency == the reference's own utils.load_adj (pygcn/utils.py:93-132) on the
    reference-generated fixture: same float32 matrix bit for bit (both accumulate in float64)."""
    c = golden("load_adj.npz")
    avg, adj = O.cbg_adjacency(c["hours"])
    assert adj.dtype == np.float32 and adj.shape == c["adj"].shape
    assert O.normwise_err(adj, c["adj"]) < 1e-7
    assert np.array_equal(adj, adj.T)


BN_CASES = ["cbg32", "odd7", "plain16", "deadcol", "two_rows"]


@pytest.mark.parametrize("name", BN_CASES)
def test_fresh_batchnorm_oracle_matches_reference_apply_bn(golden, name):
    """oracle.fresh_batchnorm_* == the reference's own GCN.apply_bn(F.relu(y)) (pygcn/models.py:41-45, 49, 53) and its
    autograd backward on the reference-generated fixture (apply_bn.npz): <= 1e-5 norm-wise, fp32 and fp64 restatement
    (torch's own fp32 result sits 3e-6 from exact arithmetic on these cases; the 2-row batch differentiates through a
    catastrophic cancellation -- var ~ the squared half-difference -- where torch's fp32 gradient is 4e-5 off)."""
    c = golden("apply_bn.npz")
    y, g, relu = c[name + "/y"], c[name + "/g"], bool(c[name + "/relu"])
    for dtype, tol in ((np.float32, TOL), (np.float64, 1e-4 if name == "two_rows" else TOL)):
        out, mean, rstd = O.fresh_batchnorm_forward(y, relu, 1e-5, dtype)
        assert out.shape == y.shape and mean.shape == (y.shape[1],) and rstd.shape == (y.shape[1],)
        assert O.normwise_err(out, c[name + "/out"]) < tol
        assert O.normwise_err(O.fresh_batchnorm_backward(y, g, relu, 1e-5, dtype), c[name + "/dy"]) < tol
    if name == "deadcol":  # ReLU zeroes column 3 entirely: var = 0, the output and the gradient of that column are 0
        assert not c[name + "/out"][:, 3].any() and not c[name + "/dy"][:, 3].any()
        assert not O.fresh_batchnorm_forward(y, relu)[0][:, 3].any()
