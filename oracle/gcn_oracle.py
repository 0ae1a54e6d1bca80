"""CPU oracle for the GCN-layer hot path of LinChen-65/pygcn.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pygcn_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or the timed baseline -- never as the thing shipped.

It restates, in plain numpy (and, for the float kernels at full size, in the C
file next to it), the algorithm the reference executes on this path:

  * adjacency pipeline   pygcn/utils.py:360-368 (edge list -> symmetric A + I),
                         pygcn/utils.py:390-397 (``normalize``: D^-1 M in fp64),
                         pygcn/utils.py:407-414 (``sparse_mx_to_torch_sparse_tensor``:
                         COO, int64 [2,nnz] indices, fp32 values)
  * layer forward        pygcn/layers.py:32-38  (mm, spmm, + bias)
  * layer backward       autograd of the above
  * parameter init       pygcn/layers.py:23-29
  * ReLU -> fresh BatchNorm  pygcn/models.py:41-45, 49, 53 (``GCN.apply_bn``)

The arithmetic of the reference lives in third-party libraries that are not
vendored under /root/reference: PyTorch (``torch.mm`` / ``torch.spmm``; unpinned
in setup.py:12-15, 2.11.0+cu128 in this image) and scipy/numpy (unpinned;
scipy 1.18.1 / numpy 2.3.5 here).  Parity pinning: the reference has no golden
vectors of its own, so this oracle is pinned against outputs of
the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (imports /root/reference/pygcn/layers.py and
utils.py) and committed under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function here against them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------
# adjacency pipeline (index / integer / fp64 work: bit-exact bar)
# --------------------------------------------------------------------------
def edges_to_counts(src, dst, n):
    """Edge list -> canonical CSR of multiplicities (fp32), rows/cols ascending.

    Follows pygcn/utils.py:360-362: a COO matrix of ones whose duplicate
    entries are summed when it is converted to CSR.
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    key = src * n + dst
    uniq, cnt = np.unique(key, return_counts=True)
    return (uniq // n).astype(np.int64), (uniq % n).astype(np.int64), cnt.astype(np.float32)


def symmetrize_max(row, col, val, n):
    """A <- A + A^T*[A^T > A] - A*[A^T > A]  ==  element-wise max(A, A^T).

    Follows pygcn/utils.py:365.  Entries whose result is zero are not stored
    (scipy's binary CSR ops drop zeros), output is (row, col) ascending.
    """
    key = row * n + col
    tkey = col * n + row
    allk = np.union1d(key, tkey)
    a = np.zeros(allk.shape[0], dtype=np.float32)
    at = np.zeros(allk.shape[0], dtype=np.float32)
    a[np.searchsorted(allk, key)] = val
    at[np.searchsorted(allk, tkey)] = val
    m = at > a
    # same operation order as the reference expression: (A + A^T.M) - A.M, in fp32
    out = (a + np.where(m, at, np.float32(0))).astype(np.float32) - np.where(m, a, np.float32(0))
    keep = out != 0
    allk, out = allk[keep], out[keep]
    return allk // n, allk % n, out.astype(np.float32)


def add_identity(row, col, val, n):
    """M <- A + I in fp64 (``sp.eye`` is float64, so the sum upcasts), utils.py:368."""
    key = row * n + col
    dkey = np.arange(n, dtype=np.int64) * (n + 1)
    allk = np.union1d(key, dkey)
    out = np.zeros(allk.shape[0], dtype=np.float64)
    out[np.searchsorted(allk, key)] += val.astype(np.float64)
    out[np.searchsorted(allk, dkey)] += 1.0
    keep = out != 0
    allk, out = allk[keep], out[keep]
    return allk // n, allk % n, out


def normalize_rows(row, col, val, n):
    """Row-normalise: D^-1 M with D = row sums, fp64 throughout (utils.py:390-397).

    rowsum accumulates the stored entries of each row in storage order;
    r_inv = rowsum ** -1 with inf -> 0; value = r_inv[row] * value.
    """
    val = np.asarray(val, dtype=np.float64)
    rowsum = np.bincount(row, weights=val, minlength=n)  # sequential, storage order
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -1.0)
    r_inv[np.isinf(r_inv)] = 0.0
    return row, col, r_inv[row] * val


def to_torch_coo_layout(row, col, val):
    """(int64 [2, nnz] indices, fp32 [nnz] values), utils.py:407-414."""
    idx = np.vstack((row, col)).astype(np.int64)
    return idx, np.asarray(val).astype(np.float32)


def build_normalized_adjacency(src, dst, n):
    """Whole pipeline the (commented) Cora loader runs: utils.py:360-368,376."""
    r, c, v = edges_to_counts(src, dst, n)
    r, c, v = symmetrize_max(r, c, v, n)
    r, c, v = add_identity(r, c, v, n)
    r, c, v = normalize_rows(r, c, v, n)
    return to_torch_coo_layout(r, c, v)


def cbg_adjacency(hourly_visits):
    """pygcn/utils.py:108-129 (`load_adj`): average the hourly POI x CBG visit matrices
    (`avg_array += poi_cbg_visits_list[i]; avg_array /= num_hours`, :117-120), then
    `adj[i][j] = np.sum(avg_array[:, i] * avg_array[:, j])` (:124-128) in float64, returned as
    `torch.FloatTensor(adj)` (:131), i.e. float32.  hourly_visits: [hours, n_poi, n_cbg]."""
    hv = np.asarray(hourly_visits, dtype=np.float64)
    avg = np.zeros(hv.shape[1:], dtype=np.float64)
    for h in range(hv.shape[0]):
        avg += hv[h]
    avg /= hv.shape[0]
    adj = avg.T @ avg  # the reference's double loop, one dot product per (i, j)
    return avg, adj.astype(np.float32)


def coo_to_csr(idx, n_rows):
    """Row pointer of a row-sorted COO (what the device build must reproduce)."""
    counts = np.bincount(idx[0], minlength=n_rows)
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr


def transpose_csr(idx, val, n_rows, n_cols):
    """CSR of A^T: entries ordered by (col, row) ascending, stable."""
    order = np.lexsort((idx[0], idx[1]))
    t_row = idx[1][order]
    t_col = idx[0][order]
    t_val = val[order]
    counts = np.bincount(t_row, minlength=n_cols)
    t_rowptr = np.zeros(n_cols + 1, dtype=np.int64)
    np.cumsum(counts, out=t_rowptr[1:])
    return t_rowptr, t_col, t_val


def degree_bins(rowptr, edges=(0, 1, 9, 33, 1025)):
    """Row-length histogram bins the device schedule uses.

    bin b holds rows with edges[b] <= deg < edges[b+1]; last bin is open.
    Returns (bin id per row, count per bin).  The bin boundaries are part of
    the C-ABI (include/gcnb200.h: GCNB_BIN_*).
    """
    deg = np.diff(rowptr)
    b = np.searchsorted(np.asarray(edges[1:]), deg, side="right")
    return b.astype(np.int32), np.bincount(b, minlength=len(edges)).astype(np.int64)


# --------------------------------------------------------------------------
# parameter init (layers.py:23-29) -- bounds only; the draws themselves come
# from torch's generator and are pinned by the golden fixture.
# --------------------------------------------------------------------------
def init_bounds(in_features, out_features):
    """kaiming_uniform_(a=0) on a [in,out] tensor uses fan_in = size(1) = out."""
    w_bound = np.sqrt(2.0) * np.sqrt(3.0 / out_features)
    b_bound = 1.0 / np.sqrt(out_features)
    return w_bound, b_bound


# --------------------------------------------------------------------------
# layer arithmetic (floating point: tolerance bar, see tests)
# --------------------------------------------------------------------------
def layer_forward(x, w, b, idx, val, n_rows, dtype=np.float32):
    """support = x @ w ; out = A @ support ; + b   (layers.py:33-36).

    The sparse product is the per-stored-entry axpy the reference's COO path
    performs: out[row] += val * support[col], in storage order.
    """
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    support = x @ w
    out = np.zeros((n_rows, w.shape[1]), dtype=dtype)
    np.add.at(out, idx[0], np.asarray(val, dtype=dtype)[:, None] * support[idx[1]])
    if b is not None:
        out = out + np.asarray(b, dtype=dtype)
    return support, out


def layer_backward(x, w, has_bias, idx, val, n_cols, g, dtype=np.float32):
    """Autograd of layers.py:33-36.

    db = sum_rows g ; d_support = A^T g ; dW = x^T d_support ; dX = d_support w^T
    """
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    g = np.asarray(g, dtype=dtype)
    db = g.sum(axis=0) if has_bias else None
    ds = np.zeros((n_cols, g.shape[1]), dtype=dtype)
    np.add.at(ds, idx[1], np.asarray(val, dtype=dtype)[:, None] * g[idx[0]])
    dw = x.T @ ds
    dx = ds @ w.T
    return dw, db, dx, ds


def relu_backward(g, out):
    """threshold_backward of the caller's F.relu (models.py:49,53,56)."""
    return np.where(out > 0, g, np.zeros_like(g))


def fresh_batchnorm_forward(y, relu=True, eps=1e-5, dtype=np.float64):
    """`GCN.apply_bn` as the models call it (pygcn/models.py:41-45, 49, 53): a NEW nn.BatchNorm1d(F) per call --
    affine weight 1 / bias 0, training mode, so batch statistics with the BIASED variance -- on F.relu(y):
    a = relu(y); out = (a - mean_rows(a)) / sqrt(var_rows(a) + eps).  Returns (out, mean, rstd)."""
    a = np.asarray(y, dtype=dtype)
    if relu:
        a = np.maximum(a, 0)
    mean = a.mean(axis=0)
    var = ((a - mean) ** 2).mean(axis=0)
    rstd = 1.0 / np.sqrt(var + dtype(eps))
    return (a - mean) * rstd, mean, rstd


def fresh_batchnorm_backward(y, g, relu=True, eps=1e-5, dtype=np.float64):
    """Autograd of the above: with xh = out, da = rstd * (g - mean_rows(g) - xh * mean_rows(g * xh)) (the batch
    statistics depend on the input), then threshold_backward of the F.relu: dy = da * [y > 0]."""
    xh, _mean, rstd = fresh_batchnorm_forward(y, relu, eps, dtype)
    g = np.asarray(g, dtype=dtype)
    da = rstd * (g - g.mean(axis=0) - xh * (g * xh).mean(axis=0))
    if relu:
        da = np.where(np.asarray(y) > 0, da, np.zeros_like(da))
    return da


def normwise_err(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY.md 8(d)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.max(np.abs(b))
    if denom == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / denom)


# --------------------------------------------------------------------------
# C restatement (same formulas, sized for the full BASELINE shapes)
# --------------------------------------------------------------------------
_C_LIB = None


def c_lib_path():
    return os.path.join(_HERE, "_build", "libgcn_oracle.so")


def build_c(force=False):
    """gcc -O2 -fopenmp oracle/gcn_oracle.c -> oracle/_build/libgcn_oracle.so"""
    out = c_lib_path()
    src = os.path.join(_HERE, "gcn_oracle.c")
    if (not force) and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(
        ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c11", "-o", out, src, "-lm"]
    )
    return out


def c_lib():
    global _C_LIB
    if _C_LIB is None:
        _C_LIB = ctypes.CDLL(build_c())
        i64, vp = ctypes.c_int64, ctypes.c_void_p
        _C_LIB.oracle_layer_forward_f32.argtypes = [i64, i64, i64, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp]
        _C_LIB.oracle_layer_backward_f32.argtypes = [i64, i64, i64, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        _C_LIB.oracle_layer_forward_f64.argtypes = _C_LIB.oracle_layer_forward_f32.argtypes
        _C_LIB.oracle_layer_backward_f64.argtypes = _C_LIB.oracle_layer_backward_f32.argtypes
        _C_LIB.oracle_csr_layer_fwdbwd_f32.argtypes = [i64, i64, i64] + [vp] * 16
        _C_LIB.oracle_csr_layer_fwdbwd_f32.restype = ctypes.c_int
    return _C_LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def c_layer_forward(x, w, b, idx, val, n_rows, dtype=np.float32):
    """C version of :func:`layer_forward` (fp32 or fp64 arithmetic)."""
    lib = c_lib()
    x = np.ascontiguousarray(x, dtype=dtype)
    w = np.ascontiguousarray(w, dtype=dtype)
    bb = None if b is None else np.ascontiguousarray(b, dtype=dtype)
    row = np.ascontiguousarray(idx[0], dtype=np.int64)
    col = np.ascontiguousarray(idx[1], dtype=np.int64)
    v = np.ascontiguousarray(val, dtype=dtype)
    n_cols, fin = x.shape
    fout = w.shape[1]
    support = np.empty((n_cols, fout), dtype=dtype)
    out = np.empty((n_rows, fout), dtype=dtype)
    fn = lib.oracle_layer_forward_f32 if dtype == np.float32 else lib.oracle_layer_forward_f64
    fn(n_rows, n_cols, fin, fout, _p(x), row.shape[0], _p(row), _p(col), _p(v), _p(w), _p(bb), _p(support), _p(out))
    return support, out


def c_layer_backward(x, w, has_bias, idx, val, n_cols, g, dtype=np.float32):
    """C version of :func:`layer_backward`."""
    lib = c_lib()
    x = np.ascontiguousarray(x, dtype=dtype)
    w = np.ascontiguousarray(w, dtype=dtype)
    g = np.ascontiguousarray(g, dtype=dtype)
    row = np.ascontiguousarray(idx[0], dtype=np.int64)
    col = np.ascontiguousarray(idx[1], dtype=np.int64)
    v = np.ascontiguousarray(val, dtype=dtype)
    n_rows, fout = g.shape
    fin = x.shape[1]
    ds = np.empty((n_cols, fout), dtype=dtype)
    dw = np.empty((fin, fout), dtype=dtype)
    db = np.empty((fout,), dtype=dtype)
    dx = np.empty((n_cols, fin), dtype=dtype)
    fn = lib.oracle_layer_backward_f32 if dtype == np.float32 else lib.oracle_layer_backward_f64
    fn(n_rows, n_cols, fin, fout, _p(x), row.shape[0], _p(row), _p(col), _p(v), _p(w), _p(g), _p(ds), _p(dw), _p(db), _p(dx))
    return dw, (db if has_bias else None), dx, ds
