"""World-size-2 (and 3) CPU tests of the multi-GPU host logic in pygcn_b200/dist.py over gloo:
nnz-balanced row partition, diagonal / remote column blocks in the all-gather layout, the padded
panel all-gather, accumulation, dW/db all-reduce.  The arithmetic backend injected here is numpy (the oracle's formulas);
the product backend (CudaOps) is exercised on GPUs by bench.py --gpus N and test_gpu_parity.py."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gcn_oracle as O
from pygcn_b200 import dist as D


class HostBlock:
    """rows [r0, r1) of a COO matrix; columns either local to [c0, c1) (diagonal block) or remapped
    to the padded all-gather layout with part `exclude` left out (remote block)."""

    def __init__(self, idx, val, r0, r1, bounds=None, pad=None, exclude=None, c0=None, c1=None, only=None):
        rows = (idx[0] >= r0) & (idx[0] < r1)
        if bounds is None:
            m = rows & (idx[1] >= c0) & (idx[1] < c1)
            self.col = idx[1][m] - c0
        else:
            part = np.searchsorted(np.asarray(bounds), idx[1], side="right") - 1
            m = rows & ((part != exclude) if only is None else np.isin(part, list(only)))
            self.col = part[m] * pad + (idx[1][m] - np.asarray(bounds)[part[m]])
        self.row = idx[0][m] - r0
        self.val = val[m]
        self.nnz = int(m.sum())
        self.n_rows = r1 - r0


class NumpyOps:
    def gemm(self, a, b, out=None):
        r = torch.from_numpy(a.numpy() @ b.numpy())
        if out is None:
            return r
        out[: r.shape[0]] = r
        return out

    def gemm_bias_act(self, a, b, bias=None, relu=False):
        r = a.numpy() @ b.numpy()
        if bias is not None:
            r = r + bias.numpy()
        if relu:
            r = np.maximum(r, 0)
        return torch.from_numpy(r.astype(np.float32))

    def spmm_block(self, block, dense, out, accumulate, bias=None, relu=False):
        acc = out.numpy().copy() if accumulate else np.zeros(out.shape, np.float32)
        np.add.at(acc, block.row, block.val[:, None] * dense.numpy()[block.col])
        if bias is not None:
            acc = acc + bias.numpy()
        if relu:
            acc = np.maximum(acc, 0)
        out.copy_(torch.from_numpy(acc.astype(np.float32)))
        return out

    def colsum(self, g, y=None, gm=None):
        m = g if y is None else torch.where(y > 0, g, torch.zeros_like(g))
        if gm is not None:
            gm[: m.shape[0]] = m
        return m.sum(0), m

    def empty(self, shape, like):
        return torch.full(shape, float("nan"), dtype=torch.float32)  # padding must never be read

    # build-time helpers of the halo exchange (dist.HaloPlan)
    def block_csr(self, block):
        order = np.argsort(block.row, kind="stable")
        n_rows = block.n_rows
        rowptr = np.zeros(n_rows + 1, np.int64)
        np.add.at(rowptr, block.row + 1, 1)
        return (torch.from_numpy(np.cumsum(rowptr)), torch.from_numpy(block.col[order].astype(np.int64)),
                torch.from_numpy(block.val[order].astype(np.float32)))

    def block_from_csr(self, rowptr, col, val, n_rows, n_cols):
        b = HostBlock.__new__(HostBlock)
        counts = np.diff(rowptr.numpy())
        b.row = np.repeat(np.arange(n_rows), counts)
        b.col, b.val, b.nnz, b.n_rows = col.numpy(), val.numpy(), int(col.numel()), n_rows
        assert b.col.size == 0 or (b.col.min() >= 0 and b.col.max() < n_cols)
        return b

    def selection_block(self, ids, n_cols):
        k = ids.numel()
        return self.block_from_csr(torch.arange(k + 1), ids.long(), torch.ones(k), k, n_cols)


def _problem(n=300, seed=3, fout=5, fin=12):
    rs = np.random.default_rng(seed)
    src = (n * rs.random(4000) ** 2).astype(np.int64)  # skewed degrees: nnz balance != row balance
    dst = rs.integers(0, n, 4000)
    idx, val = O.build_normalized_adjacency(src, dst, n)
    x = rs.standard_normal((n, 12)).astype(np.float32)
    g = rs.standard_normal((n, fout)).astype(np.float32)
    w = rs.standard_normal((12, fout)).astype(np.float32)
    b = rs.standard_normal(fout).astype(np.float32)
    return n, idx, val, x, g, w, b


def _worker(rank, world, port, outdir, relu, split=True, exchange="nccl", agg=False, fin=12, fout=5, chunk_min=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if chunk_min is not None:  # pipelined exchange in column chunks of >= chunk_min columns (the product default is 128)
        D.CHUNK_MIN_COLS = chunk_min
    try:
        n, idx, val, x, g, w, b = _problem(fin=fin, fout=fout)
        bounds = D.partition_rows_by_nnz(O.coo_to_csr(idx, n), world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        tidx = np.vstack([idx[1], idx[0]])
        pad = D.DistGraph.padded_rows(bounds)
        blocks = []
        for ii in (idx, tidx):
            blocks.append(HostBlock(ii, val, r0, r1, c0=r0, c1=r1) if split else None)
            blocks.append(HostBlock(ii, val, r0, r1, bounds=bounds, pad=pad, exclude=rank if split else -1))
        dg = D.DistGraph(rank, world, bounds, pad, *blocks, 0, idx.shape[1], split)
        ops = NumpyOps()
        xt, gt = torch.from_numpy(x[r0:r1].copy()), torch.from_numpy(g[r0:r1].copy())
        wt, bt = torch.from_numpy(w), torch.from_numpy(b)
        ef = eb = exchange
        if exchange == "pipelined":  # one column block per source rank, consumed in the order p, p+1, ... (exchange "peer")
            dg.phases = D.exchange_phases(rank, world)
            dg.fwd_blocks = [HostBlock(idx, val, r0, r1, bounds=bounds, pad=pad, only=qs) for qs in dg.phases]
            dg.bwd_blocks = [HostBlock(tidx, val, r0, r1, bounds=bounds, pad=pad, only=qs) for qs in dg.phases]
            ef = D.CollectiveExchange(rank, world, pad, w.shape[1], xt)
            eb = D.CollectiveExchange(rank, world, pad, w.shape[1], xt)
        out, ax = D.dist_layer_forward(ops, dg, xt, wt, bt, relu=relu, exch=ef, agg=agg)
        assert (ax is not None) == agg
        dx, dw, db = D.dist_layer_backward(ops, dg, None if agg else xt, wt, gt, out if relu else None, True, True, exch=eb,
                                           agg=agg, ax=ax)
        frac = D.halo_fraction(ops, dg) if exchange == "halo" else -1.0
        np.savez(os.path.join(outdir, "r%d.npz" % rank), out=out.numpy(), dx=dx.numpy(), dw=dw.numpy(), db=db.numpy(),
                 bounds=np.array(bounds), frac=frac)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_column_chunks_of_the_pipelined_exchange():
    assert D.column_chunks(256) == [(0, 128), (128, 256)]
    assert D.column_chunks(100) == [(0, 100)] and D.column_chunks(47) == [(0, 47)] and D.column_chunks(128) == [(0, 128)]
    assert D.column_chunks(260) == [(0, 132), (132, 260)]
    old = D.CHUNK_MIN_COLS
    try:
        D.CHUNK_MIN_COLS = 4
        assert D.column_chunks(12) == [(0, 8), (8, 12)] and D.column_chunks(8) == [(0, 4), (4, 8)]
        assert D.column_chunks(6) == [(0, 6)]  # not a multiple of 4: one chunk
    finally:
        D.CHUNK_MIN_COLS = old


@pytest.mark.parametrize("world,relu,exchange,agg,chunk_min", [
    (2, True, "halo", False, None), (4, False, "halo", False, None), (3, True, "nccl", False, None),
    # the pipelined exchange: X W / the staged masked G / X / G W^T produced, packed and sent in two column chunks
    (2, True, "halo", False, 4), (3, False, "halo", False, 4), (4, True, "halo", True, 4), (2, False, "halo", True, 4)])
def test_unsplit_halo_exchange_with_the_panel_produced_in_the_compact_buffer(world, relu, exchange, agg, chunk_min):
    """Widths that need no padding (8): X W, the staged masked G (and G W^T in the aggregate-first order, fin = 12
    above) are written straight into the own slot of the compact panel (dist.halo_compact) -- same results; and the
    same through dist_spmm_halo_chunked."""
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, relu, False, exchange, agg, 12, 8, chunk_min), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, "r%d.npz" % r)) for r in range(world)]
    n, idx, val, x, g, w, b = _problem(fout=8)
    _, o_ref = O.c_layer_forward(x, w, b, idx, val, n)
    gm = g
    if relu:
        gm = O.relu_backward(g, o_ref)
        o_ref = np.maximum(o_ref, 0)
    dw, db, dx, _ = O.c_layer_backward(x, w, True, idx, val, n, gm)
    assert O.normwise_err(np.concatenate([p["out"] for p in parts]), o_ref) < 1e-5
    assert O.normwise_err(np.concatenate([p["dx"] for p in parts]), dx) < 1e-5
    for p in parts:
        assert O.normwise_err(p["dw"], dw) < 1e-5 and O.normwise_err(p["db"], db) < 1e-5


@pytest.mark.parametrize("world,relu,split,exchange,agg", [
    (2, False, True, "nccl", False), (2, True, False, "nccl", False), (3, False, True, "nccl", False),
    (2, True, False, "pipelined", False), (3, False, False, "pipelined", False), (4, True, False, "pipelined", False),
    # the needed-rows-only exchange: unsplit (own slot at the front of the compact panel) and split row blocks
    (2, True, False, "halo", False), (3, False, True, "halo", False), (4, True, False, "halo", False),
    # the aggregate-first order (A X) W + b: the exchanged panel is X (forward) and G W^T (backward, for dX); dW needs
    # no exchange
    (2, True, True, "nccl", True), (3, False, False, "nccl", True), (3, True, True, "halo", True), (4, False, False, "halo", True)])
def test_row_partitioned_layer_matches_single_process_oracle(world, relu, split, exchange, agg):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, relu, split, exchange, agg), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, "r%d.npz" % r)) for r in range(world)]
    n, idx, val, x, g, w, b = _problem()
    _, o_ref = O.c_layer_forward(x, w, b, idx, val, n)
    gm = g
    if relu:
        gm = O.relu_backward(g, o_ref)
        o_ref = np.maximum(o_ref, 0)
    dw, db, dx, _ = O.c_layer_backward(x, w, True, idx, val, n, gm)
    out = np.concatenate([p["out"] for p in parts])
    dxs = np.concatenate([p["dx"] for p in parts])
    assert out.shape == o_ref.shape
    assert O.normwise_err(out, o_ref) < 1e-5
    assert O.normwise_err(dxs, dx) < 1e-5
    for p in parts:  # all-reduced: every rank holds the full gradient
        assert O.normwise_err(p["dw"], dw) < 1e-5 and O.normwise_err(p["db"], db) < 1e-5
    assert list(parts[0]["bounds"]) == list(parts[-1]["bounds"])
    if exchange == "halo":  # every rank agrees on the largest share of remote rows any rank reads
        assert len({float(p["frac"]) for p in parts}) == 1 and 0.0 < float(parts[0]["frac"]) <= 1.0


def test_aggregate_first_rule_of_the_partitioned_layer():
    """(A X) W when that moves fewer panel columns through the exchanges: in_features < out_features with an input
    that needs no gradient halves again what the reference order costs (no backward exchange at all)."""
    assert D.aggregate_first(100, 256, need_dx=False) and D.aggregate_first(100, 256, need_dx=True)
    assert not D.aggregate_first(256, 256, True) and not D.aggregate_first(256, 47, True) and not D.aggregate_first(602, 256, False)
    assert D.aggregate_first(64, 48, need_dx=False) and not D.aggregate_first(64, 32, need_dx=False)  # 64 < 2 * 48; tie keeps the reference order
    assert not D.aggregate_first(5, 9, False)  # rows of X must be 16-byte aligned


def test_balanced_node_partition_gives_equal_rows_and_entries():
    """dist.balanced_node_partition on power-law row lengths: a permutation; every rank owns N / P rows (+- 1) and a
    share of the stored entries within one long row of the mean; nodes of a rank stay in ascending id order."""
    rs = np.random.default_rng(9)
    n = 5000
    rl = (rs.pareto(1.2, n) * 3).astype(np.int64) + 1
    rl[7] = 20000
    for world in (1, 2, 3, 8):
        new_id, bounds = D.balanced_node_partition(torch.from_numpy(rl), world)
        new_id = new_id.numpy()
        assert sorted(new_id.tolist()) == list(range(n)) and bounds[0] == 0 and bounds[-1] == n and len(bounds) == world + 1
        rows = np.diff(bounds)
        assert rows.max() - rows.min() <= 1
        owner = np.searchsorted(np.asarray(bounds), new_id, side="right") - 1
        ent = np.bincount(owner, weights=rl, minlength=world)
        assert ent.max() - ent.min() <= rl.max()  # the one 20000-entry row cannot be split: everything else evens out
        rl2 = np.minimum(rl, 300)  # without a row that outweighs a whole share, the entries even out to a few rows' worth
        nid2, b2 = D.balanced_node_partition(torch.from_numpy(rl2), world)
        own2 = np.searchsorted(np.asarray(b2), nid2.numpy(), side="right") - 1
        ent2 = np.bincount(own2, weights=rl2, minlength=world)
        assert ent2.max() - ent2.min() <= 2 * rl2.max()
        for p in range(world):
            ids = np.nonzero(owner == p)[0]
            assert (np.diff(new_id[ids]) > 0).all()  # ascending ids -> ascending rows inside a rank


def test_exchange_phases_cover_every_source_once_own_slot_first():
    for world in (1, 2, 3, 4, 5, 8, 16):
        for rank in range(world):
            ph = D.exchange_phases(rank, world)
            flat = [q for qs in ph for q in qs]
            assert ph[0] == [rank] and sorted(flat) == list(range(world))
            assert flat == [(rank + k) % world for k in range(world)]  # arrival order of the rotated pushes
            assert all(1 <= len(qs) <= 2 for qs in ph) and (world < 3 or len(ph[-1]) == 1 or world % 2 == 1)


def test_partition_balances_nnz_and_exchange_is_a_matching():
    n, idx, val, *_ = _problem()
    rowptr = O.coo_to_csr(idx, n)
    for world in (1, 2, 4, 8):
        b = D.partition_rows_by_nnz(rowptr, world)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
        per = [rowptr[b[i + 1]] - rowptr[b[i]] for i in range(world)]
        assert max(per) - min(per) <= 2 * np.diff(rowptr).max()  # within one (long) row of perfect balance
        pad = D.DistGraph.padded_rows(b)
        assert pad % 8 == 0 and pad >= max(b[i + 1] - b[i] for i in range(world))
    # row_weight: cost = stored entries + weight * rows.  Degrees sorted descending (hubs first, like the low ids of an
    # R-MAT graph): entries alone give the first block few rows and the last one most of them; weighing rows by the
    # average degree bounds the largest block (= the padded all-gather slot) at the price of some nnz imbalance
    deg = np.sort((3000 * np.random.default_rng(5).random(4000) ** 8).astype(np.int64) + 1)[::-1]
    rp = np.concatenate([[0], np.cumsum(deg)])
    plain = D.partition_rows_by_nnz(rp, 8)
    mixed = D.partition_rows_by_nnz(rp, 8, row_weight=float(deg.mean()))
    rows = lambda b: [b[i + 1] - b[i] for i in range(8)]
    assert mixed[0] == 0 and mixed[-1] == 4000 and all(mixed[i] <= mixed[i + 1] for i in range(8))
    assert max(rows(mixed)) < 0.5 * max(rows(plain))
    assert max(rows(mixed)) <= 2 * 4000 // 8 + 1   # at most twice the even share when both terms weigh the same
    # more ranks than rows: empty blocks allowed
    assert D.partition_rows_by_nnz(np.array([0, 2, 5]), 4)[-1] == 2


@pytest.mark.parametrize("association", ["reference", "aggregate_first", "auto"])
def test_autograd_function_of_the_partitioned_layer_at_world_1(association):
    """_DistGCNLayerFn (what DistGraphConvolution.forward calls) end to end through torch autograd with the numpy
    backend at world size 1 -- no process group needed: argument / gradient arity, ctx plumbing, needs_input_grad,
    the fused-ReLU mask, both association orders."""
    n, idx, val, x, g, w, b = _problem()
    tidx = np.vstack([idx[1], idx[0]])
    bounds = [0, n]
    pad = D.DistGraph.padded_rows(bounds)
    dg = D.DistGraph(0, 1, bounds, pad, HostBlock(idx, val, 0, n, c0=0, c1=n), None, HostBlock(tidx, val, 0, n, c0=0, c1=n),
                     None, idx.shape[1], idx.shape[1], True)
    xt = torch.from_numpy(x.copy()).requires_grad_(True)
    wt = torch.from_numpy(w.copy()).requires_grad_(True)
    bt = torch.from_numpy(b.copy()).requires_grad_(True)
    out = D._DistGCNLayerFn.apply(xt, wt, bt, dg, True, NumpyOps(), None, None, None, association)
    out.backward(torch.from_numpy(g))
    _, o_ref = O.c_layer_forward(x, w, b, idx, val, n)
    dw, db, dx, _ = O.c_layer_backward(x, w, True, idx, val, n, O.relu_backward(g, o_ref))
    tol = 1e-5
    assert O.normwise_err(out.detach().numpy(), np.maximum(o_ref, 0)) < tol
    assert O.normwise_err(wt.grad.numpy(), dw) < tol and O.normwise_err(xt.grad.numpy(), dx) < tol
    assert O.normwise_err(bt.grad.numpy(), db) < tol
    x2 = torch.from_numpy(x.copy())  # input without grad: no dX is computed
    out2 = D._DistGCNLayerFn.apply(x2, wt, None, dg, False, NumpyOps(), None, None, None, association)
    out2.sum().backward()
    assert x2.grad is None and out2.shape == (n, w.shape[1])
