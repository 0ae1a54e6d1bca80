"""One rank's share of the ogbn-papers100M-shaped layer (BASELINE configs[4]) on ONE GPU.

    python tools/papers_block.py [--world 8] [--rank 0] [--n 111059956] [--edges 752000000] [--f 128]

The row-partitioned layer (pygcn_b200/dist.py) gives every rank a row block of A-hat and of A-hat^T whose
columns index the all-gathered [world * pad_rows, F] panel.  This tool builds rank `--rank`'s blocks with
dist._partition_counts / _partition_blocks (the partitioned build: the rank only ever sees the edges incident
to its rows), allocates the full-size gathered panel (57 GB at F = 128: it fits one B200), fills the slots of
the other ranks with random data instead of receiving them, and times the rank's kernels of one layer
forward + backward with CUDA events: X_p W, the row-block SpMM, colsum(G), the A^T row-block SpMM, dW.
What is NOT measured here is the exchange itself (7/8 of the panel per SpMM arrives over NVLink).

The other ranks' row sums, which the real build all-gathers, are replaced by 1 + incident-edge counts (exact
unless an edge is duplicated or a self-edge; ~30 such edges are expected at this size).  They only enter the
VALUES of the A^T block.  Checks: rows of the A block sum to 1 (SpMM against a panel of ones), and 2000
sampled rows of both SpMMs against an fp64 evaluation of the same CSR rows with torch index ops (<= 1e-5).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygcn_b200 as P
from pygcn_b200 import dist as D


def timed(fn, reps=3):
    ts = []
    for it in range(reps + 1):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if it >= 1:
            ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


def gb():
    return torch.cuda.max_memory_allocated() / 1e9


def sample_check(block, panel, out, n_sample=2000, seed=0):
    """rows of `out` vs the same CSR rows evaluated in fp64 with index ops"""
    rowptr, col, val = block.csr()
    gen = torch.Generator(device=out.device).manual_seed(seed)
    rows = torch.randint(0, out.shape[0], (n_sample,), generator=gen, device=out.device)
    worst = 0.0
    scale = out.abs().max().item()
    for r in rows.tolist():
        e0, e1 = int(rowptr[r]), int(rowptr[r + 1])
        want = (val[e0:e1].double()[:, None] * panel[col[e0:e1].long()].double()).sum(0)
        worst = max(worst, (out[r].double() - want).abs().max().item())
    return worst / scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--n", type=int, default=111059956)
    ap.add_argument("--edges", type=int, default=752000000, help="raw undirected edges (stored entries ~ 2x + n)")
    ap.add_argument("--f", type=int, default=128)
    ap.add_argument("--fin", type=int, default=128)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    n, world, rank, f, fin = args.n, args.world, args.rank, args.f, args.fin
    say = lambda *a: print(*a, flush=True)
    t0 = time.perf_counter()
    gen = torch.Generator(device=dev).manual_seed(7)
    src = torch.randint(0, n, (args.edges,), generator=gen, device=dev, dtype=torch.int32)
    dst = torch.randint(0, n, (args.edges,), generator=gen, device=dev, dtype=torch.int32)
    # incident-edge counts (+ self loop): the partition estimate of build_partitioned, and the stand-in for
    # the other ranks' row sums
    deg = torch.ones(n, dtype=torch.int64, device=dev)
    one = torch.ones(1, dtype=torch.int64, device=dev)
    chunk = 1 << 27
    for e0 in range(0, args.edges, chunk):
        for t in (src, dst):
            idx = t[e0:e0 + chunk].long()
            deg.index_add_(0, idx, one.expand(idx.numel()))
    rp = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), deg.cumsum(0)])
    bounds = D.partition_rows_by_nnz(rp, world)
    del rp
    pad = D.DistGraph.padded_rows(bounds)
    r0, r1 = bounds[rank], bounds[rank + 1]
    say("n %d, raw edges %d, world %d: rank %d owns rows [%d, %d) = %d, pad_rows %d  (%.1f s, peak %.1f GB)" % (
        n, args.edges, world, rank, r0, r1, r1 - r0, pad, time.perf_counter() - t0, gb()))
    t1 = time.perf_counter()
    lrp, lcol, a, rows, rowsum = D._partition_counts(src, dst, n, r0, r1)
    del src, dst
    rowsum_global = deg.to(torch.float64)
    rowsum_global[r0:r1] = rowsum          # own rows: exact
    del deg
    dg = D._partition_blocks(rank, world, bounds, pad, lrp, lcol, a, rows, rowsum_global, 0)
    del lrp, lcol, a, rows, rowsum_global
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    fwd, bwd = dg.fwd_remote, dg.bwd_remote
    say("partitioned build: %.1f s, block %s, stored entries %d (x %d ranks ~ %.2f G), peak %.1f GB" % (
        time.perf_counter() - t1, fwd, fwd.nnz, world, fwd.nnz * world / 1e9, gb()))
    ops = D.CudaOps()
    nloc = r1 - r0

    # rows of the A block sum to 1
    ones = torch.ones(world * pad, 4, device=dev)
    o4 = torch.empty(nloc, 4, device=dev)
    ops.spmm_block(fwd, ones, o4, False)
    say("rows of the A block sum to 1: max |sum - 1| = %.2e" % (o4 - 1).abs().max().item())
    del ones, o4

    panel = torch.empty(world * pad, f, device=dev)
    for q in range(world):  # in place, slot by slot (no 57 GB temporary)
        panel[q * pad:(q + 1) * pad].normal_(generator=gen)
    x = torch.empty(nloc, fin, device=dev).normal_(generator=gen)
    g = torch.empty(nloc, f, device=dev).normal_(generator=gen)
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, f).to(dev)
    w, b = layer.weight.detach(), layer.bias.detach()
    out = torch.empty(nloc, f, device=dev)
    ds = torch.empty(nloc, f, device=dev)
    my_slot = panel[rank * pad: rank * pad + nloc]
    say("panel %.1f GB, allocated %.1f GB" % (panel.numel() * 4 / 1e9, torch.cuda.memory_allocated() / 1e9))

    t_xw = timed(lambda: ops.gemm(x, w, out=my_slot))
    t_f = timed(lambda: ops.spmm_block(fwd, panel, out, False, b, False))
    say("forward SpMM sampled rows vs fp64: %.2e" % sample_check(fwd, panel, out - b))
    # backward: the rank's G goes into its slot (colsum stages it), the other slots hold the peers' G
    t_c = timed(lambda: ops.colsum(g, None, my_slot))
    t_b = timed(lambda: ops.spmm_block(bwd, panel, ds, False))
    say("A^T SpMM sampled rows vs fp64: %.2e" % sample_check(bwd, panel, ds, seed=1))
    t_dw = timed(lambda: ops.gemm(x.t(), ds))
    e = fwd.nnz
    bytes_spmm = e * 8 + (nloc + 1) * 4 + world * pad * f * 4 + nloc * f * 4
    say("rank %d of %d, papers100M shape, F %d -> %d, per step:" % (rank, world, fin, f))
    say("  X_p W %.2f ms | SpMM %.2f ms | colsum(G) %.2f ms | A^T SpMM %.2f ms | dW %.2f ms | total %.2f ms" % (
        t_xw, t_f, t_c, t_b, t_dw, t_xw + t_f + t_c + t_b + t_dw))
    say("  SpMM: %d stored entries, gathered rows %d B: %.1f GB of gathers -> %.2f TB/s; algorithmic bytes %.1f GB "
        "(whole panel counted once) -> %.2f TB/s" % (e, f * 4, e * f * 4 / 1e9, e * f * 4 / t_f / 1e9,
                                                     bytes_spmm / 1e9, bytes_spmm / t_f / 1e9))
    say("  edges/s of this rank's compute: %.3e ; x %d ranks = %.3e (exchange not included: %.1f GB in per SpMM)" % (
        e / ((t_xw + t_f + t_c + t_b + t_dw) * 1e-3), world, world * e / ((t_xw + t_f + t_c + t_b + t_dw) * 1e-3),
        (world - 1) * pad * f * 4 / 1e9))
    say("peak memory %.1f GB" % gb())


if __name__ == "__main__":
    main()
