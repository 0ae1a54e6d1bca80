"""What does the column-chunked exchange (dist_spmm_chunked) cost in SpMM time?  ONE GPU, no communicator:
rank 0's unsplit row block of the WORLD x CBG weak-scaled graph (the block bench.py --gpus WORLD multiplies)
against a gathered panel of WORLD * pad rows, once as one SpMM of width F and once as `chunks` SpMMs of
width F / chunks into column slices of the same output (what runs between the chunked all-gathers).

    python tools/chunk_probe.py [WORLD=8] [F=32]

Also checks that the chunked result equals the one-pass result bit for bit (same per-row order).
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as B
import pygcn_b200 as P
from pygcn_b200 import _lib
from pygcn_b200 import dist as D


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    f = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    wl = dict(B.WORKLOADS["cbg"])
    wl["n"] *= world
    full = B.make_graph(P, torch, wl, dev)
    dg = D.DistGraph.from_graph(full, 0, world, split=False)
    del full
    blk = dg.fwd_remote
    ops = D.CudaOps()
    flush_buf = torch.empty(B.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def flush():
        _lib.check(lib.gcnb_l2_flush(ctypes.c_void_p(flush_buf.data_ptr()), flush_buf.numel(),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "l2_flush")

    def timed(fn, reps=10):
        ts = []
        for it in range(reps + 3):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        return 1e3 * sum(ts) / len(ts)

    panel = torch.randn(world * dg.pad_rows, f, device=dev)
    bias = torch.randn(f, device=dev)
    out1 = torch.empty(dg.n_rows(), f, device=dev)
    print("world %d: row block %d x %d, %d stored entries, F = %d" % (world, dg.n_rows(), blk.n_cols, blk.nnz, f))
    t1 = timed(lambda: ops.spmm_block(blk, panel, out1, False, bias, True))
    print("one SpMM of width %d: %.1f us" % (f, t1))
    for chunks in (2, 4):
        cc = D.chunk_columns(f, chunks)
        if len(cc) < 2:
            continue
        parts = [panel[:, c0:c1].contiguous() for c0, c1 in cc]
        out2 = torch.full((dg.n_rows(), f), float("nan"), device=dev)

        def run():
            for (c0, c1), part in zip(cc, parts):
                ops.spmm_block(blk, part, out2[:, c0:c1], False, bias[c0:c1], True)
        t2 = timed(run)
        tslice = timed(lambda: [panel[:, c0:c1].contiguous() for c0, c1 in cc])
        same = torch.equal(out1, out2)
        print("%d SpMMs of widths %s: %.1f us (+ %.1f us to slice the panel); bit-identical to one pass: %s" % (
            len(cc), [c1 - c0 for c0, c1 in cc], t2, tslice, same))


if __name__ == "__main__":
    main()
