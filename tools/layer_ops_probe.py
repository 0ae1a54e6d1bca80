"""Every operation of one layer step on a bench workload's graph, timed alone with CUDA events (L2 flushed): the dense
products through gcnb_gemm / gcnb_gemm_ex, the two SpMMs, the masked column sum -- which kernel of a layer is slow.

    python tools/layer_ops_probe.py [workload=products] [fin=256] [fout=47] [reps=5]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as B
import pygcn_b200 as P
from pygcn_b200 import _lib
from pygcn_b200.dist import CudaOps


def main():
    a = sys.argv[1:]
    wl = B.WORKLOADS[a[0] if a else "products"]
    fin, fout = (int(a[1]), int(a[2])) if len(a) > 2 else (256, 47)
    reps = int(a[3]) if len(a) > 3 else 5
    dev = torch.device("cuda:0")
    lib = _lib.load()
    graph = B.make_graph(P, torch, wl, dev, edges_on=dev)
    n = graph.n_rows
    timer = B.Timer(torch, lib, _lib, dev)
    ops = CudaOps()
    ld4 = lambda f: (f + 3) // 4 * 4
    x = torch.randn(n, fin, device=dev)
    w = torch.randn(fin, fout, device=dev)
    bias = torch.randn(fout, device=dev)
    g = torch.randn(n, fout, device=dev)
    support = torch.zeros(n, ld4(fout), device=dev)
    out = torch.empty(n, fout, device=dev)
    ds = torch.zeros(n, ld4(fout), device=dev)

    def spmm(flags, dense, f, dst, ldo):
        ws = torch.empty(max(lib.gcnb_spmm_workspace_bytes(graph._h, flags, f), 256), dtype=torch.uint8, device=dev)

        def run():
            _lib.check(lib.gcnb_spmm(graph._h, flags, dense.data_ptr(), dense.stride(0), f, None, dst.data_ptr(), ldo,
                                     ws.data_ptr(), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "spmm")
        return run
    y = torch.relu(torch.randn(n, fout, device=dev))
    items = [
        ("X W            [%d,%d]x[%d,%d] -> ld %d" % (n, fin, fin, fout, ld4(fout)), lambda: ops.gemm(x, w)),
        ("(A X) W + b, relu (gemm_ex)", lambda: ops.gemm_bias_act(x, w, bias, True)),
        ("SpMM fwd f=%d" % fout, spmm(0, support, fout, out, fout)),
        ("colsum + relu mask", lambda: ops.colsum(g, y)),
        ("SpMM A^T f=%d" % fout, spmm(_lib.SPMM_TRANSPOSE, g, fout, ds, ld4(fout))),
        ("dW = X^T dS    [%d,%d]^T x [%d,%d(ld %d)]" % (n, fin, n, fout, ld4(fout)), lambda: ops.gemm(x.t(), ds[:, :fout])),
        ("dX = dS W^T    [%d,%d(ld %d)] x [%d,%d]" % (n, fout, ld4(fout), fout, fin), lambda: ops.gemm(ds[:, :fout], w.t())),
    ]
    for name, fn in items:
        ms, all_ms = timer.time(fn, reps)
        print("%-60s %8.3f ms (min %.3f)" % (name, ms, min(all_ms)), flush=True)


if __name__ == "__main__":
    main()
