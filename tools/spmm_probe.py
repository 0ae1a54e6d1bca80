"""SpMM alone on one of bench.py's workloads: checks it against torch's CUDA CSR spmm, then times
forward / transposed launches with the L2 flushed.  Development tool; also the command ncu wraps.

    python tools/spmm_probe.py <workload> [width] [--check] [--reps N]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as B
import pygcn_b200 as P
from pygcn_b200 import _lib


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    name = args[0] if args else "cbg"
    if ":" in name:  # custom uniform graph  n:avg_deg
        n_, d_ = name.split(":")
        wl = dict(n=int(n_), avg_deg=int(d_), fin=64, fout=32, name=name)
    else:
        wl = B.WORKLOADS[name]
    f = int(args[1]) if len(args) > 1 else wl["fout"]
    reps = 5
    if "--reps" in sys.argv:
        reps = int(sys.argv[sys.argv.index("--reps") + 1])
    dev = torch.device("cuda:0")
    lib = _lib.load()
    graph = B.make_graph(P, torch, wl, dev)
    n = graph.n_rows
    s = torch.randn(n, f, device=dev)
    out = torch.empty(n, f, device=dev)
    flush_buf = torch.empty(B.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def spmm(tflag):
        wsb = lib.gcnb_spmm_workspace_bytes(graph._h, tflag, f)
        ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
        _lib.check(lib.gcnb_spmm(graph._h, tflag, ctypes.c_void_p(s.data_ptr()), f, f, None,
                                 ctypes.c_void_p(out.data_ptr()), f, ctypes.c_void_p(ws.data_ptr()), ws.numel(), st), "spmm")

    if "--check" in sys.argv:
        csr = graph.to_sparse_coo().coalesce().to_sparse_csr()
        spmm(0)
        ref = torch.sparse.mm(csr, s)
        err = ((out - ref).abs().max() / ref.abs().max()).item()
        print("check vs torch CUDA CSR spmm: max|d|/max|ref| = %.3e" % err)
        assert err < 1e-5
        del csr, ref
    alg = B.algorithmic_bytes_spmm(graph.nnz, n, f)
    if "--bf16" in sys.argv:  # the bf16 panel tier: same product, panel rounded to bf16 once
        ld8 = (f + 7) // 8 * 8
        panel = torch.empty(n, ld8, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.gcnb_to_bf16(n, f, ctypes.c_void_p(s.data_ptr()), f, ctypes.c_void_p(panel.data_ptr()), ld8, st), "to_bf16")
        assert torch.equal(panel[:, :f], s.to(torch.bfloat16)), "gcnb_to_bf16 != torch's round-to-nearest-even"
        spmm(0)
        ref32 = out.clone()

        def spmm(tflag):  # noqa: F811
            wsb = lib.gcnb_spmm_workspace_bytes(graph._h, tflag, f)
            ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
            _lib.check(lib.gcnb_spmm_bf16(graph._h, tflag, ctypes.c_void_p(panel.data_ptr()), ld8, f, None,
                                          ctypes.c_void_p(out.data_ptr()), f, ctypes.c_void_p(ws.data_ptr()), ws.numel(), st),
                       "spmm_bf16")
        spmm(0)
        e = ((out - ref32).abs().max() / ref32.abs().max()).item()
        # against the fp32 kernel on the SAME rounded panel the two must agree to fp32 rounding
        s_keep = s
        s = panel[:, :f].float().contiguous()
        out2 = out.clone()
        wsb = lib.gcnb_spmm_workspace_bytes(graph._h, 0, f)
        ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
        _lib.check(lib.gcnb_spmm(graph._h, 0, ctypes.c_void_p(s.data_ptr()), f, f, None, ctypes.c_void_p(out.data_ptr()), f,
                                 ctypes.c_void_p(ws.data_ptr()), ws.numel(), st), "spmm")
        e2 = ((out - out2).abs().max() / out.abs().max()).item()
        print("bf16 panel: vs fp32 panel %.2e (tier bound 2e-2), vs fp32 kernel on the rounded panel %.2e" % (e, e2))
        assert e < 2e-2 and e2 < 1e-5
        s = s_keep
        alg = graph.nnz * 8 + (n + 1) * 4 + n * f * 2 + n * f * 4
    for tflag, name in ((0, "fwd"), (_lib.SPMM_TRANSPOSE, "A^T")):
        ts = []
        for it in range(2 + reps):
            lib.gcnb_l2_flush(ctypes.c_void_p(flush_buf.data_ptr()), flush_buf.numel(), st)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            spmm(tflag)
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        ms = sum(ts) / len(ts)
        print("%s n=%d nnz=%d f=%d %s: %.3f ms  alg %.1f GB/s  gather %.2f TB/s  %.2f Gedges/s" % (
            args[0] if args else "cbg", n, graph.nnz, f, name, ms, alg / ms / 1e6, graph.nnz * f * (2 if "--bf16" in sys.argv else 4) / ms / 1e9,
            graph.nnz / ms / 1e6))


if __name__ == "__main__":
    main()
