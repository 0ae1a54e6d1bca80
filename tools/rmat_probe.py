"""SpMM on the products-shaped R-MAT graph (BASELINE configs[3]) at several panel widths, fp32 and bf16 panels,
with the vertex labels as generated, relabelled by descending degree, or shuffled -- what the locality of the row /
column order is worth.  Development tool; also the command ncu wraps (one width, one relabelling).

    python tools/rmat_probe.py [--widths 256,100,48] [--relabel none,degree,random,rows] [--bf16] [--reps 5] [--workload products]
                               [--sweep stream:hot_mb:hint:batch,...]   e.g. 0:0:0:0,2:48:0:0,2:96:2:8  (batch = gathered rows in flight per lane: 0 auto, 2 / 4 / 8)
--sweep: every configuration of the streaming kernel (gcnb_set_tuning: GCNB_TUNE_SPMM_STREAM, _HOT_MB, _HINT, _BATCH) in
one process on the same graph; --slices 1,2,4: the panel in K column slices, one launch each (what narrower gathered
rows are worth in L2 hits: profiles/r02_rmat_probe_column_slices_old_kernels.txt); --fwd-only skips the transposed launch; --check compares with torch's CUDA CSR product.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as B
import pygcn_b200 as P
from pygcn_b200 import _lib


def opt(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def main():
    widths = [int(v) for v in opt("--widths", "256,100,48").split(",")]
    relabels = opt("--relabel", "none").split(",")
    reps = int(opt("--reps", "5"))
    wl = B.WORKLOADS[opt("--workload", "products")]
    bf16 = "--bf16" in sys.argv
    dev = torch.device("cuda:0")
    lib = _lib.load()
    flush_buf = torch.empty(B.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # --cpu-edges: the bench's own seeded edge list (CPU generator) instead of a device-drawn graph of the same family
    src0, dst0, n = B.make_edges(torch, wl, 0, device="cpu" if "--cpu-edges" in sys.argv else dev)
    src0, dst0 = src0.to(dev), dst0.to(dev)
    for how in relabels:
        src, dst = src0.long(), dst0.long()
        if how == "degree":
            deg = torch.bincount(src, minlength=n) + torch.bincount(dst, minlength=n)
            order = torch.argsort(deg, descending=True, stable=True)
            newid = torch.empty_like(order)
            newid[order] = torch.arange(n, device=dev)
            src, dst = newid[src], newid[dst]
        elif how == "random":
            newid = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
            src, dst = newid[src], newid[dst]
        graph = P.Graph.from_edges(src.int(), dst.int(), n)
        if how == "rows":  # rows processed in order of descending length, columns (= the panel) left as they are
            coo = graph.to_sparse_coo()
            idx, val = coo._indices(), coo._values()
            deg = torch.bincount(idx[0], minlength=n)
            order = torch.argsort(deg, descending=True, stable=True)
            newrow = torch.empty_like(order)
            newrow[order] = torch.arange(n, device=dev)
            graph = P.Graph.from_torch(torch.sparse_coo_tensor(torch.stack([newrow[idx[0]], idx[1]]), val, (n, n)))
            del coo, idx, val
        print("relabel=%s graph %r bins %s long_chunks %d max_degree %d" % (how, graph, graph.bin_rows, graph.n_long_chunks,
                                                                            graph.max_degree), flush=True)
        for f in widths:
            s = torch.randn(n, f, device=dev)
            out = torch.empty(n, f, device=dev)
            ref = None
            if "--check" in sys.argv:
                csr = graph.to_sparse_coo().coalesce().to_sparse_csr()
                ref = torch.sparse.mm(csr, s)
                del csr
            for cfg in opt("--sweep", "default").split(",") + ["gv%s" % v for v in opt("--group-variants", "").split(",") if v]:
              if cfg.startswith("gv"):  # --group-variants 0,2,13,-2: the group-per-row kernel's variants (narrow panels), streaming
                  # kernel off; -2 = the warp-per-row kernel
                  _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_STREAM, 0), "set_tuning")
                  _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_KERNEL, 2 if int(cfg[2:]) >= -1 else 1), "set_tuning")
                  _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_GROUP_VARIANT, max(int(cfg[2:]), -1)), "set_tuning")
              elif cfg != "default":
                  vals = [int(v) for v in cfg.split(":")]
                  for key, v in zip((_lib.TUNE_SPMM_STREAM, _lib.TUNE_STREAM_HOT_MB, _lib.TUNE_STREAM_HINT, _lib.TUNE_STREAM_BATCH), vals):
                      _lib.check(lib.gcnb_set_tuning(key, v), "set_tuning")
              print(" config stream:hot_mb:hint:batch (gvN = group variant N) = %s" % cfg)
              for slices in [int(v) for v in opt("--slices", "1").split(",")]:
               if slices > 1:
                  print(" slices %d" % slices)
               for prec in (("fp32", "bf16") if bf16 else ("fp32",)):
                  if prec == "bf16":
                      ld8 = (f + 7) // 8 * 8
                      panel = torch.empty(n, ld8, dtype=torch.bfloat16, device=dev)
                      _lib.check(lib.gcnb_to_bf16(n, f, ctypes.c_void_p(s.data_ptr()), f, ctypes.c_void_p(panel.data_ptr()), ld8, st),
                                 "to_bf16")
                  for tflag, name in (((0, "fwd"),) if "--fwd-only" in sys.argv else ((0, "fwd"), (_lib.SPMM_TRANSPOSE, "A^T"))):
                      ws = torch.empty(max(lib.gcnb_spmm_workspace_bytes(graph._h, tflag, f), 256), dtype=torch.uint8, device=dev)
                      ts = []
                      for it in range(2 + reps):
                          lib.gcnb_l2_flush(ctypes.c_void_p(flush_buf.data_ptr()), flush_buf.numel(), st)
                          a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                          a.record()
                          if prec == "bf16":
                              _lib.check(lib.gcnb_spmm_bf16(graph._h, tflag, ctypes.c_void_p(panel.data_ptr()), ld8, f, None,
                                                            ctypes.c_void_p(out.data_ptr()), f, ctypes.c_void_p(ws.data_ptr()),
                                                            ws.numel(), st), "spmm_bf16")
                          else:
                              wsl = f // slices  # --slices K: the panel in K column slices, one launch each (L2 locality)
                              for k in range(slices):
                                  _lib.check(lib.gcnb_spmm(graph._h, tflag, ctypes.c_void_p(s.data_ptr() + 4 * k * wsl), f, wsl, None,
                                                           ctypes.c_void_p(out.data_ptr() + 4 * k * wsl), f,
                                                           ctypes.c_void_p(ws.data_ptr()), ws.numel(), st), "spmm")
                          b.record()
                          torch.cuda.synchronize()
                          if it >= 2:
                              ts.append(a.elapsed_time(b))
                      if ref is not None and tflag == 0 and prec == "fp32":
                          print("   check vs torch CSR spmm: %.2e" % ((out - ref).abs().max() / ref.abs().max()).item())
                      ms = sum(ts) / len(ts)
                      es = 2 if prec == "bf16" else 4
                      alg = graph.nnz * 8 + (n + 1) * 4 + n * f * es + n * f * 4
                      print("  f=%d %s %s: %.3f ms (min %.3f)  alg %.0f GB/s (frac %.3f of 6534.8)  gather %.2f TB/s  %.2f Gedges/s" % (
                          f, prec, name, ms, min(ts), alg / ms / 1e6, alg / ms / 1e6 / 6534.8, graph.nnz * f * es / ms / 1e9,
                          graph.nnz / ms / 1e6), flush=True)
            del s, out
        del graph


if __name__ == "__main__":
    main()
