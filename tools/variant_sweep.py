"""All group-per-row SpMM variants (gcnb_set_tuning) on the same graphs in ONE process: parity with torch's
CUDA CSR spmm, then forward / transposed launch times with the L2 flushed.

    python tools/variant_sweep.py [variants=0,1,2,...] [reps=20] [--bf16]

--bf16: the same sweep through gcnb_spmm_bf16 (panel rounded once with gcnb_to_bf16; parity bound 2e-2 against the
fp32 product).
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
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    bf16 = "--bf16" in sys.argv
    variants = [int(v) for v in (argv[0] if len(argv) > 0 else "0,2,13,14").split(",")]
    reps = int(argv[1]) if len(argv) > 1 else 20
    dev = torch.device("cuda:0")
    lib = _lib.load()
    flush_buf = torch.empty(B.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def flush():
        _lib.check(lib.gcnb_l2_flush(ctypes.c_void_p(flush_buf.data_ptr()), flush_buf.numel(), st), "l2_flush")

    for name, widths in (("cbg", (32, 16, 64)), ("400000:25", (32,))):
        if ":" in name:
            n_, d_ = name.split(":")
            wl = dict(n=int(n_), avg_deg=int(d_), fin=64, fout=32, name=name)
        else:
            wl = B.WORKLOADS[name]
        graph = B.make_graph(P, torch, wl, dev)
        n = graph.n_rows
        csr = graph.to_sparse_coo().coalesce().to_sparse_csr()
        for f in widths:
            s = torch.randn(n, f, device=dev)
            out = torch.empty(n, f, device=dev)
            ref = torch.sparse.mm(csr, s)
            ws = torch.empty(max(lib.gcnb_spmm_workspace_bytes(graph._h, 0, f), 256), dtype=torch.uint8, device=dev)

            if bf16:
                ld8 = (f + 7) // 8 * 8
                panel = torch.empty(n, ld8, dtype=torch.bfloat16, device=dev)
                _lib.check(lib.gcnb_to_bf16(n, f, ctypes.c_void_p(s.data_ptr()), f, ctypes.c_void_p(panel.data_ptr()), ld8, st),
                           "to_bf16")

            def spmm(tflag):
                if bf16:
                    _lib.check(lib.gcnb_spmm_bf16(graph._h, tflag, ctypes.c_void_p(panel.data_ptr()), ld8, f, None,
                                                  ctypes.c_void_p(out.data_ptr()), f, ctypes.c_void_p(ws.data_ptr()),
                                                  ws.numel(), st), "spmm_bf16")
                    return
                _lib.check(lib.gcnb_spmm(graph._h, tflag, ctypes.c_void_p(s.data_ptr()), f, f, None,
                                         ctypes.c_void_p(out.data_ptr()), f, ctypes.c_void_p(ws.data_ptr()), ws.numel(), st),
                           "spmm")
            for rnd in range(2):  # two interleaved rounds: the spread between them is the noise
                for v in variants:
                    _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_GROUP_VARIANT, v), "set_tuning")
                    out.fill_(float("nan"))
                    spmm(0)
                    err = ((out - ref).abs().max() / ref.abs().max()).item()
                    ts = {}
                    for tflag in (0, _lib.SPMM_TRANSPOSE):
                        evs = []
                        for it in range(reps + 3):
                            flush()
                            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            a.record()
                            spmm(tflag)
                            b.record()
                            if it >= 3:
                                evs.append((a, b))
                        torch.cuda.synchronize()
                        ts[tflag] = 1e3 * sum(a.elapsed_time(b) for a, b in evs) / len(evs)
                    print("%-10s f=%-3d variant %d round %d: fwd %6.1f us  A^T %6.1f us  err %.1e%s" % (
                        name, f, v, rnd, ts[0], ts[_lib.SPMM_TRANSPOSE], err, "" if err < (2e-2 if bf16 else 1e-5) else "  PARITY FAIL"), flush=True)
            _lib.check(lib.gcnb_set_tuning(_lib.TUNE_SPMM_GROUP_VARIANT, -1), "set_tuning")
        del graph, csr


if __name__ == "__main__":
    main()
