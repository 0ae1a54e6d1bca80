"""Rows-mode tcgen05 product C[M,N] = A[M,K] B[K,N] at M = 2 449 029 over a sweep of K and N: where the time of a tile
goes when the mainloop is short (small K) -- the epilogue (TMEM -> registers -> staging -> global) or the pipeline.

    python tools/gemm_k_sweep.py [M]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygcn_b200 as P

m = int(sys.argv[1]) if len(sys.argv) > 1 else 2449029
dev = torch.device("cuda:0")
torch.manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for n in (256, 128, 64):
    for k in (32, 48, 64, 100, 128, 256):
        a, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev)
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c = P.mm(a, b, precision="tf32x3")
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        by = (m * k + m * n) * 4
        tiles = (m + 127) // 128
        print("M=%d K=%-3d N=%-3d  %.3f ms  %.0f GB/s  %.2f us per 128-row tile per SM (%.0f tiles per SM)" % (
            m, k, n, ms, by / ms / 1e6, ms * 1e3 / (tiles / 148), tiles / 148), flush=True)
        del a, b, c
