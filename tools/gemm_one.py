"""One tcgen05 product on one shape (the command ncu wraps):  python tools/gemm_one.py rows|tn M N K"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygcn_b200 as P

mode, m, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dev = torch.device("cuda:0")
torch.manual_seed(0)
if mode == "rows":   # C[M,N] = A[M,K] B[K,N]
    a, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev)
    for _ in range(3):
        c = P.mm(a, b, precision="tf32x3")
else:                # C[M,N] = X[K,M]^T Y[K,N]  (K = reduction rows)
    x, y = torch.randn(k, m, device=dev), torch.randn(k, n, device=dev)
    for _ in range(3):
        c = P.mm(x.t(), y, precision="tf32x3")
torch.cuda.synchronize()
print(mode, m, n, k, float(c.abs().max()))
