"""dW-shaped tcgen05 products C[M,N] = X[R,M]^T Y[R,N] at R = 2 449 029 (the products model's three layers): time per launch.

    python tools/gemm_tn_probe.py [R]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygcn_b200 as P

r = int(sys.argv[1]) if len(sys.argv) > 1 else 2449029
dev = torch.device("cuda:0")
torch.manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for m, n in ((256, 256), (100, 256), (256, 47), (128, 128), (256, 64)):
    # widths that are not a multiple of 4 as the layer holds them: in a buffer whose rows are padded to 16 bytes
    x, y = torch.randn(r, m, device=dev), torch.randn(r, (n + 3) // 4 * 4, device=dev)[:, :n]
    ref = (x[:200000].double().t() @ y[:200000].double())
    got = P.mm(x[:200000].t(), y[:200000], precision="tf32x3").double()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c = P.mm(x.t(), y, precision="tf32x3")
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    print("R=%d M=%-3d N=%-3d  %.3f ms  %.0f GB/s  err (200 K rows, vs fp64) %.2e" % (r, m, n, ms, (r * (m + n)) * 4 / ms / 1e6, err), flush=True)
    del x, y
