"""Development check of the tcgen05 3xTF32 GEMMs against fp64 (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pygcn_b200 as P

dev = torch.device("cuda:0")
torch.manual_seed(0)


def err(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max()).item()


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


ok = True
# mode R: C = A[M,K] B[K,N]
for (m, n, k, bt) in [(128, 32, 32, False), (128, 16, 64, False), (1000, 32, 64, False), (100000, 32, 64, False),
                      (100000, 64, 32, True), (5000, 47, 100, False), (3000, 600, 16, False), (2708, 16, 1432, False),
                      (70000, 256, 256, False), (333, 7, 16, True)]:
    a = torch.randn(m, k, device=dev)
    b = torch.randn(n, k, device=dev).t() if bt else torch.randn(k, n, device=dev)
    ref = a.double() @ b.double()
    c = P.mm(a, b, precision="tf32x3")
    c32 = P.mm(a, b, precision="fp32")
    e, e32 = err(c, ref), err(c32, ref)
    t = timeit(lambda: P.mm(a, b, precision="tf32x3"))
    t32 = timeit(lambda: P.mm(a, b, precision="fp32"))
    tt = timeit(lambda: torch.mm(a, b))
    print(f"R m={m} n={n} k={k} bt={bt}: err tf32x3 {e:.2e} fp32 {e32:.2e} | us tc {t:.1f} simt {t32:.1f} torch {tt:.1f}")
    ok &= e < 1e-5
# mode T: C[M,N] = X[R,M]^T Y[R,N]
for (r, m, n) in [(32, 128, 32), (64, 64, 32), (100000, 64, 32), (100000, 32, 32), (2708, 1432, 16), (50000, 256, 256),
                  (4097, 100, 48)]:
    x = torch.randn(r, m, device=dev)
    y = torch.randn(r, n, device=dev)
    ref = x.double().t() @ y.double()
    c = P.mm(x.t(), y, precision="tf32x3")
    c32 = P.mm(x.t(), y, precision="fp32")
    e, e32 = err(c, ref), err(c32, ref)
    t = timeit(lambda: P.mm(x.t(), y, precision="tf32x3"))
    t32 = timeit(lambda: P.mm(x.t(), y, precision="fp32"))
    tt = timeit(lambda: torch.mm(x.t(), y))
    print(f"T r={r} m={m} n={n}: err tf32x3 {e:.2e} fp32 {e32:.2e} | us tc {t:.1f} simt {t32:.1f} torch {tt:.1f}")
    ok &= e < 1e-5
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
