"""Device time of one layer fwd+bwd (CUDA-graph replay, L2 flushed) for ours and for the reference's
three lines run on CUDA by PyTorch (cuBLAS / cuSPARSE), on a few workloads.  Development tool."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pygcn_b200 as P
from pygcn_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load()
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def flush():
    lib.gcnb_l2_flush(ctypes.c_void_p(flush_buf.data_ptr()), flush_buf.numel(),
                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))


def graph_time(fn, steps=20):
    for _ in range(3):
        fn()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ev = []
    for _ in range(3 + steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in ev[3:]]
    return sum(t) / len(t) * 1e3


def run(name, adj_graph, adj_torch, n, fin, fout, need_dx=False):
    x = torch.randn(n, fin, device=dev, requires_grad=need_dx)
    g = torch.randn(n, fout, device=dev)
    layer = P.GraphConvolution(fin, fout).to(dev)

    def ours():
        layer.weight.grad = None
        layer.bias.grad = None
        o = layer(x, adj_graph)
        o.backward(g)

    w = layer.weight.detach().clone().requires_grad_(True)
    b = layer.bias.detach().clone().requires_grad_(True)

    def ref():
        w.grad = None
        b.grad = None
        o = torch.spmm(adj_torch, torch.mm(x, w)) + b
        o.backward(g)

    t_ours = graph_time(ours)
    if adj_torch.layout == torch.strided:
        t_ref = graph_time(ref)
    else:  # torch's sparse ops synchronise / allocate: not capturable, time eager launches
        for _ in range(3):
            ref()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            ref()
        b_.record()
        torch.cuda.synchronize()
        t_ref = a.elapsed_time(b_) / 10 * 1e3
        name += " (ref eager)"
    print("%-60s ours %8.1f us   torch-on-CUDA reference lines %8.1f us   x%.2f" % (name, t_ours, t_ref, t_ref / t_ours))


def run_batched(name, adj, n, bsz, fin, fout):
    """fork's evaluator step shape: the same layer on B samples -- one batched call vs the per-sample loop."""
    x = torch.randn(bsz, n, fin, device=dev, requires_grad=True)
    g = torch.randn(bsz, n, fout, device=dev)
    layer = P.GraphConvolution(fin, fout, fuse_relu=True).to(dev)

    def batched():
        layer.weight.grad = None
        layer.bias.grad = None
        layer(x, adj).backward(g)

    def loop():
        layer.weight.grad = None
        layer.bias.grad = None
        torch.stack([layer(x[i], adj) for i in range(bsz)]).backward(g)

    tb, tl = graph_time(batched), graph_time(loop)
    print("%-60s batched %8.1f us   per-sample loop %8.1f us   x%.2f" % (name, tb, tl, tl / tb))


if __name__ == "__main__":
    which = sys.argv[1:] or ["fork", "cbg"]
    if "batched" in which:
        n = 2943
        v = torch.rand(40, n, device=dev)
        adj = P.Graph.from_torch((v.t() @ v) / 40)
        run_batched("fork evaluator: dense adj N=2943, B=20, 8->32 (+relu)", adj, n, 20, 8, 32)
        run_batched("fork evaluator: dense adj N=2943, B=20, 32->32 (+relu)", adj, n, 20, 32, 32)
        n = 100_000
        src = torch.randint(0, n, (n * 50,), device=dev, dtype=torch.int32)
        dst = torch.randint(0, n, (n * 50,), device=dev, dtype=torch.int32)
        run_batched("CBG graph N=100000, B=8, 64->32 (+relu)", P.Graph.from_edges(src, dst, n), n, 8, 64, 32)
    if "fork" in which:
        n = 2943
        v = torch.rand(40, n, device=dev)
        adj = (v.t() @ v) / 40
        run("fork shape: dense adj N=2943, 8->32", P.Graph.from_torch(adj), adj, n, 8, 32, need_dx=True)
        run("fork shape: dense adj N=2943, 32->32", P.Graph.from_torch(adj), adj, n, 32, 32, need_dx=True)
    if "cbg" in which:
        n = 100_000
        src = torch.randint(0, n, (n * 50,), device=dev, dtype=torch.int32)
        dst = torch.randint(0, n, (n * 50,), device=dev, dtype=torch.int32)
        gr = P.Graph.from_edges(src, dst, n)
        coo = gr.to_sparse_coo()
        run("CBG: N=100000 nnz=%d 64->32, ref adj = COO as built" % gr.nnz, gr, coo, n, 64, 32)
        run("CBG: same, ref adj = CSR", gr, coo.coalesce().to_sparse_csr(), n, 64, 32)
    if "cora" in which:
        import numpy as np
        g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cora_pipeline.npz"))
        n = int(g["n"])
        e = torch.from_numpy(g["edges"]).to(dev)
        gr = P.Graph.from_edges(e[:, 0].contiguous(), e[:, 1].contiguous(), n)
        run("Cora L1: N=2708 nnz=13264 1433->16", gr, gr.to_sparse_coo(), n, 1433, 16)
        run("Cora L2: 16->7", gr, gr.to_sparse_coo(), n, 16, 7, need_dx=True)
