"""Programmatic dependent launch of the layer's kernel chain (GCNB_TUNE_PDL, csrc/common.cuh): the CBG step with
the knob off and on in ONE process -- results must be bit-identical (same kernels, same order), then the step is
timed as a CUDA-graph replay with the L2 flushed, like bench.py.

    python tools/pdl_probe.py [workload=cbg] [reps=40]

The PDL instantiations were written in round 1 after the GPU budget was spent: run this before enabling the knob
anywhere.
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
    wl = B.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cbg"]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    dev = torch.device("cuda:0")
    lib = _lib.load()
    graph = B.make_graph(P, torch, wl, dev)
    n, fin, fout = wl["n"], wl["fin"], wl["fout"]
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, fin, generator=gen, device=dev)
    g = torch.randn(n, fout, generator=gen, device=dev)
    flush_buf = torch.empty(B.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def flush():
        _lib.check(lib.gcnb_l2_flush(ctypes.c_void_p(flush_buf.data_ptr()), flush_buf.numel(),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "l2_flush")

    def step():
        layer.weight.grad = None
        layer.bias.grad = None
        out = layer(x, graph)
        out.backward(g)
        return out

    results = {}
    for pdl in (0, 1, 0, 1):
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_PDL, pdl), "set_tuning")
        for _ in range(3):
            out = step()
        torch.cuda.synchronize()
        snap = (out.detach().clone(), layer.weight.grad.clone(), layer.bias.grad.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(s)
        layer.weight.grad = None
        layer.bias.grad = None
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            o_static = layer(x, graph)
            o_static.backward(g)
        evs = []
        for it in range(reps + 5):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            cg.replay()
            b.record()
            if it >= 5:
                evs.append((a, b))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        replay = (o_static.detach().clone(), layer.weight.grad.clone(), layer.bias.grad.clone())
        same_replay = all(torch.equal(p, q) for p, q in zip(snap, replay))
        print("PDL %d: %.4f ms per step (graph replay, L2 flushed); replay == eager: %s" % (pdl, ms, same_replay), flush=True)
        if pdl in results:
            continue
        results[pdl] = snap
        del cg
    same = all(torch.equal(p, q) for p, q in zip(results[0], results[1]))
    print("PDL on == PDL off (out, dW, db bit for bit): %s" % same)
    _lib.check(lib.gcnb_set_tuning(_lib.TUNE_PDL, 0), "set_tuning")
    assert same


if __name__ == "__main__":
    main()
