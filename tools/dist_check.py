"""Multi-GPU parity + timing check of the row-partitioned layer (run under torchrun, one rank per GPU):

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [n] [raw_edges]

Every rank also computes the single-GPU layer on the whole (R-MAT, power-law) graph and compares its row block of
out / dX and the all-reduced dW / db, for every exchange (GCNB_DIST_CHECK_EXCHANGES, default nccl,halo,peer), split and
unsplit row blocks, both association orders (64 -> 32: A (X W); 32 -> 96: (A X) W), over several steps and under
CUDA-graph replay."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench as B
import pygcn_b200 as P
from pygcn_b200 import dist as D


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    n_raw = int(sys.argv[2]) if len(sys.argv) > 2 else 900000
    gen = torch.Generator(device=dev).manual_seed(0)
    src, dst = B.rmat_edges(torch, n, n_raw, gen, dev)
    full = P.Graph.from_edges(src, dst, n)
    ok = True
    # the partitioned build (no rank holds the whole graph) under the real all-gather: bit-identical blocks
    dgu = D.DistGraph.from_graph(full, rank, world, split=False, per_source=True, row_weight=8.0)
    dg2 = D.build_partitioned(src, dst, n, rank, world, bounds=dgu.bounds)
    same = all(torch.equal(a_, b_) for got, want in ((dg2.fwd_remote, dgu.fwd_remote), (dg2.bwd_remote, dgu.bwd_remote))
               for a_, b_ in zip(got.csr(), want.csr()))
    ok = ok and same and dg2.nnz_global == full.nnz
    if rank == 0:
        print("partitioned build == cut of the full graph: %s (nnz_global %d, bounds %s)" % (same, dg2.nnz_global, dgu.bounds), flush=True)
    dgs = D.DistGraph.from_graph(full, rank, world, split=True, bounds=dgu.bounds)
    r0, r1 = dgu.bounds[rank], dgu.bounds[rank + 1]
    for exchange in os.environ.get("GCNB_DIST_CHECK_EXCHANGES", "nccl,halo,peer").split(","):
        for fin, fout in ((64, 32), (32, 96), (64, 47)):  # (47: panel rows padded to 16 bytes for the exchange)
            for dgraph, what in ((dgu, "unsplit"), (dgs, "split")):
                if (fin, fout) == (64, 47) and what == "unsplit":
                    continue
                if exchange == "peer" and (what == "split" or fin < fout or fout % 4):
                    continue  # the peer exchange consumes per-source blocks of the unsplit row block, reference order
                x = torch.randn(n, fin, generator=gen, device=dev)
                g = torch.randn(n, fout, generator=gen, device=dev)
                torch.manual_seed(42)
                ref = P.GraphConvolution(fin, fout, fuse_relu=True).to(dev)
                xr = x.clone().requires_grad_(True)
                o_ref = ref(xr, full)
                o_ref.backward(g)
                torch.manual_seed(42)
                layer = D.DistGraphConvolution(fin, fout, fuse_relu=True, exchange=exchange).to(dev)
                xl = x[r0:r1].clone().requires_grad_(True)
                gl = g[r0:r1].contiguous()
                for step in range(3):
                    layer.inner.weight.grad = None
                    layer.inner.bias.grad = None
                    xl.grad = None
                    out = layer(xl, dgraph)
                    out.backward(gl)
                torch.cuda.synchronize()
                errs = {
                    "out": ((out - o_ref[r0:r1]).abs().max() / o_ref.abs().max()).item(),
                    "dX": ((xl.grad - xr.grad[r0:r1]).abs().max() / xr.grad.abs().max()).item(),
                    "dW": ((layer.inner.weight.grad - ref.weight.grad).abs().max() / ref.weight.grad.abs().max()).item(),
                    "db": ((layer.inner.bias.grad - ref.bias.grad).abs().max() / ref.bias.grad.abs().max()).item(),
                }
                # out / dX to the layer's bar.  dW / db: a ReLU output within rounding of zero may flip its mask between
                # the two evaluations, and one flipped entry of G moves a column sum of ~sqrt(N) by ~1 / sqrt(N): allow
                # a few flips (they show as 1e-3-sized errors; a wrong exchange or a missing all-reduce shows as O(1))
                flips = ((out > 0) != (o_ref[r0:r1] > 0)).sum()
                dist.all_reduce(flips)
                slack = 2e-5 + float(flips.item()) * 4.0 / (n ** 0.5)
                good = errs["out"] < 2e-5 and errs["dX"] < max(2e-5, slack) and errs["dW"] < slack and errs["db"] < slack
                errs["mask_flips"] = float(flips.item())
                ok = ok and good
                # CUDA-graph replay of the step (kernels + NCCL sends / receives / all-reduce), then timing (fresh leaf:
                # an AccumulateGrad node born on the default stream would pull the legacy stream into the capture)
                torch.manual_seed(42)
                layer = D.DistGraphConvolution(fin, fout, fuse_relu=True, exchange=exchange).to(dev)
                s_ = torch.cuda.Stream()
                s_.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s_):
                    xl = x[r0:r1].clone()
                    for _ in range(2):
                        layer.inner.weight.grad = None
                        layer.inner.bias.grad = None
                        layer(xl, dgraph).backward(gl)
                torch.cuda.current_stream().wait_stream(s_)
                torch.cuda.synchronize()
                dist.barrier()
                layer.inner.weight.grad = None
                layer.inner.bias.grad = None
                e_graph, ms = float("nan"), float("nan")
                try:
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg):
                        o_static = layer(xl, dgraph)
                        o_static.backward(gl)
                    torch.cuda.synchronize()
                    dist.barrier()
                    for _ in range(3):
                        cg.replay()
                    torch.cuda.synchronize()
                    dist.barrier()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(20):
                        cg.replay()
                    b.record()
                    torch.cuda.synchronize()
                    t = torch.tensor([a.elapsed_time(b) / 20], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = t.item()
                    e_graph = ((o_static - o_ref[r0:r1]).abs().max() / o_ref.abs().max()).item()
                    ok = ok and e_graph < 2e-5
                    del cg
                except Exception as e:  # a capture failure is reported, not fatal: bench.py falls back to eager launches
                    print("rank %d: graph capture failed for exchange=%s: %r" % (rank, exchange, e), flush=True)
                    torch.cuda.synchronize()
                if rank == 0:
                    print("exchange=%s %s %d->%d world=%d n=%d nnz=%d: errs %s graph-replay out err %.2e  step %.3f ms (max over "
                          "ranks, no L2 flush) %s" % (exchange, what, fin, fout, world, n, full.nnz,
                                                      {k: "%.1e" % v for k, v in errs.items()}, e_graph, ms,
                                                      "OK" if good else "FAIL"), flush=True)
                del layer
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ALL OK" if flag.item() == 1 else "MISMATCH", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
