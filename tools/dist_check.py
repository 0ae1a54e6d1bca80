"""Multi-GPU parity + timing check of the row-partitioned layer (run under torchrun, one rank per GPU):

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [n] [deg]

Every rank also computes the single-GPU layer on the whole graph and compares its row block of
out / dX and the all-reduced dW / db, for exchange = peer (own push kernel over NVLink peer memory)
and exchange = nccl (all-gather), over several steps (epochs / acks) and under CUDA-graph replay."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pygcn_b200 as P
from pygcn_b200 import dist as D


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    fin, fout = 64, 32
    gen = torch.Generator(device=dev).manual_seed(0)
    src = torch.randint(0, n, (n * deg // 2,), generator=gen, device=dev, dtype=torch.int32)
    dst = torch.randint(0, n, (n * deg // 2,), generator=gen, device=dev, dtype=torch.int32)
    full = P.Graph.from_edges(src, dst, n)
    x = torch.randn(n, fin, generator=gen, device=dev)
    g = torch.randn(n, fout, generator=gen, device=dev)
    torch.manual_seed(42)
    ref = P.GraphConvolution(fin, fout, fuse_relu=True).to(dev)
    xr = x.clone().requires_grad_(True)
    o_ref = ref(xr, full)
    o_ref.backward(g)
    dgraph = D.DistGraph.from_graph(full, rank, world, per_source=True)
    r0, r1 = dgraph.bounds[rank], dgraph.bounds[rank + 1]
    ok = True
    # the partitioned build (no rank holds the whole graph) under the real all-gather: bit-identical blocks
    dg2 = D.build_partitioned(src, dst, n, rank, world, bounds=dgraph.bounds)
    same = dgraph.split or all(torch.equal(a_, b_) for got, want in ((dg2.fwd_remote, dgraph.fwd_remote),
                                                                       (dg2.bwd_remote, dgraph.bwd_remote))
                               for a_, b_ in zip(got.csr(), want.csr()))
    ok = ok and same and dg2.nnz_global == full.nnz
    if rank == 0:
        print("partitioned build == cut of the full graph: %s (nnz_global %d)" % (same, dg2.nnz_global), flush=True)
    # GCNB_DIST_CHECK_EXCHANGES=peer,nccl,nvls adds the NVLS multicast exchange (written at the end of round 1, not yet run)
    # "nccl-bf16": the bf16 panel tier of the row-partitioned layer (dist_spmm_bf16), checked against the single-GPU layer
    # of the same tier (same two panels rounded: equal to summation order)
    fp32_ref = (o_ref, xr.grad, ref.weight.grad, ref.bias.grad)
    for exchange in os.environ.get("GCNB_DIST_CHECK_EXCHANGES", "peer,nccl").split(","):
        precision = "auto"
        o_ref, dx_ref, dw_ref, db_ref = fp32_ref
        if exchange.endswith("-bf16"):
            exchange, precision = exchange[:-5], "bf16"
            torch.manual_seed(42)
            ref16 = P.GraphConvolution(fin, fout, fuse_relu=True, precision="bf16").to(dev)
            x16 = x.clone().requires_grad_(True)
            o_ref = ref16(x16, full)
            o_ref.backward(g)
            o_ref, dx_ref, dw_ref, db_ref = o_ref.detach(), x16.grad, ref16.weight.grad, ref16.bias.grad
        torch.manual_seed(42)
        layer = D.DistGraphConvolution(fin, fout, fuse_relu=True, exchange=exchange, precision=precision).to(dev)
        xl = x[r0:r1].clone().requires_grad_(True)
        gl = g[r0:r1].contiguous()
        for step in range(4):
            layer.inner.weight.grad = None
            layer.inner.bias.grad = None
            xl.grad = None
            out = layer(xl, dgraph)
            out.backward(gl)
        torch.cuda.synchronize()
        errs = {
            "out": ((out - o_ref[r0:r1]).abs().max() / o_ref.abs().max()).item(),
            "dX": ((xl.grad - dx_ref[r0:r1]).abs().max() / dx_ref.abs().max()).item(),
            "dW": ((layer.inner.weight.grad - dw_ref).abs().max() / dw_ref.abs().max()).item(),
            "db": ((layer.inner.bias.grad - db_ref).abs().max() / db_ref.abs().max()).item(),
        }
        good = all(v < 1e-5 for v in errs.values())
        ok = ok and good
        # CUDA-graph replay of the step (kernels + push / wait / ack or NCCL), then timing
        # (fresh leaf: an AccumulateGrad node born on the default stream would pull the legacy stream into
        # the capture)
        torch.manual_seed(42)
        layer = D.DistGraphConvolution(fin, fout, fuse_relu=True, exchange=exchange, precision=precision).to(dev)
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
        ms = torch.tensor([a.elapsed_time(b) / 20], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        e_graph = ((o_static - o_ref[r0:r1]).abs().max() / o_ref.abs().max()).item()
        ok = ok and e_graph < 1e-5
        if rank == 0:
            print("exchange=%s precision=%s world=%d n=%d nnz=%d: errs %s graph-replay out err %.2e  step %.3f ms (max over ranks, no L2 flush) %s" % (
                exchange, precision, world, n, full.nnz, {k: "%.1e" % v for k, v in errs.items()}, e_graph, ms.item(),
                "OK" if good else "FAIL"), flush=True)
        del cg, layer
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ALL OK" if flag.item() == 1 else "MISMATCH", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
