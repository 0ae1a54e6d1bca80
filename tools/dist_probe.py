"""Where does a multi-GPU SpMM exchange spend its time?  Under torchrun: per rank, CUDA-event timings of
(1) the exchange alone (peer ce / sm push, NCCL all-gather), (2) the per-phase SpMMs alone, (3) the whole
row-block SpMM, (4) the pipelined exchange + SpMM.  Weak-scaled CBG graph like bench.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench as B
import pygcn_b200 as P
from pygcn_b200 import dist as D


def timed(fn, reps=10, pre=None):
    ts = []
    for it in range(reps + 3):
        if pre:
            pre()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(a.elapsed_time(b))
    t = torch.tensor([sum(ts) / len(ts)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    wl = dict(B.WORKLOADS["cbg"])
    wl["n"] = wl["n"] * world
    full = B.make_graph(P, torch, wl, dev)
    dg = D.DistGraph.from_graph(full, rank, world, per_source=True)
    del full
    f = 32
    ops = D.CudaOps()
    ex = D.PeerExchange(rank, world, dg.pad_rows, f, dev)
    ex.my_slot.normal_()
    out = torch.empty(dg.n_rows(), f, device=dev)
    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)
    say("world %d, rows/rank %d, pad %d, phases %s, nnz per phase block %s" % (
        world, dg.n_rows(), dg.pad_rows, dg.phases, [b.nnz for b in dg.fwd_blocks]))

    def consume_all():
        for q in range(world):
            ex.wait(q)
            ex.done(q)

    for mode in ("ce", "sm"):
        ex.PUSH_MODE = mode

        def xchg():
            ex.start()
            consume_all()
            ex.finish()
        say("exchange alone, push=%s: %.1f us" % (mode, timed(xchg)))
    gathered = torch.empty(world * dg.pad_rows, f, device=dev)

    def ag():
        dist.all_gather_into_tensor(gathered, ex.my_slot)
    say("NCCL all-gather alone: %.1f us" % timed(ag))
    # make every slot hold data (one full exchange), then time the SpMMs with nothing to wait for
    ex.PUSH_MODE = "ce"
    ex.start(); consume_all(); ex.finish()
    torch.cuda.synchronize()
    for i, blk in enumerate(dg.fwd_blocks):
        say("  SpMM phase %d (sources %s, nnz %d) alone: %.1f us" % (
            i, dg.phases[i], blk.nnz, timed(lambda: ops.spmm_block(blk, ex.gathered, out, i > 0))))

    def all_phases():
        for i, blk in enumerate(dg.fwd_blocks):
            ops.spmm_block(blk, ex.gathered, out, i > 0)
    say("all phase SpMMs back to back: %.1f us" % timed(all_phases))
    say("whole row block, one SpMM: %.1f us" % timed(lambda: ops.spmm_block(dg.fwd_remote, ex.gathered, out, False)))
    for mode in ("ce", "sm"):
        ex.PUSH_MODE = mode
        say("pipelined exchange + SpMM, push=%s: %.1f us" % (
            mode, timed(lambda: D.dist_spmm_pipelined(ops, dg, dg.fwd_blocks, ex))))

    def nccl_path():
        dist.all_gather_into_tensor(gathered, ex.my_slot)
        ops.spmm_block(dg.fwd_remote, gathered, out, False)
    say("NCCL all-gather + one SpMM: %.1f us" % timed(nccl_path))
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
