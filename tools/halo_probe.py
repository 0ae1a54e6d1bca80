"""The halo exchange of the row-partitioned products-shaped graph alone (run under torchrun): pack + grouped ncclSend /
ncclRecv of a [rows, F] panel through gcnb_halo_*, its bandwidth per rank, and the whole exchanged SpMM beside it.
NCCL's P2P channel count is an environment matter (NCCL_MIN_P2P_NCHANNELS / NCCL_MAX_P2P_NCHANNELS): run it once per setting.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/halo_probe.py [F=256] [reps=10]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench as B
import pygcn_b200 as P
from pygcn_b200 import _lib
from pygcn_b200 import dist as D


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    f = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    wl = B.WORKLOADS["products"]
    src, dst, n = B.make_edges(torch, wl, 0, device=dev)
    full = P.Graph.from_edges(src, dst, n)
    dg = D.DistGraph.from_graph(full, rank, world, split=True, row_weight=12.0)
    del full
    ops = D.CudaOps()
    plan = D.halo_plans(ops, dg)[0]
    n_p = dg.n_rows()
    panel = torch.randn(n_p, f, device=dev)
    compact = torch.empty(plan.n_compact, f, device=dev)
    lib = _lib.load()
    timer = B.Timer(torch, lib, _lib, dev)

    def exchange_only():
        done = ops.halo_exchange_async(plan.c, panel, compact)
        torch.cuda.current_stream().wait_event(done)

    def whole():
        D.exchanged_spmm(ops, dg, False, panel, "halo")
    for name, fn in (("exchange only (pack + send/recv)", exchange_only), ("exchanged SpMM (diag overlapped)", whole)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        ms = timer.time(fn, reps)[0]
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows = torch.tensor([plan.rows_received, plan.c["send_rows"]], device=dev, dtype=torch.float64)
        dist.all_reduce(rows, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("world=%d F=%d %s: %.3f ms (max over ranks); largest receive %d rows = %.0f MB -> %.0f GB/s in; largest send %d rows"
                  % (world, f, name, t.item(), rows[0].item(), rows[0].item() * f * 4 / 1e6, rows[0].item() * f * 4 / t.item() / 1e6,
                     rows[1].item()), flush=True)
    if rank == 0:
        print("NCCL_MIN_P2P_NCHANNELS=%s NCCL_MAX_P2P_NCHANNELS=%s" % (os.environ.get("NCCL_MIN_P2P_NCHANNELS"), os.environ.get("NCCL_MAX_P2P_NCHANNELS")))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
