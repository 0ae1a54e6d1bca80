"""Time NCCL all_gather_into_tensor of the CBG panel slots (12.8 MB per rank) under the current NCCL_* environment."""
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_048
    slot = torch.randn(rows, 32, device=dev)
    out = torch.empty(world * rows, 32, device=dev)
    for _ in range(5):
        dist.all_gather_into_tensor(out, slot)
    torch.cuda.synchronize()
    res = {}
    for mode in ("eager", "graph"):
        if mode == "graph":
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                dist.all_gather_into_tensor(out, slot)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                dist.all_gather_into_tensor(out, slot)
            fn = g.replay
        else:
            fn = lambda: dist.all_gather_into_tensor(out, slot)
        ts = []
        for it in range(23):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        t = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = t.item() * 1e3
    if rank == 0:
        env = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
        mb = slot.numel() * 4 / 1e6
        print("all-gather %d x %.1f MB: eager %.1f us, graph %.1f us (%.0f GB/s in per GPU)  env %s" % (
            world, mb, res["eager"], res["graph"], (world - 1) * mb / res["graph"] * 1e3, env), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
