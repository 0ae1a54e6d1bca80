"""Bandwidth of the peer push kernel (P2P stores over NVLink) vs the copy engine, 2+ ranks under torchrun."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pygcn_b200 import dist as D


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    pad, f = 400_000, 32  # 51.2 MB slot
    ex = D.PeerExchange(rank, world, pad, f, dev)
    ex.my_slot.normal_()
    torch.cuda.synchronize()
    dist.barrier()
    for ctas in (16, 32, 64, 128, 296):
        ex.PUSH_CTAS = ctas
        ts = []
        for it in range(6):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ex.start()
            ex.finish()
            b.record()
            # consumer side: wait + ack so the next push may proceed
            for q in range(world):
                ex.wait(q)
                ex.done(q)
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        ms = sum(ts) / len(ts)
        if rank == 0:
            print("push kernel ctas=%3d: %.3f ms for %d x %.1f MB -> %.1f GB/s egress per GPU" % (
                ctas, ms, world - 1, ex.slot_bytes / 1e6, (world - 1) * ex.slot_bytes / ms / 1e6), flush=True)
    # copy engine: cudaMemcpyAsync to the mapped peer pointers
    cudart = ctypes.CDLL("libcudart.so")
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for it in range(6):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(1, world):
            r = (rank - k) % world
            rc = cudart.cudaMemcpyAsync(ctypes.c_void_p(ex.peer_ptr[r] + rank * ex.slot_bytes),
                                        ctypes.c_void_p(ex.ptr + rank * ex.slot_bytes), ctypes.c_size_t(ex.slot_bytes),
                                        ctypes.c_int(3), ctypes.c_void_p(st))
            assert rc == 0, rc
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    if rank == 0:
        print("copy engine      : %.3f ms -> %.1f GB/s egress per GPU" % (ms, (world - 1) * ex.slot_bytes / ms / 1e6), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
