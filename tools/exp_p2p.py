"""Developer tool (torchrun, 2 ranks): latency of one neighbour hand-off through dist.PeerLink against NCCL send/recv,
host time per call included."""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from infinitevl_b200.dist import PeerLink

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
for nbytes in (2 << 20, 8 << 20):
    buf = torch.zeros(nbytes // 4, dtype=torch.float32, device=dev)
    src = torch.ones(nbytes // 4, dtype=torch.float32, device=dev)
    link = PeerLink(rank, world, None, dev)
    link.open({"X": buf})
    N = 50
    for phase in ("warm", "timed"):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        host = 0.0
        for i in range(N):
            h0 = time.perf_counter()
            if rank == 0:
                link.before_overwrite("X")
                link.send("X", src)
            else:
                link.wait("X")
                buf.add_(0)          # a kernel that reads the buffer
                link.release("X")
            host += time.perf_counter() - h0
        link.drain()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if phase == "timed":
            print(f"rank {rank} p2p  {nbytes >> 20} MiB: {1e6 * (t1 - t0) / N:8.1f} us per hand-off, host {1e6 * host / N:8.1f} us per call", flush=True)
    for phase in ("warm", "timed"):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(N):
            if rank == 0:
                w = dist.isend(src, dst=1)
            else:
                w = dist.irecv(buf, src=0)
                w.wait(); buf.add_(0)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if phase == "timed":
            print(f"rank {rank} nccl {nbytes >> 20} MiB: {1e6 * (t1 - t0) / N:8.1f} us per hand-off", flush=True)
    dist.barrier()
dist.destroy_process_group()
