import os, sys, traceback
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from infinitevl_b200.dist import PeerLink
from infinitevl_b200 import _lib
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
try:
    buf = torch.zeros(1 << 19, dtype=torch.float32, device=dev)
    src = torch.full((1 << 19,), 3.0, dtype=torch.float32, device=dev)
    link = PeerLink(rank, world, None, dev)
    link.open({"X": buf})
    print(rank, "opened; peer access", torch.cuda.can_device_access_peer(local, 1 - local), flush=True)
    if rank == 0:
        print("peer buf ptr", hex(link.peer_ptr["X"]), "flags", hex(link.next_flags), flush=True)
        link.send("X", src)
        torch.cuda.synchronize()
        print(0, "sent ok", flush=True)
    else:
        link.wait("X")
        torch.cuda.synchronize()
        print(1, "waited ok, sum", float(buf.sum()), "expect", 3.0 * (1 << 19), flush=True)
        link.release("X")
        torch.cuda.synchronize()
        print(1, "released ok", flush=True)
    dist.barrier()
    if rank == 0:
        torch.cuda.synchronize()
        print(0, "ack flag", link.flags.tolist(), flush=True)
except Exception:
    traceback.print_exc()
    print("last cuda error:", _lib.load().ivl_last_cuda_error(), flush=True)
    os._exit(1)
dist.destroy_process_group()
