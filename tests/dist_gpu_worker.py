"""torchrun entry of tests/test_dist_gpu.py: one rank per GPU over NCCL; the last rank prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from infinitevl_b200.dist import operator_parity_check, sharded_parity_check  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Ts = [int(x) for x in (sys.argv[1:] or ["32768"])]
    for T in Ts:
        res = sharded_parity_check(T=T, num_layers=8)
        if res:
            print("DIST_PARITY " + json.dumps(res), flush=True)
        for transport in ("nccl", "p2p"):
            op = operator_parity_check(T=T, transport=transport)
            if dist.get_rank() == 0:
                print("OP_PARITY " + json.dumps(op), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
