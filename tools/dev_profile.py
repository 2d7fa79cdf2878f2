"""Developer tool: launch the hot-path kernels a few times (target of `ncu -k regex:...`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(0)
hp = bench.HotPath(T, 0, 1, torch.device("cuda", 0))
for _ in range(iters):
    hp.gdn_prep()
    hp.gdn_scan(hp.h0)
    if hp.has_swa:
        hp.swa_layer()
torch.cuda.synchronize()
print("done")
