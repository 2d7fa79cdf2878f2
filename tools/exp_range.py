"""Developer tool: one overlapped GDN operator call inside a cudaProfilerStart/Stop range, for
`ncu --replay-mode range` (range replay keeps the two kernels concurrent; kernel replay cannot)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
T = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
torch.cuda.set_device(0)
hp = bench.HotPath(T, 0, 1, torch.device("cuda", 0))
for _ in range(3):
    hp.gdn_fwd(hp.h0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
hp.gdn_fwd(hp.h0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
