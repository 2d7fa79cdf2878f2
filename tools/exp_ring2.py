import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
T = 131072
torch.cuda.set_device(0)
hp = bench.HotPath(T, 0, 1, torch.device("cuda", 0))
hp.gdn_fwd(hp.h0); torch.cuda.synchronize()
os.environ["IVL_GDN_PIPE"] = "1"
for nowait in (0, 1):
    for ring in (32, 128, 512, 2047, 4096):
        os.environ["IVL_GDN_RING"] = str(ring); os.environ["IVL_GDN_NOWAIT"] = str(nowait)
        ts = bench.time_events(lambda: hp.gdn_fwd(hp.h0), 6)
        print(f"nowait={nowait} ring={ring:5d}: {sorted(ts)[len(ts)//2]:.3f} ms", flush=True)
