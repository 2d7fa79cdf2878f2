import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
torch.cuda.set_device(0)
hp = bench.HotPath(131072, 0, 1, torch.device("cuda", 0))
hp.gdn_fwd(hp.h0); torch.cuda.synchronize()
ref = None
for rnd in range(3):
    for cl in ("0", "2"):
        os.environ["IVL_GDN_CLUSTER"] = cl
        hp.o.zero_()
        ts = bench.time_events(lambda: hp.gdn_fwd(hp.h0), 8)
        if ref is None: ref = hp.o.clone()
        print(f"cluster={cl}: {sorted(ts)[len(ts)//2]:.3f} ms identical={torch.equal(ref, hp.o)}", flush=True)
for cl in ("0", "2"):
    os.environ["IVL_GDN_CLUSTER"] = cl; os.environ["IVL_GDN_BV"] = "64"
    ts = bench.time_events(lambda: hp.gdn_scan(hp.h0), 6)
    print(f"scan alone bv=64 cluster={cl}: {sorted(ts)[len(ts)//2]:.3f} ms", flush=True)
