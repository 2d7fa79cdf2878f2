"""Developer experiment: overlapped GDN operator time as a function of the image ring length."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
T = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
torch.cuda.set_device(0)
hp = bench.HotPath(T, 0, 1, torch.device("cuda", 0))
hp.gdn_fwd(hp.h0); torch.cuda.synchronize()
os.environ["IVL_GDN_PIPE"] = "0"
hp.gdn_fwd(hp.h0); torch.cuda.synchronize()
ref_o, ref_s = hp.o.clone(), hp.ht.clone()
os.environ["IVL_GDN_PIPE"] = "1"
for bv in (64, 32):
    for ring in (8, 12, 16, 24, 32, 48, 64, 128, 4096):
        os.environ["IVL_GDN_RING"] = str(ring); os.environ["IVL_GDN_BV"] = str(bv)
        hp.o.zero_(); hp.ht.zero_()
        ts = bench.time_events(lambda: hp.gdn_fwd(hp.h0), 8)
        ok = torch.equal(hp.o, ref_o) and torch.equal(hp.ht, ref_s)
        print(f"bv={bv} ring={ring:5d}: {sorted(ts)[len(ts)//2]:.3f} ms  identical={ok}", flush=True)
        if bv == 32 and ring >= 16:
            break
