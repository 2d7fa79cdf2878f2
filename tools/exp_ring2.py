"""Developer tool: the overlapped GDN operator with the L2 image ring (IVL_GDN_RING chunks per head): time and
bit-identity against the ring-less form at T = 131072."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
T = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
torch.cuda.set_device(0)
hp = bench.HotPath(T, 0, 1, torch.device("cuda", 0))
def med(fn, n=9, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
ref = None
for ring in [0] + [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "16,24,32,48,64,96".split(","))]:
    os.environ["IVL_GDN_RING"] = str(ring)
    t = med(lambda: hp.gdn_fwd(hp.h0))
    o, s = hp.o.clone(), hp.ht.clone()
    if ref is None:
        ref = (o, s)
    same = torch.equal(o, ref[0]) and torch.equal(s, ref[1])
    print(f"ring {ring:4d}: {t:.3f} ms per operator call, bit-identical to ring 0: {same}", flush=True)
