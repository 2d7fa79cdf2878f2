"""Developer experiment: per-layer device time inside a full bench step (events between layers), to compare with
the per-kernel timings taken in isolation (sustained clocks under the power cap differ from burst clocks)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
torch.cuda.set_device(0)
hp = bench.HotPath(131072, 0, 1, torch.device("cuda", 0))
for _ in range(4):
    hp.step()
torch.cuda.synchronize()
for rep in range(2):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(37)]
    evs[0].record()
    for layer in range(36):
        if layer % 4 == 0:
            hp.swa_layer()
        else:
            hp.gdn_fwd(hp.h0)
        evs[layer + 1].record()
    torch.cuda.synchronize()
    d = [evs[i].elapsed_time(evs[i + 1]) for i in range(36)]
    swa = [d[i] for i in range(36) if i % 4 == 0]; gdn = [d[i] for i in range(36) if i % 4 != 0]
    print(f"step {sum(d):.1f} ms | SWA layers: mean {sum(swa)/9:.3f} min {min(swa):.3f} max {max(swa):.3f} | "
          f"GDN layers: mean {sum(gdn)/27:.3f} min {min(gdn):.3f} max {max(gdn):.3f}", flush=True)
    print("   first 8 layers:", " ".join(f"{x:.2f}" for x in d[:8]), flush=True)
