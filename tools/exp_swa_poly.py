"""Developer experiment: SWA forward time and error as a function of IVL_SWA_POLY (pairs of 8 on the FMA pipe)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from infinitevl_b200 import swa
from oracle import err_ratio, swa_attention_ref
T = 131072
gen = torch.Generator().manual_seed(1)
q = torch.randn(1, T, 16, 128, generator=gen).bfloat16().cuda(); k = torch.randn(1, T, 2, 128, generator=gen).bfloat16().cuda()
v = torch.randn(1, T, 2, 128, generator=gen).bfloat16().cuda()
o = torch.empty_like(q)
qs, ks, vs = (torch.randn(1, h, 600, 128, generator=gen).bfloat16() for h in (16, 2, 2))
ref_small = swa_attention_ref(qs, ks, vs, window=200)
import collections
res = collections.defaultdict(list)
fn = lambda: swa.swa_attention_bthd(q, k, v, window=8192, out=o)
for _ in range(10): fn()      # reach the sustained (power-capped) clocks first
torch.cuda.synchronize()
for rnd in range(6):
    for poly in (0, 2, 3, 4):
        os.environ["IVL_SWA_POLY"] = str(poly)
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); fn(); b.record(); torch.cuda.synchronize()
        res[poly].append(a.elapsed_time(b) / 2)
W = 8192
flops = 4 * 16 * 128 * (W * (W + 1) / 2 + (T - W) * W)
for poly, ts in res.items():
    ms = sorted(ts)[len(ts) // 2]
    print(f"poly={poly}/8: median {ms:.3f} ms ({flops / ms / 1e9:.1f} TFLOP/s)  all: " + " ".join(f"{t:.2f}" for t in ts), flush=True)
