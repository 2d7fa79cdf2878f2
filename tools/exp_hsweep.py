"""Developer experiment: scan / prep time per chunk as a function of the number of heads (= number of scan
CTAs and L2->SM operand traffic).  If cycles per chunk drop with fewer heads the scan is L2-fabric limited."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from inputs import gdn_inputs
from infinitevl_b200 import _lib, ops

lib = _lib.load()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for H in (1, 2, 4, 8, 12, 16):
    q, k, v, g, beta, h0 = gdn_inputs(T=T, H=H, seed=0, device="cuda")
    o = torch.empty(1, T, H, 256, dtype=torch.bfloat16, device="cuda")
    ht = torch.empty(1, H, 128, 256, dtype=torch.float32, device="cuda")
    ws = ops.gdn_workspace(1, T, H, "cuda")
    st = torch.cuda.current_stream().cuda_stream
    def prep():
        _lib.check(lib.ivl_gdn_chunk_prep(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(),
                                          1, T, H, 0.0, 1, ws.data_ptr(), ws.numel(), st), "prep")
    def scan():
        _lib.check(lib.ivl_gdn_chunk_scan(v.data_ptr(), h0.data_ptr(), 0, o.data_ptr(), ht.data_ptr(), 0, 1, T, H,
                                          ws.data_ptr(), ws.numel(), st), "scan")
    res = {}
    for name, fn in (("prep", prep), ("scan", scan)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[name] = sorted(ts)[len(ts) // 2]
    nt = T // 64
    print(f"H={H:2d} T={T}: prep {res['prep']:.3f} ms  scan {res['scan']:.3f} ms = {res['scan']*1e6/nt:.0f} ns/chunk", flush=True)
