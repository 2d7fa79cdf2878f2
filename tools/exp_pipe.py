"""Developer experiment: parity and timing of the scan variants (IVL_GDN_BV) and of the overlapped chunk
operator (IVL_GDN_PIPE) against the stand-alone kernels."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from inputs import gdn_inputs
from infinitevl_b200 import _lib, ops
from oracle import err_ratio, gdn_chunk_ref

lib = _lib.load()
torch.set_num_threads(os.cpu_count())


def med(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


# ---- parity of every mode at small sizes
for T, H in ((64, 2), (200, 2), (1024, 16), (4096, 4)):
    q, k, v, g, beta, h0 = gdn_inputs(T=T, H=H, seed=3)
    ro, rs = gdn_chunk_ref(q, k, v, g, beta, initial_state=h0)
    dq, dk, dv, dg, db, dh = (x.cuda() for x in (q, k, v, g, beta, h0))
    for pipe, bv in ((0, 32), (0, 64), (0, 128), (1, 128), (1, 64), (1, 32)):
        os.environ["IVL_GDN_PIPE"] = str(pipe); os.environ["IVL_GDN_BV"] = str(bv)
        o, s = ops.chunk_gated_delta_rule(dq, dk, dv, dg, db, initial_state=dh, output_final_state=True,
                                          use_qk_l2norm_in_kernel=True)
        torch.cuda.synchronize()
        print(f"T={T} H={H} pipe={pipe} bv={bv}: err o={err_ratio(ro, o.float().cpu()):.2e} S={err_ratio(rs, s.cpu()):.2e}", flush=True)

# ---- timing at full size
T = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
H = 16
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


def fwd():
    _lib.check(lib.ivl_gdn_chunk_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(),
                                     h0.data_ptr(), 0, o.data_ptr(), ht.data_ptr(), 0, 1, T, H, 128, 256, 0.0, 1,
                                     ws.data_ptr(), ws.numel(), st), "fwd")


print(f"T={T} prep alone: {med(prep):.3f} ms", flush=True)
ref_o = None
for bv in (32, 64, 128):
    os.environ["IVL_GDN_BV"] = str(bv)
    t = med(scan)
    if ref_o is None:
        ref_o, ref_s = o.clone(), ht.clone()
    print(f"T={T} scan alone bv={bv}: {t:.3f} ms = {t*1e6/(T//64):.0f} ns/chunk   vs bv32: o={err_ratio(ref_o.float(), o.float()):.1e} S={err_ratio(ref_s, ht):.1e}", flush=True)
for pipe, bv in ((0, 32), (1, 128), (1, 64), (1, 32)):
    os.environ["IVL_GDN_PIPE"] = str(pipe); os.environ["IVL_GDN_BV"] = str(bv)
    o.zero_(); ht.zero_()
    t = med(fwd)
    print(f"T={T} fwd pipe={pipe} bv={bv}: {t:.3f} ms  ({3.238002688e9 * (T / 131072) / t / 1e6:.0f} GB/s algorithmic)  vs ref: o={err_ratio(ref_o.float(), o.float()):.1e} S={err_ratio(ref_s, ht):.1e}", flush=True)
