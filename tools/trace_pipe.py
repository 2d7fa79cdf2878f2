"""Developer tool: wait statistics of the overlapped GDN operator (trace build: -DIVL_TRACE).
How long prep CTAs wait for ring slots, how long the scan's copy warps wait for ready flags, the scan's chunk period."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
CS = os.path.join(ROOT, "infinitevl_b200", "csrc")
out = "/tmp/libivl_trace.so"
srcs = [os.path.join(CS, f) for f in ("ivl_abi.cu", "gdn_prep.cu", "gdn_scan.cu", "gdn_recurrent.cu", "gdn_decode.cu", "gdn_fused.cu", "swa_fwd.cu", "swa_misc.cu")]
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-DIVL_BUILDING_DLL", "-DIVL_TRACE", "-shared", "-I", os.path.join(ROOT, "include"), "-o", out] + srcs, check=True)
lib = ctypes.CDLL(out)
from inputs import gdn_inputs
T, H = 131072, 16
q, k, v, g, beta, h0 = gdn_inputs(T=16384, H=H, seed=0)
rep = T // 16384
tile = lambda x: x.repeat(1, rep, *([1] * (x.dim() - 2))).contiguous().cuda()
q, k, v, g, beta = (tile(x) for x in (q, k, v, g, beta)); h0 = h0.cuda()
lib.ivl_gdn_chunk_workspace_bytes.restype = ctypes.c_size_t
need = lib.ivl_gdn_chunk_workspace_bytes(1, T, H)
ws = torch.empty(need + 1024, dtype=torch.uint8, device="cuda"); off = (-ws.data_ptr()) % 1024; ws = ws[off:off + need]
o = torch.empty(1, T, H, 256, dtype=torch.bfloat16, device="cuda"); ht = torch.empty(1, H, 128, 256, device="cuda")
P = ctypes.c_void_p
lib.ivl_gdn_chunk_fwd.argtypes = [P] * 6 + [ctypes.c_int, P, P, ctypes.c_int] + [ctypes.c_int] * 5 + [ctypes.c_float, ctypes.c_int, P, ctypes.c_size_t, P]
st = torch.cuda.current_stream().cuda_stream
def fwd():
    assert lib.ivl_gdn_chunk_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(), h0.data_ptr(), 0,
                                 o.data_ptr(), ht.data_ptr(), 0, 1, T, H, 128, 256, 0.0, 1, ws.data_ptr(), need, st) == 0
fwd(); torch.cuda.synchronize()
U4 = ctypes.c_ulonglong * 4
lib.ivl_debug_read_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
os.environ["IVL_GDN_PIPE"] = "1"
for ring in [int(x) for x in (sys.argv[1:] or ["4096", "128", "32"])]:
    os.environ["IVL_GDN_RING"] = str(ring)
    fwd(); torch.cuda.synchronize()
    pw, sw = U4(), U4()
    lib.ivl_debug_read_prep_wait(pw, 1); lib.ivl_debug_read_scan_wait(sw, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fwd(); b.record(); torch.cuda.synchronize()
    lib.ivl_debug_read_prep_wait(pw, 1); lib.ivl_debug_read_scan_wait(sw, 1)
    buf = (ctypes.c_longlong * (64 * 16))()
    lib.ivl_debug_read_trace(buf, 64 * 16)
    t = np.array(buf[:]).reshape(64, 16).astype(np.int64)
    period = np.median(np.diff(t[:, 0]))
    TL = ctypes.c_ulonglong * (16 * 2048 * 4)
    ptl, stl = TL(), TL()
    lib.ivl_debug_read_prep_tl(ptl); lib.ivl_debug_read_scan_tl(stl)
    pt = np.array(ptl[:], dtype=np.float64).reshape(16, 2048, 4) / 1e3
    sl = np.array(stl[:], dtype=np.float64).reshape(16, 2048, 4) / 1e3
    t0 = pt[:, 0, 0].min()
    pt -= t0; sl -= t0
    print(f"--- ring={ring}: per 128-chunk block (us): prep start of head 0 | prep wait (released - start), min/median/max over heads | "
          f"scan lag behind publication (issue - published), min/median/max over heads | slowest and fastest scan head")
    for c in range(0, 2048, 128):
        wait = np.median(pt[:, c:c + 128, 1] - pt[:, c:c + 128, 0], axis=1)
        lag = np.median(sl[:, c:c + 128, :].max(axis=2) - pt[:, c:c + 128, 2], axis=1)
        issue = sl[:, c, 0]
        print(f"  c={c:5d} | {pt[0, c, 0]:8.1f} | {wait.min():6.1f} {np.median(wait):6.1f} {wait.max():6.1f} | {lag.min():7.1f} {np.median(lag):7.1f} {lag.max():7.1f} | "
              f"issue time of chunk c: min {issue.min():8.1f} (h{issue.argmin()}) max {issue.max():8.1f} (h{issue.argmax()})")
    print(f"ring={ring}: {a.elapsed_time(b):.3f} ms | prep: {pw[1]} CTAs waited, avg {pw[0] / max(pw[1], 1):.0f} cycles, {pw[2] / max(pw[1], 1):.1f} polls | "
          f"scan copy warps: {sw[1]} waits, avg {sw[0] / max(sw[1], 1):.0f} cycles, total {sw[0] / 64 / 1e6:.2f} Mcycles per CTA | scan chunk period {period:.0f} cycles", flush=True)
