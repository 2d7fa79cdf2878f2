"""Developer tool: per-chunk timeline of the scan kernel (clock64 stamps of CTA 0, chunks 1000..1063).
Builds a private copy of the library with -DIVL_TRACE; the product build never contains the probe."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
CS = os.path.join(ROOT, "infinitevl_b200", "csrc")
out = "/tmp/libivl_trace.so"
from infinitevl_b200 import build as B
srcs = [os.path.join(CS, f) for f in B.SOURCES]
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-DIVL_BUILDING_DLL", "-DIVL_TRACE", "-shared", "-I", os.path.join(ROOT, "include"), "-o", out] + srcs, check=True)
lib = ctypes.CDLL(out)
from inputs import gdn_inputs
T = 131072
q, k, v, g, beta, h0 = gdn_inputs(T=16384, H=16, seed=0)
rep = T // 16384
tile = lambda x: x.repeat(1, rep, *([1] * (x.dim() - 2))).contiguous().cuda()
q, k, v, g, beta = (tile(x) for x in (q, k, v, g, beta)); h0 = h0.cuda()
lib.ivl_gdn_chunk_workspace_bytes.restype = ctypes.c_size_t
need = lib.ivl_gdn_chunk_workspace_bytes(1, T, 16)
ws = torch.empty(need + 1024, dtype=torch.uint8, device="cuda"); off = (-ws.data_ptr()) % 1024; ws = ws[off:off + need]
o = torch.empty(1, T, 16, 256, dtype=torch.bfloat16, device="cuda"); ht = torch.empty(1, 16, 128, 256, device="cuda")
P = ctypes.c_void_p
lib.ivl_gdn_chunk_prep.argtypes = [P] * 5 + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, P, ctypes.c_size_t, P]
lib.ivl_gdn_chunk_scan.argtypes = [P, P, ctypes.c_int, P, P, ctypes.c_int] + [ctypes.c_int] * 3 + [P, ctypes.c_size_t, P]
os.environ.setdefault("IVL_GDN_TSCAN", "0")   # the row-major scan carries this tool's scan probe
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    assert lib.ivl_gdn_chunk_prep(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(), 1, T, 16, 0.0, 1, ws.data_ptr(), need, st) == 0
    assert lib.ivl_gdn_chunk_scan(v.data_ptr(), h0.data_ptr(), 0, o.data_ptr(), ht.data_ptr(), 0, 1, T, 16, ws.data_ptr(), need, st) == 0
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 16))()
lib.ivl_debug_read_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
assert lib.ivl_debug_read_trace(buf, 64 * 16) == 0
t = np.array(buf[:]).reshape(64, 16).astype(np.int64)
names = ["M:ready", "M:A issued", "M:vn seen", "M:st seen", "M:BC issued", "V:a seen", "V:ld done", "V:vn arrived", "S:s seen", "S:ld done",
         "S:sb arrived", "S:st arrived", "O:o seen"]
base = t[:, 0:1]
rel = t[:, :13] - base
period = np.diff(t[:, 0])
print("chunk period (cycles): median", np.median(period), "min", period.min(), "max", period.max())
med = np.median(rel[4:60], axis=0)
order = np.argsort(med)
for i in order:
    print(f"{names[i]:14s} +{med[i]:8.0f}")
nxt = np.median(t[5:60, 0] - t[4:59, 0])
print("next M:ready at +", nxt)

pb = (ctypes.c_longlong * (16 * 8))()
lib.ivl_debug_read_prep_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
assert lib.ivl_debug_read_prep_trace(pb, 16 * 8) == 0
pt = np.array(pb[:]).reshape(16, 8).astype(np.int64)
d = np.median(np.diff(pt, axis=1), axis=0)
print("prep phases (cycles): loads+cumsum", d[0], "| norm+images", d[1], "| KK/QK+L,P", d[2], "| zero T + diag solve", d[3],
      "| tile products", d[4], "| T->Aw,Au", d[5], "| W,U mma+stores", d[6], "| total", np.median(pt[:, 7] - pt[:, 0]))
