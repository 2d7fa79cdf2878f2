"""Developer tool: per-chunk timeline of the transposed scan kernel (clock64 stamps of CTA (0,0), chunks 1000..1063).
Builds a private copy of the library with -DIVL_TRACE; the product build never contains the probe."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from infinitevl_b200 import build as B
CS = os.path.join(ROOT, "infinitevl_b200", "csrc")
out = "/tmp/libivl_trace.so"
srcs = [os.path.join(CS, f) for f in B.SOURCES]
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-DIVL_BUILDING_DLL", "-DIVL_TRACE", *os.environ.get("IVL_NVCC_EXTRA", "").split(), "-shared", "-I", os.path.join(ROOT, "include"), "-o", out] + srcs, check=True)
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
st = torch.cuda.current_stream().cuda_stream
MODE = sys.argv[1] if len(sys.argv) > 1 else "2"
os.environ["IVL_GDN_TSCAN"] = MODE
for _ in range(3):
    assert lib.ivl_gdn_chunk_prep(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(), 1, T, 16, 0.0, 1, ws.data_ptr(), need, st) == 0
    assert lib.ivl_gdn_chunk_scan(v.data_ptr(), h0.data_ptr(), 0, o.data_ptr(), ht.data_ptr(), 0, 1, T, 16, ws.data_ptr(), need, st) == 0
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(5):
    assert lib.ivl_gdn_chunk_scan(v.data_ptr(), h0.data_ptr(), 0, o.data_ptr(), ht.data_ptr(), 0, 1, T, 16, ws.data_ptr(), need, st) == 0
ev[1].record(); torch.cuda.synchronize()
print(f"scan alone (traced build, {os.environ.get('IVL_NVCC_EXTRA', '')}): {ev[0].elapsed_time(ev[1]) / 5:.3f} ms")
buf = (ctypes.c_longlong * (64 * 16))()
lib.ivl_debug_read_ttrace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
assert lib.ivl_debug_read_ttrace(buf, 64 * 16) == 0
t = np.array(buf[:]).reshape(64, 16).astype(np.int64)
names2 = ["M1:vb(k) seen", "M1:B+RC issued", "M1:B issued", "M2:sb(k) seen", "M2:XO(k) issued", "M2:U(k+2) issued", "V:v_new(k) acc seen",
          "V:vb(k) arrived", "S:ds(k) seen", "S:ld+fma done", "S:xo(k) seen", "S:sb(k+1) arrived", "S:loop top", "S:fullS seen", "M1:dsfree seen"]
names = ["M:sb seen", "M:W+O issued", "M:vb seen", "M:B+C issued", "M:U(c+1) issued", "V:dv seen", "V:ld done", "V:vb arrived",
         "V:output done", "S:ds seen", "S:ld+fma done", "S:sb arrived"]
names3 = ["M:sb[0] seen", "M:W issued", "M:vb[0] seen", "M:B issued", "M:O+C issued", "M:U(c+1) issued", "V:dv seen", "V:vb[0] arrived",
          "V:vb[1] arrived", "S:ds seen", "S:ld+fma done", "S:sb arrived"]
if MODE == "2":
    names = names2
if MODE == "3":
    names = names3
period = np.diff(t[:, 0])
print("chunk period (cycles): median", np.median(period), "min", period.min(), "max", period.max())
rel = t[:, :len(names)] - t[:, 0:1]
med = np.median(rel[4:60], axis=0)
for i in np.argsort(med):
    print(f"{names[i]:18s} +{med[i]:8.0f}")
