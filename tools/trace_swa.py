"""Developer tool: per-tile timeline of the SWA kernel (clock64 stamps of one CTA, tiles 40..71)."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "infinitevl_b200", "csrc")
out = "/tmp/libivl_trace.so"
srcs = [os.path.join(CS, f) for f in ("ivl_abi.cu", "gdn_prep.cu", "gdn_scan.cu", "gdn_recurrent.cu", "gdn_decode.cu", "gdn_fused.cu", "swa_fwd.cu", "swa_misc.cu")]
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-DIVL_BUILDING_DLL", "-DIVL_TRACE", "-shared", "-I", os.path.join(ROOT, "include"), "-o", out] + srcs, check=True)
lib = ctypes.CDLL(out)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
gen = torch.Generator().manual_seed(1)
q = torch.randn(1, T, 16, 128, generator=gen).bfloat16().cuda(); k = torch.randn(1, T, 2, 128, generator=gen).bfloat16().cuda()
v = torch.randn(1, T, 2, 128, generator=gen).bfloat16().cuda(); o = torch.empty_like(q)
P = ctypes.c_void_p
lib.ivl_swa_fwd.argtypes = [P] * 8 + [ctypes.c_int] * 7 + [ctypes.c_float, P]
st3 = lambda t: (ctypes.c_int64 * 3)(t.stride(0), t.stride(1), t.stride(2))
for _ in range(3):
    assert lib.ivl_swa_fwd(q.data_ptr(), st3(q), k.data_ptr(), st3(k), v.data_ptr(), st3(v), o.data_ptr(), st3(o), 1, T, T, 16, 2, 128, 8192, 0.0,
                           torch.cuda.current_stream().cuda_stream) == 0
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (32 * 16))()
lib.ivl_debug_read_swa_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
assert lib.ivl_debug_read_swa_trace(buf, 32 * 16) == 0
t = np.array(buf[:]).reshape(32, 16).astype(np.int64)
names = ["M:S(t+1) go", "M:S(t+1) issued", "M:p(t) seen", "M:PV(t) issued", "X:loop top", "X:s(t) seen", "X:ld done", "X:exps done", "X:pv(t-1) seen", "X:p(t) arrived"]
base = t[:, 4:5]
rel = t[:, :10] - base
print("tile period (cycles): median", np.median(np.diff(t[:, 4])))
med = np.median(rel[2:30], axis=0)
for i in np.argsort(med):
    print(f"{names[i]:18s} +{med[i]:8.0f}")
