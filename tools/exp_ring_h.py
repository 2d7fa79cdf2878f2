import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from inputs import gdn_inputs
from infinitevl_b200 import ops
T, H, ring = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
q, k, v, g, beta, h0 = gdn_inputs(T=T, H=H, seed=1, device="cuda")
os.environ["IVL_GDN_PIPE"] = "0"
o0, s0 = ops.chunk_gated_delta_rule(q, k, v, g, beta, initial_state=h0, output_final_state=True, use_qk_l2norm_in_kernel=True)
torch.cuda.synchronize()
os.environ["IVL_GDN_PIPE"] = "1"; os.environ["IVL_GDN_BV"] = "64"; os.environ["IVL_GDN_RING"] = ring
t0 = time.time()
try:
    o1, s1 = ops.chunk_gated_delta_rule(q, k, v, g, beta, initial_state=h0, output_final_state=True, use_qk_l2norm_in_kernel=True)
    torch.cuda.synchronize()
    print(f"T={T} H={H} ring={ring}: ok {1e3*(time.time()-t0):.1f} ms identical={torch.equal(o0,o1) and torch.equal(s0,s1)}", flush=True)
except Exception as e:
    print(f"T={T} H={H} ring={ring}: FAIL after {time.time()-t0:.1f}s", flush=True)
