timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_gdn_gpu.py tests/test_stream_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -6
python - <<'P'
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from infinitevl_b200 import modeling as M
cfg = M.HybridTextConfig(num_hidden_layers=4)
mod = M.GatedDeltaNet(cfg, 1).bfloat16().cuda()
x = torch.randn(1, 131072, 2048, device="cuda").bfloat16() * 0.5
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for fused in ("1", "0"):
    os.environ["IVL_GDN_FUSED_PREFILL"] = fused
    with torch.no_grad():
        ms = t(lambda: mod(x))
    print(f"GatedDeltaNet.forward at T=131072 (projections included), fused prefill={fused}: {ms:.3f} ms", flush=True)
P
