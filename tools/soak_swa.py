"""Developer tool: randomised soak of the SWA kernels (prefill tcgen05 kernel and split-KV decode) against
flash-attn (the reference's own GPU path) on random shapes: batch, GQA group, Tq <= Tk, window or none."""
import os, random, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from infinitevl_b200 import swa
from oracle import err_ratio
from flash_attn import flash_attn_func

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = random.Random(4321)
t_end = time.time() + budget
trials = bad = 0
worst = 0.0
while time.time() < t_end:
    trials += 1
    B = rng.choice([1, 1, 2, 3])
    Hkv = rng.choice([1, 2, 4])
    Hq = Hkv * rng.choice([1, 2, 4, 8])
    Tk = rng.choice([rng.randint(1, 200), rng.randint(200, 3000), rng.randint(3000, 20000)])
    Tq = rng.choice([1, 1, Tk, Tk, rng.randint(1, Tk)])
    window = rng.choice([None, None, rng.randint(1, 300), rng.randint(300, 9000)])
    g = torch.Generator(device="cuda").manual_seed(trials)
    q = torch.randn(B, Tq, Hq, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Tk, Hkv, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Tk, Hkv, 128, device="cuda", generator=g).bfloat16()
    out = swa.swa_attention_bthd(q, k, v, window=window)
    use_w = window is not None and Tk > window            # the HF glue passes the window only when key_len > W
    ref = flash_attn_func(q, k, v, causal=True, window_size=((window - 1, window - 1) if use_w else (-1, -1)))
    torch.cuda.synchronize()
    e = err_ratio(ref.float(), out.float())
    worst = max(worst, e)
    if not (e < 1e-2) or not bool(torch.isfinite(out).all()):
        bad += 1
        print(f"MISMATCH trial {trials}: B={B} Hq={Hq} Hkv={Hkv} Tq={Tq} Tk={Tk} window={window} err={e:.3e}", flush=True)
print(f"swa soak: {trials} trials, {bad} mismatches, worst err-ratio {worst:.2e}", flush=True)
sys.exit(1 if bad else 0)
