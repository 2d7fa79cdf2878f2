"""Developer script: quick parity + timing of the SWA kernel on the GPU box."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from infinitevl_b200 import swa
from oracle import err_ratio, swa_attention_ref

def main():
    for Tq, Tk, W in ((64, 64, None), (128, 128, None), (130, 130, None), (257, 257, 100), (200, 1000, 512), (1, 300, None)):
        gen = torch.Generator().manual_seed(Tq + Tk)
        q = torch.randn(1, 16, Tq, 128, generator=gen).bfloat16(); k = torch.randn(1, 2, Tk, 128, generator=gen).bfloat16()
        v = torch.randn(1, 2, Tk, 128, generator=gen).bfloat16()
        ref = swa_attention_ref(q, k, v, window=W)
        out = swa.swa_attention(q.cuda(), k.cuda(), v.cuda(), window=W)
        torch.cuda.synchronize()
        print(f"swa Tq={Tq} Tk={Tk} W={W}: err={err_ratio(ref, out.float().cpu()):.2e} finite={bool(torch.isfinite(out).all())}", flush=True)
    for T in (32768, 131072):
        gen = torch.Generator().manual_seed(1)
        q = torch.randn(1, T, 16, 128, generator=gen).bfloat16().cuda(); k = torch.randn(1, T, 2, 128, generator=gen).bfloat16().cuda()
        v = torch.randn(1, T, 2, 128, generator=gen).bfloat16().cuda()
        o = torch.empty_like(q)
        fn = lambda: swa.swa_attention_bthd(q, k, v, window=8192, out=o)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = sorted(ts)[2]; W = 8192
        flops = 4 * 16 * 128 * (W * (W + 1) / 2 + (T - W) * W)
        print(f"swa T={T}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
        try:
            from flash_attn import flash_attn_func
            fo = flash_attn_func(q, k, v, causal=True, window_size=(8191, 8191))
            print(f"   vs flash-attn: err={err_ratio(fo.float(), o.float()):.2e}", flush=True)
        except Exception as e:
            print("flash_attn compare failed", repr(e))

if __name__ == "__main__":
    main()
