"""Developer tool: run the reference's GPU operator stack on the B200 box.

The reference's hot path executes inside two pip dependencies
(requirements.txt:18-20): flash-linear-attention (Triton) and flash-attn.
This script (1) confirms the installed builds run on sm_100, (2) times them at
the BASELINE.json sizes so DESIGN.md can quote the GPU reference beside our
kernels, and (3) writes a small fixture of the Triton operator's outputs on
seeded inputs (tests/golden/make_golden.py documents the input recipe) into
gpurun_out/ for committing under tests/golden/.
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def make_gdn_inputs(T, H=16, K=128, V=256, seed=0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(1, T, H, K, generator=g).to(torch.bfloat16)
    k = torch.randn(1, T, H, K, generator=g).to(torch.bfloat16)
    v = torch.randn(1, T, H, V, generator=g).to(torch.bfloat16)
    beta = torch.sigmoid(torch.randn(1, T, H, generator=g)).to(torch.bfloat16)
    a = torch.log(torch.empty(H).uniform_(1e-3, 16, generator=g))
    dt = torch.exp(torch.empty(H).uniform_(math.log(1e-3), math.log(1e-1), generator=g))
    dt_bias = dt + torch.log(-torch.expm1(-dt))
    gate = -torch.exp(a) * torch.nn.functional.softplus(torch.randn(1, T, H, generator=g) + dt_bias)
    h0 = torch.randn(1, H, K, V, generator=g)
    return [x.to(device) for x in (q, k, v, gate.float(), beta, h0)]


def time_fn(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__}
    try:
        import fla
        from fla.ops.gated_delta_rule import chunk_gated_delta_rule, fused_recurrent_gated_delta_rule
        res["fla"] = getattr(fla, "__version__", "?")
        for T in (1024, 32768, 131072):
            q, k, v, g, beta, h0 = make_gdn_inputs(T)
            fn = lambda: chunk_gated_delta_rule(q, k, v, g, beta, initial_state=h0, output_final_state=True,
                                                use_qk_l2norm_in_kernel=True)
            ms = time_fn(fn)
            res[f"fla_chunk_gdn_T{T}_ms"] = ms
            print(f"fla chunk_gated_delta_rule T={T}: {ms:.3f} ms", flush=True)
        # decode step
        q, k, v, g, beta, h0 = make_gdn_inputs(1)
        fn = lambda: fused_recurrent_gated_delta_rule(q=q, k=k, v=v, g=g, beta=beta, initial_state=h0, output_final_state=True,
                                                      use_qk_l2norm_in_kernel=True)
        res["fla_recurrent_T1_ms"] = time_fn(fn, iters=50)
        # fixture
        T, H = 256, 2
        q, k, v, g, beta, h0 = make_gdn_inputs(T, H=H, seed=7)
        o, ht = chunk_gated_delta_rule(q, k, v, g, beta, initial_state=h0, output_final_state=True,
                                       use_qk_l2norm_in_kernel=True)
        o2, ht2 = fused_recurrent_gated_delta_rule(q=q, k=k, v=v, g=g, beta=beta, initial_state=h0, output_final_state=True,
                                                   use_qk_l2norm_in_kernel=True)
        np.savez_compressed(os.path.join(OUT, "fla_triton_gdn_T256_H2_seed7.npz"),
                            o_chunk=o.float().cpu().numpy().astype(np.float16),
                            ht_chunk=ht.float().cpu().numpy(),
                            o_recurrent=o2.float().cpu().numpy().astype(np.float16),
                            ht_recurrent=ht2.float().cpu().numpy())
        res["fixture"] = "fla_triton_gdn_T256_H2_seed7.npz"
    except Exception as e:  # noqa: BLE001
        res["fla_error"] = repr(e)
        print("fla failed:", repr(e), flush=True)
    try:
        from flash_attn import flash_attn_func
        for T in (32768, 131072):
            g = torch.Generator(device="cpu").manual_seed(1)
            q = torch.randn(1, T, 16, 128, generator=g).to(torch.bfloat16).cuda()
            k = torch.randn(1, T, 2, 128, generator=g).to(torch.bfloat16).cuda()
            v = torch.randn(1, T, 2, 128, generator=g).to(torch.bfloat16).cuda()
            fn = lambda: flash_attn_func(q, k, v, causal=True, window_size=(8191, 8191))
            ms = time_fn(fn, iters=5)
            W = 8192
            flops = 4 * 16 * 128 * (W * (W + 1) / 2 + (T - W) * W)
            res[f"fa2_swa_T{T}_ms"] = ms
            res[f"fa2_swa_T{T}_tflops"] = flops / ms / 1e9
            print(f"flash_attn SWA T={T}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    except Exception as e:  # noqa: BLE001
        res["fa2_error"] = repr(e)
        print("flash_attn failed:", repr(e), flush=True)
    with open(os.path.join(OUT, "ref_gpu_probe.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
