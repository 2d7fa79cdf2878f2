"""Developer script: quick parity + timing of the GDN kernels on the GPU box."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from inputs import gdn_inputs  # noqa: E402

from infinitevl_b200 import _lib, ops  # noqa: E402
from oracle import err_ratio, gdn_chunk_ref, gdn_recurrent_ref  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    for T, H in ((64, 2), (256, 2), (200, 2), (1024, 16)):
        q, k, v, g, beta, h0 = gdn_inputs(T=T, H=H, seed=3)
        ro, rs = gdn_chunk_ref(q, k, v, g, beta, initial_state=h0)
        dq, dk, dv, dg, db, dh = (x.cuda() for x in (q, k, v, g, beta, h0))
        o, s = ops.chunk_gated_delta_rule(dq, dk, dv, dg, db, initial_state=dh, output_final_state=True,
                                          use_qk_l2norm_in_kernel=True)
        torch.cuda.synchronize()
        print(f"chunk  T={T} H={H}: err o={err_ratio(ro, o.float().cpu()):.2e} S={err_ratio(rs, s.cpu()):.2e}", flush=True)
        if T <= 256:
            o2, s2 = ops.fused_recurrent_gated_delta_rule(dq, dk, dv, dg, db, initial_state=dh,
                                                          output_final_state=True, use_qk_l2norm_in_kernel=True)
            torch.cuda.synchronize()
            rr, rrs = gdn_recurrent_ref(q, k, v, g, beta, initial_state=h0)
            print(f"recur  T={T} H={H}: err o={err_ratio(rr, o2.float().cpu()):.2e} S={err_ratio(rrs, s2.cpu()):.2e}", flush=True)
    lib = _lib.load()
    for T in (32768, 131072):
        q, k, v, g, beta, h0 = gdn_inputs(T=T, H=16, seed=0, device="cuda")
        o = torch.empty(1, T, 16, 256, dtype=torch.bfloat16, device="cuda")
        ht = torch.empty(1, 16, 128, 256, dtype=torch.float32, device="cuda")
        ws = ops.gdn_workspace(1, T, 16, "cuda")
        st = torch.cuda.current_stream().cuda_stream

        def prep():
            _lib.check(lib.ivl_gdn_chunk_prep(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(),
                                              1, T, 16, 0.0, 1, ws.data_ptr(), ws.numel(), st), "prep")

        def scan():
            _lib.check(lib.ivl_gdn_chunk_scan(v.data_ptr(), h0.data_ptr(), 0, o.data_ptr(), ht.data_ptr(), 0, 1, T, 16,
                                              ws.data_ptr(), ws.numel(), st), "scan")

        for name, fn in (("prep", prep), ("scan", scan)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            print(f"T={T} {name}: {sorted(ts)[len(ts)//2]:.3f} ms", flush=True)
        try:
            from fla.ops.gated_delta_rule import chunk_gated_delta_rule as fla_chunk
            fo, fs = fla_chunk(q, k, v, g, beta, initial_state=h0, output_final_state=True, use_qk_l2norm_in_kernel=True)
            prep(); scan(); torch.cuda.synchronize()
            print(f"T={T} vs fla triton: o={err_ratio(fo.float(), o.float()):.2e} S={err_ratio(fs, ht):.2e}", flush=True)
        except Exception as e:  # noqa: BLE001
            print("fla compare failed", repr(e))


if __name__ == "__main__":
    main()
