"""Developer tool: parity (vs the fp32 oracle) and timing of the transposed scan (IVL_GDN_TSCAN=1) against the
row-major scan (IVL_GDN_TSCAN=0), stand-alone kernels and the whole chunk operator."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from inputs import gdn_inputs
from infinitevl_b200 import _lib, ops
from oracle import err_ratio, gdn_chunk_ref

lib = _lib.load()
torch.set_num_threads(os.cpu_count())


def med(fn, n=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for T, H, st in ((64, 2, "f32"), (65, 2, "none"), (200, 3, "bf16"), (1024, 16, "f32"), (4160, 4, "f32")):
    q, k, v, g, beta, h0 = gdn_inputs(T=T, H=H, seed=3)
    h0 = None if st == "none" else (h0.bfloat16() if st == "bf16" else h0)
    ro, rs = gdn_chunk_ref(q, k, v, g, beta, initial_state=h0)
    dq, dk, dv, dg, db = (x.cuda() for x in (q, k, v, g, beta)); dh = None if h0 is None else h0.cuda()
    for ts, pipe in (("3", 0), ("3", 1), ("1", 0), ("0", 0)):
        os.environ["IVL_GDN_TSCAN"] = ts; os.environ["IVL_GDN_PIPE"] = str(pipe)
        o, s = ops.chunk_gated_delta_rule(dq, dk, dv, dg, db, initial_state=dh, output_final_state=True,
                                          use_qk_l2norm_in_kernel=True)
        torch.cuda.synchronize()
        print(f"T={T} H={H} h0={st} tscan={ts} pipe={pipe}: err o={err_ratio(ro, o.float().cpu()):.2e} S={err_ratio(rs, s.cpu()):.2e}", flush=True)

T = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
H = 16
q, k, v, g, beta, h0 = gdn_inputs(T=16384, H=H, seed=0)
rep = T // 16384
tile = lambda x: x.repeat(1, rep, *([1] * (x.dim() - 2))).contiguous().cuda()
q, k, v, g, beta = (tile(x) for x in (q, k, v, g, beta)); h0 = h0.cuda()
o = torch.empty(1, T, H, 256, dtype=torch.bfloat16, device="cuda")
ht = torch.empty(1, H, 128, 256, dtype=torch.float32, device="cuda")
ws = ops.gdn_workspace(1, T, H, "cuda")
st = torch.cuda.current_stream().cuda_stream
prep = lambda: _lib.check(lib.ivl_gdn_chunk_prep(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(), 1, T, H, 0.0, 1, ws.data_ptr(), ws.numel(), st), "prep")
scan = lambda: _lib.check(lib.ivl_gdn_chunk_scan(v.data_ptr(), h0.data_ptr(), 0, o.data_ptr(), ht.data_ptr(), 0, 1, T, H, ws.data_ptr(), ws.numel(), st), "scan")
fwd = lambda: _lib.check(lib.ivl_gdn_chunk_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(), h0.data_ptr(), 0, o.data_ptr(), ht.data_ptr(), 0, 1, T, H, 128, 256, 0.0, 1, ws.data_ptr(), ws.numel(), st), "fwd")
res = {}
for ts in ("3", "1", "0"):
    os.environ["IVL_GDN_TSCAN"] = ts
    os.environ["IVL_GDN_PIPE"] = "0"
    tp = med(prep); prep(); tsn = med(scan)
    res[ts] = (o.clone(), ht.clone())
    os.environ["IVL_GDN_PIPE"] = "1"
    tf = med(fwd)
    same = torch.equal(o, res[ts][0]) and torch.equal(ht, res[ts][1])
    print(f"T={T} tscan={ts}: prep {tp:.3f} ms, scan {tsn:.3f} ms = {tsn * 1e6 / (T // 64):.0f} ns/chunk, overlapped operator {tf:.3f} ms "
          f"({3.238002688e9 * (T / 131072) / tf / 1e6:.0f} GB/s algorithmic), overlapped == back-to-back: {same}", flush=True)
print(f"pipelined vs row-major: o {err_ratio(res['0'][0].float(), res['3'][0].float()):.2e} S {err_ratio(res['0'][1], res['3'][1]):.2e}; "
      f"pipelined == form 1: {torch.equal(res['1'][0], res['3'][0]) and torch.equal(res['1'][1], res['3'][1])}")
