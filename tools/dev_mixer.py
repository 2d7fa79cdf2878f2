import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from infinitevl_b200 import modeling as M, ops
from oracle import err_ratio, short_conv_ref, gdn_chunk_ref, rmsnorm_gated_ref
from oracle.gdn import gdn_gate_ref
from test_modules_gpu import _init, _params, gen
cfg = M.HybridTextConfig(num_hidden_layers=4)
mod = _init(M.GatedDeltaNet(cfg, 1), 11).bfloat16().cuda()
p = _params(mod)
x = torch.randn(1, 700, 2048, generator=gen(12)).bfloat16()
xf = x.float()
lin = lambda n: xf @ p[n + ".weight"].t()
xc = x.cuda()
for n in ("q", "k", "v"):
    ry, _ = short_conv_ref(lin(n + "_proj"), p[n + "_conv1d.weight"])
    proj = getattr(mod, n + "_proj")(xc)
    print(n, "proj err", err_ratio(lin(n + "_proj"), proj.float().cpu()))
    y, _ = getattr(mod, n + "_conv1d")(proj)
    print(n, "conv err", err_ratio(ry, y.float().cpu()))
rg, rb = gdn_gate_ref(lin("a_proj"), lin("b_proj"), p["A_log"], p["dt_bias"])
g, beta = M.gdn_gates(mod.a_proj(xc), mod.b_proj(xc), mod.A_log, mod.dt_bias)
print("g err", err_ratio(rg, g.cpu()), "beta err", err_ratio(rb, beta.float().cpu()), g.min().item(), rg.min().item())
q, _ = short_conv_ref(lin("q_proj"), p["q_conv1d.weight"]); k, _ = short_conv_ref(lin("k_proj"), p["k_conv1d.weight"]); v, _ = short_conv_ref(lin("v_proj"), p["v_conv1d.weight"])
q, k, v = q.view(1, 700, 16, 128), k.view(1, 700, 16, 128), v.view(1, 700, 16, 256)
ro, rs = gdn_chunk_ref(q, k, v, rg, rb)
o, s = ops.chunk_gated_delta_rule(q.bfloat16().cuda(), k.bfloat16().cuda(), v.bfloat16().cuda(), rg.cuda(), rb.bfloat16().cuda(), output_final_state=True, use_qk_l2norm_in_kernel=True)
print("chunk err", err_ratio(ro, o.float().cpu()), err_ratio(rs, s.cpu()), "|o| rms", ro.pow(2).mean().sqrt().item(), "|v|", v.pow(2).mean().sqrt().item(), "|q|", q.pow(2).mean().sqrt().item())
gate = lin("g_proj").view(1, 700, 16, 256)
rn = rmsnorm_gated_ref(ro, gate, p["o_norm.weight"], 1e-5)
n = M.rmsnorm_gated(ro.bfloat16().cuda(), gate.bfloat16().cuda(), mod.o_norm.weight, 1e-5)
print("norm err", err_ratio(rn, n.float().cpu()))
out, _ = mod(xc)
from oracle import gdn_mixer_ref
ref, _, _ = gdn_mixer_ref(x, p)
print("mixer err", err_ratio(ref, out.float().cpu()))
