"""Developer tool (CPU): how much precision does the segment-map pass of a sequence-parallel GDN scan need?
Compares the start state of the last segment (exact: fp64 serial scan) with the composed one when the per-chunk maps
are applied with bf16-rounded operands, in two forms:
  full : S <- bf16(A_c) S_bf16 + B_c               with A_c = gamma_c I - Kt_c^T Wg_c rounded as a whole
  split: S <- gamma_c S + (-bf16(N_c)) S_bf16 + B_c   with N_c = Kt_c^T Wg_c (the identity part stays exact fp32)
(the state operand is rounded to bf16 in both, as in today's scan kernel)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from inputs import gdn_inputs
from oracle import err_ratio, l2norm_ref

T, H, C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 2, 64
q, k, v, g, beta, h0 = gdn_inputs(T=T, H=H, seed=5)
dt = torch.float64
kf = l2norm_ref(k.to(dt), dtype=dt); vf, gf, bf = v.to(dt), g.to(dt), beta.to(dt)
NT = T // C
kc, vc = (x.permute(0, 2, 1, 3).reshape(1, H, NT, C, -1) for x in (kf, vf))
gc = gf.permute(0, 2, 1).reshape(1, H, NT, C); bc = bf.permute(0, 2, 1).reshape(1, H, NT, C)
G = gc.cumsum(-1)
Gam = (G[..., :, None] - G[..., None, :]).tril().exp().tril()
eye = torch.eye(C, dtype=dt)
L = ((kc * bc[..., None]) @ kc.transpose(-1, -2) * Gam).tril(-1)
A = torch.linalg.solve_triangular(eye + L, eye.expand_as(L).contiguous(), upper=False, unitriangular=True)
Wg = A @ (kc * (bc * G.exp())[..., None]); U = A @ (vc * bc[..., None])
Kt = kc * (G[..., -1:] - G).exp()[..., None]; gamma = G[..., -1].exp()
N = Kt.transpose(-1, -2) @ Wg
Bc = Kt.transpose(-1, -2) @ U
eyeK = torch.eye(128, dtype=dt)
r = lambda x: x.to(torch.bfloat16).to(dt)
S_exact = h0.to(dt).clone(); S_full = S_exact.clone(); S_split = S_exact.clone(); S_scan = S_exact.clone()
for c in range(NT):
    gm = gamma[:, :, c][..., None, None]
    S_exact = gm * S_exact - N[:, :, c] @ S_exact + Bc[:, :, c]
    S_full = r(gm * eyeK - N[:, :, c]) @ r(S_full) + Bc[:, :, c]
    S_split = gm * S_split - r(N[:, :, c]) @ r(S_split) + Bc[:, :, c]
    Vn = r(U[:, :, c] - r(Wg[:, :, c]) @ r(S_scan))                       # today's kernel: bf16 Wg, S shadow, v_new
    S_scan = gm * S_scan + r(Kt[:, :, c]).transpose(-1, -2) @ Vn
print(f"T={T}: state error after {NT} chunks vs fp64 --  today's scan (bf16 operands) {err_ratio(S_exact, S_scan):.2e} | "
      f"composed, A_c rounded whole {err_ratio(S_exact, S_full):.2e} | composed, split (gamma exact, N_c bf16) {err_ratio(S_exact, S_split):.2e}")
