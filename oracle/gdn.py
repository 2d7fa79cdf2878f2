"""fp32 CPU restatement of the Gated DeltaNet token mixer (SURVEY.md rows a-G0 .. a-G8).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Citations are relative to
/root/reference (``fla/`` = src/llamafactory/model/fla/, ``std`` =
infinitevl/infinitevl_standard/modeling_infinitevl.py) or, for the un-vendored
dependency flash-linear-attention (requirements.txt:19-20 pins 0.4.0; the image
ships 0.5.1), to ``site-packages/fla/...``.

All functions take tensors laid out the way the reference operators do
(time-first: [B, T, H, D]) and compute in ``dtype`` (float32 by default;
float64 is accepted for tighter self-checks).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def err_ratio(ref: torch.Tensor, x: torch.Tensor) -> float:
    """RMS(ref - x) / RMS(ref): the reference's own comparison convention
    (fla/ops/utils/testing.py:11-14)."""
    ref = ref.detach().double().flatten()
    x = x.detach().double().flatten()
    den = ref.square().mean().sqrt().item()
    num = (ref - x).square().mean().sqrt().item()
    return num / (den + 1e-12)


def l2norm_ref(x: torch.Tensor, eps: float = 1e-6, dtype=torch.float32) -> torch.Tensor:
    """y = x / sqrt(sum(x^2) + eps) over the last axis (fla/modules/l2norm.py:34-42)."""
    x = x.to(dtype)
    return x * torch.rsqrt(x.square().sum(-1, keepdim=True) + eps)


def short_conv_ref(
    x: torch.Tensor,
    weight: torch.Tensor,
    cache: Optional[torch.Tensor] = None,
    activation: Optional[str] = "silu",
    dtype=torch.float32,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Depthwise causal convolution with a carried input tail.

    x [B, T, D]; weight [D, W] (the module stores [D, 1, W]); cache [B, D, W]
    holds the last W *inputs*, newest in column W-1.  Returns (y [B, T, D],
    new_cache [B, D, W]).  y[t, d] = act(sum_i w[d, i] * xpad[t + i, d]) where
    xpad is the cache tail followed by x (zeros when there is no cache).
    Follows fla/modules/convolution.py:224-293 with the carried-state
    semantics of the pip package that actually executes at inference
    (site-packages/fla/modules/conv/short_conv.py:187-199,
    .../conv/triton/kernels.py:85-126): the cache IS left context.
    """
    B, T, D = x.shape
    if weight.dim() == 3:
        weight = weight.squeeze(1)
    W = weight.shape[-1]
    xf = x.to(dtype)
    wf = weight.to(dtype)
    left = torch.zeros(B, W, D, dtype=dtype) if cache is None else cache.to(dtype).transpose(1, 2)
    xpad = torch.cat([left, xf], dim=1)  # [B, W + T, D]; xpad[W + t] = x[t]
    y = torch.zeros(B, T, D, dtype=dtype)
    for i in range(W):
        # tap i multiplies the input (W - 1 - i) steps in the past
        y = y + wf[:, i] * xpad[:, i + 1 : i + 1 + T, :]
    if activation in ("silu", "swish"):
        y = F.silu(y)
    new_cache = xpad[:, -W:, :].transpose(1, 2).contiguous()
    return y, new_cache


def gdn_gate_ref(a: torch.Tensor, b: torch.Tensor, A_log: torch.Tensor, dt_bias: torch.Tensor,
                 dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """g = -exp(A_log) * softplus(a + dt_bias) (log-decay, fp32) and beta = sigmoid(b)
    from the two H-wide projections (std:1293-1294)."""
    g = -A_log.to(dtype).exp() * F.softplus(a.to(dtype) + dt_bias.to(dtype))
    beta = torch.sigmoid(b.to(dtype))
    return g, beta


def gdn_recurrent_ref(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    g: torch.Tensor,
    beta: torch.Tensor,
    scale: Optional[float] = None,
    initial_state: Optional[torch.Tensor] = None,
    use_qk_l2norm: bool = True,
    dtype=torch.float32,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """The definition of the operator: one token at a time.

        S <- exp(g_t) S ;  v' = beta_t (v_t - S^T k_t) ;  S <- S + k_t v'^T ;  o_t = scale S^T q_t

    q, k [B, T, H, K]; v [B, T, H, V]; g, beta [B, T, H]; state [B, H, K, V].
    Follows fla/ops/gated_delta_rule/fused_recurrent.py:85-101 (the kernel the
    model runs for q_len <= 64, std:1230) and
    site-packages/fla/ops/gated_delta_rule/naive.py:13-63.
    """
    B, T, H, K = k.shape
    V = v.shape[-1]
    if scale is None:
        scale = K ** -0.5
    qf, kf = q.to(dtype), k.to(dtype)
    if use_qk_l2norm:
        qf, kf = l2norm_ref(qf, dtype=dtype), l2norm_ref(kf, dtype=dtype)
    vf, gf, bf = v.to(dtype), g.to(dtype), beta.to(dtype)
    S = torch.zeros(B, H, K, V, dtype=dtype) if initial_state is None else initial_state.to(dtype).clone()
    o = torch.empty(B, T, H, V, dtype=dtype)
    for t in range(T):
        S = S * gf[:, t].exp()[..., None, None]
        pred = torch.einsum("bhkv,bhk->bhv", S, kf[:, t])
        vp = (vf[:, t] - pred) * bf[:, t][..., None]
        S = S + kf[:, t][..., None] * vp[..., None, :]
        o[:, t] = torch.einsum("bhkv,bhk->bhv", S, qf[:, t]) * scale
    return o, S


def gdn_chunk_ref(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    g: torch.Tensor,
    beta: torch.Tensor,
    scale: Optional[float] = None,
    initial_state: Optional[torch.Tensor] = None,
    use_qk_l2norm: bool = True,
    chunk_size: int = 64,
    dtype=torch.float32,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Chunkwise form (C = 64) -- the algorithm the CUDA kernels implement.

    Per chunk, with G the in-chunk inclusive cumsum of g and Gamma_ij = exp(G_i - G_j):
        A  = (I + tril_-1(diag(beta) (K K^T * Gamma)))^-1            fla/ops/gated_delta_rule/wy_fast.py:164-210
        Wg = A (beta K exp(G)) ;  U = A (beta V)                      wy_fast.py:287-320
        Vn = U - Wg S                                                 fla/ops/common/chunk_delta_h.py:109-117
        O  = scale [ (Q exp(G)) S + tril(Q K^T * Gamma) Vn ]          fla/ops/common/chunk_o.py:92-113
        S <- exp(G_C) S + (K exp(G_C - G))^T Vn                       chunk_delta_h.py:109-120
    (the gated single-A formulation; equivalent to the vendored two-matrix
    Aw/Au form).  Ragged tails are zero-padded with
    g = 0, beta = 0, which leaves the state untouched.
    """
    B, T, H, K = k.shape
    V = v.shape[-1]
    C = chunk_size
    if scale is None:
        scale = K ** -0.5
    qf, kf = q.to(dtype), k.to(dtype)
    if use_qk_l2norm:
        qf, kf = l2norm_ref(qf, dtype=dtype), l2norm_ref(kf, dtype=dtype)
    vf, gf, bf = v.to(dtype), g.to(dtype), beta.to(dtype)
    pad = (-T) % C
    if pad:
        qf, kf, vf = (F.pad(x, (0, 0, 0, 0, 0, pad)) for x in (qf, kf, vf))
        gf, bf = F.pad(gf, (0, 0, 0, pad)), F.pad(bf, (0, 0, 0, pad))
    NT = (T + pad) // C
    # [B, H, NT, C, D]
    qc, kc, vc = (x.permute(0, 2, 1, 3).reshape(B, H, NT, C, -1) for x in (qf, kf, vf))
    gc = gf.permute(0, 2, 1).reshape(B, H, NT, C)
    bc = bf.permute(0, 2, 1).reshape(B, H, NT, C)
    G = gc.cumsum(-1)
    Gam = (G[..., :, None] - G[..., None, :]).tril().exp().tril()  # Gamma_ij, i >= j
    eye = torch.eye(C, dtype=dtype)
    L = ((kc * bc[..., None]) @ kc.transpose(-1, -2) * Gam).tril(-1)
    A = torch.linalg.solve_triangular(eye + L, eye.expand_as(L).contiguous(), upper=False, unitriangular=True)
    Wg = A @ (kc * (bc * G.exp())[..., None])
    U = A @ (vc * bc[..., None])
    P = (qc @ kc.transpose(-1, -2) * Gam).tril()
    Qg = qc * G.exp()[..., None]
    Kt = kc * (G[..., -1:] - G).exp()[..., None]
    gamma = G[..., -1].exp()

    S = torch.zeros(B, H, K, V, dtype=dtype) if initial_state is None else initial_state.to(dtype).clone()
    o = torch.empty(B, H, NT, C, V, dtype=dtype)
    for c in range(NT):
        Vn = U[:, :, c] - Wg[:, :, c] @ S
        o[:, :, c] = (Qg[:, :, c] @ S + P[:, :, c] @ Vn) * scale
        S = S * gamma[:, :, c][..., None, None] + Kt[:, :, c].transpose(-1, -2) @ Vn
    o = o.reshape(B, H, NT * C, V)[:, :, :T].permute(0, 2, 1, 3).contiguous()
    return o, S


def gdn_chunk_segmented_ref(q, k, v, g, beta, segments: int = 4, scale: Optional[float] = None,
                            initial_state: Optional[torch.Tensor] = None, use_qk_l2norm: bool = True,
                            chunk_size: int = 64, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """The chunk form with PARALLELISM ALONG THE SEQUENCE (DESIGN.md section 6, item 1 -- the restructuring the serial
    scan kernel needs to get past its per-chunk hand-off latencies; test infrastructure like the rest of oracle/).

    Per chunk the state map is affine:  S_{c+1} = A_c S_c + B_c  with
        A_c = gamma_c I - Kt_c^T Wg_c   (K x K)        B_c = Kt_c^T U_c   (K x V)
    (from Vn = U - Wg S and S <- gamma S + Kt^T Vn of `gdn_chunk_ref`).  The sequence is cut into `segments` runs of
    chunks; pass 1 composes, independently per segment, the segment map (T_p, R_p) -- the same recurrence run on the
    augmented state [I | 0] -- pass 2 is the short serial fix-up  S_start[p+1] = T_p S_start[p] + R_p,  pass 3 runs the
    ordinary scan of every segment from its true start state (independent again).  Must equal `gdn_chunk_ref`."""
    B, T, H, K = k.shape
    V = v.shape[-1]
    C = chunk_size
    if scale is None:
        scale = K ** -0.5
    qf, kf = q.to(dtype), k.to(dtype)
    if use_qk_l2norm:
        qf, kf = l2norm_ref(qf, dtype=dtype), l2norm_ref(kf, dtype=dtype)
    vf, gf, bf = v.to(dtype), g.to(dtype), beta.to(dtype)
    pad = (-T) % C
    if pad:
        qf, kf, vf = (F.pad(x, (0, 0, 0, 0, 0, pad)) for x in (qf, kf, vf))
        gf, bf = F.pad(gf, (0, 0, 0, pad)), F.pad(bf, (0, 0, 0, pad))
    NT = (T + pad) // C
    qc, kc, vc = (x.permute(0, 2, 1, 3).reshape(B, H, NT, C, -1) for x in (qf, kf, vf))
    gc = gf.permute(0, 2, 1).reshape(B, H, NT, C)
    bc = bf.permute(0, 2, 1).reshape(B, H, NT, C)
    G = gc.cumsum(-1)
    Gam = (G[..., :, None] - G[..., None, :]).tril().exp().tril()
    eye = torch.eye(C, dtype=dtype)
    L = ((kc * bc[..., None]) @ kc.transpose(-1, -2) * Gam).tril(-1)
    A = torch.linalg.solve_triangular(eye + L, eye.expand_as(L).contiguous(), upper=False, unitriangular=True)
    Wg = A @ (kc * (bc * G.exp())[..., None])
    U = A @ (vc * bc[..., None])
    P = (qc @ kc.transpose(-1, -2) * Gam).tril()
    Qg = qc * G.exp()[..., None]
    Kt = kc * (G[..., -1:] - G).exp()[..., None]
    gamma = G[..., -1].exp()
    # the affine map of every chunk (parallel over chunks: prep-side work)
    eyeK = torch.eye(K, dtype=dtype)
    Ac = gamma[..., None, None] * eyeK - Kt.transpose(-1, -2) @ Wg          # [B, H, NT, K, K]
    Bc = Kt.transpose(-1, -2) @ U                                           # [B, H, NT, K, V]
    bounds = [round(i * NT / segments) for i in range(segments + 1)]
    # pass 1: segment maps, independent per segment
    Tp, Rp = [], []
    for p in range(segments):
        Tm = eyeK.expand(B, H, K, K).clone()
        Rm = torch.zeros(B, H, K, V, dtype=dtype)
        for c in range(bounds[p], bounds[p + 1]):
            Tm = Ac[:, :, c] @ Tm
            Rm = Ac[:, :, c] @ Rm + Bc[:, :, c]
        Tp.append(Tm)
        Rp.append(Rm)
    # pass 2: start state of every segment (serial over the few segments)
    S0 = torch.zeros(B, H, K, V, dtype=dtype) if initial_state is None else initial_state.to(dtype).clone()
    starts = [S0]
    for p in range(segments):
        starts.append(Tp[p] @ starts[p] + Rp[p])
    # pass 3: the ordinary scan per segment from its true start state (independent per segment)
    o = torch.empty(B, H, NT, C, V, dtype=dtype)
    for p in range(segments):
        S = starts[p].clone()
        for c in range(bounds[p], bounds[p + 1]):
            Vn = U[:, :, c] - Wg[:, :, c] @ S
            o[:, :, c] = (Qg[:, :, c] @ S + P[:, :, c] @ Vn) * scale
            S = S * gamma[:, :, c][..., None, None] + Kt[:, :, c].transpose(-1, -2) @ Vn
    o = o.reshape(B, H, NT * C, V)[:, :, :T].permute(0, 2, 1, 3).contiguous()
    return o, starts[-1]


def rmsnorm_gated_ref(x: torch.Tensor, gate: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5,
                      dtype=torch.float32) -> torch.Tensor:
    """y = x * rsqrt(mean(x^2) + eps) * w * gate * sigmoid(gate) over the last axis
    (fla/modules/fused_norm_gate.py:59-92; module use std:1210,1338)."""
    xf, gf = x.to(dtype), gate.to(dtype)
    y = xf * torch.rsqrt(xf.square().mean(-1, keepdim=True) + eps) * weight.to(dtype)
    return y * gf * torch.sigmoid(gf)


def rmsnorm_ref(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6, dtype=torch.float32) -> torch.Tensor:
    """Decoder RMSNorm (Qwen2RMSNorm, std:50 import; eps 1e-6)."""
    xf = x.to(dtype)
    return weight.to(dtype) * (xf * torch.rsqrt(xf.square().mean(-1, keepdim=True) + eps))


def gdn_mixer_ref(
    hidden: torch.Tensor,
    params: dict,
    conv_cache: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
    state: Optional[torch.Tensor] = None,
    H: int = 16,
    K: int = 128,
    V: int = 256,
    norm_eps: float = 1e-5,
    mode: Optional[str] = None,
    dtype=torch.float32,
    proj_dtype=None,
):
    """Whole GatedDeltaNet.forward (std:1215-1347) on [B, T, hidden].

    ``proj_dtype=torch.bfloat16`` rounds every projection output to bf16, which is what the
    reference's bf16 model does -- the decay g is very sensitive to
    the rounding of a_proj's output, so mixer-level comparisons of a bf16 model need it.

    ``params`` uses the reference's parameter names: q_proj.weight,
    k_proj.weight, v_proj.weight, a_proj.weight, b_proj.weight, A_log, dt_bias,
    {q,k,v}_conv1d.weight, g_proj.weight, o_norm.weight, o_proj.weight.
    Returns (out, (conv_q, conv_k, conv_v), state).
    """
    B, T, _ = hidden.shape
    x = hidden.to(dtype)
    def lin(name):
        y = x @ params[name + ".weight"].to(dtype).t()
        return y if proj_dtype is None else y.to(proj_dtype).to(dtype)
    cq, ck, cv = conv_cache if conv_cache is not None else (None, None, None)
    q, ncq = short_conv_ref(lin("q_proj"), params["q_conv1d.weight"], cq, dtype=dtype)
    k, nck = short_conv_ref(lin("k_proj"), params["k_conv1d.weight"], ck, dtype=dtype)
    v, ncv = short_conv_ref(lin("v_proj"), params["v_conv1d.weight"], cv, dtype=dtype)
    q, k, v = q.view(B, T, H, K), k.view(B, T, H, K), v.view(B, T, H, V)
    g, beta = gdn_gate_ref(lin("a_proj"), lin("b_proj"), params["A_log"], params["dt_bias"], dtype=dtype)
    if mode is None:
        mode = "fused_recurrent" if T <= 64 else "chunk"  # std:1230
    fn = gdn_recurrent_ref if mode == "fused_recurrent" else gdn_chunk_ref
    o, S = fn(q, k, v, g, beta, initial_state=state, use_qk_l2norm=True, dtype=dtype)
    gate = lin("g_proj").view(B, T, H, V)
    o = rmsnorm_gated_ref(o, gate, params["o_norm.weight"], eps=norm_eps, dtype=dtype)
    out = o.reshape(B, T, H * V) @ params["o_proj.weight"].to(dtype).t()
    return out, (ncq, nck, ncv), S
