"""Seeded synthetic inputs shared by the parity tests, bench.py and smoke()."""
import math

import torch


def gdn_inputs(B=1, T=1024, H=16, K=128, V=256, seed=0, device="cpu", with_state=True):
    """q,k ~ N(0,1) pre-norm bf16; v ~ N(0,1) bf16; beta = sigmoid(N(0,1)) bf16;
    g = -exp(a) softplus(N(0,1) + b) fp32 with a = log U(1e-3,16) (std:1167-1170) and b the
    inverse softplus of log-U(1e-3, 1e-1) (std:1176-1183); h0 ~ N(0,1) fp32."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(B, T, H, K, generator=gen).to(torch.bfloat16)
    k = torch.randn(B, T, H, K, generator=gen).to(torch.bfloat16)
    v = torch.randn(B, T, H, V, generator=gen).to(torch.bfloat16)
    beta = torch.sigmoid(torch.randn(B, T, H, generator=gen)).to(torch.bfloat16)
    a = torch.log(torch.empty(H).uniform_(1e-3, 16, generator=gen))
    dt = torch.exp(torch.empty(H).uniform_(math.log(1e-3), math.log(1e-1), generator=gen))
    dt_bias = dt + torch.log(-torch.expm1(-dt))
    g = -torch.exp(a) * torch.nn.functional.softplus(torch.randn(B, T, H, generator=gen) + dt_bias)
    h0 = torch.randn(B, H, K, V, generator=gen) if with_state else None
    out = [q, k, v, g.float(), beta, h0]
    return [x.to(device) if x is not None else None for x in out]
