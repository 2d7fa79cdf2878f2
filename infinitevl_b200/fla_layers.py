"""Training-side mixer with the constructor and parameter names of the reference's `fla.layers.GatedDeltaNet`
(src/llamafactory/model/fla/layers/gated_deltanet.py:81-316), which src/llamafactory/model/convert.py:79-153 looks up
by name (`getattr(layers, "GatedDeltaNet")`) and subclasses when it swaps the attention of a Qwen2.5-VL checkpoint
for linear mixers.

Training (`module.training`): every piece is differentiable -- the delta-rule operator through
`ops.chunk_gated_delta_rule` (CUDA forward + CUDA backward, ivl_gdn_chunk_fwd / ivl_gdn_bwd), the short convolutions,
gates and the gated RMS norm through plain torch ops (cuDNN / element-wise: token-local, not part of the hot path).
Evaluation: the inference kernels of `modeling.py`.  The cache protocol is the fla one the reference layer uses
(`past_key_values[layer_idx]` -> {"conv_state", "recurrent_state"}, `.update(recurrent_state=, conv_state=,
layer_idx=, offset=)`)."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import modeling, ops


def _conv_silu_train(x: torch.Tensor, weight: torch.Tensor, cache: Optional[torch.Tensor]):
    """Differentiable depthwise causal conv (kernel 4) + SiLU; cache [B, D, 4] = previous inputs (left context)."""
    B, T, D = x.shape
    W = weight.shape[-1]
    xt = x.transpose(1, 2)
    left = cache[..., 1:].to(x.dtype) if cache is not None else xt.new_zeros(B, D, W - 1)
    full = torch.cat([left, xt], dim=-1)
    y = F.conv1d(full.float(), weight.float().reshape(D, 1, W), groups=D)
    new_cache = full[..., -W:] if full.shape[-1] >= W else F.pad(full, (W - full.shape[-1], 0))
    return F.silu(y).to(x.dtype).transpose(1, 2), new_cache


def _rmsnorm_gated_train(x, gate, weight, eps):
    xf, gf = x.float(), gate.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * weight.float()
    return (y * gf * torch.sigmoid(gf)).to(x.dtype)


class GatedDeltaNet(nn.Module):
    def __init__(self, hidden_size: int = 2048, expand_v: float = 2, head_dim: int = 256, num_heads: int = 6,
                 mode: str = "chunk", use_gate: bool = True, use_short_conv: bool = True, conv_size: int = 4,
                 conv_bias: bool = False, layer_idx: int = None, norm_eps: float = 1e-5, mimic_init: bool = True,
                 **kwargs):
        super().__init__()
        self.mode, self.mimic_init = mode, mimic_init
        self.hidden_size, self.expand_v = hidden_size, expand_v
        self.use_gate, self.use_short_conv, self.conv_size, self.conv_bias = use_gate, use_short_conv, conv_size, conv_bias
        self.head_dim, self.num_heads = head_dim, num_heads
        self.key_dim = int(num_heads * head_dim)
        self.value_dim = int(self.key_dim * expand_v)
        self.head_k_dim, self.head_v_dim = head_dim, int(head_dim * expand_v)
        self.layer_idx = layer_idx
        self.norm_eps = norm_eps
        if not math.isclose(self.key_dim * expand_v, self.value_dim, rel_tol=1e-5):
            raise ValueError(f"expand_v={expand_v} does not produce an integer value when multiplied by "
                             f"key_dim={self.key_dim}.")
        assert mode in ["chunk", "fused_recurrent"], f"Not suppoerted mode `{mode}`."
        if (self.head_k_dim, self.head_v_dim, conv_size) != (128, 256, 4) or not (use_gate and use_short_conv) or conv_bias:
            raise NotImplementedError("the B200 kernels implement head_dim=128, expand_v=2, conv_size=4 with the output "
                                      "gate and bias-free short convolutions (the InfiniteVL configuration)")
        self.q_proj = nn.Linear(hidden_size, self.key_dim, bias=False)
        self.k_proj = nn.Linear(hidden_size, self.key_dim, bias=False)
        self.v_proj = nn.Linear(hidden_size, self.value_dim, bias=False)
        self.a_proj = nn.Linear(hidden_size, num_heads, bias=False)
        self.b_proj = nn.Linear(hidden_size, num_heads, bias=False)
        if mimic_init:      # start as (almost) the attention layer it replaces: no decay, no write gate dynamics
            A = torch.ones(num_heads, dtype=torch.float32)
            nn.init.constant_(self.a_proj.weight, 0.0)
            nn.init.constant_(self.b_proj.weight, 0.0)
        else:
            A = torch.empty(num_heads, dtype=torch.float32).uniform_(0, 16)
        self.A_log = nn.Parameter(torch.log(A))
        self.A_log._no_weight_decay = True
        dt = torch.clamp(torch.exp(torch.rand(num_heads) * (math.log(0.001) - math.log(0.001)) + math.log(0.001)), min=1e-4)
        self.dt_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
        self.dt_bias._no_weight_decay = True
        self.q_conv1d = modeling.ShortConvolution(self.key_dim, conv_size, activation="silu")
        self.k_conv1d = modeling.ShortConvolution(self.key_dim, conv_size, activation="silu")
        self.v_conv1d = modeling.ShortConvolution(self.value_dim, conv_size, activation="silu")
        if mimic_init:
            with torch.no_grad():
                for c in (self.q_conv1d, self.k_conv1d, self.v_conv1d):
                    c.weight.zero_()
                    c.weight[:, 0, 3] = 1
        self.g_proj = nn.Linear(hidden_size, self.value_dim, bias=False)
        self.o_norm = modeling.FusedRMSNormGated(self.head_v_dim, eps=norm_eps)
        self.o_proj = nn.Linear(self.value_dim, hidden_size, bias=False)

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, past_key_values=None,
                use_cache: Optional[bool] = False, output_attentions: Optional[bool] = False, **kwargs):
        # the reference drops the padding mask on entry (gated_deltanet.py:212)
        B, T, _ = hidden_states.shape
        grad = torch.is_grad_enabled() and (self.training or hidden_states.requires_grad)
        mode = "fused_recurrent" if (T <= 64 and not grad) else "chunk"
        last_state = None
        if past_key_values is not None and len(past_key_values) > self.layer_idx:
            last_state = past_key_values[self.layer_idx]
        cu_seqlens = kwargs.get("cu_seqlens", None)
        if cu_seqlens is not None:
            raise NotImplementedError("packed sequences (cu_seqlens) are not supported by the B200 GatedDeltaNet layer")
        cq = ck = cv = None
        if last_state is not None:
            cq, ck, cv = last_state["conv_state"]
        xq, xk, xv = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
        if grad:
            q, cq = _conv_silu_train(xq, self.q_conv1d.weight, cq)
            k, ck = _conv_silu_train(xk, self.k_conv1d.weight, ck)
            v, cv = _conv_silu_train(xv, self.v_conv1d.weight, cv)
        else:
            q, cq = self.q_conv1d(xq, cache=cq, output_final_state=use_cache)
            k, ck = self.k_conv1d(xk, cache=ck, output_final_state=use_cache)
            v, cv = self.v_conv1d(xv, cache=cv, output_final_state=use_cache)
        q = q.view(B, T, self.num_heads, self.head_k_dim)
        k = k.view(B, T, self.num_heads, self.head_k_dim)
        v = v.view(B, T, self.num_heads, self.head_v_dim)
        beta = self.b_proj(hidden_states).sigmoid()
        g = -self.A_log.float().exp() * F.softplus(self.a_proj(hidden_states).float() + self.dt_bias)
        state = last_state["recurrent_state"] if last_state is not None else None
        fn = ops.chunk_gated_delta_rule if mode == "chunk" else ops.fused_recurrent_gated_delta_rule
        o, state = fn(q=q, k=k, v=v, g=g, beta=beta, initial_state=state, output_final_state=bool(use_cache),
                      use_qk_l2norm_in_kernel=True)
        if past_key_values is not None:
            past_key_values.update(recurrent_state=state, conv_state=(cq, ck, cv), layer_idx=self.layer_idx, offset=T)
        gate = self.g_proj(hidden_states).view(B, T, self.num_heads, self.head_v_dim)
        if grad:
            o = _rmsnorm_gated_train(o, gate, self.o_norm.weight, self.norm_eps)
        else:
            o = self.o_norm(o, gate)
        o = self.o_proj(o.reshape(B, T, self.value_dim))
        return o, None
