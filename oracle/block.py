"""fp32 CPU restatement of the decoder-block wiring around the mixers (SURVEY.md rows a-S, a-T).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  ``std`` =
/root/reference/infinitevl/infinitevl_standard/modeling_infinitevl.py.  Parameters are addressed by
the reference's state-dict names (``layers.{i}.self_attn.q_proj.weight`` ...).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .cache import LinearCacheRef, SlidingWindowCacheRef
from .gdn import gdn_mixer_ref, rmsnorm_ref
from .swa import mrope_apply_ref, mrope_cos_sin_ref, swa_attention_ref


def swa_mixer_ref(hidden: torch.Tensor, p: Dict[str, torch.Tensor], cos: torch.Tensor, sin: torch.Tensor,
                  cache: Optional[SlidingWindowCacheRef] = None, Hq: int = 16, Hkv: int = 2, D: int = 128,
                  window: Optional[int] = 8192, mrope_section=(16, 24, 24), dtype=torch.float32,
                  proj_dtype=None) -> torch.Tensor:
    """InfiniteVLSelfAttention.forward (std:1032-1113): q/k/v projections with bias, M-RoPE on the new
    tokens, cache update returning [previous tail ; new], windowed causal attention, o_proj."""
    B, T, _ = hidden.shape
    x = hidden.to(dtype)
    def lin(n):
        y = x @ p[n + ".weight"].to(dtype).t() + (p[n + ".bias"].to(dtype) if n + ".bias" in p else 0)
        return y if proj_dtype is None else y.to(proj_dtype).to(dtype)
    q = lin("q_proj").view(B, T, Hq, D).transpose(1, 2)
    k = lin("k_proj").view(B, T, Hkv, D).transpose(1, 2)
    v = lin("v_proj").view(B, T, Hkv, D).transpose(1, 2)
    q, k = mrope_apply_ref(q, k, cos.to(dtype), sin.to(dtype), mrope_section)
    if cache is not None:
        k, v = cache.update(k, v)
    o = swa_attention_ref(q, k, v, scale=D ** -0.5, window=window, dtype=dtype)  # [B, T, Hq, D]
    return o.reshape(B, T, Hq * D) @ p["o_proj.weight"].to(dtype).t()


def mlp_ref(x: torch.Tensor, p: Dict[str, torch.Tensor], dtype=torch.float32) -> torch.Tensor:
    x = x.to(dtype)
    return (F.silu(x @ p["gate_proj.weight"].to(dtype).t()) * (x @ p["up_proj.weight"].to(dtype).t())) \
        @ p["down_proj.weight"].to(dtype).t()


def _sub(params: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}


def hybrid_decoder_ref(hidden: torch.Tensor, params: Dict[str, torch.Tensor], layer_types: List[str],
                       position_ids: torch.Tensor, caches: Optional[list] = None, window: int = 8192,
                       rms_eps: float = 1e-6, mixers_only: bool = False, final_norm: bool = True,
                       theta: float = 1e6, dtype=torch.float32, proj_dtype=None) -> torch.Tensor:
    """InfiniteVLTextModel's layer loop (std:1549-1576) over ``layer_types`` at the 3B head shapes.
    position_ids [3, B, T].  ``caches`` (one SlidingWindowCacheRef / dict per layer) enables streaming."""
    cos, sin = mrope_cos_sin_ref(position_ids, 128, theta, out_dtype=dtype)
    h = hidden.to(dtype)
    for i, lt in enumerate(layer_types):
        lp = _sub(params, f"layers.{i}.")
        x = rmsnorm_ref(h, lp["input_layernorm.weight"], rms_eps, dtype=dtype)
        if lt == "sliding_attention":
            c = caches[i] if caches is not None else None
            y = swa_mixer_ref(x, _sub(lp, "self_attn."), cos, sin, cache=c, window=window, dtype=dtype,
                              proj_dtype=proj_dtype)
        else:
            st = caches[i] if caches is not None else None
            conv = st.get("conv") if st else None
            state = st.get("state") if st else None
            y, nconv, nstate = gdn_mixer_ref(x, _sub(lp, "self_attn."), conv_cache=conv, state=state, dtype=dtype,
                                             proj_dtype=proj_dtype)
            if st is not None:
                st["conv"], st["state"] = nconv, nstate
        h = h + y
        if not mixers_only:
            h = h + mlp_ref(rmsnorm_ref(h, lp["post_attention_layernorm.weight"], rms_eps, dtype=dtype),
                            _sub(lp, "mlp."), dtype=dtype)
    if final_norm:
        h = rmsnorm_ref(h, params["norm.weight"], rms_eps, dtype=dtype)
    return h
