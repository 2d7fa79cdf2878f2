"""fp32 CPU restatement of the Sliding-Window Attention mixer (SURVEY.md rows a-S, a-S1, a-S2).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  ``std`` =
/root/reference/infinitevl/infinitevl_standard/modeling_infinitevl.py.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch


def mrope_cos_sin_ref(position_ids: torch.Tensor, head_dim: int = 128, theta: float = 1e6,
                      out_dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin tables [3, B, T, head_dim] from the three position rows (t, h, w).

    inv_freq_j = theta^(-2j/head_dim); freqs = pos * inv_freq computed in fp32,
    duplicated to the full head_dim, then cast to ``out_dtype`` -- the model casts
    to bf16 before the multiply (std:916-930).
    """
    assert position_ids.dim() == 3 and position_ids.shape[0] == 3
    half = head_dim // 2
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    assert inv_freq.numel() == half
    freqs = position_ids[..., None].float() * inv_freq  # [3, B, T, half]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos().to(out_dtype), emb.sin().to(out_dtype)


def _select_sections(table: torch.Tensor, mrope_section: Sequence[int]) -> torch.Tensor:
    """Channel section i of the doubled section list takes position row i % 3 (std:972-978)."""
    sec = list(mrope_section) * 2
    parts, start = [], 0
    for i, n in enumerate(sec):
        parts.append(table[i % 3, ..., start:start + n])
        start += n
    return torch.cat(parts, dim=-1)  # [B, T, head_dim]


def _rot_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)  # std:521-525


def mrope_apply_ref(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
                    mrope_section: Sequence[int] = (16, 24, 24)) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rotate q [B, Hq, T, D] and k [B, Hkv, T, D] (std:949-984).  The arithmetic runs in
    the dtype of the inputs, as in the model (bf16 there)."""
    c = _select_sections(cos, mrope_section)[:, None]
    s = _select_sections(sin, mrope_section)[:, None]
    return q * c + _rot_half(q) * s, k * c + _rot_half(k) * s


def swa_visible_mask(Tq: int, Tk: int, window: Optional[int]) -> torch.Tensor:
    """Boolean [Tq, Tk]: key j is visible to query i (bottom-right aligned causal,
    optionally windowed): 0 <= (i + Tk - Tq) - j <= window - 1.

    The window is only enforced when Tk > window, as the HF flash-attention
    glue does (site-packages/transformers/modeling_flash_attention_utils.py:627-632);
    for Tk <= window the two rules coincide.
    """
    i = torch.arange(Tq)[:, None] + (Tk - Tq)
    j = torch.arange(Tk)[None, :]
    vis = j <= i
    if window is not None and Tk > window:
        vis &= (i - j) <= (window - 1)
    return vis


def swa_attention_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None,
                      window: Optional[int] = 8192, dtype=torch.float32) -> torch.Tensor:
    """softmax(scale q k^T + mask) v with GQA.

    q [B, Hq, Tq, D], k/v [B, Hkv, Tk, D] -> [B, Tq, Hq, D] (the layout the HF
    attention interface returns, std:557-580 for the eager form; q-head h reads
    kv-head h // (Hq / Hkv), std:545-554).
    """
    B, Hq, Tq, D = q.shape
    Hkv, Tk = k.shape[1], k.shape[2]
    rep = Hq // Hkv
    if scale is None:
        scale = D ** -0.5
    qf = q.to(dtype)
    kf = k.to(dtype).repeat_interleave(rep, dim=1)
    vf = v.to(dtype).repeat_interleave(rep, dim=1)
    vis = swa_visible_mask(Tq, Tk, window)
    out = torch.empty(B, Hq, Tq, D, dtype=dtype)
    blk = 1024
    for s in range(0, Tq, blk):  # query blocks keep the score matrix small
        e = min(Tq, s + blk)
        sc = qf[:, :, s:e] @ kf.transpose(-1, -2) * scale
        sc = sc.masked_fill(~vis[s:e], float("-inf"))
        out[:, :, s:e] = torch.softmax(sc, dim=-1) @ vf
    return out.transpose(1, 2).contiguous()
