"""Sliding-window attention operator on top of the C ABI (ivl_swa_fwd).

``sliding_window_attention_forward`` has the signature of the HF attention-interface
callables the reference looks up by name (``ALL_ATTENTION_FUNCTIONS["flash_attention_2"]``,
infinitevl_standard/modeling_infinitevl.py:1092-1108): it takes query [B,Hq,Tq,D] and
key/value [B,Hkv,Tk,D] and returns (attn_output [B,Tq,Hq,D], None).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib


_decode_ws: dict = {}


def _strides3(t: torch.Tensor):
    """(batch, time, head) element strides of a [B, T, H, D] view as a ctypes int64[3]."""
    assert t.stride(3) == 1, "innermost (head_dim) axis must be contiguous"
    return (ctypes.c_int64 * 3)(t.stride(0), t.stride(1), t.stride(2))


def swa_attention_bthd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, window: Optional[int] = None,
                       scale: Optional[float] = None, out: Optional[torch.Tensor] = None,
                       key_pos0: int = 0) -> torch.Tensor:
    """q [B,Tq,Hq,128], k/v [B,Tk,Hkv,128] (bf16; arbitrary batch/time/head strides) -> out [B,Tq,Hq,128].
    Causal with bottom-right alignment; ``window`` keys visible (self included) once Tk > window.
    ``key_pos0``: position of k[:, 0] in its sequence (cache + new tokens, halo + shard): the kernel tiles the keys at
    absolute multiples of 64, so the output is bit-identical to the one-shot prefill of the whole sequence."""
    if not q.is_cuda:
        raise _lib.IvlError("infinitevl_b200 operators run on CUDA tensors only (no CPU fallback)")
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16, "SWA kernel computes in bf16"
    B, Tq, Hq, D = q.shape
    Tk, Hkv = k.shape[1], k.shape[2]
    assert k.shape == v.shape and k.shape[0] == B and k.shape[3] == D
    fix = lambda t: t if (t.stride(3) == 1 and all(s % 8 == 0 for s in t.stride()[:3]) and t.data_ptr() % 16 == 0) \
        else t.contiguous()
    q, k, v = fix(q), fix(k), fix(v)
    if out is None:
        out = torch.empty(B, Tq, Hq, D, dtype=torch.bfloat16, device=q.device)
    lib = _lib.load()
    if Tq == 1 and Hq // Hkv <= 8 and out.is_contiguous():
        # decode step: split-KV kernel (HBM-bound) instead of the tensor-core prefill kernel
        q = q.contiguous()
        need = lib.ivl_swa_decode_workspace_bytes(B, Tk, Hq)
        key = (q.device.index, torch.cuda.current_stream(q.device).cuda_stream)
        ws = _decode_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=q.device)
            _decode_ws[key] = ws
        code = lib.ivl_swa_decode_fwd(q.data_ptr(), k.data_ptr(), _strides3(k), v.data_ptr(), _strides3(v),
                                      out.data_ptr(), B, Tk, Hq, Hkv, D, int(window or 0), float(scale or 0.0),
                                      ws.data_ptr(), ws.numel(), torch.cuda.current_stream(q.device).cuda_stream)
        _lib.check(code, "ivl_swa_decode_fwd")
        return out
    code = lib.ivl_swa_fwd_pos(q.data_ptr(), _strides3(q), k.data_ptr(), _strides3(k), v.data_ptr(), _strides3(v),
                               out.data_ptr(), _strides3(out), B, Tq, Tk, Hq, Hkv, D, int(window or 0),
                               float(scale or 0.0), int(key_pos0), torch.cuda.current_stream(q.device).cuda_stream)
    _lib.check(code, "ivl_swa_fwd_pos")
    return out


def swa_attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, dout: torch.Tensor,
                           window: Optional[int] = None, scale: Optional[float] = None, block: int = 512):
    """Gradient of `swa_attention_bthd` (attention half; the reference trains through flash-attn's
    backward): recomputes the probabilities query block by query block from q, k, v and the saved output -- the
    flash-attention backward identities, dS = P * (dP - rowsum(dO * O)) -- with plain matrix products over the
    visible key range of the block, fp32.  q, out, dout [B,Tq,Hq,D]; k, v [B,Tk,Hkv,D] -> (dq, dk, dv) in fp32.
    Works on any device (the formula is layout-only torch code); O(block * window) memory."""
    B, Tq, Hq, D = q.shape
    Tk, Hkv = k.shape[1], k.shape[2]
    G = Hq // Hkv
    scale = float(scale) if scale else D ** -0.5
    W = int(window) if (window and Tk > int(window)) else None
    shift = Tk - Tq
    dq = torch.zeros(B, Tq, Hq, D, dtype=torch.float32, device=q.device)
    dk = torch.zeros(B, Tk, Hkv, D, dtype=torch.float32, device=q.device)
    dv = torch.zeros_like(dk)
    grp = lambda t: t.float().reshape(t.shape[0], t.shape[1], Hkv, G, D).permute(0, 2, 3, 1, 4)   # [B,Hkv,G,n,D]
    kv = lambda t: t.float().permute(0, 2, 1, 3)                                                      # [B,Hkv,m,D]
    for i0 in range(0, Tq, block):
        i1 = min(Tq, i0 + block)
        pos = torch.arange(i0, i1, device=q.device) + shift
        j_lo = max(0, i0 + shift - W + 1) if W else 0
        j_hi = min(Tk, i1 + shift)
        if j_hi <= j_lo:
            continue
        j = torch.arange(j_lo, j_hi, device=q.device)
        vis = j[None, :] <= pos[:, None]
        if W:
            vis &= (pos[:, None] - j[None, :]) <= W - 1
        qs, do, o = grp(q[:, i0:i1]), grp(dout[:, i0:i1]), grp(out[:, i0:i1])
        ks, vs = kv(k[:, j_lo:j_hi]), kv(v[:, j_lo:j_hi])
        S = torch.einsum("bhgnd,bhmd->bhgnm", qs, ks) * scale
        P = torch.softmax(S.masked_fill(~vis, float("-inf")), dim=-1)
        P = torch.nan_to_num(P)            # a query with no visible key (does not happen for causal self-attention)
        dP = torch.einsum("bhgnd,bhmd->bhgnm", do, vs)
        dS = P * (dP - (do * o).sum(-1, keepdim=True)) * scale
        dq[:, i0:i1] = torch.einsum("bhgnm,bhmd->bhgnd", dS, ks).permute(0, 3, 1, 2, 4).reshape(B, i1 - i0, Hq, D)
        dk[:, j_lo:j_hi] += torch.einsum("bhgnm,bhgnd->bhmd", dS, qs).permute(0, 2, 1, 3)
        dv[:, j_lo:j_hi] += torch.einsum("bhgnm,bhgnd->bhmd", P, do).permute(0, 2, 1, 3)
    return dq, dk, dv


class SlidingWindowAttentionFunction(torch.autograd.Function):
    """Differentiable form of the attention operator: CUDA forward (ivl_swa_fwd_pos), recomputing backward."""

    @staticmethod
    def forward(ctx, q, k, v, window, scale, key_pos0):
        out = swa_attention_bthd(q.detach(), k.detach(), v.detach(), window=window, scale=scale, key_pos0=key_pos0)
        ctx.save_for_backward(q, k, v, out)
        ctx.window, ctx.scale = window, scale
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out = ctx.saved_tensors
        dq, dk, dv = swa_attention_backward(q, k, v, out, dout, window=ctx.window, scale=ctx.scale)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None, None


def swa_attention_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens, window: Optional[int] = None,
                         scale: Optional[float] = None) -> torch.Tensor:
    """Packed batch: q [1,T,Hq,128], k/v [1,T,Hkv,128] holding len(cu_seqlens) - 1 sequences back to back
    (cu_seqlens: boundaries, tensor or list); every sequence attends to itself only.  -> [1,T,Hq,128]; tokens outside
    every sequence come out zero."""
    if not q.is_cuda:
        raise _lib.IvlError("infinitevl_b200 operators run on CUDA tensors only (no CPU fallback)")
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16 and q.shape[0] == 1 and k.shape[:2] == q.shape[:2] == v.shape[:2]
    _, T, Hq, D = q.shape
    Hkv = k.shape[2]
    bounds = [int(x) for x in (cu_seqlens.tolist() if hasattr(cu_seqlens, "tolist") else cu_seqlens)]
    tok0, lo, hi = [], [], []
    for s0, s1 in zip(bounds[:-1], bounds[1:]):
        for t in range(s0, s1, 128):
            tok0.append(t); lo.append(s0); hi.append(s1)
    out = torch.zeros(1, T, Hq, D, dtype=torch.bfloat16, device=q.device)
    if not tok0:
        return out
    fix = lambda t: t if (t.stride(3) == 1 and all(s % 8 == 0 for s in t.stride()[:3]) and t.data_ptr() % 16 == 0) \
        else t.contiguous()
    q, k, v = fix(q), fix(k), fix(v)
    tab = torch.tensor([tok0, lo, hi], dtype=torch.int32).to(q.device)
    code = _lib.load().ivl_swa_fwd_varlen(q.data_ptr(), _strides3(q), k.data_ptr(), _strides3(k), v.data_ptr(), _strides3(v),
                                          out.data_ptr(), _strides3(out), T, Hq, Hkv, D, int(window or 0), float(scale or 0.0),
                                          tab[0].data_ptr(), tab[1].data_ptr(), tab[2].data_ptr(), len(tok0),
                                          torch.cuda.current_stream(q.device).cuda_stream)
    _lib.check(code, "ivl_swa_fwd_varlen")
    return out


def swa_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, window: Optional[int] = None,
                  scale: Optional[float] = None, key_pos0: int = 0) -> torch.Tensor:
    """HF head-first layout: q [B,Hq,Tq,D], k/v [B,Hkv,Tk,D] -> [B,Tq,Hq,D].  Differentiable (training through the HF
    attention interface): when a gradient is required the call goes through SlidingWindowAttentionFunction."""
    qb, kb, vb = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        return SlidingWindowAttentionFunction.apply(qb, kb, vb, window, scale, key_pos0)
    return swa_attention_bthd(qb, kb, vb, window, scale, key_pos0=key_pos0)


def sliding_window_attention_forward(module, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                                     attention_mask: Optional[torch.Tensor] = None, dropout: float = 0.0,
                                     scaling: Optional[float] = None, sliding_window: Optional[int] = None,
                                     **kwargs) -> Tuple[torch.Tensor, None]:
    """Drop-in for the HF attention interface (same arguments as transformers'
    flash_attention_forward).  Padding masks are not supported -- the reference's FA2 path gets
    ``attention_mask=None`` for unpadded batches, and its GDN layers ignore padding anyway
    (modeling_infinitevl.py:1223)."""
    if attention_mask is not None and attention_mask.dim() == 2 and not bool(attention_mask.all()):
        raise NotImplementedError("padded batches are not supported by the B200 SWA kernel")
    if dropout:
        raise NotImplementedError("attention dropout is not supported (inference / dropout=0 training only)")
    if scaling is None:
        scaling = query.shape[-1] ** -0.5
    cu = kwargs.get("cu_seqlens", kwargs.get("cu_seq_lens_q"))
    if cu is not None:   # packed sequences (the HF glue's flash_attn_varlen path): self-attention per sequence
        return swa_attention_varlen(query.transpose(1, 2), key.transpose(1, 2), value.transpose(1, 2), cu,
                                    window=sliding_window, scale=scaling), None
    # key_position_offset (extension keyword): position of key[:, :, 0] in the sequence, see swa_attention_bthd
    out = swa_attention(query, key, value, window=sliding_window, scale=scaling,
                        key_pos0=int(kwargs.get("key_position_offset", 0) or 0))
    return out, None


def register_attention_interface(name: str = "ivl_b200_swa") -> str:
    """Register `sliding_window_attention_forward` with the HF attention interface, so that a model whose
    `config._attn_implementation` is `name` (the reference forces "flash_attention_2" at std:1028 and looks the callable
    up in ALL_ATTENTION_FUNCTIONS at std:1092-1108) reaches the B200 kernel.  Returns the name."""
    from transformers import AttentionInterface
    AttentionInterface.register(name, sliding_window_attention_forward)
    return name
