"""Operator layer: the reference's operator signatures on top of the C ABI.

``chunk_gated_delta_rule`` and ``fused_recurrent_gated_delta_rule`` keep the names,
argument meaning and error behaviour of
src/llamafactory/model/fla/ops/gated_delta_rule/chunk.py:273-392 and
.../fused_recurrent.py:218-335 (reference checkout), so the model code that calls
them (infinitevl_standard/modeling_infinitevl.py:1297-1320) needs no change.
PyTorch is used for device memory and the current stream only.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import IVL_DTYPE_BF16, IVL_DTYPE_F32

_workspaces: dict = {}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return IVL_DTYPE_F32
    if t.dtype == torch.bfloat16:
        return IVL_DTYPE_BF16
    raise TypeError(f"state must be float32 or bfloat16, got {t.dtype}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def gdn_workspace(B: int, T: int, H: int, device) -> torch.Tensor:
    """Cached scratch buffer for the chunk operator (grown on demand, reused across calls and
    layers so that steady-state calls allocate nothing -- a CUDA-graph requirement)."""
    lib = _lib.load()
    need = lib.ivl_gdn_chunk_workspace_bytes(B, T, H)
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need + 1024:
        if torch.cuda.is_current_stream_capturing():
            # a buffer allocated during capture lives in the graph's private pool; keeping it in this module-level
            # cache would outlive the graph.  Pre-size with gdn_workspace(B, T, H, device) before capturing.
            raise _lib.IvlError(f"GDN workspace of {need} bytes must be allocated before CUDA-graph capture: call "
                                f"infinitevl_b200.ops.gdn_workspace({B}, {T}, {H}, device) on this stream first")
        ws = None
        _workspaces.pop(key, None)   # drop the old buffer before allocating the larger one
        ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    off = (-ws.data_ptr()) % 1024
    return ws[off:off + need]


def release_workspaces() -> None:
    """Free the cached scratch buffers of the chunk operator (22.5 KB per token per call shape: 2.95 GB after one
    128K-token prefill, held per (device, stream) until this is called) and of the SWA decode kernel."""
    _workspaces.clear()
    from . import swa as _swa
    _swa._decode_ws.clear()


def _check_common(q, k, v, g, beta, initial_state, cu_seqlens, head_first):
    if not q.is_cuda:
        raise _lib.IvlError("infinitevl_b200 operators run on CUDA tensors only (no CPU fallback)")
    assert q.dtype == k.dtype == v.dtype
    assert len(beta.shape) == 3, "beta must be of shape [B, T, H]."
    # The kernels take B, T, H, K from q alone: every other operand must really have that geometry (a k / v with
    # fewer heads would be read past its end, not broadcast).
    if q.dim() != 4 or k.shape != q.shape:
        raise ValueError(f"k must have q's shape [B, T, H, K]: q {tuple(q.shape)}, k {tuple(k.shape)}")
    if v.dim() != 4 or v.shape[:3] != q.shape[:3]:
        raise ValueError(f"v must be [B, T, H, V] with q's B, T, H: q {tuple(q.shape)}, v {tuple(v.shape)}")
    lead = (q.shape[0], q.shape[2], q.shape[1]) if head_first else tuple(q.shape[:3])
    if tuple(g.shape) != tuple(beta.shape) or tuple(beta.shape) not in (tuple(q.shape[:3]), lead):
        raise ValueError(f"g and beta must be [B, T, H] like q: q {tuple(q.shape)}, g {tuple(g.shape)}, "
                         f"beta {tuple(beta.shape)}")
    if cu_seqlens is not None:
        if int(cu_seqlens[0]) != 0:
            raise ValueError(f"cu_seqlens must start at 0, got {int(cu_seqlens[0])}")
        if q.shape[0] != 1:
            raise ValueError(
                f"The batch size is expected to be 1 rather than {q.shape[0]} when using `cu_seqlens`."
                f"Please flatten variable-length inputs before processing.")
        if head_first:
            raise RuntimeError("Sequences with variable lengths are not supported for head-first mode")
        if initial_state is not None and initial_state.shape[0] != len(cu_seqlens) - 1:
            raise ValueError(
                f"The number of initial states is expected to be equal to the number of input sequences, "
                f"i.e., {len(cu_seqlens) - 1} rather than {initial_state.shape[0]}.")


def _run_chunk(q, k, v, g, beta, scale, h0, ht, o, l2norm):
    lib = _lib.load()
    B, T, H, K = q.shape
    V = v.shape[-1]
    ws = gdn_workspace(B, T, H, q.device)
    code = lib.ivl_gdn_chunk_fwd(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(),
        _ptr(h0), _dtype_code(h0) if h0 is not None else 0, o.data_ptr(),
        _ptr(ht), _dtype_code(ht) if ht is not None else 0,
        B, T, H, K, V, float(scale), int(l2norm), ws.data_ptr(), ws.numel(), _stream_ptr(q.device))
    _lib.check(code, "ivl_gdn_chunk_fwd")


def _run_recurrent(q, k, v, g, beta, scale, h0, ht, o, l2norm):
    lib = _lib.load()
    B, T, H, K = q.shape
    V = v.shape[-1]
    code = lib.ivl_gdn_recurrent_fwd(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(),
        _ptr(h0), _dtype_code(h0) if h0 is not None else 0, o.data_ptr(),
        _ptr(ht), _dtype_code(ht) if ht is not None else 0,
        B, T, H, K, V, float(scale), int(l2norm), _stream_ptr(q.device))
    _lib.check(code, "ivl_gdn_recurrent_fwd")


def _gated_delta_rule(runner, q, k, v, g, beta, scale, initial_state, output_final_state, cu_seqlens, l2norm,
                      state_out=None):
    q, k, v = (x.to(torch.bfloat16).contiguous() for x in (q, k, v))
    g = g.to(torch.float32).contiguous()
    beta = beta.to(torch.bfloat16).contiguous()
    B, T, H, K = q.shape
    V = v.shape[-1]
    o = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=q.device)
    if initial_state is not None:
        initial_state = initial_state.contiguous()
    if cu_seqlens is None:
        ht = None
        if output_final_state:
            ht = state_out if state_out is not None else torch.empty(B, H, K, V, dtype=torch.float32, device=q.device)
        runner(q, k, v, g, beta, scale, initial_state, ht, o, l2norm)
        return o, ht
    # packed variable-length sequences: every sequence is an independent scan over a contiguous
    # token range of the flattened batch (fla/ops/gated_delta_rule/chunk.py:355-369)
    bounds = [int(x) for x in cu_seqlens.tolist()]
    N = len(bounds) - 1
    ht = torch.empty(N, H, K, V, dtype=torch.float32, device=q.device) if output_final_state else None
    if runner is _run_chunk:
        _run_chunk_varlen(q, k, v, g, beta, scale, initial_state, ht, o, l2norm, bounds)
        return o, ht
    for n in range(N):
        s, e = bounds[n], bounds[n + 1]
        if e <= s:
            if ht is not None:
                ht[n].copy_(initial_state[n] if initial_state is not None else torch.zeros_like(ht[n]))
            continue
        runner(q[:, s:e], k[:, s:e], v[:, s:e], g[:, s:e], beta[:, s:e], scale,
               None if initial_state is None else initial_state[n:n + 1],
               None if ht is None else ht[n:n + 1], o[:, s:e], l2norm)
    return o, ht


def varlen_chunk_tables(bounds, device):
    """The chunk geometry of a packed batch, as the C ABI wants it (include/ivl_b200.h,
    ivl_gdn_chunk_fwd_varlen): sequences are cut into 64-token chunks that never straddle a
    boundary -- the reference's prepare_chunk_indices (fla/ops/gated_delta_rule/chunk.py:211-214)."""
    tok0, valid, begin = [], [], [0]
    for n in range(len(bounds) - 1):
        s, e = bounds[n], bounds[n + 1]
        for t in range(s, e, 64):
            tok0.append(t)
            valid.append(min(64, e - t))
        begin.append(len(tok0))
    as_dev = lambda x: torch.tensor(x, dtype=torch.int32).to(device, non_blocking=False)
    return as_dev(tok0), as_dev(valid), as_dev(begin), len(tok0)


def _run_chunk_varlen(q, k, v, g, beta, scale, h0, ht, o, l2norm, bounds):
    """One launch pair for the whole packed batch."""
    lib = _lib.load()
    _, T, H, K = q.shape
    V = v.shape[-1]
    N = len(bounds) - 1
    tok0, valid, begin, num_chunks = varlen_chunk_tables(bounds, q.device)
    if bounds[-1] < T:
        o[:, bounds[-1]:].zero_()  # tokens past the last sequence belong to nobody
    if num_chunks == 0:
        if ht is not None:
            ht.copy_(h0.to(ht.dtype) if h0 is not None else torch.zeros_like(ht))
        return
    ws = gdn_workspace(1, 64 * num_chunks, H, q.device)
    code = lib.ivl_gdn_chunk_fwd_varlen(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), g.data_ptr(), beta.data_ptr(),
        _ptr(h0), _dtype_code(h0) if h0 is not None else 0, o.data_ptr(),
        _ptr(ht), _dtype_code(ht) if ht is not None else 0,
        T, H, K, V, float(scale), int(l2norm), tok0.data_ptr(), valid.data_ptr(), num_chunks, begin.data_ptr(), N,
        ws.data_ptr(), ws.numel(), _stream_ptr(q.device))
    _lib.check(code, "ivl_gdn_chunk_fwd_varlen")


def chunk_gated_delta_rule_fused(xq: torch.Tensor, xk: torch.Tensor, v: torch.Tensor, a: torch.Tensor, b: torch.Tensor,
                                 conv_weight_q: torch.Tensor, conv_weight_k: torch.Tensor, A_log: torch.Tensor,
                                 dt_bias: torch.Tensor, conv_state_q: Optional[torch.Tensor] = None,
                                 conv_state_k: Optional[torch.Tensor] = None, output_conv_state: bool = False,
                                 scale: Optional[float] = None, initial_state: Optional[torch.Tensor] = None,
                                 output_final_state: bool = False, state_out: Optional[torch.Tensor] = None):
    """Prefill-side fusion (`ivl_gdn_chunk_fwd_fused`): the chunk operator on the RAW q / k projection
    outputs -- the depthwise causal conv + SiLU of q and k, the gate math and the L2 norm all run inside the operator's
    pre-pass.  Replaces std:1263-1266 (q / k ShortConvolution), std:1293-1294 (g, beta) and std:1298-1308 in one call.
        xq, xk [B,T,H*128] bf16 (before the conv); v [B,T,H,256] bf16 (after ITS conv); a, b [B,T,H] bf16;
        conv_weight_* [H*128, 4] (or [H*128, 1, 4]) bf16; A_log, dt_bias [H]; conv_state_* [B, H*128, 4] or None.
    Returns (o, final_state | None, (conv_state_q, conv_state_k) | None); bit-identical to the unfused chain."""
    if not xq.is_cuda:
        raise _lib.IvlError("infinitevl_b200 operators run on CUDA tensors only (no CPU fallback)")
    B, T, D = xq.shape
    H = a.shape[-1]
    K = D // H
    V = v.shape[-1]
    if xk.shape != xq.shape or v.shape[:3] != (B, T, H) or a.shape != (B, T, H) or b.shape != (B, T, H) or K != 128:
        raise ValueError("chunk_gated_delta_rule_fused: inconsistent operand shapes")
    bf = lambda t: t.to(torch.bfloat16).contiguous()
    xq, xk, v, a, b = bf(xq), bf(xk), bf(v), bf(a), bf(b)
    wq, wk = bf(conv_weight_q.reshape(D, -1)), bf(conv_weight_k.reshape(D, -1))
    if wq.shape[1] != 4 or wk.shape != wq.shape:
        raise ValueError("chunk_gated_delta_rule_fused: conv kernel size must be 4")
    A32, dt32 = A_log.detach().float().contiguous(), dt_bias.detach().float().contiguous()
    cq = bf(conv_state_q) if conv_state_q is not None else None
    ck = bf(conv_state_k) if conv_state_k is not None else None
    cq_out = torch.empty(B, D, 4, dtype=torch.bfloat16, device=xq.device) if output_conv_state else None
    ck_out = torch.empty_like(cq_out) if output_conv_state else None
    o = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=xq.device)
    h0 = initial_state.contiguous() if initial_state is not None else None
    ht = None
    if output_final_state:
        ht = state_out if state_out is not None else torch.empty(B, H, K, V, dtype=torch.float32, device=xq.device)
    ws = gdn_workspace(B, T, H, xq.device)
    code = _lib.load().ivl_gdn_chunk_fwd_fused(
        xq.data_ptr(), xk.data_ptr(), v.data_ptr(), a.data_ptr(), b.data_ptr(), wq.data_ptr(), wk.data_ptr(),
        _ptr(cq), _ptr(ck), _ptr(cq_out), _ptr(ck_out), A32.data_ptr(), dt32.data_ptr(),
        _ptr(h0), _dtype_code(h0) if h0 is not None else 0, o.data_ptr(), _ptr(ht), _dtype_code(ht) if ht is not None else 0,
        B, T, H, K, V, float(scale or K ** -0.5), ws.data_ptr(), ws.numel(), _stream_ptr(xq.device))
    _lib.check(code, "ivl_gdn_chunk_fwd_fused")
    return o, ht, ((cq_out, ck_out) if output_conv_state else None)


def chunk_gated_delta_rule(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    g: torch.Tensor,
    beta: torch.Tensor,
    scale: Optional[float] = None,
    initial_state: Optional[torch.Tensor] = None,
    output_final_state: bool = False,
    cu_seqlens: Optional[torch.LongTensor] = None,
    head_first: bool = False,
    use_qk_l2norm_in_kernel: bool = False,
    state_out: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """q,k [B,T,H,K]; v [B,T,H,V]; g (log decay, fp32) and beta [B,T,H];
    initial_state [N,H,K,V] (fp32 or bf16).  Returns (o [B,T,H,V] in q.dtype, final_state fp32 | None).

    ``state_out`` (extension): write the final state into this preallocated fp32/bf16 buffer
    instead of allocating -- what a CUDA-graph-captured cache update needs."""
    _check_common(q, k, v, g, beta, initial_state, cu_seqlens, head_first)
    assert q.dtype != torch.float32, "ChunkGatedDeltaRuleFunction does not support float32. Please use bfloat16."
    if head_first:
        q, k, v = (x.transpose(1, 2) for x in (q, k, v))
        beta, g = beta.transpose(1, 2), g.transpose(1, 2)
    if scale is None:
        scale = k.shape[-1] ** -0.5
    else:
        assert scale > 0, "Scale must be positive."
    out_dtype = q.dtype
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (q, k, v, g, beta, initial_state))
    if needs_grad:
        # training: the reference's ChunkGatedDeltaRuleFunction (fla/ops/gated_delta_rule/chunk.py:180-269)
        if state_out is not None:
            raise ValueError("state_out (in-place final state) is an inference extension; not available with autograd")
        o, ht = _chunk_autograd(q, k, v, g, beta, scale, initial_state, output_final_state, cu_seqlens,
                                use_qk_l2norm_in_kernel)
        o = o.to(out_dtype)
        return (o.transpose(1, 2) if head_first else o), ht
    with torch.no_grad():
        o, ht = _gated_delta_rule(_run_chunk, q, k, v, g, beta, scale, initial_state, output_final_state, cu_seqlens,
                                  use_qk_l2norm_in_kernel, state_out)
    o = o.to(out_dtype)
    if head_first:
        o = o.transpose(1, 2)
    return o, ht


@torch.no_grad()
def fused_recurrent_gated_delta_rule(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    g: torch.Tensor,
    beta: Optional[torch.Tensor] = None,
    scale: Optional[float] = None,
    initial_state: Optional[torch.Tensor] = None,
    output_final_state: bool = False,
    cu_seqlens: Optional[torch.LongTensor] = None,
    head_first: bool = False,
    use_qk_l2norm_in_kernel: bool = False,
    state_out: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Token-serial form of the same operator (decode / q_len <= 64).  Time-first layout by
    default, as the model calls it (the vendored copy defaulted to head_first=True but the
    model never passes the flag and the pip package it runs against is time-first)."""
    if beta is None:
        beta = torch.ones_like(q[..., 0])
    _check_common(q, k, v, g, beta, initial_state, cu_seqlens, head_first)
    if head_first:
        q, k, v = (x.transpose(1, 2) for x in (q, k, v))
        beta, g = beta.transpose(1, 2), g.transpose(1, 2)
    if scale is None:
        scale = k.shape[-1] ** -0.5
    else:
        assert scale > 0, "scale must be positive"
    out_dtype = q.dtype
    o, ht = _gated_delta_rule(_run_recurrent, q, k, v, g, beta, scale, initial_state, output_final_state,
                              cu_seqlens, use_qk_l2norm_in_kernel, state_out)
    o = o.to(out_dtype)
    if head_first:
        o = o.transpose(1, 2)
    return o, ht


# ------------------------------------------------------------------------------------------------
# autograd (training drop-in)
# ------------------------------------------------------------------------------------------------
def _normalised_rows(x: torch.Tensor, l2norm: bool):
    """The rows the forward kernels multiply with: fp32 L2-normalisation rounded to bf16 (l2norm.py:42), as fp32.
    Returns (rows, rstd | None)."""
    xf = x.float()
    if not l2norm:
        return xf.contiguous(), None
    rstd = torch.rsqrt(xf.pow(2).sum(-1, keepdim=True) + 1e-6)
    return (xf * rstd).to(torch.bfloat16).float().contiguous(), rstd


class ChunkGatedDeltaRuleFunction(torch.autograd.Function):
    """Forward: the chunk kernels (ivl_gdn_chunk_fwd).  Backward: ivl_gdn_bwd -- the exact fp32 gradient of the
    recurrence with recomputed states (the reference recomputes w, u, h and runs its chunked backward kernels,
    fla/ops/gated_delta_rule/chunk.py:237-269)."""

    @staticmethod
    def forward(ctx, q, k, v, g, beta, scale, initial_state, output_final_state, l2norm):
        with torch.no_grad():
            o, ht = _gated_delta_rule(_run_chunk, q, k, v, g, beta, scale, initial_state, output_final_state, None,
                                      l2norm, None)
        ctx.save_for_backward(q, k, v, g, beta, initial_state)
        ctx.scale, ctx.l2norm = float(scale), bool(l2norm)
        return o, ht

    @staticmethod
    def backward(ctx, d_o, d_ht):
        q, k, v, g, beta, h0 = ctx.saved_tensors
        lib = _lib.load()
        B, T, H, K = q.shape
        V = v.shape[-1]
        dev = q.device
        qn, rq = _normalised_rows(q, ctx.l2norm)
        kn, rk = _normalised_rows(k, ctx.l2norm)
        vb = v.to(torch.bfloat16).contiguous()
        gf = g.float().contiguous()
        bf = beta.float().contiguous()
        dob = d_o.to(torch.bfloat16).contiguous()
        h0f = None if h0 is None else h0.float().contiguous()
        dht = None if d_ht is None else d_ht.float().contiguous()
        dqn = torch.empty(B, T, H, K, dtype=torch.float32, device=dev)
        dkn = torch.empty_like(dqn)
        dv = torch.empty(B, T, H, V, dtype=torch.float32, device=dev)
        dg = torch.empty(B, T, H, dtype=torch.float32, device=dev)
        db = torch.empty_like(dg)
        dh0 = torch.empty(B, H, K, V, dtype=torch.float32, device=dev) if h0 is not None else None
        ws = torch.empty(lib.ivl_gdn_bwd_workspace_bytes(B, T, H), dtype=torch.uint8, device=dev)
        code = lib.ivl_gdn_bwd(qn.data_ptr(), kn.data_ptr(), vb.data_ptr(), gf.data_ptr(), bf.data_ptr(), dob.data_ptr(),
                               _ptr(h0f), _ptr(dht), dqn.data_ptr(), dkn.data_ptr(), dv.data_ptr(), dg.data_ptr(),
                               db.data_ptr(), _ptr(dh0), B, T, H, K, V, ctx.scale, ws.data_ptr(), ws.numel(),
                               _stream_ptr(dev))
        _lib.check(code, "ivl_gdn_bwd")

        def through_norm(dy, x, rstd):      # y = x rsqrt(sum x^2 + eps):  dx = rstd (dy - y (y . dy))
            if rstd is None:
                return dy
            y = x.float() * rstd
            return rstd * (dy - y * (y * dy).sum(-1, keepdim=True))

        dq = through_norm(dqn, q, rq).to(q.dtype)
        dk = through_norm(dkn, k, rk).to(k.dtype)
        return (dq, dk, dv.to(v.dtype), dg.to(g.dtype), db.to(beta.dtype), None,
                None if h0 is None else dh0.to(h0.dtype), None, None)


def _chunk_autograd(q, k, v, g, beta, scale, initial_state, output_final_state, cu_seqlens, l2norm):
    if cu_seqlens is None:
        return ChunkGatedDeltaRuleFunction.apply(q, k, v, g, beta, scale, initial_state, output_final_state, l2norm)
    # packed sequences: one differentiable call per sequence (the backward kernel takes dense rows)
    bounds = [int(x) for x in cu_seqlens.tolist()]
    outs, states = [], []
    for n in range(len(bounds) - 1):
        s0, e0 = bounds[n], bounds[n + 1]
        h0 = None if initial_state is None else initial_state[n:n + 1]
        if e0 <= s0:
            states.append(h0.float() if h0 is not None else
                          torch.zeros(1, q.shape[2], q.shape[3], v.shape[3], device=q.device))
            continue
        o, ht = ChunkGatedDeltaRuleFunction.apply(q[:, s0:e0], k[:, s0:e0], v[:, s0:e0], g[:, s0:e0], beta[:, s0:e0],
                                                  scale, h0, output_final_state, l2norm)
        outs.append(o)
        states.append(ht)
    o = torch.cat(outs, dim=1)
    if bounds[-1] < q.shape[1]:
        o = torch.cat([o, o.new_zeros(1, q.shape[1] - bounds[-1], *o.shape[2:])], dim=1)
    return o, (torch.cat(states, dim=0) if output_final_state else None)
