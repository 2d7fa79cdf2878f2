"""Host-side mirror of the reference's mixer / decoder-block interface for the hot path.

Same class names, constructor arguments, parameter (state-dict) names and forward signatures as
infinitevl/infinitevl_standard/modeling_infinitevl.py (GatedDeltaNet :1116-1347,
InfiniteVLSelfAttention :986-1113, InfiniteVLDecoderLayer :1350-1429, InfiniteVLRotaryEmbedding
:895-930) and the fla modules they use (ShortConvolution, FusedRMSNormGated), so a checkpoint of
the reference loads with load_state_dict and the reference's callers need no change.  The
token-mixing math runs in libivl_b200.so; projections are plain nn.Linear (cuBLAS).

Everything here is inference-only (no autograd through the custom kernels): the backward pass is
row f-1 of SURVEY.md section 8, not built yet.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops, swa
from .cache import StaticCachePrealloc


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


# ------------------------------------------------------------------------------------------------
# functional wrappers of the element-wise kernels
# ------------------------------------------------------------------------------------------------
def short_conv_silu(x: torch.Tensor, weight: torch.Tensor, cache: Optional[torch.Tensor] = None,
                    output_final_state: bool = False, activation: Optional[str] = "silu",
                    cache_out: Optional[torch.Tensor] = None):
    """x bf16 [B,T,D]; weight [D,1,4] or [D,4]; cache [B,D,4] (left context) -> (y, final_state | None)."""
    if not x.is_cuda:
        raise _lib.IvlError("infinitevl_b200 operators run on CUDA tensors only (no CPU fallback)")
    assert x.dtype == torch.bfloat16
    B, T, D = x.shape
    x = x.contiguous()
    w = weight.reshape(D, -1).to(torch.bfloat16).contiguous()
    assert w.shape[1] == 4, "the B200 short-conv kernel is specialised for kernel_size 4"
    y = torch.empty_like(x)
    if cache is not None:
        cache = cache.to(torch.bfloat16).contiguous()
    out_state = None
    if output_final_state:
        out_state = cache_out if cache_out is not None else torch.empty(B, D, 4, dtype=torch.bfloat16, device=x.device)
    code = _lib.load().ivl_short_conv_fwd(x.data_ptr(), w.data_ptr(), None if cache is None else cache.data_ptr(),
                                          y.data_ptr(), None if out_state is None else out_state.data_ptr(),
                                          B, T, D, 1 if activation in ("silu", "swish") else 0, _stream(x))
    _lib.check(code, "ivl_short_conv_fwd")
    return y, out_state


def _packed_linear(owner: nn.Module, attr: str, mods, x: torch.Tensor):
    """y_i = mods[i](x) for several nn.Linear over the same input as ONE matmul: the weights (and biases) are stacked
    row-wise once and cached on `owner` (keyed on storage, version counter and device of every parameter, so a
    load_state_dict / .to() / in-place update rebuilds the stack).  Costs one extra copy of the weights."""
    params = [p for m in mods for p in (m.weight, m.bias) if p is not None]
    key = tuple((p.data_ptr(), p._version, p.device, p.dtype) for p in params)
    cached = getattr(owner, attr, None)
    if cached is None or cached[0] != key:
        if x.is_cuda and torch.cuda.is_current_stream_capturing():
            # never build the stack inside a graph capture (it would live in the graph's private pool): this call goes
            # through the separate projections, an eager warm-up step builds it
            return tuple(m(x) for m in mods)
        with torch.no_grad():
            W = torch.cat([m.weight for m in mods], dim=0).contiguous()
            bias = None
            if any(m.bias is not None for m in mods):
                bias = torch.cat([m.bias if m.bias is not None else m.weight.new_zeros(m.weight.shape[0]) for m in mods])
        cached = (key, W, bias, [m.weight.shape[0] for m in mods])
        object.__setattr__(owner, attr, cached)     # (not a registered buffer: it must stay out of the state dict)
    return torch.nn.functional.linear(x, cached[1], cached[2]).split(cached[3], dim=-1)


def left_context_table(cu_seqlens, T: int, device) -> torch.Tensor:
    """uint8 [T]: min(3, tokens of the token's own sequence before it); tokens outside every sequence get 0."""
    cu = torch.as_tensor(cu_seqlens, device=device).to(torch.int64)
    t = torch.arange(T, device=device)
    seq = torch.bucketize(t, cu, right=True) - 1
    start = cu[seq.clamp(0, cu.numel() - 1)]
    return (t - start).clamp(0, 3).to(torch.uint8).contiguous()


def short_conv_silu_varlen(x: torch.Tensor, weight: torch.Tensor, cu_seqlens, activation: Optional[str] = "silu"):
    """Packed batch x bf16 [1,T,D]: the conv window never crosses a sequence boundary (ivl_short_conv_fwd_varlen)."""
    if not x.is_cuda:
        raise _lib.IvlError("infinitevl_b200 operators run on CUDA tensors only (no CPU fallback)")
    assert x.dtype == torch.bfloat16 and x.shape[0] == 1
    _, T, D = x.shape
    x = x.contiguous()
    w = weight.reshape(D, -1).to(torch.bfloat16).contiguous()
    assert w.shape[1] == 4, "the B200 short-conv kernel is specialised for kernel_size 4"
    lctx = left_context_table(cu_seqlens, T, x.device)
    y = torch.empty_like(x)
    code = _lib.load().ivl_short_conv_fwd_varlen(x.data_ptr(), w.data_ptr(), y.data_ptr(), lctx.data_ptr(), T, D,
                                                 1 if activation in ("silu", "swish") else 0, _stream(x))
    _lib.check(code, "ivl_short_conv_fwd_varlen")
    return y


def gdn_gates(a: torch.Tensor, b: torch.Tensor, A_log: torch.Tensor, dt_bias: torch.Tensor):
    """a, b bf16 [..., H] -> (g fp32 [..., H], beta bf16 [..., H])  (std:1293-1294)."""
    H = a.shape[-1]
    a, b = a.to(torch.bfloat16).contiguous(), b.to(torch.bfloat16).contiguous()
    g = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    beta = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    # keep the fp32 copies alive until the launch is enqueued: a temporary freed between the two
    # conversions would hand the same allocator block to both
    A32, dt32 = A_log.float().contiguous(), dt_bias.float().contiguous()
    code = _lib.load().ivl_gdn_gate_fwd(a.data_ptr(), b.data_ptr(), A32.data_ptr(), dt32.data_ptr(), g.data_ptr(),
                                        beta.data_ptr(), a.numel() // H, H, _stream(a))
    _lib.check(code, "ivl_gdn_gate_fwd")
    return g, beta


def rmsnorm_gated(x: torch.Tensor, gate: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    assert x.shape == gate.shape and x.shape[-1] == 256 and x.dtype == torch.bfloat16
    x, gate = x.contiguous(), gate.to(torch.bfloat16).contiguous()
    y = torch.empty_like(x)
    w = weight.to(torch.bfloat16).contiguous()
    code = _lib.load().ivl_rmsnorm_gated_fwd(x.data_ptr(), gate.data_ptr(), w.data_ptr(), y.data_ptr(),
                                             x.numel() // 256, 256, float(eps), _stream(x))
    _lib.check(code, "ivl_rmsnorm_gated_fwd")
    return y


def mrope_select(cos: torch.Tensor, sin: torch.Tensor, mrope_section) -> Tuple[torch.Tensor, torch.Tensor]:
    """[3,B,T,D] tables -> [B,T,D]: channel section i of the doubled list takes position row i % 3 (std:972-978)."""
    sec = list(mrope_section) * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(cos.split(sec, dim=-1))], dim=-1)
    sin = torch.cat([m[i % 3] for i, m in enumerate(sin.split(sec, dim=-1))], dim=-1)
    return cos.contiguous(), sin.contiguous()


def mrope_apply_(x_bthd: torch.Tensor, cos_btd: torch.Tensor, sin_btd: torch.Tensor) -> torch.Tensor:
    """In-place rotation of a [B,T,H,128] bf16 view (any batch/time/head strides)."""
    import ctypes
    assert x_bthd.dtype == torch.bfloat16 and x_bthd.stride(3) == 1
    B, T, Hn, D = x_bthd.shape
    strides = (ctypes.c_int64 * 3)(x_bthd.stride(0), x_bthd.stride(1), x_bthd.stride(2))
    c, s = cos_btd.to(torch.bfloat16).contiguous(), sin_btd.to(torch.bfloat16).contiguous()
    code = _lib.load().ivl_mrope_apply(x_bthd.data_ptr(), strides, c.data_ptr(), s.data_ptr(), B, T, Hn, D,
                                       _stream(x_bthd))
    _lib.check(code, "ivl_mrope_apply")
    return x_bthd


# ------------------------------------------------------------------------------------------------
# modules (reference parameter names)
# ------------------------------------------------------------------------------------------------
class ShortConvolution(nn.Module):
    """Depthwise causal conv + SiLU; weight [hidden_size, 1, kernel_size] like the nn.Conv1d the
    reference subclasses (fla/modules/convolution.py:124-293)."""

    def __init__(self, hidden_size: int, kernel_size: int = 4, bias: bool = False, activation: Optional[str] = "silu",
                 **kwargs):
        super().__init__()
        if bias:
            raise NotImplementedError("conv bias is not used by InfiniteVL (conv_bias=False)")
        self.hidden_size, self.kernel_size, self.activation = hidden_size, kernel_size, activation
        self.weight = nn.Parameter(torch.empty(hidden_size, 1, kernel_size))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        self.bias = None

    def forward(self, x: torch.Tensor, cache: Optional[torch.Tensor] = None, output_final_state: bool = False,
                cu_seqlens=None, **kwargs):
        if cu_seqlens is not None:
            # packed sequences (fla/modules/convolution.py:224-251): no carried tail, the window stops at sequence starts
            if cache is not None or output_final_state:
                raise ValueError("ShortConvolution: cu_seqlens excludes cache / output_final_state (as in the reference)")
            return short_conv_silu_varlen(x, self.weight, cu_seqlens, self.activation), None
        return short_conv_silu(x, self.weight, cache, output_final_state, self.activation)


class FusedRMSNormGated(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-5, **kwargs):
        super().__init__()
        self.hidden_size, self.eps = hidden_size, eps
        self.weight = nn.Parameter(torch.ones(hidden_size))

    def forward(self, x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
        return rmsnorm_gated(x, g, self.weight, self.eps)


class InfiniteVLRMSNorm(nn.Module):
    """Qwen2RMSNorm (fp32 inside, eps 1e-6); token-local, not part of the hot path (plain torch)."""

    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, hidden_states):
        dt = hidden_states.dtype
        h = hidden_states.to(torch.float32)
        h = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + self.variance_epsilon)
        return self.weight * h.to(dt)


class InfiniteVLRotaryEmbedding(nn.Module):
    """cos/sin tables [3,B,T,head_dim] from M-RoPE position ids [3,B,T] (std:895-930)."""

    def __init__(self, config, device=None):
        super().__init__()
        head_dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
        rs = getattr(config, "rope_scaling", None) or {}
        theta = rs.get("rope_theta", getattr(config, "rope_theta", 1e6))
        inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
        self.register_buffer("inv_freq", inv_freq.to(device) if device is not None else inv_freq, persistent=False)
        self.attention_scaling = 1.0

    @torch.no_grad()
    def forward(self, x, position_ids):
        inv = self.inv_freq[None, None, :, None].float().expand(3, position_ids.shape[1], -1, 1).to(x.device)
        pos = position_ids[:, :, None, :].float()
        freqs = (inv @ pos).transpose(2, 3)
        emb = torch.cat((freqs, freqs), dim=-1)
        return (emb.cos() * self.attention_scaling).to(x.dtype), (emb.sin() * self.attention_scaling).to(x.dtype)


class GatedDeltaNet(nn.Module):
    """Gated DeltaNet mixer (std:1116-1347).  forward(hidden_states, attention_mask=None,
    past_key_values=None, cache_position=None, **kw) -> (o, None)."""

    def __init__(self, config, layer_idx: int):
        super().__init__()
        self.mode = getattr(config, "mode", "chunk")
        self.hidden_size = config.hidden_size
        self.expand_v = config.expand_v
        self.norm_eps = getattr(config, "norm_eps", 1e-5)
        self.use_gate = getattr(config, "use_gate", True)
        self.use_short_conv = getattr(config, "use_short_conv", True)
        self.conv_size = getattr(config, "conv_size", 4)
        self.num_heads = config.num_linear_heads
        self.num_key_value_heads = getattr(config, "num_linear_key_value_heads", self.num_heads)
        self.head_dim = getattr(config, "linear_head_dim", None) or config.hidden_size // config.num_attention_heads
        self.key_dim = int(self.num_key_value_heads * self.head_dim)
        self.value_dim = int(self.key_dim * self.expand_v)
        self.head_k_dim = self.head_dim
        self.head_v_dim = int(self.head_dim * self.expand_v)
        self.layer_idx = layer_idx
        if not math.isclose(self.head_dim * self.expand_v, self.head_v_dim, rel_tol=1e-5):
            raise ValueError(f"expand_v={self.expand_v} does not produce an integer head_v_dim")
        assert self.mode in ["chunk", "fused_recurrent"], f"Not suppoerted mode `{self.mode}`."
        if not (self.use_short_conv and self.use_gate):
            raise NotImplementedError("the B200 path implements the shipped configuration (short conv + output gate)")
        if (self.head_k_dim, self.head_v_dim, self.conv_size) != (128, 256, 4):
            raise NotImplementedError("the B200 kernels are specialised for head dims K=128, V=256 and conv_size=4")
        if self.num_key_value_heads != self.num_heads:
            # the reference views k / v with num_linear_key_value_heads and lets the Triton kernels index q with the
            # same head count (std:1284-1286), which only works when the two are equal -- as in the shipped config
            raise NotImplementedError("num_linear_key_value_heads != num_linear_heads is not supported "
                                      f"({self.num_key_value_heads} vs {self.num_heads})")
        self.q_proj = nn.Linear(self.hidden_size, self.num_heads * self.head_dim, bias=False)
        self.k_proj = nn.Linear(self.hidden_size, self.key_dim, bias=False)
        self.v_proj = nn.Linear(self.hidden_size, self.value_dim, bias=False)
        self.a_proj = nn.Linear(self.hidden_size, self.num_heads, bias=False)
        self.b_proj = nn.Linear(self.hidden_size, self.num_heads, bias=False)
        A = torch.empty(self.num_heads, dtype=torch.float32).uniform_(0, 16)
        self.A_log = nn.Parameter(torch.log(A))
        self.A_log._no_weight_decay = True
        dt = torch.exp(torch.rand(self.num_heads) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
        dt = torch.clamp(dt, min=1e-4)
        self.dt_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
        self.dt_bias._no_weight_decay = True
        self.q_conv1d = ShortConvolution(self.num_heads * self.head_dim, self.conv_size, activation="silu")
        self.k_conv1d = ShortConvolution(self.key_dim, self.conv_size, activation="silu")
        self.v_conv1d = ShortConvolution(self.value_dim, self.conv_size, activation="silu")
        self.g_proj = nn.Linear(self.hidden_size, self.num_heads * self.head_v_dim, bias=False)
        self.o_norm = FusedRMSNormGated(self.head_v_dim, eps=self.norm_eps)
        self.o_proj = nn.Linear(self.num_heads * self.head_v_dim, self.hidden_size, bias=False)

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                past_key_values=None, cache_position: Optional[torch.LongTensor] = None, **kwargs):
        # padding masks are ignored, exactly as the reference does (std:1223)
        cu_seqlens = kwargs.get("cu_seqlens")
        if cu_seqlens is not None:
            # packed sequences: the reference forwards cu_seqlens to the convs and the operator (std:1234,1267,1306)
            return self._forward_packed(hidden_states, cu_seqlens, past_key_values)
        B, q_len, _ = hidden_states.shape
        mode = "fused_recurrent" if q_len <= 64 else self.mode
        prev_q = prev_k = prev_v = recurrent_state = None
        use_cache = past_key_values is not None
        if use_cache:
            (prev_q, prev_k, prev_v), recurrent_state = past_key_values.update(
                layer_idx=self.layer_idx, key_states=None, value_states=None, conv_state=None, recurrent_state=None,
                cache_kwargs={"op": "get", "cache_position": cache_position})
        if q_len == 1 and self._fused_decode_ok(prev_q, prev_k, prev_v, recurrent_state):
            return self._decode_step(hidden_states, past_key_values, cache_position, prev_q, prev_k, prev_v,
                                     recurrent_state), None
        if mode == "chunk" and self._fused_prefill_ok(hidden_states):
            # prefill-side fusion: q / k conv + SiLU, gate math and L2 norm run inside the operator's
            # pre-pass (ivl_gdn_chunk_fwd_fused); bit-identical to the kernel-by-kernel chain below
            v, new_v = self.v_conv1d(self.v_proj(hidden_states), cache=prev_v, output_final_state=use_cache)
            o, next_state, convs = ops.chunk_gated_delta_rule_fused(
                self.q_proj(hidden_states), self.k_proj(hidden_states), v.view(B, q_len, self.num_heads, self.head_v_dim),
                self.a_proj(hidden_states), self.b_proj(hidden_states), self.q_conv1d.weight, self.k_conv1d.weight,
                self.A_log, self.dt_bias, conv_state_q=prev_q, conv_state_k=prev_k, output_conv_state=use_cache,
                initial_state=recurrent_state, output_final_state=use_cache)
            if use_cache:
                past_key_values.update(layer_idx=self.layer_idx, key_states=None, value_states=None,
                                       conv_state=(convs[0], convs[1], new_v), recurrent_state=next_state,
                                       cache_kwargs={"op": "set", "delta_len": q_len, "cache_position": cache_position})
            gate = self.g_proj(hidden_states).view(B, q_len, self.num_heads, self.head_v_dim)
            o = self.o_norm(o, gate)
            return self.o_proj(o.reshape(B, q_len, self.num_heads * self.head_v_dim)), None
        q, new_q = self.q_conv1d(self.q_proj(hidden_states), cache=prev_q, output_final_state=use_cache)
        k, new_k = self.k_conv1d(self.k_proj(hidden_states), cache=prev_k, output_final_state=use_cache)
        v, new_v = self.v_conv1d(self.v_proj(hidden_states), cache=prev_v, output_final_state=use_cache)
        q = q.view(B, q_len, self.num_heads, self.head_k_dim)
        k = k.view(B, q_len, self.num_key_value_heads, self.head_k_dim)
        v = v.view(B, q_len, self.num_key_value_heads, self.head_v_dim)
        g, beta = gdn_gates(self.a_proj(hidden_states), self.b_proj(hidden_states), self.A_log, self.dt_bias)
        fn = ops.chunk_gated_delta_rule if mode == "chunk" else ops.fused_recurrent_gated_delta_rule
        o, next_state = fn(q=q, k=k, v=v, g=g, beta=beta, initial_state=recurrent_state,
                           output_final_state=use_cache, use_qk_l2norm_in_kernel=True)
        if use_cache:
            past_key_values.update(layer_idx=self.layer_idx, key_states=None, value_states=None,
                                   conv_state=(new_q, new_k, new_v), recurrent_state=next_state,
                                   cache_kwargs={"op": "set", "delta_len": q_len, "cache_position": cache_position})
        gate = self.g_proj(hidden_states).view(B, q_len, self.num_heads, self.head_v_dim)
        o = self.o_norm(o, gate)
        o = self.o_proj(o.reshape(B, q_len, self.num_heads * self.head_v_dim))
        return o, None


    def _forward_packed(self, hidden_states, cu_seqlens, past_key_values):
        """Packed variable-length batch [1, T, hidden]: every sequence is an independent scan (no cache)."""
        if past_key_values is not None:
            raise ValueError("cu_seqlens and a cache exclude each other (the reference's packed path is training-time)")
        B, T, _ = hidden_states.shape
        if B != 1:
            raise ValueError("packed sequences come as one row [1, T, hidden] (fla/ops/gated_delta_rule/chunk.py:355-359)")
        q, _ = self.q_conv1d(self.q_proj(hidden_states), cu_seqlens=cu_seqlens)
        k, _ = self.k_conv1d(self.k_proj(hidden_states), cu_seqlens=cu_seqlens)
        v, _ = self.v_conv1d(self.v_proj(hidden_states), cu_seqlens=cu_seqlens)
        g, beta = gdn_gates(self.a_proj(hidden_states), self.b_proj(hidden_states), self.A_log, self.dt_bias)
        o, _ = ops.chunk_gated_delta_rule(
            q=q.view(1, T, self.num_heads, self.head_k_dim), k=k.view(1, T, self.num_key_value_heads, self.head_k_dim),
            v=v.view(1, T, self.num_key_value_heads, self.head_v_dim), g=g, beta=beta, cu_seqlens=cu_seqlens,
            use_qk_l2norm_in_kernel=True)
        gate = self.g_proj(hidden_states).view(1, T, self.num_heads, self.head_v_dim)
        return self.o_proj(self.o_norm(o, gate).reshape(1, T, self.num_heads * self.head_v_dim)), None

    FUSED_PREFILL_MAX_T = 4096

    def _fused_prefill_ok(self, hidden_states) -> bool:
        # IVL_GDN_FUSED_PREFILL: 1 = always, 0 = never, unset = for calls of at most FUSED_PREFILL_MAX_T tokens.  Measured
        # (profiles/r02_summary.md): the conv + gate work moves into the pre-pass; at 128K tokens that makes the pre-pass
        # the longer side of the overlapped operator (whole mixer 10.29 vs 9.93 ms), for streamed frames it removes three
        # launches per layer.
        knob = os.environ.get("IVL_GDN_FUSED_PREFILL", "")
        if knob == "0" or os.environ.get("IVL_GDN_TSCAN", "3") not in ("1", "3"):
            return False
        if knob != "1" and hidden_states.shape[1] > self.FUSED_PREFILL_MAX_T:
            return False
        ws = (self.q_proj.weight, self.q_conv1d.weight, self.k_conv1d.weight)
        return (hidden_states.is_cuda and self.num_key_value_heads == self.num_heads and self.head_k_dim == 128
                and self.head_v_dim == 256 and getattr(self, "conv_size", 4) == 4
                and all(w.dtype == torch.bfloat16 for w in ws))

    # -- single-token decode: the whole mixer core in one launch (ivl_gdn_decode_step) -------------------------
    def _fused_decode_ok(self, prev_q, prev_k, prev_v, state) -> bool:
        """The fused step updates the cache buffers in place, so it needs a started cache whose conv tails are
        bf16 (the cache dtype of a bf16 model); anything else takes the kernel-by-kernel path."""
        if os.environ.get("IVL_GDN_FUSED_DECODE", "1") == "0":
            return False
        ts = (prev_q, prev_k, prev_v, state)
        ws = (self.q_proj.weight, self.q_conv1d.weight, self.k_conv1d.weight, self.v_conv1d.weight, self.o_norm.weight)
        return (all(t is not None and t.is_cuda and t.is_contiguous() for t in ts)
                and all(t.dtype == torch.bfloat16 for t in ts[:3]) and state.dtype in (torch.bfloat16, torch.float32)
                and self.num_key_value_heads == self.num_heads
                and all(w.dtype == torch.bfloat16 and w.is_contiguous() for w in ws))

    def _decode_step(self, hidden_states, past_key_values, cache_position, conv_q, conv_k, conv_v, state):
        B = hidden_states.shape[0]
        H = self.num_heads
        if B == 1 and os.environ.get("IVL_DECODE_PACKED_PROJ", "1") != "0":
            # one GEMV over the six input projections stacked row-wise instead of six (a decode step is launch-bound:
            # 2.79 -> see profiles/r02_summary.md); a one-row output splits into contiguous pieces, no copies
            mods = (self.q_proj, self.k_proj, self.v_proj, self.a_proj, self.b_proj, self.g_proj)
            xq, xk, xv, a, b, gate = _packed_linear(self, "_packed_in_proj", mods, hidden_states)
        else:
            xq, xk, xv = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
            a, b, gate = self.a_proj(hidden_states), self.b_proj(hidden_states), self.g_proj(hidden_states)
        # fp32 copies of the two per-head gate parameters, refreshed whenever the parameters change (in-place
        # update, load_state_dict, .to()): keyed on storage, version counter and device
        key = (self.A_log.data_ptr(), self.A_log._version, self.dt_bias.data_ptr(), self.dt_bias._version, xq.device)
        cached = getattr(self, "_gate_params_f32", None)
        if cached is None or cached[0] != key:
            cached = (key, self.A_log.detach().float().contiguous(), self.dt_bias.detach().float().contiguous())
            self._gate_params_f32 = cached
        f32 = cached[1:]
        out = torch.empty(B, 1, H * self.head_v_dim, dtype=torch.bfloat16, device=xq.device)
        wq, wk, wv = (m.weight for m in (self.q_conv1d, self.k_conv1d, self.v_conv1d))
        nw = self.o_norm.weight
        xq, xk, xv, a, b, gate = (t.contiguous() for t in (xq, xk, xv, a, b, gate))
        code = _lib.load().ivl_gdn_decode_step(
            xq.data_ptr(), xk.data_ptr(), xv.data_ptr(), a.data_ptr(), b.data_ptr(), gate.data_ptr(),
            wq.data_ptr(), wk.data_ptr(), wv.data_ptr(), f32[0].data_ptr(), f32[1].data_ptr(), nw.data_ptr(),
            conv_q.data_ptr(), conv_k.data_ptr(), conv_v.data_ptr(), state.data_ptr(),
            _lib.IVL_DTYPE_F32 if state.dtype == torch.float32 else _lib.IVL_DTYPE_BF16, out.data_ptr(),
            B, H, self.head_k_dim, self.head_v_dim, float(self.head_k_dim ** -0.5), float(self.norm_eps), _stream(xq))
        _lib.check(code, "ivl_gdn_decode_step")
        # the buffers were updated in place; "set" with the same tensors only advances the cache's bookkeeping
        past_key_values.update(layer_idx=self.layer_idx, key_states=None, value_states=None,
                               conv_state=(conv_q, conv_k, conv_v), recurrent_state=state,
                               cache_kwargs={"op": "set", "delta_len": 1, "cache_position": cache_position})
        return self.o_proj(out)


class InfiniteVLSelfAttention(nn.Module):
    """Sliding-window attention mixer (std:986-1113)."""

    def __init__(self, config, layer_idx: Optional[int] = None):
        super().__init__()
        self.config = config
        self.layer_idx = layer_idx
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.is_causal = True
        self.attention_dropout = getattr(config, "attention_dropout", 0.0)
        self.rope_scaling = getattr(config, "rope_scaling", None) or {"mrope_section": [16, 24, 24]}
        self.scaling = self.head_dim ** -0.5
        if self.head_dim * self.num_heads != self.hidden_size:
            raise ValueError(f"hidden_size must be divisible by num_heads (got `hidden_size`: {self.hidden_size}"
                             f" and `num_heads`: {self.num_heads}).")
        self.q_proj = nn.Linear(self.hidden_size, self.num_heads * self.head_dim, bias=True)
        self.k_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=True)
        self.v_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=True)
        self.o_proj = nn.Linear(self.num_heads * self.head_dim, self.hidden_size, bias=False)
        layer_types = getattr(config, "layer_types", None)
        is_sliding = layer_types is None or layer_idx is None or layer_types[layer_idx] == "sliding_attention"
        self.sliding_window = getattr(config, "sliding_window", None) if is_sliding else None
        self.rotary_emb = InfiniteVLRotaryEmbedding(config=config)

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.LongTensor] = None, past_key_values=None, output_attentions: bool = False,
                use_cache: bool = False, cache_position: Optional[torch.LongTensor] = None,
                position_embeddings: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, **kwargs):
        B, q_len, _ = hidden_states.shape
        if B == 1 and q_len == 1 and os.environ.get("IVL_DECODE_PACKED_PROJ", "1") != "0":
            q, k, v = _packed_linear(self, "_packed_qkv", (self.q_proj, self.k_proj, self.v_proj), hidden_states)
        else:
            q, k, v = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
        q = q.view(B, q_len, self.num_heads, self.head_dim)
        k = k.view(B, q_len, self.num_key_value_heads, self.head_dim)
        v = v.view(B, q_len, self.num_key_value_heads, self.head_dim)
        cos, sin = position_embeddings
        if cos.dim() == 4:  # [3,B,T,D] M-RoPE tables -> merged [B,T,D]
            cos, sin = mrope_select(cos, sin, self.rope_scaling["mrope_section"])
        mrope_apply_(q, cos, sin)
        mrope_apply_(k, cos, sin)
        if kwargs.get("cu_seqlens") is not None:
            # packed sequences (the HF glue's flash_attn_varlen path): self-attention per sequence, no cache
            if past_key_values is not None:
                raise ValueError("cu_seqlens and a cache exclude each other")
            out = swa.swa_attention_varlen(q, k, v, kwargs["cu_seqlens"], window=self.sliding_window, scale=self.scaling)
            return self.o_proj(out.reshape(B, q_len, self.num_heads * self.head_dim)), None
        key_states, value_states = k.transpose(1, 2), v.transpose(1, 2)  # [B,H,T,D] views, as the cache expects
        layer = past_key_values.layers[self.layer_idx] if past_key_values is not None else None
        if attention_mask is not None and attention_mask.dim() != 2:
            attention_mask = None   # 4-D additive masks are an eager/SDPA concept; the FA2 path builds none
        if layer is not None and getattr(layer, "is_ring", False) and q.is_cuda:
            # ring-buffer window cache: append + attention without concatenating or rolling the window
            if attention_mask is not None and not bool(attention_mask.all()):
                raise NotImplementedError("padded batches are not supported by the B200 SWA kernel")
            out = layer.attend(q, k, v, self.scaling, self.sliding_window)
            out = self.o_proj(out.reshape(B, q_len, self.num_heads * self.head_dim))
            return out, None
        key_pos0 = 0
        if past_key_values is not None:
            # position of the first visible key = tokens seen before this call - cached keys (Python integers of
            # the cache, as in the reference): anchors the kernel's key tiles at absolute positions
            key_pos0 = max(0, int(getattr(layer, "cumulative_length", 0)) - int(getattr(layer, "size", 0)))
            key_states, value_states = past_key_values.update(
                layer_idx=self.layer_idx, key_states=key_states, value_states=value_states, conv_state=None,
                recurrent_state=None, cache_kwargs={"sin": sin, "cos": cos, "cache_position": cache_position})
        # crop a 2-D padding mask to the visible keys exactly as the reference does before FA2 (std:1080-1090);
        # the operator then refuses masks that really pad (it has no unpad path)
        if past_key_values is not None and self.sliding_window is not None and attention_mask is not None:
            kv_len, kv_offset = past_key_values.layers[self.layer_idx].get_mask_sizes(cache_position)
            attention_mask = None if kv_offset != 0 else attention_mask[:, kv_offset:kv_offset + kv_len]
        out, _ = swa.sliding_window_attention_forward(self, q.transpose(1, 2), key_states, value_states,
                                                      attention_mask, dropout=0.0, scaling=self.scaling,
                                                      sliding_window=self.sliding_window, key_position_offset=key_pos0)
        out = self.o_proj(out.reshape(B, q_len, self.num_heads * self.head_dim))
        return out, None


class InfiniteVLTextMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.down_proj = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class InfiniteVLDecoderLayer(nn.Module):
    """RMSNorm -> mixer -> +res -> RMSNorm -> MLP -> +res (std:1350-1429)."""

    def __init__(self, config, layer_idx: int):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.layer_type = config.layer_types[layer_idx]
        if self.layer_type == "linear_attention":
            self.self_attn = GatedDeltaNet(config, layer_idx)
        elif self.layer_type in ("full_attention", "sliding_attention"):
            self.self_attn = InfiniteVLSelfAttention(config, layer_idx)
        else:
            raise ValueError(f"unknown layer type {self.layer_type}")
        self.mlp = InfiniteVLTextMLP(config)
        self.input_layernorm = InfiniteVLRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = InfiniteVLRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.attention_type = self.layer_type

    @torch.no_grad()
    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_values=None,
                output_attentions=False, use_cache=False, cache_position=None, position_embeddings=None, **kwargs):
        residual = hidden_states
        hidden_states = self.input_layernorm(hidden_states)
        hidden_states, _ = self.self_attn(hidden_states=hidden_states, attention_mask=attention_mask,
                                          position_ids=position_ids, past_key_values=past_key_values,
                                          output_attentions=output_attentions, use_cache=use_cache,
                                          cache_position=cache_position, position_embeddings=position_embeddings,
                                          **kwargs)
        hidden_states = residual + hidden_states
        residual = hidden_states
        hidden_states = self.mlp(self.post_attention_layernorm(hidden_states))
        return (residual + hidden_states,)


class HybridTextConfig:
    """Minimal stand-in for InfiniteVLTextConfig (configuration_infinitevl.py:208-286) with the shipped
    3B values (config.json:15-47); any object with these attributes works in its place."""

    def __init__(self, **kw):
        d = dict(hidden_size=2048, intermediate_size=11008, num_hidden_layers=36, num_attention_heads=16,
                 num_key_value_heads=2, rms_norm_eps=1e-6, use_sliding_window=True, sliding_window=8192,
                 attention_dropout=0.0, rope_theta=1e6,
                 rope_scaling={"rope_type": "default", "mrope_section": [16, 24, 24], "rope_theta": 1e6},
                 num_linear_heads=16, num_linear_key_value_heads=16, linear_head_dim=128, expand_v=2, mode="chunk",
                 use_gate=True, use_short_conv=True, conv_size=4, conv_bias=False, norm_eps=1e-5,
                 max_position_embeddings=128000, layer_types=None)
        d.update(kw)
        for k_, v_ in d.items():
            setattr(self, k_, v_)
        if self.layer_types is None:
            self.layer_types = ["sliding_attention" if i % 4 == 0 else "linear_attention"
                                for i in range(self.num_hidden_layers)]
        if not self.use_sliding_window:
            self.sliding_window = None
        self.head_dim = self.hidden_size // self.num_attention_heads


def normalize_position_ids(position_ids, cache_position, batch_size):
    """Position-id normalisation of InfiniteVLTextModel.forward (std:1512-1525): None -> the cache positions on all
    three M-RoPE rows; [B, T] -> the same row three times; [4, B, T] (the packed FA2 form) -> text positions
    (row 0) split from the three M-RoPE rows.  Returns (mrope_position_ids [3, B, T], text_position_ids | None).

    The reference uses the text row to build block-diagonal masks for packed sequences; the sliding-window kernel
    here is causal over the whole row, so a text row that restarts inside a sequence is refused, not ignored."""
    text_position_ids = None
    if position_ids is None:
        position_ids = cache_position.view(1, 1, -1).expand(3, batch_size, -1)
    elif position_ids.dim() == 2:
        position_ids = position_ids[None].expand(3, position_ids.shape[0], -1)
    if position_ids.dim() == 3 and position_ids.shape[0] == 4:
        text_position_ids = position_ids[0]
        position_ids = position_ids[1:]
    if position_ids.dim() != 3 or position_ids.shape[0] != 3:
        raise ValueError(f"position_ids must be [B, T], [3, B, T] or [4, B, T], got {tuple(position_ids.shape)}")
    return position_ids, text_position_ids


def packed_cu_seqlens(text_position_ids):
    """A text position row that restarts inside the row marks packed sequences (the reference turns it into a
    block-diagonal mask / flash_attn_varlen boundaries): returns cu_seqlens (int32 tensor on the same device) for a
    [1, T] row with restarts, None for an ordinary row."""
    if text_position_ids is None or text_position_ids.shape[-1] <= 1:
        return None
    restart = text_position_ids[:, 1:] <= text_position_ids[:, :-1]
    if not bool(restart.any()):
        return None
    if text_position_ids.shape[0] != 1:
        raise ValueError("packed sequences come as one row (batch size 1)")
    T = text_position_ids.shape[-1]
    cuts = (torch.nonzero(restart[0]).flatten() + 1).to(torch.int32)
    return torch.cat([cuts.new_zeros(1), cuts, cuts.new_full((1,), T)])


class HybridDecoder(nn.Module):
    """The decoder stack of InfiniteVLTextModel without embeddings / lm_head (std:1455-1591): what the hot
    path lives in.  forward(inputs_embeds [B,T,hidden], position_ids [3,B,T], [4,B,T] or [B,T]) -> hidden states."""

    def __init__(self, config, mixers_only: bool = False):
        super().__init__()
        self.config = config
        self.layers = nn.ModuleList([InfiniteVLDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = InfiniteVLRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.rotary_emb = InfiniteVLRotaryEmbedding(config=config)
        self.mixers_only = mixers_only

    def allocate_inference_cache(self, batch_size: int, device=None, dtype=None, state_dtype=None, **kw):
        p = next(self.parameters())
        return StaticCachePrealloc(config=self.config, batch_size=batch_size, device=device or p.device,
                                   dtype=dtype or p.dtype, state_dtype=state_dtype, **kw)

    @torch.no_grad()
    def forward(self, inputs_embeds, position_ids=None, past_key_values=None, cache_position=None, use_cache=None):
        B, T, _ = inputs_embeds.shape
        if cache_position is None:
            past = past_key_values.get_seq_length() if past_key_values is not None else 0
            cache_position = torch.arange(past, past + T, device=inputs_embeds.device)
        position_ids, _ = normalize_position_ids(position_ids, cache_position, B)
        cos, sin = self.rotary_emb(inputs_embeds, position_ids)
        cos, sin = mrope_select(cos, sin, self.config.rope_scaling["mrope_section"])
        h = inputs_embeds
        for layer in self.layers:
            if self.mixers_only:
                h = h + layer.self_attn(hidden_states=layer.input_layernorm(h), past_key_values=past_key_values,
                                        cache_position=cache_position, position_embeddings=(cos, sin))[0]
            else:
                h = layer(h, position_ids=position_ids, past_key_values=past_key_values,
                          cache_position=cache_position, position_embeddings=(cos, sin))[0]
        return self.norm(h)


# ------------------------------------------------------------------------------------------------
# model surface: InfiniteVLTextModel / InfiniteVLModel / InfiniteVLQwen2_5_VLForConditionalGeneration
# (std:1430-1591, 1595-1936, 1980-2322).  Same constructor arguments, module tree (= state-dict keys of the
# checkpoint: model.language_model.*, lm_head.weight), forward signatures and output fields, so that
# inference_examples/demo_streaming_inference.py's calls work unchanged.  The vision tower is outside the hot
# path: a caller attaches its own module as `model.visual`, or passes `inputs_embeds`.
# ------------------------------------------------------------------------------------------------
class _Output(dict):
    """Attribute + key + index access, like transformers' ModelOutput (None fields are skipped when indexing)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if isinstance(k, int):
            return [v for v in self.values() if v is not None][k]
        return dict.__getitem__(self, k)

    def to_tuple(self):
        return tuple(v for v in self.values() if v is not None)


class InfiniteVLConfig:
    """Minimal stand-in for InfiniteVLConfig (configuration_infinitevl.py): `text_config` (HybridTextConfig or any
    object with its attributes), the multimodal token ids and the two vision_config fields the text side reads.
    `from_dict` / `from_json` accept the checkpoint's config.json (text fields at top level or under text_config)."""

    def __init__(self, text_config=None, vision_config=None, image_token_id=151655, video_token_id=151656,
                 vision_start_token_id=151652, vision_end_token_id=151653, tie_word_embeddings=True, **kw):
        self.text_config = text_config if text_config is not None else HybridTextConfig()
        vc = dict(spatial_merge_size=2, tokens_per_second=2)
        vc.update(vision_config or {})
        self.vision_config = _Output(vc)
        self.image_token_id, self.video_token_id = image_token_id, video_token_id
        self.vision_start_token_id, self.vision_end_token_id = vision_start_token_id, vision_end_token_id
        self.tie_word_embeddings = tie_word_embeddings
        for k_, v_ in kw.items():
            setattr(self, k_, v_)

    def get_text_config(self, decoder=False):
        return self.text_config

    @classmethod
    def from_dict(cls, d):
        d = dict(d)
        text = dict(d.pop("text_config", None) or {})
        names = set(vars(HybridTextConfig()).keys()) | {"vocab_size", "pad_token_id", "use_cache"}
        for k_ in list(d):
            if k_ in names and k_ not in text:
                text[k_] = d[k_]
        if isinstance(text.get("rope_scaling"), dict):
            text["rope_scaling"] = {"rope_theta": text.get("rope_theta", 1e6), **text["rope_scaling"]}
        keep = {k_: d[k_] for k_ in ("image_token_id", "video_token_id", "vision_start_token_id", "vision_end_token_id",
                                     "tie_word_embeddings") if k_ in d}
        return cls(text_config=HybridTextConfig(**text), vision_config=d.get("vision_config"), **keep)

    @classmethod
    def from_json(cls, path):
        import json
        with open(path) as f:
            return cls.from_dict(json.load(f))


def _rope_index_vectorized(config, input_ids, image_grid_thw, video_grid_thw, second_per_grid_ts, attention_mask):
    """get_rope_index without a Python loop over the vision blocks (the reference walks the blocks of
    every row on the CPU, std:1664-1742 -- thousands of tiny tensor operations for a long stream): every step below is
    a tensor operation over all blocks / all tokens of a row ON THE DEVICE of `input_ids`, with one shape-dependent
    host synchronisation per row.  Same integer arithmetic as the reference (so the same bits); returns None for rows
    that are not well formed (a vision-start token not followed by exactly t*h*w placeholders), which then take the
    block-by-block path."""
    dev = input_ids.device
    merge = int(config.vision_config.spatial_merge_size)
    tps = config.vision_config.tokens_per_second
    IMG, VID, VS = config.image_token_id, config.video_token_id, config.vision_start_token_id
    B, T = input_ids.shape
    as_grid = lambda g: (torch.zeros(1, 3, dtype=torch.long, device=dev) if g is None or g.numel() == 0
                         else g.to(dev, torch.long).reshape(-1, 3))
    img_g, vid_g = as_grid(image_grid_thw), as_grid(video_grid_thw)
    spg = None if second_per_grid_ts is None else second_per_grid_ts.to(dev).reshape(-1)
    out = torch.ones(3, B, T, dtype=input_ids.dtype, device=dev)
    deltas = torch.zeros(B, dtype=torch.long, device=dev)
    n_img = n_vid = 0
    for b in range(B):
        keep = (attention_mask[b] == 1).to(dev) if attention_mask is not None else None
        row = input_ids[b][keep] if keep is not None else input_ids[b]
        L = row.numel()
        if L < 2:
            return None
        is_start = (row[:-1] == VS) & ((row[1:] == IMG) | (row[1:] == VID))
        s = torch.nonzero(is_start).squeeze(1)
        nb = int(s.numel())
        if nb == 0:
            llm = torch.arange(L, device=dev).view(1, -1).expand(3, -1)
        else:
            kind = row[s + 1]
            is_img = kind == IMG
            img_rank = torch.cumsum(is_img.long(), 0) - 1
            vid_rank = torch.cumsum((~is_img).long(), 0) - 1
            ni, nv = int(is_img.sum()), nb - int(is_img.sum())
            if n_img + ni > (0 if image_grid_thw is None else img_g.shape[0]) or \
               n_vid + nv > (0 if video_grid_thw is None else vid_g.shape[0]):
                return None
            thw = torch.where(is_img[:, None], img_g[(n_img + img_rank).clamp(0, img_g.shape[0] - 1)],
                              vid_g[(n_vid + vid_rank).clamp(0, vid_g.shape[0] - 1)])
            if spg is not None and nv:
                sec = torch.where(is_img, torch.zeros((), device=dev, dtype=spg.dtype),
                                  spg[(n_vid + vid_rank).clamp(0, spg.numel() - 1)])
            else:
                sec = torch.where(is_img, 0.0, 1.0).to(dev)
            sec_i = sec.to(torch.long)                          # the reference multiplies in the integer dtype of arange
            t, gh, gw = thw[:, 0], thw[:, 1] // merge, thw[:, 2] // merge
            n = t * gh * gw
            ed = s + 1
            st = torch.cat([ed.new_zeros(1), (ed + n)[:-1]])
            text_len = ed - st
            t_max = ((t - 1) * sec_i * tps).long()
            span = torch.maximum(torch.maximum(t_max, gh - 1), gw - 1) + 1
            base = torch.cumsum(text_len + span, 0) - span      # first coordinate of block b's grid
            nxt = base - text_len                               # first position of the text in front of block b
            tail_st, tail_nxt = ed[-1] + n[-1], base[-1] + span[-1]
            p = torch.arange(L, device=dev)
            bnd = torch.cat([torch.stack([st, ed], 1).reshape(-1), tail_st.view(1)])
            idx = torch.bucketize(p, bnd, right=True) - 1       # 2b: text of block b, 2b + 1: its placeholders, 2nb: tail
            blk = (idx // 2).clamp(max=nb - 1)
            vis = (idx % 2 == 1) & (idx < 2 * nb)
            tail = idx == 2 * nb
            # well-formedness (one host sync): runs in order, exactly the placeholders where the grids say, none elsewhere
            ok = (text_len >= 0).all() & (tail_st <= L) & (n > 0).all() & \
                 (((row == IMG) | (row == VID)) == vis).all() & (row[vis] == kind[blk[vis]]).all()
            if not bool(ok):
                return None
            text_pos = torch.where(tail, tail_nxt + (p - tail_st), nxt[blk] + (p - st[blk]))
            kk = (p - ed[blk]).clamp(min=0)
            ghb, gwb = gh[blk].clamp(min=1), gw[blk].clamp(min=1)
            t_idx = ((kk // (ghb * gwb)) * sec_i[blk] * tps).long()
            h_idx, w_idx = (kk // gwb) % ghb, kk % gwb
            grid = torch.stack([t_idx, h_idx, w_idx]) + base[blk]
            llm = torch.where(vis[None], grid, text_pos[None].expand(3, -1))
            n_img, n_vid = n_img + ni, n_vid + nv
        if keep is not None:
            out[:, b, keep] = llm.to(out.dtype)
        else:
            out[:, b, :] = llm.to(out.dtype)
        deltas[b] = llm.max() + 1 - T
    return out, deltas.unsqueeze(1)


def get_rope_index(config, input_ids=None, image_grid_thw=None, video_grid_thw=None, second_per_grid_ts=None,
                   attention_mask=None):
    """M-RoPE position ids [3, B, T] and per-row deltas [B, 1] for text with image / video placeholders
    (InfiniteVLModel.get_rope_index, std:1623-1758; index arithmetic: bit-exact).  Text tokens advance all three rows
    together; the placeholders of a vision block get (t * seconds_per_grid * tokens_per_second, h, w) grid
    coordinates offset by the position the block starts at; the next text token continues after the block's
    largest coordinate."""
    merge = int(config.vision_config.spatial_merge_size)
    tps = config.vision_config.tokens_per_second
    if input_ids is None or (image_grid_thw is None and video_grid_thw is None):
        if attention_mask is not None:
            pos = attention_mask.long().cumsum(-1) - 1
            pos.masked_fill_(attention_mask == 0, 1)
            pos = pos.unsqueeze(0).expand(3, -1, -1).to(attention_mask.device)
            mx = pos.max(0, keepdim=False)[0].max(-1, keepdim=True)[0]
            return pos, mx + 1 - attention_mask.shape[-1]
        B, T = input_ids.shape
        pos = torch.arange(T, device=input_ids.device).view(1, 1, -1).expand(3, B, -1)
        return pos, torch.zeros([B, 1], device=input_ids.device, dtype=input_ids.dtype)
    B, T = input_ids.shape
    fast = _rope_index_vectorized(config, input_ids, image_grid_thw, video_grid_thw, second_per_grid_ts, attention_mask)
    if fast is not None:
        return fast
    ids_cpu = input_ids.cpu()
    mask_cpu = (attention_mask == 1).cpu() if attention_mask is not None else None
    img = image_grid_thw.cpu().tolist() if image_grid_thw is not None else []
    vid = video_grid_thw.cpu().tolist() if video_grid_thw is not None else []
    spg = second_per_grid_ts.cpu().tolist() if second_per_grid_ts is not None else None
    out = torch.ones(3, B, T, dtype=input_ids.dtype)
    deltas = []
    n_img = n_vid = 0
    for b in range(B):
        row = ids_cpu[b][mask_cpu[b]] if mask_cpu is not None else ids_cpu[b]
        L = row.numel()
        starts = torch.nonzero(row == config.vision_start_token_id).squeeze(1)
        starts = starts[starts + 1 < L]
        kinds = row[starts + 1]
        blocks = [(int(s), int(kd)) for s, kd in zip(starts.tolist(), kinds.tolist())
                  if kd in (config.image_token_id, config.video_token_id)]
        pieces, st, nxt = [], 0, 0
        for _, kind in blocks:
            tok = row[st:]
            hit = torch.nonzero(tok == kind)
            ed = st + int(hit[0]) if hit.numel() else L + 1      # first placeholder of this block at or after st
            if kind == config.image_token_id:
                t, hh, ww = img[n_img]; n_img += 1
                sec = 0
            else:
                t, hh, ww = vid[n_vid]
                sec = spg[n_vid] if spg is not None else 1.0
                n_vid += 1
            gh, gw = hh // merge, ww // merge
            text_len = ed - st
            pieces.append(torch.arange(text_len).view(1, -1).expand(3, -1) + nxt)
            base = text_len + nxt
            # the time coordinate goes through the same integer-tensor arithmetic as the reference
            # (arange (int64) * seconds * tokens_per_second, then .long())
            rng = torch.arange(t).view(-1, 1).expand(-1, gh * gw)
            t_idx = (rng * torch.as_tensor(sec, dtype=rng.dtype) * tps).long().flatten()
            h_idx = torch.arange(gh).view(1, -1, 1).expand(t, -1, gw).flatten()
            w_idx = torch.arange(gw).view(1, 1, -1).expand(t, gh, -1).flatten()
            grid = torch.stack([t_idx, h_idx, w_idx]) + base
            pieces.append(grid)
            nxt = int(grid.max()) + 1 if grid.numel() else (int(pieces[-2].max()) + 1 if text_len else nxt)
            st = ed + t * gh * gw
        if st < L:
            pieces.append(torch.arange(L - st).view(1, -1).expand(3, -1) + nxt)
        llm = torch.cat(pieces, dim=1).reshape(3, -1)
        if mask_cpu is not None:
            out[:, b, mask_cpu[b]] = llm.to(out.dtype)
        else:
            out[:, b, :] = llm.to(out.dtype)
        deltas.append(int(llm.max()) + 1 - T)
    return out.to(input_ids.device), torch.tensor(deltas).unsqueeze(1).to(input_ids.device)


class InfiniteVLTextModel(nn.Module):
    """Embeddings + decoder stack + final norm (std:1430-1591)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.padding_idx = getattr(config, "pad_token_id", None)
        self.vocab_size = getattr(config, "vocab_size", 151936)
        self.embed_tokens = nn.Embedding(self.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([InfiniteVLDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = InfiniteVLRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.rotary_emb = InfiniteVLRotaryEmbedding(config=config)
        self.has_sliding_layers = "sliding_attention" in config.layer_types

    def get_input_embeddings(self):
        return self.embed_tokens

    def set_input_embeddings(self, value):
        self.embed_tokens = value

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                cache_position=None, **kwargs):
        use_cache = use_cache if use_cache is not None else getattr(self.config, "use_cache", False)
        return_dict = True if return_dict is None else return_dict
        if (input_ids is None) ^ (inputs_embeds is not None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        if output_attentions:
            raise NotImplementedError("attention weights are never materialised by the B200 kernels")
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)
        # allocate the static cache on the first forward pass (std:1488-1501)
        if use_cache and (past_key_values is None or not isinstance(past_key_values, StaticCachePrealloc)):
            past_key_values = StaticCachePrealloc(config=self.config, batch_size=inputs_embeds.shape[0],
                                                  dtype=inputs_embeds.dtype, device=inputs_embeds.device)
        B, T, _ = inputs_embeds.shape
        if cache_position is None:
            past = past_key_values.get_seq_length() if past_key_values is not None else 0
            cache_position = torch.arange(past, past + T, device=inputs_embeds.device)
        position_ids, text_position_ids = normalize_position_ids(position_ids, cache_position.reshape(-1), B)
        cu_seqlens = packed_cu_seqlens(text_position_ids)
        extra = {}
        if cu_seqlens is not None:
            # packed training rows: every mixer treats the sequences separately (varlen operators); no cache
            if past_key_values is not None and past_key_values.get_seq_length() > 0:
                raise ValueError("packed sequences cannot continue a cache")
            past_key_values, use_cache = None, False
            extra["cu_seqlens"] = cu_seqlens
        # The reference builds its mask through create_causal_mask; on its FA2 path that is the 2-D padding mask when
        # something is padded and None otherwise (causality and the window live in the kernel).  Same here: an
        # all-ones mask is dropped, a real padding mask travels to the attention operator, which refuses it.
        if isinstance(attention_mask, dict):
            attention_mask = attention_mask.get("full_attention")
        if attention_mask is not None and attention_mask.dim() == 2 and bool(attention_mask.all()):
            attention_mask = None
        cos, sin = self.rotary_emb(inputs_embeds, position_ids)
        cos, sin = mrope_select(cos, sin, self.config.rope_scaling["mrope_section"])
        h = inputs_embeds
        all_hidden = () if output_hidden_states else None
        for layer in self.layers:
            if output_hidden_states:
                all_hidden += (h,)
            h = layer(h, attention_mask=attention_mask, position_ids=text_position_ids,
                      past_key_values=past_key_values, use_cache=use_cache, cache_position=cache_position,
                      position_embeddings=(cos, sin), **extra)[0]
        h = self.norm(h)
        if output_hidden_states:
            all_hidden += (h,)
        out = _Output(last_hidden_state=h, past_key_values=past_key_values if use_cache else None,
                      hidden_states=all_hidden, attentions=None)
        return out if return_dict else out.to_tuple()


class InfiniteVLModel(nn.Module):
    """language_model (+ an optional caller-supplied `visual`), M-RoPE position bookkeeping (std:1595-1936)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.visual = None           # vision tower: out of the hot path; attach a module with the reference's interface
        self.language_model = InfiniteVLTextModel(config.text_config)
        self.rope_deltas = None

    def get_input_embeddings(self):
        return self.language_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.language_model.set_input_embeddings(value)

    def get_decoder(self):
        return self.language_model

    def set_decoder(self, decoder):
        self.language_model = decoder

    def get_rope_index(self, input_ids=None, image_grid_thw=None, video_grid_thw=None, second_per_grid_ts=None,
                       attention_mask=None):
        return get_rope_index(self.config, input_ids, image_grid_thw, video_grid_thw, second_per_grid_ts, attention_mask)

    def _vision_features(self, pixels, grid_thw):
        if self.visual is None:
            raise NotImplementedError("no vision tower attached (model.visual is None): pass inputs_embeds with the "
                                      "image features already scattered in, or attach the reference's vision module")
        emb = self.visual(pixels.type(next(self.visual.parameters()).dtype), grid_thw=grid_thw)
        return emb

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None, pixel_values=None,
                pixel_values_videos=None, image_grid_thw=None, video_grid_thw=None, rope_deltas=None,
                cache_position=None, second_per_grid_ts=None, **kwargs):
        return_dict = True if return_dict is None else return_dict
        if inputs_embeds is None:
            inputs_embeds = self.get_input_embeddings()(input_ids)
        for pixels, grid, token in ((pixel_values, image_grid_thw, self.config.image_token_id),
                                    (pixel_values_videos, video_grid_thw, self.config.video_token_id)):
            if pixels is None:
                continue
            feats = self._vision_features(pixels, grid).to(inputs_embeds.device, inputs_embeds.dtype)
            if input_ids is None:
                raise NotImplementedError("scattering vision features needs input_ids to find the placeholders")
            mask = (input_ids == token).unsqueeze(-1).expand_as(inputs_embeds)
            if inputs_embeds[mask].numel() != feats.numel():
                raise ValueError(f"Image features and image tokens do not match: tokens: {int((input_ids == token).sum())}, "
                                 f"features {feats.shape[0]}")
            inputs_embeds = inputs_embeds.masked_scatter(mask, feats)
        if position_ids is None:
            prefill = (cache_position is not None and int(cache_position.reshape(-1)[0]) == 0) or \
                      (past_key_values is None or past_key_values.get_seq_length() == 0)
            if (prefill or self.rope_deltas is None) and input_ids is None and attention_mask is None:
                # embeddings only (the reference needs input_ids here): plain text positions
                B, T, _ = inputs_embeds.shape
                position_ids = torch.arange(T, device=inputs_embeds.device).view(1, 1, -1).expand(3, B, -1)
                self.rope_deltas = torch.zeros(B, 1, dtype=torch.long, device=inputs_embeds.device)
            elif prefill or self.rope_deltas is None:
                position_ids, rope_deltas = self.get_rope_index(input_ids, image_grid_thw, video_grid_thw,
                                                                second_per_grid_ts=second_per_grid_ts,
                                                                attention_mask=attention_mask)
                self.rope_deltas = rope_deltas
            else:
                B, T, _ = inputs_embeds.shape
                position_ids = torch.arange(T, device=inputs_embeds.device).view(1, 1, -1).expand(3, B, -1)
                if cache_position is not None:
                    delta = (cache_position.reshape(-1)[0] + self.rope_deltas).to(inputs_embeds.device)
                else:
                    delta = torch.zeros((B, T), device=inputs_embeds.device)
                delta = delta.repeat_interleave(B // delta.shape[0], dim=1)
                position_ids = position_ids + delta.to(position_ids.device)
        out = self.language_model(input_ids=None, position_ids=position_ids, attention_mask=attention_mask,
                                  past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                                  output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                                  return_dict=True, cache_position=cache_position, **kwargs)
        res = _Output(last_hidden_state=out.last_hidden_state, past_key_values=out.past_key_values,
                      hidden_states=out.hidden_states, attentions=None, rope_deltas=self.rope_deltas)
        return res if return_dict else res.to_tuple()


class InfiniteVLQwen2_5_VLForConditionalGeneration(nn.Module):
    """model + lm_head (std:1980-2322).  forward(...).logits, allocate_inference_cache(batch_size), a greedy
    generate(); `from_pretrained(dir)` loads the language-model weights of a checkpoint directory (config.json +
    *.safetensors), ignoring the vision tower's."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = InfiniteVLModel(config)
        tc = config.text_config
        self.lm_head = nn.Linear(tc.hidden_size, getattr(tc, "vocab_size", 151936), bias=False)
        if getattr(config, "tie_word_embeddings", True):
            self.lm_head.weight = self.model.language_model.embed_tokens.weight

    @property
    def language_model(self):
        return self.model.language_model

    @property
    def visual(self):
        return self.model.visual

    def get_input_embeddings(self):
        return self.model.get_input_embeddings()

    def get_decoder(self):
        return self.model.get_decoder()

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def allocate_inference_cache(self, batch_size, **kw):
        """std:2316-2322 (extension: keyword arguments of StaticCachePrealloc, e.g. state_dtype=torch.float32)."""
        return StaticCachePrealloc(config=self.config.text_config, batch_size=batch_size, dtype=self.dtype,
                                   device=self.device, **kw)

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, pixel_values=None,
                pixel_values_videos=None, image_grid_thw=None, video_grid_thw=None, rope_deltas=None,
                cache_position=None, second_per_grid_ts=None, logits_to_keep=0, return_dict=True, **kwargs):
        out = self.model(input_ids=input_ids, pixel_values=pixel_values, pixel_values_videos=pixel_values_videos,
                         image_grid_thw=image_grid_thw, video_grid_thw=video_grid_thw,
                         second_per_grid_ts=second_per_grid_ts, position_ids=position_ids,
                         attention_mask=attention_mask, past_key_values=past_key_values, inputs_embeds=inputs_embeds,
                         use_cache=use_cache, output_attentions=output_attentions,
                         output_hidden_states=output_hidden_states, return_dict=True, cache_position=cache_position,
                         **kwargs)
        h = out.last_hidden_state
        # logits_to_keep = 0 keeps every position (slice(-0, None)), as in the reference (std:2091-2092)
        idx = slice(-logits_to_keep, None) if isinstance(logits_to_keep, int) else logits_to_keep
        logits = self.lm_head(h[:, idx, :])
        loss = None
        if labels is not None:
            lg = logits.float()
            loss = F.cross_entropy(lg[:, :-1].reshape(-1, lg.shape[-1]), labels[:, 1:].reshape(-1), ignore_index=-100)
        res = _Output(loss=loss, logits=logits, past_key_values=out.past_key_values, hidden_states=out.hidden_states,
                      attentions=None, rope_deltas=out.rope_deltas)
        return res if return_dict else res.to_tuple()

    @torch.no_grad()
    def generate(self, input_ids=None, inputs_embeds=None, position_ids=None, max_new_tokens=32, eos_token_id=None,
                 past_key_values=None, use_cuda_graph=False, **kwargs):
        """Greedy decoding: one prefill call, then single-token steps through the static cache (what the demo's QA
        branch does by hand, demo_streaming_inference.py:389-421).  With use_cuda_graph the decode step is captured
        once and replayed (ring caches keep their counters on the device, so the replay is exact)."""
        dev = self.device
        B = (input_ids if input_ids is not None else inputs_embeds).shape[0]
        cache = past_key_values if past_key_values is not None else self.allocate_inference_cache(B)
        past = cache.get_seq_length()
        T = (input_ids if input_ids is not None else inputs_embeds).shape[1]
        cp = torch.arange(past, past + T, device=dev)
        out = self(input_ids=input_ids, inputs_embeds=inputs_embeds, position_ids=position_ids, past_key_values=cache,
                   use_cache=True, cache_position=cp, logits_to_keep=1, **kwargs)
        nxt = out.logits[:, -1].argmax(-1, keepdim=True)
        pos_next = (position_ids.max() + 1 if position_ids is not None else torch.tensor(past + T, device=dev)).reshape(())
        tokens = [nxt]
        static_tok = nxt.clone()
        static_cp = torch.zeros(1, dtype=torch.long, device=dev)
        static_pos = torch.zeros(3, B, 1, dtype=torch.long, device=dev)

        def step():
            return self(input_ids=static_tok, position_ids=static_pos, past_key_values=cache, use_cache=True,
                        cache_position=static_cp, logits_to_keep=1).logits[:, -1].argmax(-1, keepdim=True)

        graph = None
        finished = torch.zeros(B, dtype=torch.bool, device=dev)
        for i in range(1, max_new_tokens):
            if eos_token_id is not None:
                finished |= nxt.view(-1) == eos_token_id
                if bool(finished.all()):
                    break
            static_tok.copy_(nxt)
            static_cp.fill_(past + T + i - 1)
            static_pos.copy_((pos_next + (i - 1)).expand(3, B, 1))
            if use_cuda_graph and graph is None and i >= 2:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        static_out = step()
                torch.cuda.current_stream().wait_stream(side)
                graph.replay()
                nxt = static_out.clone()
            elif graph is not None:
                graph.replay()
                nxt = static_out.clone()
            else:
                nxt = step()
            tokens.append(nxt)
        return torch.cat(tokens, dim=1)

    # -- checkpoints ---------------------------------------------------------------------------------------------
    def hf_state_dict(self):
        """State dict under the checkpoint's names (model.language_model.*, lm_head.weight); the tied lm_head is
        omitted as in the published checkpoint."""
        sd = {k_: v_ for k_, v_ in self.state_dict().items() if not k_.startswith("model.visual.")}
        if getattr(self.config, "tie_word_embeddings", True):
            sd.pop("lm_head.weight", None)
        return sd

    @classmethod
    def from_pretrained(cls, path, torch_dtype=torch.bfloat16, device="cuda", **kwargs):
        import glob
        import os as _os
        cfg = InfiniteVLConfig.from_json(_os.path.join(path, "config.json"))
        model = cls(cfg)
        files = sorted(glob.glob(_os.path.join(path, "*.safetensors")))
        if not files:
            raise FileNotFoundError(f"no *.safetensors under {path}")
        from safetensors.torch import load_file
        sd = {}
        for f in files:
            sd.update(load_file(f))
        # old checkpoints store the text model directly under `model.` (std:1597 _checkpoint_conversion_mapping)
        fixed = {}
        for k_, v_ in sd.items():
            if k_.startswith("visual.") or k_.startswith("model.visual."):
                continue
            if k_.startswith("model.") and not k_.startswith("model.language_model."):
                k_ = "model.language_model." + k_[len("model."):]
            fixed[k_] = v_
        missing, unexpected = model.load_state_dict(fixed, strict=False)
        missing = [m for m in missing if not (m == "lm_head.weight" and cfg.tie_word_embeddings)
                   and not m.endswith("rotary_emb.inv_freq")]
        if missing or unexpected:
            raise RuntimeError(f"checkpoint does not match the model: missing {missing[:5]}, unexpected {unexpected[:5]}")
        return model.to(device=device, dtype=torch_dtype).eval()
