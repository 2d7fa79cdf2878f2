"""Inference caches with the reference's class names, attribute names and index semantics
(infinitevl_standard/modeling_infinitevl.py:66-443), so that callers which poke at the buffers
directly -- inference_examples/demo_streaming_inference.py:111-160 clones `_buf_keys`,
`_buf_values`, `keys`, `values`, `size`, `cumulative_length`, `capacity`, `recurrent_state`,
`conv_state_{q,k,v}`, `seq_len`, `start` -- keep working unchanged.

Nothing here allocates after construction: update() only copies into the pre-allocated
buffers, which is what CUDA-graph capture of the forward needs.
"""
from __future__ import annotations

import ctypes
from typing import Any, Optional, Tuple

import torch

try:  # the reference's classes derive from these (std:66,229,366): generate() and friends check isinstance
    from transformers.cache_utils import Cache as _HFCache, CacheLayerMixin as _HFCacheLayer
except Exception:  # noqa: BLE001  (transformers absent or too old: plain classes with the same protocol)
    _HFCache, _HFCacheLayer = object, object


def _get_decoder_cfg(config):
    if hasattr(config, "get_text_config"):
        try:
            return config.get_text_config(decoder=True)
        except TypeError:
            return config.get_text_config()
    return config


class StaticSlidingWindowLayerPrealloc(_HFCacheLayer):
    """Last (sliding_window - 1) keys/values of one SWA layer (std:66-227)."""
    is_sliding = True

    def __new__(cls, *args, ring: Optional[bool] = None, **kw):
        # ring=True (the default for bf16 caches on a CUDA device) selects the ring-buffer layer below: same
        # attributes and integer semantics, no O(W) roll, device-side token counter
        if cls is StaticSlidingWindowLayerPrealloc:
            if ring is None:
                dev = torch.device(kw.get("device", "cpu"))
                ring = dev.type == "cuda" and kw.get("dtype", torch.float32) == torch.bfloat16
            if ring:
                return object.__new__(RingSlidingWindowLayer)
        return object.__new__(cls)

    def __init__(self, *, config, batch_size: int, device="cpu", dtype=torch.float32, zero_init: bool = False,
                 ring: Optional[bool] = None, max_append: int = 256):
        cfg = _get_decoder_cfg(config)
        num_kv_heads = int(getattr(cfg, "num_key_value_heads", getattr(cfg, "num_attention_heads")))
        head_dim = int(getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads)
        W = (getattr(cfg, "sliding_window", None) or getattr(cfg, "attention_chunk_size", None)
             or int(getattr(cfg, "max_position_embeddings")))
        if W is None or int(W) <= 0:
            raise ValueError("SWA requires valid sliding_window / attention_chunk_size / max_position_embeddings")
        self.sliding_window = int(W)
        self.capacity = max(self.sliding_window - 1, 0)
        self.is_initialized = True
        self.dtype, self.device = dtype, device
        self.batch_size, self.num_kv_heads, self.head_dim = int(batch_size), num_kv_heads, head_dim
        self.num_heads = int(getattr(cfg, "num_attention_heads", num_kv_heads))
        self._allocate(zero_init, int(max_append))

    def _allocate(self, zero_init: bool, max_append: int) -> None:
        dtype, device, num_kv_heads, head_dim = self.dtype, self.device, self.num_kv_heads, self.head_dim
        self.size = 0
        self.cumulative_length = 0
        alloc = torch.zeros if zero_init else torch.empty
        if self.capacity > 0:
            shape = (self.batch_size, num_kv_heads, self.capacity, head_dim)
            self._buf_keys = alloc(shape, dtype=dtype, device=device)
            self._buf_values = alloc(shape, dtype=dtype, device=device)
            self.keys = self._buf_keys[:, :, :0, :]
            self.values = self._buf_values[:, :, :0, :]
        else:
            empty = torch.empty((self.batch_size, num_kv_heads, 0, head_dim), dtype=dtype, device=device)
            self._buf_keys = self._buf_values = None
            self.keys = self.values = empty

    def update(self, key_states, value_states, conv_state=None, recurrent_state=None,
               cache_kwargs: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns [previous tail ; new] and then keeps the newest <= capacity rows (std:126-173)."""
        assert key_states.shape == value_states.shape, "K/V shapes must match"
        B, H, Tq, D = key_states.shape
        if B != self.batch_size:
            raise ValueError(f"SWA pre-allocated batch_size={self.batch_size}, but got B={B}")
        if H != self.num_kv_heads or D != self.head_dim:
            raise ValueError(
                f"SWA head dim mismatch: got H={H},D={D}, expect H={self.num_kv_heads},D={self.head_dim}")
        full_k = torch.cat([self.keys, key_states], dim=-2)
        full_v = torch.cat([self.values, value_states], dim=-2)
        new_size = min(self.capacity, self.size + Tq)
        if self.capacity > 0 and new_size > 0:
            total = full_k.shape[-2]
            self._buf_keys[:, :, :new_size, :].copy_(full_k[:, :, total - new_size:, :])
            self._buf_values[:, :, :new_size, :].copy_(full_v[:, :, total - new_size:, :])
            self.keys = self._buf_keys[:, :, :new_size, :]
            self.values = self._buf_values[:, :, :new_size, :]
        self.size = int(new_size)
        self.cumulative_length += Tq
        return full_k, full_v

    def get_mask_sizes(self, cache_position: torch.Tensor) -> Tuple[int, int]:
        q_len = int(cache_position.shape[0])
        pre_cum = max(int(self.cumulative_length) - q_len, 0)
        kv_offset = max(pre_cum - self.sliding_window + 1, 0)
        kv_len = (self.sliding_window - 1 if pre_cum >= self.sliding_window else pre_cum) + q_len
        return kv_len, kv_offset

    def get_seq_length(self) -> int:
        return int(self.cumulative_length)

    def get_max_cache_shape(self) -> int:
        return int(self.sliding_window)

    def crop(self, max_length: int) -> None:
        if self.get_seq_length() >= self.sliding_window:
            raise ValueError("Cropping is forbidden after filling SWA window (to avoid state loss)")
        new_size = max(0, self.size - abs(max_length)) if max_length < 0 else min(self.size, max_length)
        if self.capacity > 0:
            if new_size > 0:
                self._buf_keys[:, :, :new_size, :].copy_(self._buf_keys[:, :, self.size - new_size:self.size, :].clone())
                self._buf_values[:, :, :new_size, :].copy_(
                    self._buf_values[:, :, self.size - new_size:self.size, :].clone())
            self.keys = self._buf_keys[:, :, :new_size, :]
            self.values = self._buf_values[:, :, :new_size, :]
        self.size = int(new_size)
        self.cumulative_length = int(self.size)

    def batch_repeat_interleave(self, repeats: int) -> None:
        if repeats != 1:
            raise RuntimeError("Static cache forbids changing batch size (repeat_interleave)")

    def batch_select_indices(self, indices: torch.Tensor) -> None:
        if indices.numel() != self.batch_size:
            raise RuntimeError("Static cache forbids changing batch size (select_indices)")

    def lazy_initialization(self, *args, **kwargs):
        return

    def reset(self) -> None:
        self.size = 0
        self.cumulative_length = 0
        if self.capacity > 0:
            self.keys = self._buf_keys[:, :, :0, :]
            self.values = self._buf_values[:, :, :0, :]

    # -- snapshots -------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"keys": self.keys.contiguous().clone(), "values": self.values.contiguous().clone(),
                "size": torch.tensor(int(self.size)), "cumulative_length": torch.tensor(int(self.cumulative_length))}

    def load_state_dict(self, sd: dict) -> None:
        n = int(sd["size"])
        if n > 0:
            self._buf_keys[:, :, :n, :].copy_(sd["keys"])
            self._buf_values[:, :, :n, :].copy_(sd["values"])
        self.keys, self.values = self._buf_keys[:, :, :n, :], self._buf_values[:, :, :n, :]
        self.size, self.cumulative_length = n, int(sd["cumulative_length"])


class RingSlidingWindowLayer(StaticSlidingWindowLayerPrealloc):
    """Ring-buffer form of the sliding-window cache layer.

    Same constructor, attributes (`keys`, `values`, `_buf_keys`, `_buf_values`, `size`, `cumulative_length`,
    `capacity`, ...) and integer semantics as the reference class (std:66-227), so `clone_inference_cache` of
    inference_examples/demo_streaming_inference.py:111-160 works on it unchanged -- but update() copies only the new
    tokens (the reference concatenates and re-copies the whole 8191-token window: 8.4 MB per layer and call), and
    the token counter that the kernels address by lives in device memory (`_state[0]`), so a captured CUDA graph of
    a decode step or of a streamed frame replays correctly (the reference's Python counters freeze under replay).

    Storage `_ring_k/_ring_v` [B, 2R, Hkv, D], R = capacity + max_append: token t is written to slots t % R and
    t % R + R, which makes the last n <= R tokens one contiguous run starting at (cum - n) % R.  `keys` / `values` /
    `_buf_*` are views of that run in the reference's [B, Hkv, n, D] shape.  The Python integers are kept in step for
    every eager call; writes to them from outside (the demo's clone) re-anchor the ring on the next call.
    `attend()` is the fast path InfiniteVLSelfAttention uses: append + attention with no concatenation, one launch
    for a decode step (ivl_swa_ring_decode), two for a frame of <= max_append tokens."""
    is_ring = True

    def _allocate(self, zero_init: bool, max_append: int) -> None:
        if self.capacity <= 0:
            raise ValueError("ring cache needs sliding_window >= 2")
        self.max_append = max(int(max_append), 1)
        self.R = self.capacity + self.max_append
        shape = (self.batch_size, 2 * self.R, self.num_kv_heads, self.head_dim)
        # always zero-initialised: the attention kernels fetch whole 64-key tiles of the ring, and a masked key
        # contributes 0 * v -- which must not meet a NaN bit pattern left in never-written slots
        self._ring_k = torch.zeros(shape, dtype=self.dtype, device=self.device)
        self._ring_v = torch.zeros(shape, dtype=self.dtype, device=self.device)
        self._state = torch.zeros(2 + self.batch_size * self.num_kv_heads, dtype=torch.int32, device=self.device)
        self._size = 0
        self._cum = 0
        self._pend_size = self._pend_cum = None
        self._dirty = False
        self._ws = None

    # -- the reference's attributes as views / properties --------------------------------------------------------
    def _start(self) -> int:
        return (self._cum - self._size) % self.R

    def _window(self, ring: torch.Tensor, n: int) -> torch.Tensor:
        s = self._start()
        return ring[:, s:s + n].permute(0, 2, 1, 3)          # [B, Hkv, n, D] view

    # Writes to `size` / `cumulative_length` from outside (the demo's clone protocol, the sharded hand-off) are held
    # as PENDING values until the next call re-anchors the ring: until then the `_buf_*` / `keys` views keep pointing
    # at the rows the outside writer has filled (rows [0, n) of the window as it was anchored when it wrote), exactly
    # as the reference's fixed `_buf_keys[:, :, :n]` does.
    size = property(lambda self: self._size if self._pend_size is None else self._pend_size)
    cumulative_length = property(lambda self: self._cum if self._pend_cum is None else self._pend_cum)

    @size.setter
    def size(self, v):
        self._pend_size, self._dirty = int(v), True

    @cumulative_length.setter
    def cumulative_length(self, v):
        self._pend_cum, self._dirty = int(v), True

    keys = property(lambda self: self._window(self._ring_k, self.size))
    values = property(lambda self: self._window(self._ring_v, self.size))
    _buf_keys = property(lambda self: self._window(self._ring_k, self.capacity))
    _buf_values = property(lambda self: self._window(self._ring_v, self.capacity))

    @keys.setter
    def keys(self, v):          # the demo re-assigns the views after copying into _buf_*: nothing to do
        pass

    @values.setter
    def values(self, v):
        pass

    def _resync(self) -> None:
        """After size / cumulative_length (and the window contents, through the views) were written from outside:
        rewrite both ring copies of the window and the device counter."""
        n, cum = self.size, self.cumulative_length          # the values written from outside (or the current ones)
        k = self._window(self._ring_k, n).clone()           # rows [0, n) of the window as anchored BEFORE the writes
        v = self._window(self._ring_v, n).clone()
        self._pend_size = self._pend_cum = None
        self._state.zero_()
        self._state[0] = cum - n
        self._dirty = False
        self._cum = cum - n
        self._size = 0
        if n > 0:
            self._append(k.transpose(1, 2), v.transpose(1, 2))

    # -- appending ---------------------------------------------------------------------------------------------------
    def _append(self, k_bthd: torch.Tensor, v_bthd: torch.Tensor) -> None:
        """k, v [B, Tq, Hkv, D] (any batch / time / head strides)."""
        Tq = k_bthd.shape[1]
        if k_bthd.is_cuda:
            from . import _lib
            fix = lambda t: t if (t.stride(3) == 1 and all(x % 8 == 0 for x in t.stride()[:3]) and t.data_ptr() % 16 == 0
                                  and t.dtype == torch.bfloat16) else t.to(torch.bfloat16).contiguous()
            k_bthd, v_bthd = fix(k_bthd), fix(v_bthd)
            st = lambda t: (ctypes.c_int64 * 3)(t.stride(0), t.stride(1), t.stride(2))
            code = _lib.load().ivl_swa_ring_append(
                k_bthd.data_ptr(), st(k_bthd), v_bthd.data_ptr(), st(v_bthd), self._ring_k.data_ptr(),
                self._ring_v.data_ptr(), self._state.data_ptr(), self.batch_size, Tq, self.num_kv_heads, self.head_dim,
                self.R, torch.cuda.current_stream(k_bthd.device).cuda_stream)
            _lib.check(code, "ivl_swa_ring_append")
        else:   # host-resident cache (the reference supports device="cpu"): plain index arithmetic
            first = max(0, Tq - self.R)
            slots = (self._cum + torch.arange(first, Tq)) % self.R
            for ring, src in ((self._ring_k, k_bthd), (self._ring_v, v_bthd)):
                ring[:, slots] = src[:, first:].to(ring.dtype)
                ring[:, slots + self.R] = src[:, first:].to(ring.dtype)
            self._state[0] += Tq
        self._size = min(self.capacity, self._size + Tq)
        self._cum += Tq

    def update(self, key_states, value_states, conv_state=None, recurrent_state=None,
               cache_kwargs: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """The reference's protocol: returns [previous tail ; new] (std:126-173).  Only the new tokens are copied."""
        assert key_states.shape == value_states.shape, "K/V shapes must match"
        B, H, Tq, D = key_states.shape
        if B != self.batch_size:
            raise ValueError(f"SWA pre-allocated batch_size={self.batch_size}, but got B={B}")
        if H != self.num_kv_heads or D != self.head_dim:
            raise ValueError(
                f"SWA head dim mismatch: got H={H},D={D}, expect H={self.num_kv_heads},D={self.head_dim}")
        if self._dirty:
            self._resync()
        full_k = torch.cat([self.keys, key_states], dim=-2)
        full_v = torch.cat([self.values, value_states], dim=-2)
        self._append(key_states.transpose(1, 2), value_states.transpose(1, 2))
        return full_k, full_v

    # -- fast path: append + attention, no concatenation ---------------------------------------------------------------
    def attend(self, q_bthd: torch.Tensor, k_bthd: torch.Tensor, v_bthd: torch.Tensor, scale: float,
               window: Optional[int]) -> torch.Tensor:
        """q [B,Tq,Hq,D], k/v [B,Tq,Hkv,D] of the new tokens (bf16, CUDA) -> attention output [B,Tq,Hq,D]; the new
        keys/values are appended.  Graph-safe for Tq <= max_append (all addressing comes from the device counter)."""
        from . import _lib, swa
        B, Tq, Hq, D = q_bthd.shape
        if self._dirty:
            self._resync()
        W = self.sliding_window
        if window is None or int(window) != W or Tq > self.max_append or Hq // self.num_kv_heads > 8 or W > 8192:
            # general case (long prefill into a cache, no window): the reference's concatenate-and-attend
            pos0 = max(0, int(self.cumulative_length) - int(self.size))   # position of the first cached key
            fk, fv = self.update(k_bthd.transpose(1, 2), v_bthd.transpose(1, 2))
            return swa.swa_attention_bthd(q_bthd, fk.transpose(1, 2), fv.transpose(1, 2), window=window, scale=scale,
                                          key_pos0=pos0)
        lib = _lib.load()
        stream = torch.cuda.current_stream(q_bthd.device).cuda_stream
        st3 = lambda t: (ctypes.c_int64 * 3)(t.stride(0), t.stride(1), t.stride(2))
        out = torch.empty(B, Tq, Hq, D, dtype=torch.bfloat16, device=q_bthd.device)
        if Tq == 1:
            if self._ws is None:
                self._ws = torch.empty(lib.ivl_swa_ring_workspace_bytes(B, Hq, W), dtype=torch.uint8, device=q_bthd.device)
            q = q_bthd.contiguous()
            st2 = lambda t: (ctypes.c_int64 * 2)(t.stride(0), t.stride(2))
            code = lib.ivl_swa_ring_decode(q.data_ptr(), k_bthd.data_ptr(), st2(k_bthd), v_bthd.data_ptr(), st2(v_bthd),
                                           self._ring_k.data_ptr(), self._ring_v.data_ptr(), self._state.data_ptr(),
                                           out.data_ptr(), B, Hq, self.num_kv_heads, D, W, self.R, float(scale or 0.0),
                                           self._ws.data_ptr(), self._ws.numel(), stream)
            _lib.check(code, "ivl_swa_ring_decode")
            self._size = min(self.capacity, self._size + 1)
            self._cum += 1
            return out
        self._append(k_bthd, v_bthd)
        q = q_bthd if (q_bthd.stride(3) == 1 and all(x % 8 == 0 for x in q_bthd.stride()[:3])) else q_bthd.contiguous()
        code = lib.ivl_swa_ring_fwd(q.data_ptr(), st3(q), self._ring_k.data_ptr(), self._ring_v.data_ptr(),
                                    self._state.data_ptr(), out.data_ptr(), st3(out), B, Tq, Hq, self.num_kv_heads, D, W,
                                    self.R, float(scale or 0.0), stream)
        _lib.check(code, "ivl_swa_ring_fwd")
        return out

    def crop(self, max_length: int) -> None:
        if self.get_seq_length() >= self.sliding_window:
            raise ValueError("Cropping is forbidden after filling SWA window (to avoid state loss)")
        if self._dirty:
            self._resync()
        new_size = max(0, self._size - abs(max_length)) if max_length < 0 else min(self._size, max_length)
        k = self._window(self._ring_k, self._size)[:, :, self._size - new_size:].clone()
        v = self._window(self._ring_v, self._size)[:, :, self._size - new_size:].clone()
        self.reset()
        if new_size > 0:
            self._append(k.transpose(1, 2), v.transpose(1, 2))

    def reset(self) -> None:
        self._size = self._cum = 0
        self._pend_size = self._pend_cum = None
        self._dirty = False
        self._state.zero_()

    def sync_from_device(self) -> None:
        """Refresh the Python-side integers from the device counter (one host sync).  Needed after CUDA-graph
        replays, which advance the device counter only, before anything reads `size` / `cumulative_length` /
        `keys` on the host again."""
        self._cum = int(self._state[0].item())
        self._size = min(self.capacity, self._cum)
        self._pend_size = self._pend_cum = None
        self._dirty = False

    # -- snapshots (branch a stream for a question, resume a stream from disk) ---------------------------
    def state_dict(self) -> dict:
        """The window in logical order + the counters: enough to rebuild the layer (safetensors-friendly)."""
        if self._dirty:
            self._resync()
        return {"keys": self.keys.contiguous().clone(), "values": self.values.contiguous().clone(),
                "size": torch.tensor(self._size), "cumulative_length": torch.tensor(self._cum)}

    def load_state_dict(self, sd: dict) -> None:
        self.reset()
        n = int(sd["size"])
        self._cum = int(sd["cumulative_length"]) - n
        self._state[0] = self._cum
        if n > 0:
            self._append(sd["keys"].to(self._ring_k.device).transpose(1, 2), sd["values"].to(self._ring_k.device).transpose(1, 2))


class StaticLinearLayerPrealloc(_HFCacheLayer):
    """Conv tails [B, D, conv_size] x3 and the recurrent state [B, H, K, V] of one GDN layer,
    stored in the cache dtype -- bf16 for a bf16 model, so the state is re-rounded at every
    call boundary exactly as in the reference (std:229-364)."""
    is_sliding = False

    def __init__(self, *, config, batch_size: int, device="cpu", dtype=torch.float32, zero_init: bool = False,
                 recurrent_state_shape: Optional[Tuple[int, ...]] = None, state_dtype: Optional[torch.dtype] = None):
        cfg = _get_decoder_cfg(config)
        self.num_linear_heads = int(getattr(cfg, "num_linear_heads", getattr(cfg, "num_attention_heads")))
        self.num_linear_kv_heads = int(getattr(cfg, "num_linear_key_value_heads", self.num_linear_heads))
        self.linear_head_dim = int(getattr(cfg, "linear_head_dim", None) or getattr(cfg, "head_dim"))
        self.conv_size = int(getattr(cfg, "conv_size", 1))
        self.use_short_conv = bool(getattr(cfg, "use_short_conv", True))
        self.v_head_dim = int(round(self.linear_head_dim * float(getattr(cfg, "expand_v", 1.0))))
        self.is_initialized = True
        self.dtype, self.device = dtype, device
        self.batch_size = int(batch_size)
        self.seq_len = 0
        self.start = False
        alloc = torch.zeros if zero_init else torch.empty
        B, Hq, Hk, C, Cv, K = (self.batch_size, self.num_linear_heads, self.num_linear_kv_heads, self.linear_head_dim,
                               self.v_head_dim, self.conv_size)
        if self.use_short_conv:
            self.conv_state_q = alloc((B, Hq * C, K), dtype=dtype, device=device)
            self.conv_state_k = alloc((B, Hk * C, K), dtype=dtype, device=device)
            self.conv_state_v = alloc((B, Hk * Cv, K), dtype=dtype, device=device)
        else:
            self.conv_state_q = self.conv_state_k = self.conv_state_v = None
        if recurrent_state_shape is None:
            recurrent_state_shape = (B, Hq, C, Cv)
        else:
            assert recurrent_state_shape[0] == B, "recurrent_state_shape batch dim must match pre-allocated batch_size"
        # state_dtype (extension): keep S in fp32 instead of re-rounding it to the cache dtype at
        # every call boundary (Appendix B point 11) -- what the sequence-sharded prefill hands over so that the
        # sharded run reproduces the one-GPU run.  None = the reference's behaviour (cache dtype).
        self.recurrent_state = alloc(tuple(recurrent_state_shape), dtype=state_dtype or dtype, device=device)

    def update(self, key_states=None, value_states=None, conv_state: Optional[tuple] = None,
               recurrent_state: Optional[torch.Tensor] = None, cache_kwargs: Optional[dict] = None) -> tuple:
        if cache_kwargs is None:
            cache_kwargs = {}
        op = cache_kwargs.get("op", "get" if (conv_state is None and recurrent_state is None) else "set")
        if self.start is False:  # the first call of the layer's life reports "nothing cached" (std:298-300)
            self.start = True
            return (None, None, None), None
        if op == "get":
            return (self.conv_state_q, self.conv_state_k, self.conv_state_v), self.recurrent_state
        if conv_state is not None and self.use_short_conv:
            assert isinstance(conv_state, (tuple, list)), "conv_state must be (cq, ck, cv)"
            cq, ck, cv = (tuple(conv_state) + (None, None, None))[:3]
            for name, src, dst in (("conv_q", cq, self.conv_state_q), ("conv_k", ck, self.conv_state_k),
                                   ("conv_v", cv, self.conv_state_v)):
                if src is None:
                    continue
                if tuple(src.shape) != tuple(dst.shape):
                    raise RuntimeError(f"{name} shape changed: got {tuple(src.shape)} vs prealloc {tuple(dst.shape)}")
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src)
        elif conv_state is not None and not self.use_short_conv:
            raise RuntimeError("config.use_short_conv=False, but conv_state was passed")
        if recurrent_state is not None:
            if tuple(recurrent_state.shape) != tuple(self.recurrent_state.shape):
                raise RuntimeError(f"recurrent_state shape changed: got {tuple(recurrent_state.shape)} vs prealloc "
                                   f"{tuple(self.recurrent_state.shape)}")
            if recurrent_state.data_ptr() != self.recurrent_state.data_ptr():
                self.recurrent_state.copy_(recurrent_state)
        self.seq_len += int(cache_kwargs.get("delta_len", 0))
        return (self.conv_state_q, self.conv_state_k, self.conv_state_v), self.recurrent_state

    def get_mask_sizes(self, cache_position: torch.Tensor) -> Tuple[int, int]:
        qlen = cache_position.shape[0] if cache_position is not None else 0
        return self.get_seq_length() + qlen, 0

    def get_seq_length(self) -> int:
        return int(self.seq_len)

    def get_max_cache_shape(self) -> int:
        return -1

    def crop(self, max_length: int) -> None:
        if max_length < 0:
            max_length = max(0, self.get_seq_length() - abs(max_length))
        self.seq_len = min(self.get_seq_length(), max_length)

    def batch_repeat_interleave(self, repeats: int) -> None:
        if repeats != 1:
            raise RuntimeError("Static cache forbids changing batch size (repeat_interleave)")

    def batch_select_indices(self, indices: torch.Tensor) -> None:
        if indices.numel() != self.batch_size:
            raise RuntimeError("Static cache forbids changing batch size (select_indices)")

    def lazy_initialization(self, *args, **kwargs):
        return

    def reset(self) -> None:
        self.seq_len = 0
        self.start = False

    # -- snapshots -------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        sd = {"recurrent_state": self.recurrent_state.clone(), "seq_len": torch.tensor(int(self.seq_len)),
              "start": torch.tensor(int(bool(self.start)))}
        if self.use_short_conv:
            sd.update(conv_state_q=self.conv_state_q.clone(), conv_state_k=self.conv_state_k.clone(),
                      conv_state_v=self.conv_state_v.clone())
        return sd

    def load_state_dict(self, sd: dict) -> None:
        self.recurrent_state.copy_(sd["recurrent_state"])
        if self.use_short_conv:
            for n in ("conv_state_q", "conv_state_k", "conv_state_v"):
                getattr(self, n).copy_(sd[n])
        self.seq_len, self.start = int(sd["seq_len"]), bool(int(sd["start"]))


class StaticCachePrealloc(_HFCache):
    """Per-layer caches built from config.layer_types (std:366-443)."""

    def __init__(self, *, config, batch_size: int = 1, device="cpu", dtype=torch.float32, zero_init: bool = False,
                 recurrent_state_shape: Optional[Tuple[int, ...]] = None, offloading: bool = False,
                 offload_only_non_sliding: bool = False, state_dtype: Optional[torch.dtype] = None,
                 ring: Optional[bool] = None, max_append: int = 256):
        cfg = _get_decoder_cfg(config)
        layer_types = getattr(cfg, "layer_types", None)
        if layer_types is None:
            layer_types = ["linear_attention"] * int(getattr(cfg, "num_hidden_layers"))
        if hasattr(cfg, "num_kv_shared_layers"):
            layer_types = layer_types[: -int(getattr(cfg, "num_kv_shared_layers"))]
        self.layers = []
        for lt in layer_types:
            if lt in ("sliding_attention", "chunked_attention"):
                self.layers.append(StaticSlidingWindowLayerPrealloc(config=cfg, batch_size=batch_size, device=device,
                                                                    dtype=dtype, zero_init=zero_init, ring=ring,
                                                                    max_append=max_append))
            elif lt in ("linear_attention", "delta_net", "retnet", "state_space"):
                self.layers.append(StaticLinearLayerPrealloc(config=cfg, batch_size=batch_size, device=device,
                                                             dtype=dtype, zero_init=zero_init,
                                                             recurrent_state_shape=recurrent_state_shape,
                                                             state_dtype=state_dtype))
            # full-attention layers are skipped, as in the reference (std:417-421); the shipped config has none
        layers = self.layers
        if _HFCache is not object:
            try:
                super().__init__(layers=layers)
            except TypeError:   # older / newer signatures: the attributes the base class reads
                self.layers, self.layer_class_to_replicate, self.offloading = layers, None, False

    def update(self, layer_idx: int, key_states=None, value_states=None, conv_state=None, recurrent_state=None,
               cache_kwargs: Optional[dict[str, Any]] = None):
        return self.layers[layer_idx].update(key_states, value_states, conv_state, recurrent_state, cache_kwargs)

    def get_seq_length(self, layer_idx: int = 0) -> int:
        if not self.layers:
            return 0
        return self.layers[layer_idx].get_seq_length()

    def get_mask_sizes(self, cache_position: torch.Tensor, layer_idx: int = 0) -> Tuple[int, int]:
        return self.layers[layer_idx].get_mask_sizes(cache_position)

    def __len__(self):
        return len(self.layers)

    def reset(self):
        for layer in self.layers:
            layer.reset()

    def sync_from_device(self) -> None:
        for layer in self.layers:
            if hasattr(layer, "sync_from_device"):
                layer.sync_from_device()

    # -- snapshots on disk (resume a stream, branch it for a question) -------------------------------
    def state_dict(self) -> dict:
        """Flat {"layers.<i>.<name>": tensor}: the windows in logical order, the recurrent / conv states and the
        integer bookkeeping as 0-d tensors (safetensors-friendly).  Call `sync_from_device()` first if the cache was
        advanced by CUDA-graph replays."""
        out = {}
        for i, layer in enumerate(self.layers):
            for k, v in layer.state_dict().items():
                out[f"layers.{i}.{k}"] = v.detach().contiguous()
        return out

    def load_state_dict(self, sd: dict) -> "StaticCachePrealloc":
        for i, layer in enumerate(self.layers):
            pre = f"layers.{i}."
            layer.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
        return self

    def save(self, path: str) -> None:
        from safetensors.torch import save_file
        save_file({k: v.cpu() for k, v in self.state_dict().items()}, path)

    def load(self, path: str) -> "StaticCachePrealloc":
        from safetensors.torch import load_file
        return self.load_state_dict(load_file(path))

    def copy_from(self, src: "StaticCachePrealloc") -> "StaticCachePrealloc":
        """Device-side snapshot of another cache of the same geometry (what the demo's clone_inference_cache does
        attribute by attribute, demo_streaming_inference.py:111-160) -- no host synchronisation, so it also works on
        caches that were advanced by graph replays."""
        for d, s_ in zip(self.layers, src.layers):
            if getattr(s_, "is_ring", False):
                d._ring_k.copy_(s_._ring_k); d._ring_v.copy_(s_._ring_v); d._state.copy_(s_._state)
                d._size, d._cum, d._dirty = s_._size, s_._cum, s_._dirty
                d._pend_size, d._pend_cum = s_._pend_size, s_._pend_cum
            elif getattr(s_, "is_sliding", False):
                d._buf_keys.copy_(s_._buf_keys); d._buf_values.copy_(s_._buf_values)
                d.size, d.cumulative_length = s_.size, s_.cumulative_length
                d.keys, d.values = d._buf_keys[:, :, :d.size, :], d._buf_values[:, :, :d.size, :]
            else:
                for n in ("conv_state_q", "conv_state_k", "conv_state_v", "recurrent_state"):
                    if getattr(s_, n) is not None:
                        getattr(d, n).copy_(getattr(s_, n))
                d.seq_len, d.start = s_.seq_len, s_.start
        return self

    def to_legacy_cache(self):
        return tuple((getattr(l, "keys", None), getattr(l, "values", None)) for l in self.layers)
