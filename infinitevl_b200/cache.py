"""Inference caches with the reference's class names, attribute names and index semantics
(infinitevl_standard/modeling_infinitevl.py:66-443), so that callers which poke at the buffers
directly -- inference_examples/demo_streaming_inference.py:111-160 clones `_buf_keys`,
`_buf_values`, `keys`, `values`, `size`, `cumulative_length`, `capacity`, `recurrent_state`,
`conv_state_{q,k,v}`, `seq_len`, `start` -- keep working unchanged.

Nothing here allocates after construction: update() only copies into the pre-allocated
buffers, which is what CUDA-graph capture of the forward needs.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple

import torch


def _get_decoder_cfg(config):
    if hasattr(config, "get_text_config"):
        try:
            return config.get_text_config(decoder=True)
        except TypeError:
            return config.get_text_config()
    return config


class StaticSlidingWindowLayerPrealloc:
    """Last (sliding_window - 1) keys/values of one SWA layer (std:66-227)."""
    is_sliding = True

    def __init__(self, *, config, batch_size: int, device="cpu", dtype=torch.float32, zero_init: bool = False):
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


class StaticLinearLayerPrealloc:
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


class StaticCachePrealloc:
    """Per-layer caches built from config.layer_types (std:366-443)."""

    def __init__(self, *, config, batch_size: int = 1, device="cpu", dtype=torch.float32, zero_init: bool = False,
                 recurrent_state_shape: Optional[Tuple[int, ...]] = None, offloading: bool = False,
                 offload_only_non_sliding: bool = False, state_dtype: Optional[torch.dtype] = None):
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
                                                                    dtype=dtype, zero_init=zero_init))
            elif lt in ("linear_attention", "delta_net", "retnet", "state_space"):
                self.layers.append(StaticLinearLayerPrealloc(config=cfg, batch_size=batch_size, device=device,
                                                             dtype=dtype, zero_init=zero_init,
                                                             recurrent_state_shape=recurrent_state_shape,
                                                             state_dtype=state_dtype))
            # full-attention layers are skipped, as in the reference (std:417-421); the shipped config has none

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

    def to_legacy_cache(self):
        return tuple((getattr(l, "keys", None), getattr(l, "values", None)) for l in self.layers)
