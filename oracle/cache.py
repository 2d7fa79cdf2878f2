"""CPU restatement of the inference caches' index logic (SURVEY.md rows a-C1, a-C2).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  These are integer /
copy semantics and must be matched bit-exactly.  ``std`` =
/root/reference/infinitevl/infinitevl_standard/modeling_infinitevl.py.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def swa_mask_sizes_ref(cumulative_length_after_update: int, q_len: int, window: int) -> Tuple[int, int]:
    """(kv_len, kv_offset) as StaticSlidingWindowLayerPrealloc.get_mask_sizes computes
    them *after* update() has advanced cumulative_length by q_len (std:175-184)."""
    past = max(int(cumulative_length_after_update) - int(q_len), 0)
    kv_offset = max(past - window + 1, 0)
    kv_len = (window - 1 if past >= window else past) + q_len
    return kv_len, kv_offset


class SlidingWindowCacheRef:
    """Keeps the last (window - 1) keys/values.  update() returns [previous tail, new]
    and then stores the newest <= window - 1 rows (std:126-173)."""

    def __init__(self, window: int):
        self.window = int(window)
        self.capacity = max(self.window - 1, 0)
        self.keys: Optional[torch.Tensor] = None
        self.values: Optional[torch.Tensor] = None
        self.size = 0
        self.cumulative_length = 0

    def update(self, k: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        Tq = k.shape[-2]
        if self.keys is None:
            full_k, full_v = k, v
        else:
            full_k = torch.cat([self.keys, k], dim=-2)
            full_v = torch.cat([self.values, v], dim=-2)
        keep = min(self.capacity, self.size + Tq)
        self.keys = full_k[..., full_k.shape[-2] - keep:, :].clone()
        self.values = full_v[..., full_v.shape[-2] - keep:, :].clone()
        self.size = keep
        self.cumulative_length += Tq
        return full_k, full_v

    def get_mask_sizes(self, q_len: int) -> Tuple[int, int]:
        return swa_mask_sizes_ref(self.cumulative_length, q_len, self.window)


class LinearCacheRef:
    """Conv tails + recurrent state for one GDN layer.  The very first update() of a
    layer's life returns "nothing cached" whatever the op (std:298-300); "set" copies in
    place and *rounds to the cache dtype* (std:279-284,305-338)."""

    def __init__(self, batch: int, H: int = 16, K: int = 128, V: int = 256, conv: int = 4,
                 dtype=torch.bfloat16):
        self.conv_state_q = torch.zeros(batch, H * K, conv, dtype=dtype)
        self.conv_state_k = torch.zeros(batch, H * K, conv, dtype=dtype)
        self.conv_state_v = torch.zeros(batch, H * V, conv, dtype=dtype)
        self.recurrent_state = torch.zeros(batch, H, K, V, dtype=dtype)
        self.seq_len = 0
        self.start = False

    def update(self, conv_state=None, recurrent_state=None, op: Optional[str] = None, delta_len: int = 0):
        if op is None:
            op = "get" if (conv_state is None and recurrent_state is None) else "set"
        if not self.start:
            self.start = True
            return (None, None, None), None
        if op == "set":
            if conv_state is not None:
                for dst, src in zip((self.conv_state_q, self.conv_state_k, self.conv_state_v), conv_state):
                    if src is not None:
                        if tuple(src.shape) != tuple(dst.shape):
                            raise RuntimeError("conv state shape changed")
                        dst.copy_(src)
            if recurrent_state is not None:
                if tuple(recurrent_state.shape) != tuple(self.recurrent_state.shape):
                    raise RuntimeError("recurrent_state shape changed")
                self.recurrent_state.copy_(recurrent_state)
            self.seq_len += int(delta_len)
        return (self.conv_state_q, self.conv_state_k, self.conv_state_v), self.recurrent_state
