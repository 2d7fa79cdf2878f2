"""Sequence-chunk sharding of a long prefill over the GPUs of one node.

The reference never shards a sequence (it has no SP/CP): this is the part of the
north star that goes beyond it.  Rank r owns the contiguous token range
[r T / P, (r + 1) T / P).  Everything in a decoder layer is token-local except what the layer's
inference cache carries across calls -- the DeltaNet state S and the three conv tails for a GDN
layer, the last W - 1 rotated keys/values for an SWA layer -- so the hand-off between neighbouring
ranks is exactly "send your cache layer after you have run the layer on your range":

    for each layer i:   recv cache.layers[i] from rank r-1   (r > 0)
                        run the layer on the local tokens with that cache
                        send cache.layers[i] to rank r+1     (r < P-1)

One point-to-point message per layer boundary (NCCL send/recv over NVLink on GPUs, gloo in the
CPU tests), no collective; rank r starts layer i when rank r-1 has finished it, which gives the
wavefront over (rank, layer) with ideal efficiency L / (L + P - 1).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .cache import StaticCachePrealloc, StaticLinearLayerPrealloc, StaticSlidingWindowLayerPrealloc


def shard_range(T: int, world: int, rank: int, multiple: int = 64) -> Tuple[int, int]:
    """Contiguous token range of `rank`; every boundary is a multiple of `multiple` (the GDN chunk)
    so that sharded and single-device scans use identical chunking."""
    if T % (world * multiple) != 0:
        raise ValueError(f"sequence length {T} must be a multiple of world_size * {multiple} = {world * multiple}")
    n = T // world
    return rank * n, (rank + 1) * n


def _linear_tensors(layer: StaticLinearLayerPrealloc) -> List[torch.Tensor]:
    return [layer.recurrent_state, layer.conv_state_q, layer.conv_state_k, layer.conv_state_v]


def send_cache_layer(layer, dst: int, group=None) -> None:
    """Ship the state a layer's cache carries to the rank that owns the next token range."""
    if isinstance(layer, StaticLinearLayerPrealloc) or not getattr(layer, "is_sliding", False):
        for t in _linear_tensors(layer):
            dist.send(t, dst=dst, group=group)
    else:
        n = layer.size
        dist.send(layer._buf_keys[:, :, :n, :].contiguous(), dst=dst, group=group)
        dist.send(layer._buf_values[:, :, :n, :].contiguous(), dst=dst, group=group)


def recv_cache_layer(layer, src: int, tokens_before: int, group=None) -> None:
    """Receive the predecessor's cache layer; `tokens_before` = number of tokens all earlier ranks own
    (fixes the integer bookkeeping the reference keeps in Python: size / cumulative_length / seq_len)."""
    if isinstance(layer, StaticLinearLayerPrealloc) or not getattr(layer, "is_sliding", False):
        for t in _linear_tensors(layer):
            dist.recv(t, src=src, group=group)
        layer.start = True
        layer.seq_len = int(tokens_before)
    else:
        n = min(layer.capacity, int(tokens_before))
        kbuf = torch.empty_like(layer._buf_keys[:, :, :n, :]).contiguous()
        vbuf = torch.empty_like(kbuf)
        dist.recv(kbuf, src=src, group=group)
        dist.recv(vbuf, src=src, group=group)
        layer._buf_keys[:, :, :n, :].copy_(kbuf)
        layer._buf_values[:, :, :n, :].copy_(vbuf)
        layer.keys = layer._buf_keys[:, :, :n, :]
        layer.values = layer._buf_values[:, :, :n, :]
        layer.size = n
        layer.cumulative_length = int(tokens_before)


def sharded_layer_loop(layer_fns: Sequence[Callable], hidden_local: torch.Tensor, cache: StaticCachePrealloc,
                       tokens_before: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Run `layer_fns[i](hidden, cache, i) -> hidden` for every layer with the cache hand-off above.
    `hidden_local` holds this rank's tokens only."""
    h = hidden_local
    for i, fn in enumerate(layer_fns):
        if rank > 0:
            recv_cache_layer(cache.layers[i], rank - 1, tokens_before, group)
        h = fn(h, cache, i)
        if rank < world - 1:
            send_cache_layer(cache.layers[i], rank + 1, group)
    return h


def sharded_prefill(decoder, inputs_embeds_local: torch.Tensor, position_ids_local: torch.Tensor, T_total: int,
                    rank: Optional[int] = None, world: Optional[int] = None, group=None,
                    cache: Optional[StaticCachePrealloc] = None):
    """Sequence-sharded forward of an infinitevl_b200.modeling.HybridDecoder.  Returns (hidden states of
    the local token range, cache) -- after the call the LAST rank's cache holds the state of the whole
    sequence (what decoding continues from)."""
    from .modeling import mrope_select
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    start, end = shard_range(T_total, world, rank)
    B, T_local, _ = inputs_embeds_local.shape
    assert T_local == end - start
    if cache is None:
        cache = decoder.allocate_inference_cache(B)
    cache_position = torch.arange(start, end, device=inputs_embeds_local.device)
    if position_ids_local.dim() == 2:
        position_ids_local = position_ids_local[None].expand(3, -1, -1)
    cos, sin = decoder.rotary_emb(inputs_embeds_local, position_ids_local)
    cos, sin = mrope_select(cos, sin, decoder.config.rope_scaling["mrope_section"])

    def make(layer):
        def run(h, c, i):
            if decoder.mixers_only:
                return h + layer.self_attn(hidden_states=layer.input_layernorm(h), past_key_values=c,
                                           cache_position=cache_position, position_embeddings=(cos, sin))[0]
            return layer(h, past_key_values=c, cache_position=cache_position, position_embeddings=(cos, sin))[0]
        return run

    h = sharded_layer_loop([make(l) for l in decoder.layers], inputs_embeds_local, cache, start, rank, world, group)
    return decoder.norm(h), cache
