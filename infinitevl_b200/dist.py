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

All transfers are posted asynchronously (``isend`` / ``irecv``): the receive of layer i + 1 is
posted before layer i runs, straight into the cache's own buffers (no staging copy when the
whole window is handed over), and a send only makes the communication stream wait for the
compute stream, never the reverse.  The compute stream waits for a receive only at the point
where the received state is first read -- for a GDN layer that is after its state-independent
pre-pass (`gdn_layer_sharded`).
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


class Pending:
    """A group of posted point-to-point transfers.  wait() makes the CURRENT stream wait for them (NCCL) or
    blocks the host (gloo); finish() runs the bookkeeping that has to follow the data."""

    def __init__(self, works=(), after: Optional[Callable[[], None]] = None, keep=()):
        self.works = [w for w in works if w is not None]
        self.after = after
        self.keep = keep          # tensors that must stay alive until the transfer is done

    def wait(self) -> None:
        for w in self.works:
            w.wait()
        self.works = []
        if self.after is not None:
            self.after()
            self.after = None
        self.keep = ()


def _linear_tensors(layer: StaticLinearLayerPrealloc) -> List[torch.Tensor]:
    return [layer.recurrent_state, layer.conv_state_q, layer.conv_state_k, layer.conv_state_v]


def _is_linear(layer) -> bool:
    return isinstance(layer, StaticLinearLayerPrealloc) or not getattr(layer, "is_sliding", False)


def send_cache_layer(layer, dst: int, group=None) -> Pending:
    """Ship the state a layer's cache carries to the rank that owns the next token range."""
    if _is_linear(layer):
        ts = _linear_tensors(layer)
    else:
        n = layer.size
        k, v = layer._buf_keys[:, :, :n, :], layer._buf_values[:, :, :n, :]
        ts = [k if k.is_contiguous() else k.contiguous(), v if v.is_contiguous() else v.contiguous()]
    return Pending([dist.isend(t, dst=dst, group=group) for t in ts], keep=ts)


def recv_cache_layer(layer, src: int, tokens_before: int, group=None) -> Pending:
    """Post the receive of the predecessor's cache layer; `tokens_before` = number of tokens all earlier ranks
    own (fixes the integer bookkeeping the reference keeps in Python: size / cumulative_length / seq_len).
    The data is valid, and the bookkeeping done, after .wait()."""
    if _is_linear(layer):
        works = [dist.irecv(t, src=src, group=group) for t in _linear_tensors(layer)]

        def after():
            layer.start = True
            layer.seq_len = int(tokens_before)
        return Pending(works, after)
    n = min(layer.capacity, int(tokens_before))
    kdst, vdst = layer._buf_keys[:, :, :n, :], layer._buf_values[:, :, :n, :]
    direct = kdst.is_contiguous() and vdst.is_contiguous()   # the whole window (or one kv head / batch row)
    kbuf = kdst if direct else torch.empty(kdst.shape, dtype=kdst.dtype, device=kdst.device)
    vbuf = vdst if direct else torch.empty_like(kbuf)
    works = [dist.irecv(kbuf, src=src, group=group), dist.irecv(vbuf, src=src, group=group)]

    def after():
        if not direct:
            kdst.copy_(kbuf)
            vdst.copy_(vbuf)
        layer.keys, layer.values = kdst, vdst
        layer.size = n
        layer.cumulative_length = int(tokens_before)
    return Pending(works, after, keep=(kbuf, vbuf))


def sharded_layer_loop(layer_fns: Sequence[Callable], hidden_local: torch.Tensor, cache: StaticCachePrealloc,
                       tokens_before: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Run `layer_fns[i](hidden, cache, i) -> hidden` for every layer with the cache hand-off above.
    `hidden_local` holds this rank's tokens only.  The receive of layer i + 1 is in flight while layer i runs."""
    h = hidden_local
    L = len(layer_fns)
    post = (lambda i: recv_cache_layer(cache.layers[i], rank - 1, tokens_before, group)) if rank > 0 else None
    nxt = post(0) if post and L else None
    sends: List[Pending] = []
    for i, fn in enumerate(layer_fns):
        cur, nxt = nxt, (post(i + 1) if post and i + 1 < L else None)
        if cur is not None:
            cur.wait()
        h = fn(h, cache, i)
        if rank < world - 1:
            sends.append(send_cache_layer(cache.layers[i], rank + 1, group))
    for s in sends:
        s.wait()
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


# ------------------------------------------------------------------------------------------------
# operator-level hand-off (what bench.py --gpus N times: the same protocol without the modules around it)
# ------------------------------------------------------------------------------------------------
class OperatorHandOff:
    """Neighbour hand-off of the two hot-path operators for one rank of a sequence-sharded prefill.

    GDN layer:  prep (state independent) -> wait for S from rank r-1 -> scan -> isend S to rank r+1.
    SWA layer:  isend the last W-1 local keys/values, irecv the halo straight into the first W-1 rows of a
                [W-1 + T_local] K/V buffer whose tail holds the local K/V (no concatenation), attend.
    Receives are posted one layer ahead (`post_recv`), so a transfer never waits for the host."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self._sends: List[Pending] = []

    @property
    def first(self) -> bool:
        return self.rank == 0

    @property
    def last(self) -> bool:
        return self.rank == self.world - 1

    def post_recv(self, tensors: Sequence[torch.Tensor]) -> Optional[Pending]:
        if self.first:
            return None
        return Pending([dist.irecv(t, src=self.rank - 1, group=self.group) for t in tensors], keep=tuple(tensors))

    def post_send(self, tensors: Sequence[torch.Tensor]) -> None:
        if self.last:
            return
        self._sends.append(Pending([dist.isend(t, dst=self.rank + 1, group=self.group) for t in tensors],
                                   keep=tuple(tensors)))
        if len(self._sends) > 8:   # bound the number of outstanding works; the oldest finished long ago
            self._sends.pop(0).wait()

    def drain(self) -> None:
        for s in self._sends:
            s.wait()
        self._sends = []


def gdn_layer_sharded(ho: OperatorHandOff, prep: Callable[[], None], scan: Callable[[torch.Tensor], None],
                      h0_local: torch.Tensor, state_in: torch.Tensor, state_out: torch.Tensor,
                      pending: Optional[Pending]) -> None:
    """One GDN layer of a sharded prefill at operator level: `prep()` launches the chunk pre-pass,
    `scan(h0)` the recurrence writing the final state into `state_out`.  `pending` is the posted receive
    of `state_in` (None on rank 0, which starts from `h0_local`)."""
    prep()
    h0 = h0_local
    if pending is not None:
        pending.wait()
        h0 = state_in
    scan(h0)
    ho.post_send([state_out])


# ------------------------------------------------------------------------------------------------
# parity of the sharded run against the one-GPU run (error ratio <= 1e-3)
# ------------------------------------------------------------------------------------------------
def sharded_parity_check(T: int = 32768, num_layers: int = 8, seed: int = 0, group=None, config=None) -> dict:
    """Every rank builds the same HybridDecoder (3B mixer dims, `num_layers` layers in the model's 1 SWA : 3 GDN
    pattern, mixers only) and the same inputs; the ranks run `sharded_prefill` over NCCL, the last rank also runs
    the whole sequence alone, and the rank-concatenated output and the last rank's cache are compared with that
    run.  The DeltaNet state is handed over (and cached) in fp32, as `sharded_prefill` callers should.
    Returns {"out": err, "state": max err over GDN layers, "conv": ..., "kv": max err over SWA layers} on the
    last rank, {} elsewhere."""
    from .modeling import HybridDecoder, HybridTextConfig
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    cfg = config or HybridTextConfig(num_hidden_layers=num_layers)
    torch.manual_seed(seed)
    dec = HybridDecoder(cfg, mixers_only=True)
    for p in dec.parameters():
        if p.dim() >= 2:
            torch.nn.init.normal_(p, std=0.02)
    dec = dec.to(dev, torch.bfloat16).eval()
    gen = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(1, T, cfg.hidden_size, generator=gen).to(dev, torch.bfloat16)
    pos = torch.arange(T, device=dev)[None, None].expand(3, 1, -1)
    s, e = shard_range(T, world, rank)
    cache = dec.allocate_inference_cache(1, state_dtype=torch.float32)
    out, cache = sharded_prefill(dec, x[:, s:e], pos[:, :, s:e], T, rank, world, group, cache)
    parts = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(parts, out.contiguous(), group=group)
    res = {}
    if rank == world - 1:
        def err(ref, y):
            ref, y = ref.float(), y.float()
            return float(((ref - y).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-30)).item())
        rcache = dec.allocate_inference_cache(1, state_dtype=torch.float32)
        ref = dec(x, position_ids=pos, past_key_values=rcache)
        res = {"out": err(ref, torch.cat(parts, 1)), "state": 0.0, "conv": 0.0, "kv": 0.0, "ints_equal": True,
               "T": T, "layers": num_layers, "world": world}
        for a, b in zip(cache.layers, rcache.layers):
            if a.is_sliding:
                res["ints_equal"] &= (a.size == b.size and a.cumulative_length == b.cumulative_length)
                res["kv"] = max(res["kv"], err(b.keys, a.keys), err(b.values, a.values))
            else:
                res["ints_equal"] &= a.seq_len == b.seq_len
                res["state"] = max(res["state"], err(b.recurrent_state, a.recurrent_state))
                res["conv"] = max(res["conv"], err(b.conv_state_v, a.conv_state_v))
        del rcache, ref
    del dec, x, cache, out, parts
    torch.cuda.empty_cache()
    return res
