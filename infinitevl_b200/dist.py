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

    def __init__(self, rank: int, world: int, group=None, dry: bool = False):
        self.rank, self.world, self.group = rank, world, group
        self._sends: List[Pending] = []
        self.dry = dry   # developer knob: skip every transfer (per-rank compute time of the sharded step)

    @property
    def first(self) -> bool:
        return self.rank == 0

    @property
    def last(self) -> bool:
        return self.rank == self.world - 1

    def post_recv(self, tensors: Sequence[torch.Tensor]) -> Optional[Pending]:
        if self.first or self.dry:
            return None
        return Pending([dist.irecv(t, src=self.rank - 1, group=self.group) for t in tensors], keep=tuple(tensors))

    def post_send(self, tensors: Sequence[torch.Tensor]) -> None:
        if self.last or self.dry:
            return
        self._sends.append(Pending([dist.isend(t, dst=self.rank + 1, group=self.group) for t in tensors],
                                   keep=tuple(tensors)))
        if len(self._sends) > 8:   # bound the number of outstanding works; the oldest finished long ago
            self._sends.pop(0).wait()

    def drain(self) -> None:
        for s in self._sends:
            s.wait()
        self._sends = []


_IPC_BASES: dict = {}   # (device, IPC handle) -> base address of the mapping in this process


class PeerLink:
    """Neighbour hand-off r -> r + 1 through PEER MEMORY instead of NCCL: the receive buffers of rank r + 1 (the
    DeltaNet state slots, the halo rows at the front of its K/V buffers) are opened on rank r through CUDA IPC, the
    sender stores its tensors straight into them over NVLink and then publishes a flag, in one small launch
    (`ivl_peer_put`: no rendezvous, a few CTAs for a few microseconds); the receiver's compute stream waits on that flag with a stream memory operation
    (`ivl_stream_wait_value32`).  Back-pressure is a second flag the other way round ("consumed"), so a buffer is
    never overwritten before the kernel that read it has finished.  Everything is stream-ordered: the host never
    blocks, and neither side spends an SM on the transfer -- NCCL's send / recv kernels cost the two-GPU step 12 %
    (75.0 ms against 65.8 ms of per-rank compute, profiles/r02_summary.md).

    Usage (same call order on every rank):
        link = PeerLink(rank, world, group, device)
        link.open({"S0": state_in_0, ..., "K0": kbuf0[:, :halo], ...})   # the tensors the PREVIOUS rank writes
        link.wait("S0"); <kernels reading state_in_0>; link.release("S0")  # receiver
        link.before_overwrite("S0"); <kernel writing src>; link.send("S0", src)   # sender
    """

    def __init__(self, rank: int, world: int, group=None, device=None):
        from . import _lib
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.first, self.last = rank == 0, rank == world - 1
        self._lib, self._libmod = _lib.load(), _lib
        # high priority: a put is a few CTAs that must slip in between the thousands of queued CTAs of the next layer's
        # kernels -- at equal priority the block scheduler drains those first and the neighbour waits for ~0.5 ms
        self.side = torch.cuda.Stream(self.device, priority=-1)   # sends
        self.ack = torch.cuda.Stream(self.device, priority=-1)    # "consumed" flags going back (never behind a waiting send)
        self.names: List[str] = []
        import os
        # Fork / join events between the compute stream and the two side streams are created WITH timing: measured on
        # 2 x B200 the pipelined step takes 73-74 ms with timing-disabled events and 66.5-68 ms with these (the
        # per-rank compute is 65.8 ms); profiles/r02_summary.md.  IVL_P2P_TIMING_EVENTS=0 restores the plain ones.
        self._timing = os.environ.get("IVL_P2P_TIMING_EVENTS", "1") == "1"

    def open(self, recv: dict) -> None:
        """Collective.  `recv[name]`: the local tensor rank - 1 will write (every rank passes the same names; rank 0's
        tensors are never written and may be None)."""
        self.names = list(recv.keys())
        n = len(self.names)
        self.idx = {k: i for i, k in enumerate(self.names)}
        self.recv = dict(recv)
        # flags[0:n]: "data of epoch e is in" (written by rank - 1); flags[n:2n]: "epoch e consumed" (written by rank + 1)
        self.flags = torch.zeros(2 * n, dtype=torch.int32, device=self.device)
        self.counters = torch.zeros(2 * n, dtype=torch.int32, device=self.device)   # block counters of ivl_peer_put
        torch.cuda.synchronize(self.device)

        def export(t: torch.Tensor):
            # (IPC handle of the BASE allocation the caching allocator carved this tensor from, byte offset, bytes)
            assert t.is_contiguous(), "PeerLink buffers must be contiguous"
            import ctypes
            handle = ctypes.create_string_buffer(64)
            off = ctypes.c_uint64()
            self._libmod.check(self._lib.ivl_ipc_export(t.data_ptr(), handle, ctypes.byref(off)), "ivl_ipc_export")
            return handle.raw, int(off.value), t.numel() * t.element_size()

        mine = {"flags": export(self.flags),
                "recv": {k: (export(t) if (t is not None and not self.first) else None) for k, t in recv.items()}}
        every = [None] * self.world
        dist.all_gather_object(every, mine, group=self.group)
        def map_(ex) -> int:
            # one mapping per exported allocation and process (CUDA refuses to open a handle twice): shared by every
            # link of this process, e.g. the links a sweep opens one after the other over recycled allocator blocks
            handle, off, _ = ex
            key = (self.device.index, handle)
            if key not in _IPC_BASES:
                import ctypes
                p = ctypes.c_void_p()
                self._libmod.check(self._lib.ivl_ipc_open(handle, ctypes.byref(p)), "ivl_ipc_open")
                _IPC_BASES[key] = int(p.value)
            return _IPC_BASES[key] + off

        self.peer_ptr, self.peer_bytes, self.next_flags, self.prev_flags = {}, {}, None, None
        err = None
        try:
            if not self.last:
                nxt = every[self.rank + 1]
                self.next_flags = map_(nxt["flags"])
                for k, ex in nxt["recv"].items():
                    if ex is not None:
                        self.peer_ptr[k], self.peer_bytes[k] = map_(ex), ex[2]
            if not self.first:
                self.prev_flags = map_(every[self.rank - 1]["flags"])
        except Exception as e:  # noqa: BLE001
            err = e
        # all or nothing: a rank that cannot map its neighbours must not leave them waiting on flags it never writes
        ok = torch.tensor([0.0 if err is not None else 1.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() < 1.0:
            raise RuntimeError(f"PeerLink: peer mapping failed on at least one rank ({err!r} here)")
        self.epoch_in = {k: 0 for k in self.names}
        self.epoch_out = {k: 0 for k in self.names}
        self.sent = {}
        dist.barrier(group=self.group)

    def _wait_value(self, stream: torch.cuda.Stream, index: int, value: int) -> None:
        self._libmod.check(self._lib.ivl_stream_wait_value32(stream.cuda_stream, self.flags.data_ptr() + 4 * index,
                                                           value & 0xFFFFFFFF), "ivl_stream_wait_value32")

    def wait(self, name: str) -> None:
        """The current stream waits until the previous rank's next transfer into `recv[name]` has landed."""
        if self.first:
            return
        self.epoch_in[name] += 1
        self._wait_value(torch.cuda.current_stream(self.device), self.idx[name], self.epoch_in[name])

    def release(self, name: str) -> None:
        """Everything enqueued on the current stream so far has finished reading `recv[name]`: tell the sender."""
        if self.first:
            return
        ev = torch.cuda.Event(enable_timing=self._timing)
        ev.record(torch.cuda.current_stream(self.device))
        n, i, e = len(self.names), self.idx[name], self.epoch_in[name]
        self.ack.wait_event(ev)
        self._put(None, None, self.prev_flags, n + i, e, self.ack)

    def _put(self, dst_ptr, src, flags_ptr: int, index: int, value: int, stream=None) -> None:
        nbytes = 0 if src is None else src.numel() * src.element_size()
        self._libmod.check(self._lib.ivl_peer_put(
            dst_ptr, src.data_ptr() if src is not None else None, nbytes, flags_ptr + 4 * index, value & 0xFFFFFFFF,
            self.counters.data_ptr() + 4 * index, (stream or self.side).cuda_stream), "ivl_peer_put")

    def send(self, name: str, src: torch.Tensor) -> None:
        """Copy `src` (as of everything enqueued on the current stream so far) into the next rank's `recv[name]`."""
        if self.last:
            return
        self.epoch_out[name] += 1
        n, i, e = len(self.names), self.idx[name], self.epoch_out[name]
        ev = torch.cuda.Event(enable_timing=self._timing)
        ev.record(torch.cuda.current_stream(self.device))
        assert src.is_contiguous() and src.numel() * src.element_size() == self.peer_bytes[name]
        self.side.wait_event(ev)
        if e > 1:
            self._wait_value(self.side, n + i, e - 1)      # the receiver has consumed the previous epoch
        self._put(self.peer_ptr[name], src, self.next_flags, i, e)
        done = torch.cuda.Event()
        done.record(self.side)
        self.sent[name] = done

    def before_overwrite(self, name: str) -> None:
        """The current stream waits until the last send of `name` has read its source."""
        ev = self.sent.pop(name, None)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def drain(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.side)
        torch.cuda.current_stream(self.device).wait_stream(self.ack)


def gdn_layer_sharded(ho: OperatorHandOff, prep: Callable[[], None], scan: Callable[[torch.Tensor], None],
                      h0_local: torch.Tensor, state_in: torch.Tensor, state_out: torch.Tensor,
                      pending: Optional[Pending], fwd: Optional[Callable[[torch.Tensor], None]] = None) -> None:
    """One GDN layer of a sharded prefill at operator level: `prep()` launches the chunk pre-pass,
    `scan(h0)` the recurrence writing the final state into `state_out`.  `pending` is the posted receive
    of `state_in` (None on rank 0, which starts from `h0_local`).
    `fwd(h0)` (optional): the whole chunk operator, which overlaps its pre-pass with its scan on two streams.  With
    the receive posted a layer ahead the state is normally there before the layer starts, so waiting for it first
    and then running the overlapped operator beats "pre-pass, wait, scan" (measured, profiles/r02_summary.md)."""
    h0 = h0_local
    if fwd is not None:
        if pending is not None:
            pending.wait()
            h0 = state_in
        fwd(h0)
    else:
        prep()
        if pending is not None:
            pending.wait()
            h0 = state_in
        scan(h0)
    ho.post_send([state_out])


# ------------------------------------------------------------------------------------------------
# parity of the sharded run against the one-GPU run
# ------------------------------------------------------------------------------------------------
def _err(ref: torch.Tensor, y: torch.Tensor) -> float:
    ref, y = ref.float(), y.float()
    return float(((ref - y).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-30)).item())


def operator_parity_check(T: int = 32768, seed: int = 0, group=None, transport: str = "nccl") -> dict:
    """The hot path itself (the two operators, no projections): every rank builds the SAME full-length inputs, runs
    its token range with the neighbour hand-off of `OperatorHandOff` / `gdn_layer_sharded` over the process group, and
    compares its outputs with the matching slice of the one-shot operator call on the whole sequence, which it also
    runs.  The kernels are deterministic, the state travels in fp32 and the SWA key tiles are anchored at absolute
    positions, so the sharded run must reproduce the one-shot run BIT FOR BIT.  transport: "nccl" (isend / irecv) or
    "p2p" (`PeerLink`: copy-engine writes into the neighbour's buffers + stream-ordered flags).  Same dict on every rank:
    {"gdn_o_equal", "gdn_state_equal", "swa_equal"} (logical AND over the ranks) and the worst error ratios."""
    from . import ops, swa
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    s, e = shard_range(T, world, rank)
    Tl = e - s
    H, K, V, HQ, HKV, D, W = 16, 128, 256, 16, 2, 128, 8192
    gen = torch.Generator().manual_seed(seed)
    q = torch.randn(1, T, H, K, generator=gen).bfloat16().to(dev)
    k = torch.randn(1, T, H, K, generator=gen).bfloat16().to(dev)
    v = torch.randn(1, T, H, V, generator=gen).bfloat16().to(dev)
    g = (-torch.rand(1, T, H, generator=gen) * 0.1).float().to(dev)
    beta = torch.rand(1, T, H, generator=gen).bfloat16().to(dev)
    h0 = (torch.randn(1, H, K, V, generator=gen) * 0.1).float().to(dev)
    ref_o, ref_s = ops.chunk_gated_delta_rule(q, k, v, g, beta, initial_state=h0, output_final_state=True,
                                              use_qk_l2norm_in_kernel=True)
    p2p = transport == "p2p"
    ho = OperatorHandOff(rank, world, group, dry=p2p)
    state_in, state_out = torch.empty_like(h0), torch.empty_like(h0)
    Hh = W - 1
    n_in, n_out = min(Hh, s), min(Hh, e)
    kb = torch.zeros(1, Hh + Tl, HKV, D, dtype=torch.bfloat16, device=dev)
    vb = torch.zeros_like(kb)
    link = None
    if p2p:
        link = PeerLink(rank, world, group, dev)
        link.open({"S": state_in, "K": kb[:, Hh - n_in:Hh] if n_in else None, "V": vb[:, Hh - n_in:Hh] if n_in else None})
    pend = ho.post_recv([state_in])
    if pend is not None:
        pend.wait()
    if link is not None:
        link.wait("S")
    o, _ = ops.chunk_gated_delta_rule(q[:, s:e].contiguous(), k[:, s:e].contiguous(), v[:, s:e].contiguous(),
                                      g[:, s:e].contiguous(), beta[:, s:e].contiguous(),
                                      initial_state=h0 if rank == 0 else state_in, output_final_state=True,
                                      use_qk_l2norm_in_kernel=True, state_out=state_out)
    if link is not None:
        link.release("S")
        link.send("S", state_out)
    ho.post_send([state_out])
    ho.drain()
    flags = [float(torch.equal(o, ref_o[:, s:e])), float(torch.equal(state_out, ref_s)) if rank == world - 1 else 1.0]
    errs = [_err(ref_o[:, s:e], o), _err(ref_s, state_out) if rank == world - 1 else 0.0]
    # SWA: halo of the last W-1 keys/values from the previous rank(s) in front of the local ones
    sq = torch.randn(1, T, HQ, D, generator=gen).bfloat16().to(dev)
    sk = torch.randn(1, T, HKV, D, generator=gen).bfloat16().to(dev)
    sv = torch.randn(1, T, HKV, D, generator=gen).bfloat16().to(dev)
    ref_a = swa.swa_attention_bthd(sq, sk, sv, window=W)
    kb[:, Hh:].copy_(sk[:, s:e])
    vb[:, Hh:].copy_(sv[:, s:e])

    def send_halo():
        if link is not None:
            link.send("K", kb[:, Hh + Tl - n_out:])
            link.send("V", vb[:, Hh + Tl - n_out:])
        else:
            ho.post_send([kb[:, Hh + Tl - n_out:], vb[:, Hh + Tl - n_out:]])

    pend = ho.post_recv([kb[:, Hh - n_in:Hh], vb[:, Hh - n_in:Hh]]) if n_in else None
    if n_out <= Tl:
        send_halo()          # only local rows leave: no need to wait for the incoming halo
    if pend is not None:
        pend.wait()
    if link is not None and n_in:
        link.wait("K")
        link.wait("V")
    if n_out > Tl:
        send_halo()
    a = swa.swa_attention_bthd(sq[:, s:e], kb[:, Hh - n_in:], vb[:, Hh - n_in:], window=W, key_pos0=s - n_in)
    if link is not None and n_in:
        link.release("K")
        link.release("V")
    ho.drain()
    if link is not None:
        link.drain()
    flags.append(float(torch.equal(a, ref_a[:, s:e])))
    errs.append(_err(ref_a[:, s:e], a))
    ft = torch.tensor(flags, device=dev)
    et = torch.tensor(errs, device=dev)
    dist.all_reduce(ft, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(et, op=dist.ReduceOp.MAX, group=group)
    f, er = ft.tolist(), et.tolist()
    return {"gdn_o_equal": bool(f[0]), "gdn_state_equal": bool(f[1]), "swa_equal": bool(f[2]),
            "gdn_o_err": er[0], "gdn_state_err": er[1], "swa_err": er[2], "T": T, "world": world, "transport": transport}


def sharded_parity_check(T: int = 32768, num_layers: int = 8, seed: int = 0, group=None, config=None) -> dict:
    """Every rank builds the same HybridDecoder (3B mixer dims, `num_layers` layers in the model's 1 SWA : 3 GDN
    pattern, mixers only) and the same inputs; the ranks run `sharded_prefill` over NCCL.  The last rank also runs
    (a) the same token ranges ONE AFTER THE OTHER through one cache on its own GPU -- the identical computation
    without the network: the sharded run must match it bit for bit (keys "out", "state", "conv", "kv": error ratios,
    gate 0) -- and (b) the whole sequence in one call ("oneshot_*").  The hot-path operators are bit-identical
    between (a) and (b) (`operator_parity_check`); the only thing that may differ is cuBLAS, should it pick a
    different kernel for the projections at a different number of rows.  Measured on B200 (profiles/r02_summary.md):
    0.0 everywhere at P = 2 -- the sharded prefill reproduces the one-call prefill bit for bit.  The DeltaNet state
    is handed over in fp32.
    Returns the dict on the last rank, {} elsewhere."""
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
        got = torch.cat(parts, 1)
        # (a) the same ranges, one after the other, on this GPU
        ccache = dec.allocate_inference_cache(1, state_dtype=torch.float32)
        outs = []
        for r in range(world):
            rs, re_ = shard_range(T, world, r)
            o_r, ccache = _local_range(dec, x[:, rs:re_], pos[:, :, rs:re_], rs, re_, ccache)
            outs.append(o_r)
        chunked = torch.cat(outs, 1)
        res = {"out": _err(chunked, got), "state": 0.0, "conv": 0.0, "kv": 0.0, "ints_equal": True,
               "bit_identical": bool(torch.equal(chunked, got)), "T": T, "layers": num_layers, "world": world}
        for a, b in zip(cache.layers, ccache.layers):
            if a.is_sliding:
                res["ints_equal"] &= (a.size == b.size and a.cumulative_length == b.cumulative_length)
                res["kv"] = max(res["kv"], _err(b.keys, a.keys), _err(b.values, a.values))
            else:
                res["ints_equal"] &= a.seq_len == b.seq_len
                res["state"] = max(res["state"], _err(b.recurrent_state, a.recurrent_state))
                res["conv"] = max(res["conv"], _err(b.conv_state_v, a.conv_state_v))
        # (b) one call over the whole sequence
        rcache = dec.allocate_inference_cache(1, state_dtype=torch.float32)
        ref = dec(x, position_ids=pos, past_key_values=rcache)
        res["oneshot_out"] = _err(ref, got)
        res["oneshot_out_chunked_on_one_gpu"] = _err(ref, chunked)
        res["oneshot_state"] = max(_err(b.recurrent_state, a.recurrent_state)
                                   for a, b in zip(cache.layers, rcache.layers) if not a.is_sliding)
        del rcache, ref, ccache, chunked, got
    del dec, x, cache, out, parts
    torch.cuda.empty_cache()
    return res


def _local_range(decoder, x_local, pos_local, start, end, cache):
    """One token range of a prefill through `cache` on this device (no hand-off): what a rank of `sharded_prefill`
    computes between its receive and its send."""
    from .modeling import mrope_select
    cache_position = torch.arange(start, end, device=x_local.device)
    cos, sin = decoder.rotary_emb(x_local, pos_local)
    cos, sin = mrope_select(cos, sin, decoder.config.rope_scaling["mrope_section"])
    h = x_local
    for layer in decoder.layers:
        if decoder.mixers_only:
            h = h + layer.self_attn(hidden_states=layer.input_layernorm(h), past_key_values=cache,
                                    cache_position=cache_position, position_embeddings=(cos, sin))[0]
        else:
            h = layer(h, past_key_values=cache, cache_position=cache_position, position_embeddings=(cos, sin))[0]
    return decoder.norm(h), cache
