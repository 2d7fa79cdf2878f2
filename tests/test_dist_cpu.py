"""CPU (gloo, world_size 2 and 4) test of the sequence-chunk hand-off logic in infinitevl_b200/dist.py:
the product's cache classes + send/recv protocol + wavefront loop, with the oracle standing in for the
CUDA mixers.  The rank-concatenated output must equal the single-process run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from infinitevl_b200.cache import StaticCachePrealloc
from infinitevl_b200.dist import shard_range, sharded_layer_loop
from infinitevl_b200.modeling import HybridTextConfig
from oracle import err_ratio, gdn_mixer_ref, mrope_cos_sin_ref, swa_mixer_ref

HID, H, K, V, HQ, HKV, D, W = 32, 2, 16, 32, 2, 1, 16, 24
# every rank gets 128 tokens: the oracle mixer switches to the token recurrence at q_len <= 64 like the model does,
# and mixing the two fp32 forms would blur the 1e-5 / 1e-6 comparisons below
LAYER_TYPES = ["sliding_attention", "linear_attention", "linear_attention", "sliding_attention", "linear_attention"]


def _cfg():
    return HybridTextConfig(hidden_size=HID, num_attention_heads=HQ, num_key_value_heads=HKV, sliding_window=W,
                            num_linear_heads=H, num_linear_key_value_heads=H, linear_head_dim=K, expand_v=2,
                            num_hidden_layers=len(LAYER_TYPES), layer_types=list(LAYER_TYPES))


def _params():
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g) * 0.2
    ps = []
    for lt in LAYER_TYPES:
        if lt == "linear_attention":
            ps.append({"q_proj.weight": r(H * K, HID), "k_proj.weight": r(H * K, HID), "v_proj.weight": r(H * V, HID),
                       "a_proj.weight": r(H, HID), "b_proj.weight": r(H, HID), "g_proj.weight": r(H * V, HID),
                       "o_proj.weight": r(HID, H * V), "A_log": torch.log(torch.rand(H, generator=g) * 2 + 0.1),
                       "dt_bias": r(H), "q_conv1d.weight": r(H * K, 1, 4), "k_conv1d.weight": r(H * K, 1, 4),
                       "v_conv1d.weight": r(H * V, 1, 4), "o_norm.weight": torch.ones(V)})
        else:
            ps.append({"q_proj.weight": r(HQ * D, HID), "q_proj.bias": r(HQ * D), "k_proj.weight": r(HKV * D, HID),
                       "k_proj.bias": r(HKV * D), "v_proj.weight": r(HKV * D, HID), "v_proj.bias": r(HKV * D),
                       "o_proj.weight": r(HID, HQ * D)})
    return ps


class _SwaAdapter:
    def __init__(self, layer):
        self.layer = layer

    def update(self, k, v):
        return self.layer.update(k, v)


def _layer_fns(params, start, end):
    pos = torch.arange(start, end)[None, None].expand(3, 1, -1)
    cos, sin = mrope_cos_sin_ref(pos, D, 1e6)
    fns = []
    for lt, p in zip(LAYER_TYPES, params):
        if lt == "linear_attention":
            def fn(h, cache, i, p=p):
                conv, S = cache.update(i, cache_kwargs={"op": "get"})
                y, nconv, nS = gdn_mixer_ref(h, p, conv_cache=None if conv[0] is None else conv, state=S, H=H, K=K, V=V)
                cache.update(i, conv_state=nconv, recurrent_state=nS, cache_kwargs={"op": "set", "delta_len": h.shape[1]})
                return h + y
        else:
            def fn(h, cache, i, p=p):
                y = swa_mixer_ref(h, p, cos, sin, cache=_SwaAdapter(cache.layers[i]), Hq=HQ, Hkv=HKV, D=D, window=W,
                                  mrope_section=(2, 3, 3))
                return h + y
        fns.append(fn)
    return fns


def _single(x):
    cache = StaticCachePrealloc(config=_cfg(), batch_size=1, dtype=torch.float32, zero_init=True)
    out = sharded_layer_loop(_layer_fns(_params(), 0, x.shape[1]), x, cache, 0, 0, 1)
    return out, cache


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    T = 128 * world
    x = torch.randn(1, T, HID, generator=torch.Generator().manual_seed(1))
    s, e = shard_range(T, world, rank)
    cache = StaticCachePrealloc(config=_cfg(), batch_size=1, dtype=torch.float32, zero_init=True)
    out = sharded_layer_loop(_layer_fns(_params(), s, e), x[:, s:e], cache, s, rank, world)
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    if rank == world - 1:
        ref, rcache = _single(x)
        ok = err_ratio(ref, torch.cat(gathered, 1)) < 1e-5
        # the last rank's cache == the single-process cache after the whole sequence
        for a, b in zip(cache.layers, rcache.layers):
            if a.is_sliding:
                ok &= a.size == b.size and a.cumulative_length == b.cumulative_length
                ok &= err_ratio(b.keys, a.keys) < 1e-5
            else:
                ok &= a.seq_len == b.seq_len and err_ratio(b.recurrent_state, a.recurrent_state) < 1e-5
                ok &= err_ratio(b.conv_state_v, a.conv_state_v) < 1e-6
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])   # 4: middle ranks both receive and send
def test_hand_off_equals_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_range_rules():
    assert shard_range(131072, 8, 3) == (49152, 65536)
    assert shard_range(256, 2, 1) == (128, 256)
    with pytest.raises(ValueError):
        shard_range(200, 2, 0)
