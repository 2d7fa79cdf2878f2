"""CPU tests: the product's cache classes against the reference's own trace (bit-exact integer /
copy semantics, SURVEY.md rows a-C1..a-C3) and the demo's clone protocol."""
import os

import numpy as np
import pytest
import torch

from infinitevl_b200.cache import StaticCachePrealloc, StaticLinearLayerPrealloc, StaticSlidingWindowLayerPrealloc
from infinitevl_b200.modeling import HybridTextConfig


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=True)


def test_swa_layer_matches_reference_trace(golden_dir):
    z = _load(golden_dir, "ref_swa_cache.npz")
    cfg = HybridTextConfig(sliding_window=int(z["window"]), num_key_value_heads=1, num_attention_heads=2,
                           hidden_size=8)
    layer = StaticSlidingWindowLayerPrealloc(config=cfg, batch_size=1, dtype=torch.float32, zero_init=True)
    assert layer.capacity == 7 and layer.is_sliding
    base = 0
    for step, n in enumerate(z["steps"]):
        n = int(n)
        kk = torch.arange(base, base + n, dtype=torch.float32)[None, None, :, None].expand(1, 1, n, 4).contiguous()
        buf_ptr = layer._buf_keys.data_ptr()
        fk, fv = layer.update(kk, kk * 2)
        base += n
        kv_len, kv_off = layer.get_mask_sizes(torch.arange(n))
        assert [n, kv_len, kv_off, layer.size, layer.cumulative_length, fk.shape[-2]] == [int(x) for x in z["sizes"][step]]
        assert np.array_equal(fk[0, 0, :, 0].numpy(), z["full"][step])
        assert np.array_equal(layer.keys[0, 0, :, 0].numpy(), z["tail"][step])
        assert layer._buf_keys.data_ptr() == buf_ptr and layer.keys.data_ptr() == buf_ptr  # fixed addresses
    with pytest.raises(ValueError):
        layer.crop(3)  # forbidden once the window has filled (std:192-194)
    with pytest.raises(ValueError):
        layer.update(torch.zeros(2, 1, 1, 4), torch.zeros(2, 1, 1, 4))
    with pytest.raises(RuntimeError):
        layer.batch_repeat_interleave(2)


def test_linear_layer_matches_reference(golden_dir):
    z = _load(golden_dir, "ref_linear_cache.npz")
    cfg = HybridTextConfig(num_linear_heads=2, num_linear_key_value_heads=2, linear_head_dim=4, expand_v=2, conv_size=4)
    lin = StaticLinearLayerPrealloc(config=cfg, batch_size=1, dtype=torch.bfloat16, zero_init=True)
    assert [list(lin.conv_state_q.shape) + [0], list(lin.conv_state_v.shape) + [0], list(lin.recurrent_state.shape)] \
        == [list(x) for x in z["shapes"]]
    first = lin.update(cache_kwargs={"op": "get"})
    assert first == ((None, None, None), None)
    assert lin.update(cache_kwargs={"op": "get"})[1] is lin.recurrent_state
    t = lambda n: torch.from_numpy(z[n])
    ptr = lin.recurrent_state.data_ptr()
    lin.update(conv_state=(t("cq"), t("ck"), t("cv")), recurrent_state=t("state_in"),
               cache_kwargs={"op": "set", "delta_len": 7})
    (cq, ck, cv), st = lin.update(cache_kwargs={"op": "get"})
    assert st.data_ptr() == ptr and st.dtype == torch.bfloat16
    assert torch.equal(st.float(), t("state_out")) and torch.equal(cq.float(), t("cq_out")) and torch.equal(cv.float(), t("cv_out"))
    assert lin.get_seq_length() == int(z["seq_len"]) and lin.get_mask_sizes(torch.arange(3)) == (10, 0)
    with pytest.raises(RuntimeError):
        lin.update(recurrent_state=torch.zeros(1, 2, 4, 9), cache_kwargs={"op": "set"})
    with pytest.raises(RuntimeError):
        lin.update(conv_state=(torch.zeros(1, 9, 4), None, None), cache_kwargs={"op": "set"})


def test_cache_container_and_demo_clone_protocol():
    """inference_examples/demo_streaming_inference.py:111-160 deep-copies these attributes layer by layer."""
    cfg = HybridTextConfig(num_hidden_layers=8, sliding_window=16)
    cache = StaticCachePrealloc(config=cfg, batch_size=1, dtype=torch.bfloat16, zero_init=True)
    assert len(cache.layers) == 8
    assert [l.is_sliding for l in cache.layers] == [True, False, False, False, True, False, False, False]
    total = sum(l._buf_keys.numel() * 2 for l in cache.layers if l.is_sliding) + \
        sum(l.recurrent_state.numel() for l in cache.layers if not l.is_sliding)
    assert total == 2 * 2 * 2 * 15 * 128 + 6 * 16 * 128 * 256
    sw, lin = cache.layers[0], cache.layers[1]
    for attr in ("_buf_keys", "_buf_values", "keys", "values", "size", "cumulative_length", "capacity", "sliding_window"):
        assert hasattr(sw, attr)
    for attr in ("recurrent_state", "conv_state_q", "conv_state_k", "conv_state_v", "seq_len", "start"):
        assert hasattr(lin, attr)
    k = torch.randn(1, 2, 5, 128).bfloat16()
    fk, _ = cache.update(0, k, k)
    assert fk.shape[-2] == 5 and cache.get_seq_length(0) == 5
    # clone as the demo does: copy buffers, re-slice views, copy counters
    clone = StaticCachePrealloc(config=cfg, batch_size=1, dtype=torch.bfloat16, zero_init=True)
    src, dst = cache.layers[0], clone.layers[0]
    dst._buf_keys.copy_(src._buf_keys); dst._buf_values.copy_(src._buf_values)
    dst.size, dst.cumulative_length = src.size, src.cumulative_length
    dst.keys, dst.values = dst._buf_keys[:, :, :dst.size, :], dst._buf_values[:, :, :dst.size, :]
    k2 = torch.randn(1, 2, 1, 128).bfloat16()
    a, _ = cache.update(0, k2, k2)
    b, _ = clone.update(0, k2, k2)
    assert torch.equal(a, b)
    # full-config memory: 27 x (1 MiB + 64 KiB) + 9 x 8.4 MB in bf16 (SURVEY.md a-C3)
    big = StaticCachePrealloc(config=HybridTextConfig(), batch_size=1, dtype=torch.bfloat16, device="meta")
    nbytes = 0
    for l in big.layers:
        if l.is_sliding:
            nbytes += 2 * l._buf_keys.numel() * 2
        else:
            nbytes += 2 * (l.recurrent_state.numel() + l.conv_state_q.numel() + l.conv_state_k.numel() + l.conv_state_v.numel())
    assert nbytes == 27 * (16 * 128 * 256 * 2 + (2048 + 2048 + 4096) * 4 * 2) + 9 * 2 * 2 * 8191 * 128 * 2


def test_position_id_normalisation_matches_reference_rules():
    """InfiniteVLTextModel.forward (std:1512-1525): default positions from cache_position, 2-D ids broadcast to the
    three M-RoPE rows, the 4-row packed form split into text row + M-RoPE rows; a text row that restarts marks packed
    sequences and turns into cu_seqlens (one row only)."""
    import pytest
    from infinitevl_b200.modeling import normalize_position_ids
    cp = torch.arange(5, 9)
    p, t = normalize_position_ids(None, cp, 2)
    assert p.shape == (3, 2, 4) and t is None and torch.equal(p[1, 1], cp)
    ids = torch.arange(8).view(2, 4)
    p, t = normalize_position_ids(ids, cp, 2)
    assert p.shape == (3, 2, 4) and torch.equal(p[2], ids)
    four = torch.stack([ids, ids + 1, ids + 2, ids + 3])
    p, t = normalize_position_ids(four, cp, 2)
    assert torch.equal(t, ids) and torch.equal(p, four[1:])
    packed = four.clone()
    packed[0, 0] = torch.tensor([0, 1, 0, 1])
    from infinitevl_b200.modeling import packed_cu_seqlens
    _, text = normalize_position_ids(packed, cp, 2)
    with pytest.raises(ValueError):
        packed_cu_seqlens(text)                     # two rows: packed sequences come as one row
    one = torch.tensor([[0, 1, 2, 0, 1, 0, 1, 2, 3]])
    assert packed_cu_seqlens(one).tolist() == [0, 3, 5, 9]
    assert packed_cu_seqlens(torch.arange(7)[None]) is None
    with pytest.raises(ValueError):
        normalize_position_ids(torch.zeros(2, 2, 4, dtype=torch.long), cp, 2)


def test_demo_import_alias_resolves_to_the_b200_caches():
    """demo_streaming_inference.py:36-44 imports the cache classes from a module named modeling_qwen2_5_vl;
    infinitevl_b200/compat on sys.path provides it."""
    import importlib
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "infinitevl_b200", "compat")
    sys.path.insert(0, compat)
    try:
        mod = importlib.import_module("modeling_qwen2_5_vl")
        assert mod.StaticCachePrealloc is StaticCachePrealloc
        assert mod.StaticSlidingWindowLayerPrealloc is StaticSlidingWindowLayerPrealloc
        assert mod.StaticLinearLayerPrealloc is StaticLinearLayerPrealloc
    finally:
        sys.path.remove(compat)
        sys.modules.pop("modeling_qwen2_5_vl", None)


def test_ring_layer_equals_reference_layer_on_the_reference_trace():
    """RingSlidingWindowLayer against the reference-semantics layer on the trace the golden fixture
    was made with: same returned [tail ; new] tensors, same `keys` / `_buf_keys` views, same integers -- and it is an
    HF Cache layer like the reference's classes."""
    import torch
    from infinitevl_b200.cache import RingSlidingWindowLayer, StaticCachePrealloc, StaticSlidingWindowLayerPrealloc
    from infinitevl_b200.modeling import HybridTextConfig
    cfg = HybridTextConfig(num_hidden_layers=4, sliding_window=8, num_key_value_heads=1, num_attention_heads=2, hidden_size=8)
    cfg.head_dim = 4
    mk = lambda ring: StaticSlidingWindowLayerPrealloc(config=cfg, batch_size=1, dtype=torch.float32, zero_init=True,
                                                      ring=ring, max_append=4)
    ref, rng = mk(False), mk(True)
    assert isinstance(rng, RingSlidingWindowLayer) and not isinstance(ref, RingSlidingWindowLayer) and rng.R == 11
    base = 0
    for n in [3, 1, 1, 5, 1, 9, 2, 1, 1, 20, 1]:
        kk = torch.arange(base, base + n, dtype=torch.float32)[None, None, :, None].expand(1, 1, n, 4).contiguous()
        a, b = ref.update(kk, kk * 2), rng.update(kk, kk * 2)
        base += n
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        assert (ref.size, ref.cumulative_length) == (rng.size, rng.cumulative_length)
        assert torch.equal(ref.keys, rng.keys) and torch.equal(ref.values, rng.values)
        assert torch.equal(ref._buf_keys[:, :, :ref.size], rng._buf_keys[:, :, :rng.size])
        assert ref.get_mask_sizes(torch.arange(n)) == rng.get_mask_sizes(torch.arange(n))
    assert int(rng._state[0]) == base
    # the demo's clone protocol (demo_streaming_inference.py:111-160) on ring layers
    dst = mk(True)
    dst.size, dst.cumulative_length = rng.size, rng.cumulative_length
    L = rng.size
    dst._buf_keys[:, :, :L, :].copy_(rng._buf_keys[:, :, :L, :])
    dst._buf_values[:, :, :L, :].copy_(rng._buf_values[:, :, :L, :])
    dst.keys, dst.values = dst._buf_keys[:, :, :L, :], dst._buf_values[:, :, :L, :]
    kk = torch.full((1, 1, 2, 4), 99.0)
    a, b = rng.update(kk, kk), dst.update(kk, kk)
    assert torch.equal(a[0], b[0]) and torch.equal(rng.keys, dst.keys) and int(dst._state[0]) == int(rng._state[0])
    # the other order (dist.recv_cache_layer: rows first, integers afterwards), into a layer with its own history
    dst2 = mk(True)
    dst2.update(torch.ones(1, 1, 5, 4), torch.ones(1, 1, 5, 4))
    L = rng.size
    dst2._buf_keys[:, :, :L, :].copy_(rng._buf_keys[:, :, :L, :])
    dst2._buf_values[:, :, :L, :].copy_(rng._buf_values[:, :, :L, :])
    dst2.size, dst2.cumulative_length = rng.size, rng.cumulative_length
    assert torch.equal(dst2.keys, rng.keys) and torch.equal(dst2.values, rng.values)
    a, b = rng.update(kk, kk), dst2.update(kk, kk)
    assert torch.equal(a[0], b[0]) and torch.equal(rng.keys, dst2.keys) and int(dst2._state[0]) == int(rng._state[0])
    # snapshot / restore
    again = mk(True)
    again.load_state_dict(rng.state_dict())
    assert torch.equal(again.keys, rng.keys) and again.cumulative_length == rng.cumulative_length
    with pytest.raises(ValueError):
        rng.crop(2)     # the window is full: cropping is forbidden, as in the reference
    try:
        from transformers.cache_utils import Cache, CacheLayerMixin
    except Exception:  # noqa: BLE001
        return
    cache = StaticCachePrealloc(config=cfg, batch_size=1)
    assert isinstance(cache, Cache) and all(isinstance(l, CacheLayerMixin) for l in cache.layers)


@pytest.mark.parametrize("ring", [False, True])
def test_cache_snapshot_on_disk(tmp_path, ring):
    """the whole inference cache as one safetensors file -- windows in logical order, recurrent / conv
    states, the integer bookkeeping -- and back: the restored cache is equal and continues the stream identically."""
    import torch
    from infinitevl_b200.cache import StaticCachePrealloc
    from infinitevl_b200.modeling import HybridTextConfig
    cfg = HybridTextConfig(num_hidden_layers=4, sliding_window=8, num_key_value_heads=1, num_attention_heads=2,
                           hidden_size=8, num_linear_heads=2, num_linear_key_value_heads=2, linear_head_dim=4)
    cfg.head_dim = 4

    def feed(c, n, seed):
        g = torch.Generator().manual_seed(seed)
        kv = torch.randn(1, 1, n, 4, generator=g)
        c.update(0, kv, kv * 2)
        for li in (1, 2, 3):
            lin = c.layers[li]
            lin.update(conv_state=None, recurrent_state=None, cache_kwargs={"op": "get"})
            mk = lambda t: torch.randn(t.shape, generator=g)
            lin.update(conv_state=(mk(lin.conv_state_q), mk(lin.conv_state_k), mk(lin.conv_state_v)),
                       recurrent_state=mk(lin.recurrent_state), cache_kwargs={"op": "set", "delta_len": n})

    a = StaticCachePrealloc(config=cfg, batch_size=1, ring=ring, zero_init=True)
    for i, n in enumerate([3, 5, 2, 9]):
        feed(a, n, i)
    path = str(tmp_path / "cache.safetensors")
    a.save(path)
    b = StaticCachePrealloc(config=cfg, batch_size=1, ring=ring, zero_init=True).load(path)

    def same(x, y):
        for la, lb in zip(x.layers, y.layers):
            if la.is_sliding:
                assert (la.size, la.cumulative_length) == (lb.size, lb.cumulative_length)
                assert torch.equal(la.keys, lb.keys) and torch.equal(la.values, lb.values)
            else:
                assert (la.seq_len, la.start) == (lb.seq_len, lb.start)
                assert torch.equal(la.recurrent_state, lb.recurrent_state) and torch.equal(la.conv_state_v, lb.conv_state_v)
    same(a, b)
    feed(a, 4, 99)
    feed(b, 4, 99)
    same(a, b)
    assert a.layers[0].cumulative_length == 23
