"""Decoder-level streaming: 256-token frames through HybridDecoder with the inference cache, eagerly and as ONE
CUDA graph replayed per frame -- what inference_examples/demo_streaming_inference.py:453-489 does with the whole
model (BASELINE.json config 3).

What is pinned here:
  * eager streaming == the fp32 oracle streaming the same frames through the reference's cache rules;
  * a captured forward replays bit-identically to eager calls and allocates nothing (memory flat over 64 frames);
  * the reference's cache keeps its window bookkeeping in Python integers, which a graph replay cannot advance
    (SURVEY.md 3.2 / appendix C): with `StaticSlidingWindowLayerPrealloc` in its reference mode a replayed frame
    sees the keys cached at capture time plus itself.  The test states that behaviour exactly (replay == eager
    with the counters pinned at their capture-time values); the ring mode keeps its counters on
    the device and replays like eager streaming -- covered by test_ring_cache_graph_stream_equals_eager.
"""
import pytest
import torch

from oracle import SlidingWindowCacheRef, err_ratio, hybrid_decoder_ref
from test_modules_gpu import _init, _params, gen

pytestmark = pytest.mark.gpu

FRAME = 256
WINDOW = 1024


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from infinitevl_b200 import modeling
    return modeling


def _decoder(M, layers=4, window=WINDOW):
    cfg = M.HybridTextConfig(num_hidden_layers=layers, sliding_window=window)
    return cfg, _init(M.HybridDecoder(cfg), 41).bfloat16().cuda().eval()


def _frames(n, first=300, seed=42):
    """A 300-token prompt followed by n frames of 256 tokens; positions advance by one per token."""
    T = first + n * FRAME
    x = torch.randn(1, T, 2048, generator=gen(seed)).bfloat16()
    cuts = [(0, first)] + [(first + i * FRAME, first + (i + 1) * FRAME) for i in range(n)]
    return x, cuts


@pytest.mark.parametrize("ring", [False, True])
def test_eager_stream_matches_oracle(M, ring):
    cfg, dec = _decoder(M)
    p = _params(dec)
    x, cuts = _frames(6)
    cache = dec.allocate_inference_cache(1, ring=ring)
    rcaches = [SlidingWindowCacheRef(WINDOW) if lt == "sliding_attention" else {} for lt in cfg.layer_types]
    outs, refs = [], []
    for a, b in cuts:
        pos = torch.arange(a, b)[None, None].expand(3, 1, -1)
        outs.append(dec(x[:, a:b].cuda(), position_ids=pos.cuda(), past_key_values=cache,
                        cache_position=torch.arange(a, b, device="cuda")))
        refs.append(hybrid_decoder_ref(x[:, a:b], p, cfg.layer_types, pos, caches=rcaches, window=WINDOW,
                                       proj_dtype=torch.bfloat16))
        for st in rcaches:   # the reference's cache stores S in the model dtype (Appendix B point 11)
            if isinstance(st, dict) and st.get("state") is not None:
                st["state"] = st["state"].bfloat16().float()
    assert err_ratio(torch.cat(refs, 1), torch.cat(outs, 1).float().cpu()) < 2e-2
    sw = cache.layers[0]
    assert sw.size == rcaches[0].size == WINDOW - 1 and sw.cumulative_length == rcaches[0].cumulative_length
    assert cache.layers[1].seq_len == cuts[-1][1]


def _pin(cache, snap):
    """Put the Python-side bookkeeping of every cache layer back to `snap` (what a graph replay sees)."""
    for layer, s in zip(cache.layers, snap):
        if layer.is_sliding:
            layer.size, layer.cumulative_length = s
            layer.keys = layer._buf_keys[:, :, :layer.size, :]
            layer.values = layer._buf_values[:, :, :layer.size, :]
        else:
            layer.seq_len = s


def _snapshot(cache):
    return [(l.size, l.cumulative_length) if l.is_sliding else l.seq_len for l in cache.layers]


def test_graph_replay_of_the_decoder_forward(M):
    """Capture HybridDecoder.forward for one frame with the cache, replay it for 64 frames."""
    from infinitevl_b200 import ops
    cfg, dec = _decoder(M)
    n_frames = 64
    x, cuts = _frames(n_frames + 2)
    xs = x.cuda()

    def run_eager(cache, a, b, xin=None):
        pos = torch.arange(a, b, device="cuda")[None, None].expand(3, 1, -1)
        return dec(xs[:, a:b] if xin is None else xin, position_ids=pos, past_key_values=cache,
                   cache_position=torch.arange(a, b, device="cuda"))

    caches = [dec.allocate_inference_cache(1, ring=False) for _ in range(2)]   # [0]: graph, [1]: eager twin
    for c in caches:
        for a, b in cuts[:3]:    # prompt + two frames, eagerly
            run_eager(c, a, b)
    snap = _snapshot(caches[0])
    static_x = torch.empty(1, FRAME, 2048, dtype=torch.bfloat16, device="cuda")
    static_pos = torch.empty(3, 1, FRAME, dtype=torch.long, device="cuda")
    static_cp = torch.empty(FRAME, dtype=torch.long, device="cuda")

    def fwd():
        return dec(static_x, position_ids=static_pos, past_key_values=caches[0], cache_position=static_cp)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    a0, b0 = cuts[3]
    with torch.cuda.stream(side):
        ops.gdn_workspace(1, FRAME, cfg.num_linear_heads, "cuda")   # scratch exists before capture
        static_x.copy_(xs[:, a0:b0]); static_cp.copy_(torch.arange(a0, b0, device="cuda"))
        static_pos.copy_(static_cp[None, None].expand(3, 1, -1))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            static_out = fwd()
    torch.cuda.current_stream().wait_stream(side)
    _pin(caches[0], snap)   # capture ran the Python bookkeeping once without running any kernel
    mem = []
    for i, (a, b) in enumerate(cuts[3:3 + n_frames]):
        static_x.copy_(xs[:, a:b]); static_cp.copy_(torch.arange(a, b, device="cuda"))
        static_pos.copy_(static_cp[None, None].expand(3, 1, -1))
        graph.replay()
        got = static_out.clone()
        _pin(caches[1], snap)   # the eager twin with the counters a replay sees
        want = run_eager(caches[1], a, b)
        torch.cuda.synchronize()
        assert torch.equal(got, want), f"frame {i}"
        mem.append(torch.cuda.memory_allocated())
    # GDN state and conv tails advanced in place, identically
    for lg, le in zip(caches[0].layers, caches[1].layers):
        if not lg.is_sliding:
            assert torch.equal(lg.recurrent_state, le.recurrent_state)
            assert torch.equal(lg.conv_state_v, le.conv_state_v)
        else:
            assert torch.equal(lg._buf_keys, le._buf_keys)
    assert torch.isfinite(static_out).all()
    assert mem[10] == mem[-1], "streaming must not grow memory (BASELINE.json config 3)"


def test_ring_cache_graph_stream_equals_eager(M):
    """Ring-buffer window cache: the token counter the kernels address by lives on the device, so
    ONE captured forward replayed for 64 frames of 256 tokens equals eager streaming of the same frames bit for bit
    -- the window really slides (the 1023-token window wraps the 1279-slot ring many times) -- with flat memory.
    Then decode steps continue from the replayed cache."""
    from infinitevl_b200 import ops
    cfg, dec = _decoder(M)
    n_frames = 64
    x, cuts = _frames(n_frames + 2)
    xs = x.cuda()

    def run_eager(cache, a, b):
        pos = torch.arange(a, b, device="cuda")[None, None].expand(3, 1, -1)
        return dec(xs[:, a:b], position_ids=pos, past_key_values=cache, cache_position=torch.arange(a, b, device="cuda"))

    caches = [dec.allocate_inference_cache(1, ring=True) for _ in range(2)]
    assert caches[0].layers[0].is_ring and caches[0].layers[0].R == WINDOW - 1 + 256
    for c in caches:
        for a, b in cuts[:3]:
            run_eager(c, a, b)
    static_x = torch.empty(1, FRAME, 2048, dtype=torch.bfloat16, device="cuda")
    static_pos = torch.empty(3, 1, FRAME, dtype=torch.long, device="cuda")
    static_cp = torch.empty(FRAME, dtype=torch.long, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    snap = caches[0].layers[0]._state.clone()
    with torch.cuda.stream(side):
        ops.gdn_workspace(1, FRAME, cfg.num_linear_heads, "cuda")
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            static_out = dec(static_x, position_ids=static_pos, past_key_values=caches[0], cache_position=static_cp)
    torch.cuda.current_stream().wait_stream(side)
    assert torch.equal(caches[0].layers[0]._state, snap)   # capturing ran no kernel: the device counter did not move
    mem = []
    for i, (a, b) in enumerate(cuts[3:3 + n_frames]):
        static_x.copy_(xs[:, a:b]); static_cp.copy_(torch.arange(a, b, device="cuda"))
        static_pos.copy_(static_cp[None, None].expand(3, 1, -1))
        graph.replay()
        got = static_out.clone()
        want = run_eager(caches[1], a, b)
        torch.cuda.synchronize()
        assert torch.equal(got, want), f"frame {i}"
        mem.append(torch.cuda.memory_allocated())
    assert mem[10] == mem[-1], "streaming must not grow memory (BASELINE.json config 3)"
    caches[0].sync_from_device()
    for lg, le in zip(caches[0].layers, caches[1].layers):
        if lg.is_sliding:
            assert (lg.size, lg.cumulative_length) == (le.size, le.cumulative_length) == (WINDOW - 1, cuts[2 + n_frames][1])
            assert torch.equal(lg.keys, le.keys) and torch.equal(lg.values, le.values)
        else:
            assert torch.equal(lg.recurrent_state, le.recurrent_state)
    # decode continues from the replayed cache exactly as from the eager one
    T0 = cuts[2 + n_frames][1]
    tok = torch.randn(1, 8, 2048, generator=gen(77)).bfloat16().cuda()
    for j in range(8):
        outs = []
        for c in caches:
            pos = torch.full((3, 1, 1), T0 + j, device="cuda")
            outs.append(dec(tok[:, j:j + 1], position_ids=pos, past_key_values=c,
                            cache_position=torch.tensor([T0 + j], device="cuda")))
        assert torch.equal(outs[0], outs[1]), j
