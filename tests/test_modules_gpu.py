"""GPU parity tests of the element-wise kernels, the two mixers and one hybrid decoder block
against the fp32 oracle (BASELINE.json config 1: [SWA, GDN, GDN, GDN] at the 3B dims, T = 1024)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import (SlidingWindowCacheRef, err_ratio, gdn_mixer_ref, hybrid_decoder_ref, mrope_apply_ref,
                    mrope_cos_sin_ref, rmsnorm_gated_ref, short_conv_ref, swa_mixer_ref)
from oracle.gdn import gdn_gate_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from infinitevl_b200 import modeling
    return modeling


def gen(seed):
    return torch.Generator().manual_seed(seed)


# ---------------------------------------------------------------- element-wise kernels ---------
@pytest.mark.parametrize("T", [1, 3, 4, 50, 1000])
@pytest.mark.parametrize("with_cache", [False, True])
def test_short_conv(M, T, with_cache):
    g = gen(T)
    D = 2048
    x = torch.randn(2, T, D, generator=g).bfloat16()
    w = (torch.randn(D, 1, 4, generator=g) * 0.5).bfloat16()
    cache = torch.randn(2, D, 4, generator=g).bfloat16() if with_cache else None
    ry, rs = short_conv_ref(x, w, cache)
    y, s = M.short_conv_silu(x.cuda(), w.cuda(), None if cache is None else cache.cuda(), output_final_state=True)
    assert err_ratio(ry, y.float().cpu()) < 4e-3   # bf16 output rounding only
    assert torch.equal(s.float().cpu(), rs)        # the carried tail is a copy of inputs: bit-exact


def test_short_conv_streaming_equals_one_shot(M):
    g = gen(7)
    x = torch.randn(1, 300, 4096, generator=g).bfloat16().cuda()
    w = (torch.randn(4096, 1, 4, generator=g) * 0.5).bfloat16().cuda()
    y, s = M.short_conv_silu(x, w, None, output_final_state=True)
    outs, c = [], None
    for a, b in ((0, 257), (257, 258), (258, 259), (259, 300)):
        yi, c = M.short_conv_silu(x[:, a:b], w, c, output_final_state=True)
        outs.append(yi)
    assert torch.equal(torch.cat(outs, 1), y) and torch.equal(c, s)


def test_gates_and_gated_norm(M):
    g = gen(3)
    a = torch.randn(1, 500, 16, generator=g).bfloat16()
    b = torch.randn(1, 500, 16, generator=g).bfloat16()
    A_log = torch.log(torch.empty(16).uniform_(0.01, 16, generator=g))
    dt_bias = torch.randn(16, generator=g)
    rg, rb = gdn_gate_ref(a, b, A_log, dt_bias)
    dg, db = M.gdn_gates(a.cuda(), b.cuda(), A_log.cuda(), dt_bias.cuda())
    assert dg.dtype == torch.float32 and err_ratio(rg, dg.cpu()) < 1e-5
    assert err_ratio(rb, db.float().cpu()) < 3e-3
    o = torch.randn(700, 16, 256, generator=g).bfloat16()
    gate = torch.randn(700, 16, 256, generator=g).bfloat16()
    w = torch.rand(256, generator=g).bfloat16()
    ref = rmsnorm_gated_ref(o, gate, w, eps=1e-5)
    out = M.rmsnorm_gated(o.cuda(), gate.cuda(), w.cuda(), 1e-5)
    assert err_ratio(ref, out.float().cpu()) < 4e-3


def test_mrope_bit_exact(M):
    g = gen(5)
    T = 333
    pos = torch.stack([torch.arange(T), torch.randint(0, 50, (T,), generator=g), torch.randint(0, 50, (T,), generator=g)])[:, None]
    cos, sin = mrope_cos_sin_ref(pos, 128, 1e6, out_dtype=torch.bfloat16)
    q = torch.randn(1, 16, T, 128, generator=g).bfloat16()
    k = torch.randn(1, 2, T, 128, generator=g).bfloat16()
    rq, rk = mrope_apply_ref(q, k, cos, sin, (16, 24, 24))
    cm, sm = M.mrope_select(cos.cuda(), sin.cuda(), [16, 24, 24])
    # strided [B,T,H,D] views of a [B,T,H*D] projection output, rotated in place
    qp = q.transpose(1, 2).reshape(1, T, 16 * 128).contiguous().cuda()
    kp = k.transpose(1, 2).reshape(1, T, 2 * 128).contiguous().cuda()
    M.mrope_apply_(qp.view(1, T, 16, 128), cm, sm)
    M.mrope_apply_(kp.view(1, T, 2, 128), cm, sm)
    assert torch.equal(qp.view(1, T, 16, 128).transpose(1, 2).cpu(), rq)
    assert torch.equal(kp.view(1, T, 2, 128).transpose(1, 2).cpu(), rk)


# ---------------------------------------------------------------- mixers -----------------------
def _init(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("A_log"):
                p.copy_(torch.log(torch.empty(p.shape).uniform_(0.01, 16, generator=g)))
            elif name.endswith("dt_bias"):
                dt = torch.exp(torch.empty(p.shape).uniform_(-6.9, -2.3, generator=g))
                p.copy_(dt + torch.log(-torch.expm1(-dt)))
            elif "norm" in name:
                p.fill_(1.0)
            elif "conv1d" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return module


def _params(module):
    return {k: v.detach().float().cpu() for k, v in module.state_dict().items()}


def test_gated_deltanet_mixer_and_streaming(M):
    cfg = M.HybridTextConfig(num_hidden_layers=4)
    mod = _init(M.GatedDeltaNet(cfg, 1), 11).bfloat16().cuda()
    p = _params(mod)
    x = torch.randn(1, 700, 2048, generator=gen(12)).bfloat16()
    ref, rconv, rstate = gdn_mixer_ref(x, p, proj_dtype=torch.bfloat16)
    out, _ = mod(x.cuda())
    assert err_ratio(ref, out.float().cpu()) < 1.5e-2
    # streaming with the reference's cache protocol: 300-token prefill, a 200-token frame, then decode steps
    cache = M.StaticCachePrealloc(config=cfg, batch_size=1, device="cuda", dtype=torch.bfloat16)
    outs = []
    for a, b in ((0, 300), (300, 500), (500, 501), (501, 502), (502, 700)):
        y, _ = mod(x[:, a:b].cuda(), past_key_values=cache, cache_position=torch.arange(a, b, device="cuda"))
        outs.append(y)
    assert err_ratio(ref, torch.cat(outs, 1).float().cpu()) < 1.5e-2
    lin = cache.layers[1]
    assert lin.seq_len == 700 and lin.recurrent_state.dtype == torch.bfloat16
    assert err_ratio(rstate, lin.recurrent_state.float().cpu()) < 1.5e-2
    assert err_ratio(rconv[2], lin.conv_state_v.float().cpu()) < 1e-2


def test_packed_sequences_through_both_mixers(M):
    """cu_seqlens through the mixers (short conv whose window stops at sequence starts, the packed
    chunk operator, the packed attention kernel): a packed row equals the sequences run one by one."""
    cfg = M.HybridTextConfig(num_hidden_layers=4)
    bounds = [0, 70, 71, 400, 1000]
    cu = torch.tensor(bounds, dtype=torch.int32, device="cuda")
    x = torch.randn(1, 1000, 2048, generator=gen(52)).bfloat16().cuda()
    # short conv: bit-identical to per-sequence calls
    w = (torch.randn(2048, 1, 4, generator=gen(53)) * 0.5).bfloat16().cuda()
    y = M.short_conv_silu_varlen(x, w, cu)
    for s0, s1 in zip(bounds[:-1], bounds[1:]):
        assert torch.equal(y[:, s0:s1], M.short_conv_silu(x[:, s0:s1], w)[0])
    gdn = _init(M.GatedDeltaNet(cfg, 1), 51).bfloat16().cuda()
    out, _ = gdn(x, cu_seqlens=cu)
    for s0, s1 in zip(bounds[:-1], bounds[1:]):
        one, _ = gdn(x[:, s0:s1])
        # (sequences of <= 64 tokens take the token recurrence when run alone, the chunk kernel when packed)
        assert err_ratio(one.float(), out[:, s0:s1].float()) < (1.5e-2 if s1 - s0 <= 64 else 1e-6), (s0, s1)
    attn = _init(M.InfiniteVLSelfAttention(cfg, 0), 54).bfloat16().cuda()
    pos = torch.cat([torch.arange(s1 - s0) for s0, s1 in zip(bounds[:-1], bounds[1:])]).cuda()
    cos, sin = attn.rotary_emb(x, pos[None, None].expand(3, 1, -1))     # [3, 1, T, 128]
    out, _ = attn(x, position_embeddings=(cos, sin), cu_seqlens=cu)
    for s0, s1 in zip(bounds[:-1], bounds[1:]):
        if s1 - s0 < 2:
            continue
        one, _ = attn(x[:, s0:s1], position_embeddings=(cos[:, :, s0:s1], sin[:, :, s0:s1]))
        assert err_ratio(one.float(), out[:, s0:s1].float()) < 1e-6, (s0, s1)


@pytest.mark.parametrize("B", [1, 2])
def test_fused_prefill_is_bit_identical_to_the_kernel_chain(M, monkeypatch, B):
    """Prefill-side fusion: for q_len > 64 the mixer hands the RAW q / k projections and a / b to
    ivl_gdn_chunk_fwd_fused, whose pre-pass does the depthwise conv + SiLU, the gate math and the L2 norm.  Same
    expressions, same rounding points as ivl_short_conv_fwd + ivl_gdn_gate_fwd + ivl_gdn_chunk_fwd: outputs, the
    recurrent state and the carried conv tails must be equal bit for bit -- without a cache, with a fresh cache, and
    continuing from a cache (tails and state carried in), with a ragged last chunk and a 3-token left context."""
    cfg = M.HybridTextConfig(num_hidden_layers=4)
    mod = _init(M.GatedDeltaNet(cfg, 1), 41).bfloat16().cuda()
    x = torch.randn(B, 1000, 2048, generator=gen(42)).bfloat16().cuda()
    res = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("IVL_GDN_FUSED_PREFILL", fused)   # "1" forces the fused path at any length
        outs = [mod(x[:, :333])[0]]
        cache = M.StaticCachePrealloc(config=cfg, batch_size=B, device="cuda", dtype=torch.bfloat16)
        for a, b in ((0, 67), (67, 600), (600, 1000)):
            outs.append(mod(x[:, a:b], past_key_values=cache, cache_position=torch.arange(a, b, device="cuda"))[0])
        lin = cache.layers[1]
        assert lin.seq_len == 1000
        res[fused] = outs + [lin.recurrent_state.clone(), lin.conv_state_q.clone(), lin.conv_state_k.clone(),
                             lin.conv_state_v.clone()]
    for a, b in zip(res["1"], res["0"]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("B,cache_dtype", [(1, torch.bfloat16), (3, torch.bfloat16)])
def test_fused_decode_step_is_bit_identical_to_the_kernel_chain(M, monkeypatch, B, cache_dtype):
    """q_len == 1 with a started cache runs the whole mixer core in one launch (ivl_gdn_decode_step: conv steps,
    gates, recurrence, gated norm, caches updated in place).  It repeats the arithmetic of the kernel-by-kernel
    path operation for operation, so outputs and cache contents must be equal bit for bit over many steps."""
    cfg = M.HybridTextConfig(num_hidden_layers=4)
    mod = _init(M.GatedDeltaNet(cfg, 1), 31).bfloat16().cuda()
    x = torch.randn(B, 140, 2048, generator=gen(32)).bfloat16().cuda()
    res = {}
    monkeypatch.setenv("IVL_DECODE_PACKED_PROJ", "0")     # same projection GEMMs in both paths: compare the mixer cores
    for fused in ("1", "0"):
        monkeypatch.setenv("IVL_GDN_FUSED_DECODE", fused)
        cache = M.StaticCachePrealloc(config=cfg, batch_size=B, device="cuda", dtype=cache_dtype)
        outs = [mod(x[:, :100], past_key_values=cache, cache_position=torch.arange(0, 100, device="cuda"))[0]]
        for t in range(100, 140):
            outs.append(mod(x[:, t:t + 1], past_key_values=cache, cache_position=torch.arange(t, t + 1, device="cuda"))[0])
        lin = cache.layers[1]
        assert lin.seq_len == 140
        res[fused] = (torch.cat(outs, 1), lin.recurrent_state.clone(), lin.conv_state_q.clone(),
                      lin.conv_state_k.clone(), lin.conv_state_v.clone())
    for a, b in zip(res["1"], res["0"]):
        assert torch.equal(a, b)
    # and the streamed result is the same function as the one-shot forward (different chunking: tolerance)
    full, _ = mod(x)
    assert err_ratio(full.float().cpu(), res["1"][0].float().cpu()) < 1.5e-2
    if B == 1:
        # decode steps with the six input projections stacked into one GEMV (the default for a single sequence)
        monkeypatch.setenv("IVL_DECODE_PACKED_PROJ", "1")
        monkeypatch.setenv("IVL_GDN_FUSED_DECODE", "1")
        cache = M.StaticCachePrealloc(config=cfg, batch_size=B, device="cuda", dtype=cache_dtype)
        outs = [mod(x[:, :100], past_key_values=cache, cache_position=torch.arange(0, 100, device="cuda"))[0]]
        for t in range(100, 140):
            outs.append(mod(x[:, t:t + 1], past_key_values=cache, cache_position=torch.arange(t, t + 1, device="cuda"))[0])
        assert err_ratio(res["1"][0].float(), torch.cat(outs, 1).float()) < 5e-3
        assert "_packed_in_proj" not in mod.state_dict() and hasattr(mod, "_packed_in_proj")


def test_self_attention_mixer_with_cache(M):
    cfg = M.HybridTextConfig(num_hidden_layers=4, sliding_window=256)
    mod = _init(M.InfiniteVLSelfAttention(cfg, 0), 21).bfloat16().cuda()
    p = _params(mod)
    T = 600
    x = torch.randn(1, T, 2048, generator=gen(22)).bfloat16()
    pos = torch.arange(T)[None, None].expand(3, 1, -1)
    cos, sin = mrope_cos_sin_ref(pos, 128, 1e6)
    ref = swa_mixer_ref(x, p, cos, sin, cache=None, window=256)
    cb, sb = cos.bfloat16().cuda(), sin.bfloat16().cuda()
    out, _ = mod(x.cuda(), position_embeddings=(cb, sb))
    assert err_ratio(ref, out.float().cpu()) < 1.5e-2
    # chunked prefill + decode through the sliding-window cache
    cache = M.StaticCachePrealloc(config=cfg, batch_size=1, device="cuda", dtype=torch.bfloat16)
    rc = SlidingWindowCacheRef(256)
    outs, refs = [], []
    for a, b in ((0, 200), (200, 520), (520, 521), (521, 600)):
        y, _ = mod(x[:, a:b].cuda(), past_key_values=cache, cache_position=torch.arange(a, b, device="cuda"),
                   position_embeddings=(cb[:, :, a:b], sb[:, :, a:b]))
        outs.append(y)
        refs.append(swa_mixer_ref(x[:, a:b], p, cos[:, :, a:b], sin[:, :, a:b], cache=rc, window=256))
        assert cache.layers[0].size == rc.size and cache.layers[0].cumulative_length == rc.cumulative_length
    assert err_ratio(torch.cat(refs, 1), torch.cat(outs, 1).float().cpu()) < 1.5e-2
    assert err_ratio(ref, torch.cat(outs, 1).float().cpu()) < 1.5e-2  # chunked == one-shot under the window rule


def test_hybrid_block_config1(M):
    """BASELINE.json config 1: one hybrid block (1 SWA + 3 GDN incl. norms and MLPs), T = 1024, 3B dims,
    synthetic image+text M-RoPE positions; bf16 CUDA path vs fp32 CPU oracle with identical weights."""
    cfg = M.HybridTextConfig(num_hidden_layers=4)
    dec = _init(M.HybridDecoder(cfg), 31).bfloat16().cuda()
    p = _params(dec)
    T = 1024
    # 16 text tokens, then 3 frames of 16x16 vision tokens each followed by 80 text tokens
    t_ids, h_ids, w_ids, nxt = [], [], [], 0
    def text(n):
        nonlocal nxt
        for _ in range(n):
            t_ids.append(nxt); h_ids.append(nxt); w_ids.append(nxt); nxt += 1
    text(16)
    for f in range(3):
        base = nxt
        for hh in range(16):
            for ww in range(16):
                t_ids.append(base + 2 * f * 0); h_ids.append(base + hh); w_ids.append(base + ww)
        nxt = base + 16
        text(80)
    pos = torch.tensor([t_ids, h_ids, w_ids])[:, None, :]
    assert pos.shape == (3, 1, T)
    x = torch.randn(1, T, 2048, generator=gen(32)).bfloat16()
    ref = hybrid_decoder_ref(x, p, cfg.layer_types, pos, proj_dtype=torch.bfloat16)
    out = dec(x.cuda(), position_ids=pos.cuda())
    assert torch.isfinite(out).all()
    assert err_ratio(ref, out.float().cpu()) < 2e-2
