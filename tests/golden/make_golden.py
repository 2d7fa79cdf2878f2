"""Generate the golden fixtures that pin oracle/ to the reference.

Run in the BUILD container only (needs /root/reference, which does not exist on
the GPU box):  python tests/golden/make_golden.py

The reference ships no tests or golden vectors for this path (SURVEY.md
section 4), so the fixtures are outputs of the reference's own Python code on
seeded inputs:

  ref_delta_rule.npz    fla/ops/delta_rule/naive.py  delta_rule_recurrence / delta_rule_chunkwise
                        (in-tree, src/llamafactory/model/fla)
  ref_gated_naive.npz   flash-linear-attention's naive_recurrent_gated_delta_rule /
                        naive_chunk_gated_delta_rule (the un-vendored dependency whose kernels
                        execute the gated op; requirements.txt:19-20, 0.5.1 in this image)
  ref_mrope.npz         InfiniteVLRotaryEmbedding.forward + apply_multimodal_rotary_pos_emb
                        (infinitevl_standard/modeling_infinitevl.py:916-930,949-984)
  ref_eager_attn.npz    eager_attention_forward with an additive sliding-window causal mask (:557-580)
  ref_swa_cache.npz     StaticSlidingWindowLayerPrealloc update()/get_mask_sizes() trace (:66-227)
  ref_linear_cache.npz  StaticLinearLayerPrealloc get/set trace (:229-364)

tests/golden/fla_triton_gdn_T256_H2_seed7.npz is different: it holds outputs of the
dependency's Triton kernels run on a B200 by tools/ref_gpu_probe.py.
"""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def gdn_inputs(B, T, H, K, V, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, T, H, K, generator=g)
    k = torch.randn(B, T, H, K, generator=g)
    v = torch.randn(B, T, H, V, generator=g)
    beta = torch.sigmoid(torch.randn(B, T, H, generator=g))
    a = torch.log(torch.empty(H).uniform_(1e-3, 16, generator=g))
    dt = torch.exp(torch.empty(H).uniform_(math.log(1e-3), math.log(1e-1), generator=g))
    dt_bias = dt + torch.log(-torch.expm1(-dt))
    gate = -torch.exp(a) * torch.nn.functional.softplus(torch.randn(B, T, H, generator=g) + dt_bias)
    h0 = torch.randn(B, H, K, V, generator=g)
    return q, k, v, gate, beta, h0


def main():
    torch.manual_seed(0)
    # ---- in-tree ungated delta rule -------------------------------------------------
    import importlib.util  # load the one file; importing the vendored package would shadow pip fla
    spec = importlib.util.spec_from_file_location(
        "ref_naive", os.path.join(REF, "src/llamafactory/model/fla/ops/delta_rule/naive.py"))
    naive = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(naive)
    q, k, v, _, beta, h0 = gdn_inputs(1, 128, 2, 32, 48, seed=11)
    k = torch.nn.functional.normalize(k, dim=-1)  # delta rule needs |k| <= 1 to stay bounded
    hf = lambda x: x.transpose(1, 2).contiguous()  # the in-tree functions are head-first
    o_rec, s_rec = naive.delta_rule_recurrence(hf(q), hf(k), hf(v), hf(beta), initial_state=h0)
    o_chk = naive.delta_rule_chunkwise(hf(q), hf(k), hf(v), hf(beta), chunk_size=32)
    if isinstance(o_chk, tuple):
        o_chk = o_chk[0]
    np.savez_compressed(os.path.join(HERE, "ref_delta_rule.npz"), q=q.numpy(), k=k.numpy(), v=v.numpy(),
                        beta=beta.numpy(), h0=h0.numpy(), o_recurrence_h0=o_rec.transpose(1, 2).numpy(),
                        s_recurrence_h0=s_rec.numpy(), o_chunkwise=o_chk.transpose(1, 2).numpy())

    # ---- the dependency's gated naive forms ------------------------------------------
    from fla.ops.gated_delta_rule.naive import naive_chunk_gated_delta_rule, naive_recurrent_gated_delta_rule
    q, k, v, g, beta, h0 = gdn_inputs(1, 200, 2, 64, 96, seed=12)
    qn, kn = (torch.nn.functional.normalize(x, dim=-1) for x in (q, k))
    o1, s1 = naive_recurrent_gated_delta_rule(qn, kn, v, beta, g, initial_state=h0, output_final_state=True)
    o2, s2 = naive_chunk_gated_delta_rule(qn, kn, v, g, beta, initial_state=h0, output_final_state=True)
    np.savez_compressed(os.path.join(HERE, "ref_gated_naive.npz"), q=q.numpy(), k=k.numpy(), v=v.numpy(),
                        g=g.numpy(), beta=beta.numpy(), h0=h0.numpy(), o_recurrent=o1.numpy(),
                        s_recurrent=s1.numpy(), o_chunk=o2.numpy(), s_chunk=s2.numpy())

    # ---- model file -------------------------------------------------------------------
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS

    def _default_rope(config, device=None, **kw):  # removed from transformers 5.x; the 4.57 definition
        base = config.rope_scaling.get("rope_theta", getattr(config, "rope_theta", 1e6)) \
            if getattr(config, "rope_scaling", None) else getattr(config, "rope_theta", 1e6)
        dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
        inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).to(device=device, dtype=torch.float) / dim))
        return inv, 1.0

    ROPE_INIT_FUNCTIONS.setdefault("default", _default_rope)
    sys.path.insert(0, os.path.join(REF, "infinitevl"))
    import infinitevl_standard.modeling_infinitevl as M
    from infinitevl_standard.configuration_infinitevl import InfiniteVLTextConfig

    cfg = InfiniteVLTextConfig()
    rope = M.InfiniteVLRotaryEmbedding(cfg)
    T = 96
    g = torch.Generator().manual_seed(13)
    pos = torch.stack([torch.arange(T), torch.randint(0, 40, (T,), generator=g),
                       torch.randint(0, 40, (T,), generator=g)])[:, None, :]  # [3, 1, T]
    x = torch.zeros(1, T, 8, dtype=torch.float32)
    cos, sin = rope(x, pos)
    qh = torch.randn(1, 4, T, 128, generator=g)
    kh = torch.randn(1, 2, T, 128, generator=g)
    sect = cfg.rope_scaling.get("mrope_section", [16, 24, 24]) if cfg.rope_scaling else [16, 24, 24]
    qe, ke = M.apply_multimodal_rotary_pos_emb(qh, kh, cos, sin, sect)
    cos_b, sin_b = rope(x.to(torch.bfloat16), pos)
    qeb, keb = M.apply_multimodal_rotary_pos_emb(qh.to(torch.bfloat16), kh.to(torch.bfloat16), cos_b, sin_b, sect)
    np.savez_compressed(os.path.join(HERE, "ref_mrope.npz"), pos=pos.numpy(), cos=cos.numpy(), sin=sin.numpy(),
                        q=qh.numpy(), k=kh.numpy(), q_rot=qe.numpy(), k_rot=ke.numpy(),
                        mrope_section=np.array(sect), q_rot_bf16=qeb.float().numpy(), k_rot_bf16=keb.float().numpy(),
                        theta=np.array(float(getattr(cfg, "rope_theta", 1e6))))

    # eager attention with an additive windowed causal mask
    Tq, Tk, W = 40, 72, 24
    g = torch.Generator().manual_seed(14)
    qa = torch.randn(1, 4, Tq, 32, generator=g)
    ka = torch.randn(1, 2, Tk, 32, generator=g)
    va = torch.randn(1, 2, Tk, 32, generator=g)
    i = torch.arange(Tq)[:, None] + (Tk - Tq)
    j = torch.arange(Tk)[None, :]
    vis = (j <= i) & ((i - j) <= W - 1)
    mask = torch.zeros(1, 1, Tq, Tk).masked_fill(~vis, float("-inf"))
    mod = types.SimpleNamespace(num_key_value_groups=2, training=False)
    oa, _ = M.eager_attention_forward(mod, qa, ka, va, mask, scaling=32 ** -0.5)
    np.savez_compressed(os.path.join(HERE, "ref_eager_attn.npz"), q=qa.numpy(), k=ka.numpy(), v=va.numpy(),
                        out=oa.numpy(), window=np.array(W))

    # caches: small window so the roll-over is exercised
    cfg_small = InfiniteVLTextConfig(use_sliding_window=True, sliding_window=8, num_key_value_heads=1, num_attention_heads=2, hidden_size=8,
                                     head_dim=4)
    cfg_small.head_dim = 4
    layer = M.StaticSlidingWindowLayerPrealloc(config=cfg_small, batch_size=1, dtype=torch.float32, zero_init=True)
    steps = [3, 1, 1, 5, 1, 9, 2, 1, 1, 20, 1]
    trace_sizes, trace_full, trace_tail = [], [], []
    base = 0
    for n in steps:
        kk = (torch.arange(base, base + n, dtype=torch.float32)[None, None, :, None]).expand(1, 1, n, 4).contiguous()
        fk, fv = layer.update(kk, kk * 2)
        base += n
        kv_len, kv_off = layer.get_mask_sizes(torch.arange(n))
        trace_sizes.append([n, kv_len, kv_off, layer.size, layer.cumulative_length, fk.shape[-2]])
        trace_full.append(fk[0, 0, :, 0].numpy().copy())
        trace_tail.append(layer.keys[0, 0, :, 0].numpy().copy())
    np.savez_compressed(os.path.join(HERE, "ref_swa_cache.npz"), steps=np.array(steps), sizes=np.array(trace_sizes),
                        full=np.array(trace_full, dtype=object), tail=np.array(trace_tail, dtype=object),
                        window=np.array(8), allow_pickle=True)

    cfg_lin = InfiniteVLTextConfig(num_linear_heads=2, num_linear_key_value_heads=2, linear_head_dim=4, expand_v=2, conv_size=4)
    lin = M.StaticLinearLayerPrealloc(config=cfg_lin, batch_size=1, dtype=torch.bfloat16, zero_init=True)
    first = lin.update(cache_kwargs={"op": "get"})
    second = lin.update(cache_kwargs={"op": "get"})
    g = torch.Generator().manual_seed(15)
    st = torch.randn(1, 2, 4, 8, generator=g)
    cq, ck, cv = torch.randn(1, 8, 4, generator=g), torch.randn(1, 8, 4, generator=g), torch.randn(1, 16, 4, generator=g)
    lin.update(conv_state=(cq, ck, cv), recurrent_state=st, cache_kwargs={"op": "set", "delta_len": 7})
    (gq, gk, gv), gs = lin.update(cache_kwargs={"op": "get"})
    np.savez_compressed(os.path.join(HERE, "ref_linear_cache.npz"),
                        first_is_none=np.array([first[1] is None, all(x is None for x in first[0])]),
                        second_is_tensor=np.array([second[1] is not None]),
                        state_in=st.numpy(), cq=cq.numpy(), ck=ck.numpy(), cv=cv.numpy(),
                        state_out=gs.float().numpy(), cq_out=gq.float().numpy(), cv_out=gv.float().numpy(),
                        seq_len=np.array(lin.get_seq_length()),
                        shapes=np.array([list(gq.shape) + [0], list(gv.shape) + [0], list(gs.shape)], dtype=object),
                        allow_pickle=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
