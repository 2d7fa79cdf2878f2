"""CPU tests of the model surface (SURVEY.md section 8 rows a-T, b-6, f-4): module tree = checkpoint state-dict
names, config.json parsing, get_rope_index bit-exact against the reference's own function, checkpoint round trip,
argument rules of the text model.  (The forward passes need the CUDA kernels: tests/test_model_gpu.py.)"""
import json
import os

import numpy as np
import pytest
import torch

from infinitevl_b200 import modeling as M

from golden.make_golden_rope_index import cases


def _small(layers=4, vocab=512):
    return M.InfiniteVLConfig(text_config=M.HybridTextConfig(num_hidden_layers=layers, vocab_size=vocab))


def test_module_tree_matches_checkpoint_names():
    m = M.InfiniteVLQwen2_5_VLForConditionalGeneration(_small())
    keys = set(m.state_dict())
    for k in ("model.language_model.embed_tokens.weight", "model.language_model.norm.weight", "lm_head.weight",
              "model.language_model.layers.0.self_attn.q_proj.bias", "model.language_model.layers.0.self_attn.o_proj.weight",
              "model.language_model.layers.1.self_attn.A_log", "model.language_model.layers.1.self_attn.dt_bias",
              "model.language_model.layers.1.self_attn.q_conv1d.weight", "model.language_model.layers.1.self_attn.o_norm.weight",
              "model.language_model.layers.1.self_attn.g_proj.weight", "model.language_model.layers.1.mlp.down_proj.weight",
              "model.language_model.layers.3.post_attention_layernorm.weight"):
        assert k in keys, k
    # tied head: one tensor, and the exported dict omits it as the published checkpoint does
    assert m.lm_head.weight is m.model.language_model.embed_tokens.weight
    assert "lm_head.weight" not in m.hf_state_dict()
    assert m.model.language_model.layers[0].attention_type == "sliding_attention"
    assert m.model.language_model.layers[1].attention_type == "linear_attention"


def test_config_from_reference_style_json(tmp_path):
    d = {"hidden_size": 2048, "intermediate_size": 11008, "num_hidden_layers": 36, "num_attention_heads": 16,
         "num_key_value_heads": 2, "rms_norm_eps": 1e-6, "rope_theta": 1e6, "sliding_window": 8192,
         "use_sliding_window": True, "tie_word_embeddings": True, "vocab_size": 151936, "image_token_id": 151655,
         "video_token_id": 151656, "vision_start_token_id": 151652,
         "rope_scaling": {"mrope_section": [16, 24, 24], "rope_type": "default", "type": "default"},
         "vision_config": {"spatial_merge_size": 2, "tokens_per_second": 2, "depth": 32}}
    p = tmp_path / "config.json"
    p.write_text(json.dumps(d))
    cfg = M.InfiniteVLConfig.from_json(str(p))
    tc = cfg.text_config
    assert (tc.num_hidden_layers, tc.sliding_window, tc.vocab_size, tc.num_linear_heads) == (36, 8192, 151936, 16)
    assert tc.layer_types[:5] == ["sliding_attention", "linear_attention", "linear_attention", "linear_attention",
                                  "sliding_attention"]
    assert tc.rope_scaling["mrope_section"] == [16, 24, 24] and tc.rope_scaling["rope_theta"] == 1e6
    assert cfg.vision_config.spatial_merge_size == 2 and cfg.image_token_id == 151655


def test_get_rope_index_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_rope_index.npz"))
    cfg = _small()
    for i, c in enumerate(cases()):
        t = lambda x, dt=torch.long: None if x is None else torch.tensor(x, dtype=dt)
        pos, delta = M.get_rope_index(cfg, t(c["ids"]), t(c["img"]), t(c["vid"]), t(c["spg"], torch.float32), t(c["mask"]))
        assert np.array_equal(pos.numpy(), z[f"pos{i}"]), i
        assert np.array_equal(delta.numpy(), z[f"delta{i}"]), i
    # text only: arange on all three rows, zero deltas; with a padding mask: cumsum rule
    ids = torch.arange(12).view(2, 6)
    pos, delta = M.get_rope_index(cfg, ids)
    assert torch.equal(pos, torch.arange(6).view(1, 1, 6).expand(3, 2, 6)) and int(delta.abs().sum()) == 0
    mask = torch.tensor([[0, 0, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]])
    pos, delta = M.get_rope_index(cfg, ids, attention_mask=mask)
    assert pos[0, 0].tolist() == [1, 1, 0, 1, 2, 3] and delta.view(-1).tolist() == [-2, 0]


def test_checkpoint_round_trip(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    cfg = _small(layers=4, vocab=256)
    torch.manual_seed(0)
    m = M.InfiniteVLQwen2_5_VLForConditionalGeneration(cfg).bfloat16()
    sd = {k: v.contiguous() for k, v in m.hf_state_dict().items()}
    sd["model.visual.blocks.0.attn.qkv.weight"] = torch.zeros(4, 4)      # the vision tower's weights are ignored
    st.save_file(sd, str(tmp_path / "model-00001-of-00001.safetensors"))
    d = {"num_hidden_layers": 4, "vocab_size": 256, "tie_word_embeddings": True,
         "rope_scaling": {"mrope_section": [16, 24, 24], "rope_type": "default"}}
    (tmp_path / "config.json").write_text(json.dumps(d))
    m2 = M.InfiniteVLQwen2_5_VLForConditionalGeneration.from_pretrained(str(tmp_path), device="cpu")
    a, b = m.state_dict(), m2.state_dict()
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # a key that matches nothing is an error, not silently dropped
    sd["model.language_model.layers.0.self_attn.bogus"] = torch.zeros(1)
    st.save_file(sd, str(tmp_path / "model-00001-of-00001.safetensors"))
    with pytest.raises(RuntimeError):
        M.InfiniteVLQwen2_5_VLForConditionalGeneration.from_pretrained(str(tmp_path), device="cpu")


def test_text_model_argument_rules():
    tm = M.InfiniteVLTextModel(M.HybridTextConfig(num_hidden_layers=2, vocab_size=64))
    with pytest.raises(ValueError):
        tm(input_ids=None, inputs_embeds=None)
    with pytest.raises(ValueError):
        tm(input_ids=torch.zeros(1, 4, dtype=torch.long), inputs_embeds=torch.zeros(1, 4, 2048))
    with pytest.raises(NotImplementedError):
        tm(input_ids=torch.zeros(1, 4, dtype=torch.long), output_attentions=True)
    model = M.InfiniteVLModel(_small(layers=2, vocab=64))
    with pytest.raises(NotImplementedError):   # no vision tower attached
        model(input_ids=torch.zeros(1, 4, dtype=torch.long), pixel_values=torch.zeros(4, 8),
              image_grid_thw=torch.tensor([[1, 2, 2]]))


def test_left_context_table_and_stacked_projections():
    """Host-side helpers of the packed-row conv and of the stacked decode projections (no GPU needed)."""
    import torch
    from infinitevl_b200 import modeling as M
    # left context per token of a packed row: min(3, tokens of the own sequence before it); unowned tokens get 0
    tab = M.left_context_table([0, 1, 3, 3, 9], 11, "cpu")
    assert tab.dtype == torch.uint8
    naive = []
    for t in range(11):
        start = max(s for s in [0, 1, 3, 3, 9] if s <= t)
        naive.append(min(3, t - start))
    assert tab.tolist() == naive
    # several nn.Linear over one input as one matmul; the stack follows the parameters
    torch.manual_seed(0)
    owner = torch.nn.Module()
    mods = (torch.nn.Linear(16, 8, bias=True), torch.nn.Linear(16, 4, bias=False), torch.nn.Linear(16, 12, bias=True))
    x = torch.randn(1, 1, 16)
    outs = M._packed_linear(owner, "_stack", mods, x)
    for m, y in zip(mods, outs):
        assert torch.allclose(m(x), y, atol=1e-6)
    first = owner._stack[1]
    assert M._packed_linear(owner, "_stack", mods, x)[0] is not None and owner._stack[1] is first   # cached
    with torch.no_grad():
        mods[1].weight.add_(1.0)                       # in-place update bumps the version counter: the stack is rebuilt
    outs = M._packed_linear(owner, "_stack", mods, x)
    assert owner._stack[1] is not first and torch.allclose(mods[1](x), outs[1], atol=1e-6)
    assert "_stack" not in owner.state_dict()


def test_vectorised_rope_index_is_the_reference_function(golden_dir, monkeypatch):
    """the device-side, loop-free get_rope_index (one tensor operation per step over all vision blocks
    and tokens) reproduces the reference's own outputs (golden vectors) and, on random well-formed prompts -- images,
    videos with fractional seconds per grid, batches with left padding -- the block-by-block path bit for bit; rows
    that are not well formed fall back to that path."""
    import random
    from infinitevl_b200 import modeling as MM
    z = np.load(os.path.join(golden_dir, "ref_rope_index.npz"))
    cfg = _small()
    t = lambda x, dt=torch.long: None if x is None else torch.tensor(x, dtype=dt)
    for i, c in enumerate(cases()):
        fast = MM._rope_index_vectorized(cfg, t(c["ids"]), t(c["img"]), t(c["vid"]), t(c["spg"], torch.float32), t(c["mask"]))
        assert fast is not None, i
        assert np.array_equal(fast[0].numpy(), z[f"pos{i}"]) and np.array_equal(fast[1].numpy(), z[f"delta{i}"]), i
    IMG, VID, VS, VE = 151655, 151656, 151652, 151653
    rng = random.Random(7)
    for trial in range(40):
        rows, img, vid, spg = [], [], [], []
        for _ in range(rng.choice([1, 1, 2, 3])):
            ids = [rng.randint(10, 99) for _ in range(rng.randint(0, 6))]
            for _ in range(rng.randint(1, 5)):
                tt, hh, ww = rng.choice([1, 1, 2, 3]), 2 * rng.randint(1, 4), 2 * rng.randint(1, 4)
                if rng.random() < 0.5:
                    img.append([1, hh, ww]); kind, tt = IMG, 1
                else:
                    vid.append([tt, hh, ww]); spg.append(rng.choice([0.5, 1.0, 1.5, 2.0, 3.7])); kind = VID
                ids += [VS] + [kind] * (tt * (hh // 2) * (ww // 2)) + [VE] + [rng.randint(10, 99) for _ in range(rng.randint(0, 5))]
            rows.append(ids)
        L = max(len(r) for r in rows)
        mask = [[0] * (L - len(r)) + [1] * len(r) for r in rows]
        ids = [[0] * (L - len(r)) + r for r in rows]
        args = (cfg, t(ids), t(img) if img else None, t(vid) if vid else None, t(spg, torch.float32) if spg else None, t(mask))
        fast = MM._rope_index_vectorized(*args)
        assert fast is not None, trial
        monkeypatch.setattr(MM, "_rope_index_vectorized", lambda *a, **k: None)
        slow = MM.get_rope_index(*args)
        monkeypatch.undo()
        assert torch.equal(fast[0], slow[0]) and torch.equal(fast[1], slow[1]), trial
    # a row whose placeholder run is one token short is not well formed: block-by-block path
    bad = [5, VS] + [IMG] * 5 + [VE, 6]
    assert MM._rope_index_vectorized(cfg, t([bad]), t([[1, 4, 6]]), None, None, None) is None


def test_attention_interface_registration():
    """The drop-in point of the SWA operator is the HF attention interface (std:1092-1108): registering the B200
    entry makes `ALL_ATTENTION_FUNCTIONS[name]` resolve to it, and it refuses what it cannot do instead of ignoring it."""
    from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    from infinitevl_b200 import swa
    name = swa.register_attention_interface("ivl_b200_swa_test")
    assert ALL_ATTENTION_FUNCTIONS[name] is swa.sliding_window_attention_forward
    q = torch.zeros(2, 4, 3, 128)
    with pytest.raises(NotImplementedError):     # a real padding mask
        swa.sliding_window_attention_forward(None, q, q, q, attention_mask=torch.tensor([[0, 1, 1], [1, 1, 1]]))
    with pytest.raises(NotImplementedError):     # dropout
        swa.sliding_window_attention_forward(None, q, q, q, dropout=0.1)
    with pytest.raises(Exception):               # CPU tensors: there is no CPU fallback
        swa.sliding_window_attention_forward(None, q.bfloat16(), q.bfloat16(), q.bfloat16())
