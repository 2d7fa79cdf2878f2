"""GPU tests of the model surface (SURVEY.md section 8 rows a-T, b-6): the text model against the fp32 oracle, cache
auto-allocation, logits_to_keep, and the greedy generate loop -- prefill then single-token steps through the static
cache, eagerly and with the decode step captured in a CUDA graph (demo_streaming_inference.py:389-421, 473-489)."""
import pytest
import torch

from oracle import err_ratio, hybrid_decoder_ref
from test_modules_gpu import _init, gen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from infinitevl_b200 import modeling
    return modeling


def _model(M, layers=4, vocab=512, window=1024):
    cfg = M.InfiniteVLConfig(text_config=M.HybridTextConfig(num_hidden_layers=layers, vocab_size=vocab,
                                                            sliding_window=window))
    m = _init(M.InfiniteVLQwen2_5_VLForConditionalGeneration(cfg), 51)
    with torch.no_grad():
        m.model.language_model.embed_tokens.weight.normal_(0, 1.0, generator=gen(52))
    return cfg, m.bfloat16().cuda().eval()


def test_causal_lm_forward_matches_oracle(M):
    cfg, m = _model(M)
    T = 700
    ids = torch.randint(0, 512, (1, T), generator=gen(53))
    out = m(input_ids=ids.cuda(), use_cache=True, logits_to_keep=0)
    assert isinstance(out.past_key_values, M.StaticCachePrealloc)        # allocated on the first forward (std:1488-1501)
    assert out.past_key_values.get_seq_length() == T
    assert out.logits.shape == (1, T, 512)
    lm = m.model.language_model
    p = {k: v.detach().float().cpu() for k, v in lm.state_dict().items()}
    x = p["embed_tokens.weight"][ids[0]][None]
    pos = torch.arange(T)[None, None].expand(3, 1, -1)
    h = hybrid_decoder_ref(x, p, cfg.text_config.layer_types, pos, window=1024, proj_dtype=torch.bfloat16)
    ref = h @ p["embed_tokens.weight"].t()                               # tied head
    assert err_ratio(ref, out.logits.float().cpu()) < 2.5e-2
    # logits_to_keep = 1 keeps the last position only; no cache unless asked for
    out1 = m(input_ids=ids.cuda(), logits_to_keep=1)
    assert out1.logits.shape == (1, 1, 512) and out1.past_key_values is None
    assert torch.equal(out1.logits[:, -1], out.logits[:, -1])
    # inputs_embeds instead of input_ids gives the same result
    emb = m.get_input_embeddings()(ids.cuda())
    out2 = m(inputs_embeds=emb, logits_to_keep=1)
    assert torch.equal(out2.logits, out1.logits)


def test_greedy_generate_eager_and_graph(M):
    cfg, m = _model(M, window=256)       # small window: the decode steps run with a full, wrapping ring
    ids = torch.randint(0, 512, (1, 4096), generator=gen(54)).cuda()
    a = m.generate(input_ids=ids, max_new_tokens=32)
    b = m.generate(input_ids=ids, max_new_tokens=32, use_cuda_graph=True)
    assert a.shape == (1, 32) and torch.equal(a, b)
    # the same tokens from a hand-written loop over forward() (what the demo's QA branch does)
    cache = m.allocate_inference_cache(1)
    out = m(input_ids=ids, past_key_values=cache, use_cache=True, cache_position=torch.arange(4096, device="cuda"),
            logits_to_keep=1)
    tok = out.logits[:, -1].argmax(-1, keepdim=True)
    toks = [tok]
    for i in range(31):
        pos = torch.full((3, 1, 1), 4096 + i, device="cuda")
        out = m(input_ids=tok, past_key_values=cache, use_cache=True, position_ids=pos,
                cache_position=torch.tensor([4096 + i], device="cuda"), logits_to_keep=1)
        tok = out.logits[:, -1].argmax(-1, keepdim=True)
        toks.append(tok)
    assert torch.equal(torch.cat(toks, 1), a)
    assert cache.get_seq_length() == 4096 + 31
