"""GPU parity of the caption-encoder end (include/ndit_text.h, lumina_t2x_b200/text_encoder.py) against transformers' GemmaModel
(fixture tests/golden/gemma_tiny.pt: hidden_states[-2] in fp32 and bf16) and the oracle's bf16 mode; the padded positions of a row are
whatever the causal stack makes of the pad tokens - like in the reference, the denoiser masks them out - so only the valid positions
are compared.  Tolerances as in test_model_gpu.py: <= 2e-2 against the oracle in bf16 mode, within 1.5x of transformers' own
bf16-vs-fp32 distance (+2e-3) against its fp32 output."""
import os

import pytest
import torch

from oracle import gemma_oracle as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b, valid):
    return ((a.float() - b.float())[valid].abs().max() / b.float()[valid].abs().max()).item()


def _build(cfg, W, **kw):
    from lumina_t2x_b200.text_encoder import GemmaTextEncoder
    m = GemmaTextEncoder(None, vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
                         num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim,
                         intermediate_size=cfg.intermediate_size, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, **kw)
    m.load_state_dict(W, strict=True)
    return m.eval().to("cuda", dtype=torch.bfloat16)


def test_hidden_states_m2_vs_transformers_fixture_and_oracle():
    fx = torch.load(os.path.join(GOLD, "gemma_tiny.pt"), map_location="cpu", weights_only=False)
    cfg = G.GemmaCfg(**fx["cfg"])
    W = G.synthetic_weights(cfg, seed=fx["weight_seed"])
    m = _build(cfg, W, max_tokens=256)
    for c in fx["cases"]:
        ids, mask = c["ids"], c["mask"]
        valid = mask.bool()
        out = m(input_ids=ids.cuda(), attention_mask=mask.cuda(), output_hidden_states=True).hidden_states[-2].float().cpu()
        assert out.shape == c["h_fp32"].shape and torch.isfinite(out).all()
        orc = G.hidden_states_m2(cfg, W, ids, mask, "bf16")
        floor = _rel(c["h_bf16_cpu"], c["h_fp32"], valid)
        assert _rel(out, orc, valid) < 2e-2, (_rel(out, orc, valid), floor)
        assert _rel(out, c["h_fp32"], valid) < 1.5 * floor + 2e-3, (_rel(out, c["h_fp32"], valid), floor)
    # determinism + no mask = all-ones mask
    ids, mask = fx["cases"][0]["ids"].cuda(), fx["cases"][0]["mask"].cuda()
    a = m(input_ids=ids, attention_mask=torch.ones_like(mask), output_hidden_states=True).hidden_states[-2]
    b = m(input_ids=ids, attention_mask=None, output_hidden_states=True).hidden_states[-2]
    assert torch.equal(a, b)


def test_strict_loading_and_errors():
    cfg = G.config_tiny()
    W = G.synthetic_weights(cfg, seed=0)
    from lumina_t2x_b200.text_encoder import GemmaTextEncoder
    bad = dict(W)
    bad.pop("layers.1.mlp.up_proj.weight")
    m = GemmaTextEncoder(None, **{k: getattr(cfg, k) for k in ("vocab_size", "hidden_size", "num_hidden_layers", "num_attention_heads",
                                                              "num_key_value_heads", "head_dim", "intermediate_size")})
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)
    with pytest.raises(RuntimeError):
        _build(cfg, W)(input_ids=torch.zeros(1, 8, dtype=torch.long))            # CPU ids: no CPU path


def test_gemma_2b_shape_runs():
    """google/gemma-2b sizes (18 layers, hidden 2048, 8 heads / 1 kv head of 256, MLP 16384; random weights created on the device),
    2 prompts x 128 tokens: finite, deterministic, padded keys do not reach the valid positions."""
    from lumina_t2x_b200.text_encoder import GemmaTextEncoder
    with torch.device("cuda"):
        m = GemmaTextEncoder(None, vocab_size=32000, max_tokens=512)        # a smaller vocabulary keeps the test light; the stack is full size
    with torch.no_grad():
        for k, p in m.named_parameters():
            if p.dim() == 2:
                p.normal_(std=p.shape[1] ** -0.5)
            else:
                p.normal_(std=0.1)
    m = m.eval().to("cuda", dtype=torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(0)
    ids = torch.randint(1, 32000, (2, 128), device="cuda", generator=g)
    mask = torch.ones(2, 128, dtype=torch.long, device="cuda")
    mask[1, 40:] = 0
    a = m(input_ids=ids, attention_mask=mask).hidden_states[-2]
    b = m(input_ids=ids, attention_mask=mask).hidden_states[-2]
    assert a.shape == (2, 128, 2048) and torch.isfinite(a.float()).all() and torch.equal(a, b)
    ids2 = ids.clone()
    ids2[1, 40:] = 7                                                       # different pad tokens behind the mask
    c = m(input_ids=ids2, attention_mask=mask).hidden_states[-2]
    assert torch.equal(a[1, :40], c[1, :40]) and torch.equal(a[0], c[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(input_ids=ids, attention_mask=mask)
    e1.record()
    torch.cuda.synchronize()
    print(f"gemma-2b shape, 2 x 128 tokens: {e0.elapsed_time(e1) / 5:.2f} ms per encode")
