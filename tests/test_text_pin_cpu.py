"""Pins the oracle's XLM-R text tower (oracle/functional.py::text_tower) to the third-party arithmetic the reference calls:
transformers' XLMRobertaModel (mm_backbone.py:358-359,382-390: CLS row -> Linear head -> L2 normalise; the reference pins
transformers 4.57.1, this image has 5.x: same module tree built from the same config keys)."""
import pytest
import torch


@pytest.mark.parametrize("text", ["base", "large"])
def test_text_tower_matches_hf_xlm_roberta(text):
    transformers = pytest.importorskip("transformers")
    from oracle import functional as Fn, synth
    from wedetect_b200 import schema
    size = next(k for k, v in schema.SIZES.items() if v["text"] == text)
    t = schema.TEXT[text]
    vocab, S, L = 1000, 7, 9
    sd = synth.synth_state_dict(size, seed=3, with_text=True, text_vocab=vocab, calibrate=False)
    cfg = transformers.XLMRobertaConfig(vocab_size=vocab, hidden_size=t["hidden"], num_hidden_layers=t["layers"], num_attention_heads=t["heads"],
                                        intermediate_size=t["inter"], max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=schema.TEXT_EPS,
                                        pad_token_id=schema.TEXT_PAD, hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = transformers.XLMRobertaModel(cfg, add_pooling_layer=False).eval()
    pre = "backbone.text_model.model."
    hf_sd = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    missing, unexpected = hf.load_state_dict(hf_sd, strict=False)
    assert not unexpected and all("position_ids" in m or "token_type_ids" in m for m in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(3, vocab, (S, L), generator=g)
    ids[:, 0] = 0
    for s in range(S):                      # ragged: <s> tokens </s> <pad>...
        n = 2 + s % (L - 1)
        ids[s, n - 1] = 2
        ids[s, n:] = schema.TEXT_PAD
    mask = (ids != schema.TEXT_PAD).long()
    with torch.no_grad():
        h = hf(input_ids=ids, attention_mask=mask)["last_hidden_state"][:, 0]
        want = torch.nn.functional.normalize(torch.nn.functional.linear(h, sd["backbone.text_model.head.weight"], sd["backbone.text_model.head.bias"]), dim=-1)
        got = Fn.text_tower(sd, size, ids.int(), mask.int())
    assert float((got - want).abs().max()) <= 2e-6, float((got - want).abs().max())
