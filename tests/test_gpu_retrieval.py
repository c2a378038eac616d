"""GPU parity of the object-retrieval path (BASELINE config 5) against oracle/retrieval.py, which is
pinned to the reference's own extract_embedding.head_predict / retrieval_metric.py lines (tests/test_oracle_pin.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
D = "cuda:0"


@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("B,P,K", [(3, 37, 83), (2, 300, 1203)])
def test_retrieval_scorer_matches_oracle(precise, B, P, K):
    """scale_rows -> tcgen05 GEMM -> retr_reduce on random rows: ragged proposal counts (0, 1, P), K not a multiple of 8."""
    from oracle.retrieval import image_scores_ref
    from wedetect_b200.retrieval import RetrievalScorer
    g = torch.Generator().manual_seed(K + P)
    text = torch.nn.functional.normalize(torch.randn(K, 768, generator=g), dim=-1)
    counts = [0, 1, P][:B] if B == 3 else [P, P // 3]
    embs = [torch.randn(n, 768, generator=g) for n in counts]
    scales = [torch.rand(n, generator=g) - 1.5 for n in counts]
    biases = [torch.randn(n, generator=g) * 2 - 3 for n in counts]
    sc = RetrievalScorer(text, B, P, device=D, precise=precise)
    sc.load(embs, scales, biases)
    got = sc.run().cpu()
    tol = 2e-5 if precise else 5e-3
    for b, n in enumerate(counts):
        if n == 0:
            assert float(got[b].abs().max()) == 0.0          # no proposals: every class scores 0 (never above a threshold)
            continue
        want = image_scores_ref(embs[b], text, scales[b], biases[b])
        assert float((got[b] - want).abs().max()) <= tol, (b, n, float((got[b] - want).abs().max()))
    # hqclip switch of the reference (plain sigmoid of the logits): zero scale / bias rows
    zeros = [torch.zeros(n) for n in counts]
    sc.load(embs, zeros, zeros)
    got = sc.run().cpu()
    for b, n in enumerate(counts):
        if n:
            want = image_scores_ref(embs[b], text, None, None, model="hqclip")
            assert float((got[b] - want).abs().max()) <= tol


def test_extract_variant_and_scores_match_oracle():
    """SimpleYOLOWorldDetector(extract=True): labels / scales / bias per kept proposal (extract_embedding.py:1238-1260), the
    device-side image x class scores, the saved-corpus scorer and the corpus loop (padding of the last batch)."""
    from oracle import retrieval as R, synth
    from wedetect_b200 import retrieval as WR
    from wedetect_b200.detector import SimpleYOLOWorldDetector
    B, H = 2, 320
    sd = synth.synth_state_dict("base", seed=0, uni=True, regime="sparse")
    x = synth.synth_images(3, H, H, seed=2)
    with torch.no_grad():
        ref = R.extract_ref(sd, "base", x, tv_numel_thr=20000)
    text = torch.nn.functional.normalize(torch.randn(80, 768, generator=torch.Generator().manual_seed(11)), dim=-1)
    m = SimpleYOLOWorldDetector("base", 768, 256, 300, device=D, precise=True, extract=True)
    m.load_state_dict(sd)
    out = m.forward_tensor(x[:B].to(D))
    scores = m.score_text(text).cpu()
    for b in range(B):
        r, o = ref[b], {k: v.cpu() for k, v in out[b].items()}
        n = len(r["scores"])
        assert len(o["scores"]) == n
        ka = m.last_batch_result["anchors"][b, :n].cpu().long() * 4096 + o["labels"]
        kr = r["anchors"] * 4096 + r["labels"]
        ia, ir = torch.argsort(ka), torch.argsort(kr)
        assert torch.equal(ka[ia], kr[ir]), f"image {b}: kept (anchor, prompt) sets differ"
        assert o["labels"].dtype == torch.int64
        assert torch.equal(o["scales"][ia], r["scales"][ir]) and torch.equal(o["bias"][ia], r["bias"][ir])   # exact: copied level constants
        assert float((o["scores"][ia] - r["scores"][ir]).abs().max()) <= 1e-3
        assert float((o["embeddings"][ia] - r["embeddings"][ir]).abs().max()) <= 1e-2
        want = R.image_scores_ref(r["embeddings"], text, r["scales"], r["bias"])
        assert float((scores[b] - want).abs().max()) <= 1e-3, float((scores[b] - want).abs().max())
    # corpus loop on 3 images with batches of 2 (last batch padded), single process
    corpus = WR.extract_corpus(m, x.to(D), [11, 12, 13], batch_size=2, text_embedding=text)
    assert corpus["image_ids"].tolist() == [11, 12, 13] and corpus["scores"].shape == (3, 80)
    assert [it["image_id"] for it in corpus["image_embedding"]] == [11, 12, 13]
    for i in range(3):
        r = ref[i]
        want = R.image_scores_ref(r["embeddings"], text, r["scales"], r["bias"])
        assert float((corpus["scores"][i] - want).abs().max()) <= 1e-3
        assert corpus["image_embedding"][i]["embedding"].shape == (len(r["scores"]), 768)
    # the saved payload goes through the stand-alone scorer (the reference's retrieval_metric.py loop) to the same scores
    again = WR.score_saved(corpus, device=D, batch_size=2, precise=True)
    assert float((again - corpus["scores"]).abs().max()) <= 1e-5
    classnames = [f"class_{k}" for k in range(80)]
    got = WR.predictions_from_scores(again, corpus["image_ids"].tolist(), classnames, thre=0.3)
    want = R.predictions_ref(corpus, classnames, 0.3)
    near = {(classnames[k], int(corpus["image_ids"][i])) for i in range(3) for k in range(80) if abs(float(again[i, k]) - 0.3) < 1e-4}
    for c in classnames:
        assert {(c, i) for i in got[c]} ^ {(c, i) for i in want[c]} <= near
