"""The reference's detection entry point end to end on the GPU, through the drop-in facade, against the oracle chain.

    init_detector(config, checkpoint)                          infer_wedetect.py:156
    model.reparameterize(texts)                                infer_wedetect.py:163-168   (81 prompts: 80 COCO classes + ' ')
    model.test_step(dict(inputs=uint8 BGR, data_samples=...))  infer_wedetect.py:113-116   (pad_param / scale_factor un-mapping, clamp)
    pred[pred.scores > thr], topk(max_dets), .cpu().numpy()    infer_wedetect.py:117-126

The oracle side is preprocess (BGR->RGB, /255) -> XLM-R text tower on the reference tokenizer's ids (tests/golden/tokens_coco_zh.json)
-> ConvNeXt / neck / head -> C post-process with the same rescale metadata.  Covers SURVEY §8 rows a1, a7, a8, a12 on the device."""
import os

import pytest
import torch

from util import FixtureTokenizer

pytestmark = pytest.mark.gpu
D = "cuda:0"
CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "wedetect_base_min.py")


def _letterboxed_batch(seed=11):
    """Two 640x640 uint8 BGR inputs as the mmdet test pipeline would hand them over: image 0 is a 480x600 original resized by
    640/600 and padded top / bottom with 114; image 1 is a native 640x640 picture."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(2, 3, 640, 640, generator=g) * 255).to(torch.uint8)
    x[0, :, :64] = 114
    x[0, :, 576:] = 114
    sf = 640.0 / 600.0
    metas = [dict(img_id=0, img_path="a.jpg", ori_shape=(480, 600), img_shape=(640, 640), scale_factor=(sf, sf), pad_param=(64.0, 64.0, 0.0, 0.0)),
             dict(img_id=1, img_path="b.jpg", ori_shape=(640, 640), img_shape=(640, 640), scale_factor=(1.0, 1.0), pad_param=(0.0, 0.0, 0.0, 0.0))]
    return x, metas


def test_infer_wedetect_chain_matches_oracle():
    from oracle import functional as Fn, synth
    from oracle.postprocess import postprocess_ref
    from wedetect_b200 import schema
    from wedetect_b200.api import DetDataSample, init_detector
    tok = FixtureTokenizer("coco_zh")
    texts = [[t] for t in tok.texts]                       # infer_wedetect.py:163-167: [[caption], ..., [' ']]
    K, B, H, W = len(texts), 2, 640, 640
    sd = synth.synth_state_dict("base", seed=0, with_text=True, regime="sparse")
    x, metas = _letterboxed_batch()

    # ---------------- ours: the reference's call sequence ----------------
    model = init_detector(CFG, checkpoint=dict(state_dict=sd), device=D)
    model._tokenizer = tok
    model.reparameterize(texts)
    assert tuple(model.text_feats.shape) == (1, K, schema.EMBED_DIM)
    samples = [DetDataSample(dict(m, texts=texts)) for m in metas]        # the pipeline's samples carry the prompts too (yolo_world.py:88-96)
    out = model.test_step(dict(inputs=x.to(D), data_samples=samples))
    assert len(out) == B
    plain = model.test_step(dict(inputs=x.to(D), data_samples=[DetDataSample(dict(m)) for m in metas]))   # prompts from reparameterize

    # ---------------- oracle chain ----------------
    with torch.no_grad():
        feats = Fn.text_tower(sd, "base", tok.ids.int(), tok.mask.int())
        ref = Fn.vision_forward(sd, "base", Fn.preprocess(x), text=feats)
    assert float((model.text_feats[0].cpu() - feats).abs().max()) <= 2e-6, "text tower (tokenizer ids -> 768-d class embeddings)"
    meta = torch.tensor([[m["pad_param"][2], m["pad_param"][0], m["scale_factor"][0], m["scale_factor"][1], 0.0, 0.0, 1.0, 0.0] for m in metas])
    clamp = torch.tensor([[float(m["ori_shape"][1]), float(m["ori_shape"][0])] for m in metas])
    det = postprocess_ref([lv["logits"].reshape(-1, K) for lv in ref["levels"]], [lv["dist"].reshape(-1, 4) for lv in ref["levels"]], schema.level_hw(H, W),
                          list(schema.STRIDES), K=K, B=B, score_thr=0.001, nms_pre=30000, iou_thr=0.7, max_per_img=300, nms_mode=0, img_meta=meta, clamp_wh=clamp)

    for b in range(B):
        n = int(det["counts"][b])
        p = out[b].pred_instances
        assert len(p.scores) == n and p.labels.dtype == torch.int64 and p.bboxes.dtype == torch.float32
        key = lambda bx, lb: sorted(zip(lb.tolist(), [tuple(round(v, 0) for v in r) for r in bx.tolist()]))  # noqa: E731
        # identical (label, score-order) assignment; boxes in ORIGINAL-image coordinates within 0.1 px, inside the image
        assert torch.equal(p.labels.cpu(), det["labels"][b, :n].long()), f"image {b}: labels / order differ"
        assert float((p.scores.cpu() - det["scores"][b, :n]).abs().max()) <= 1e-3
        assert float((p.bboxes.cpu() - det["boxes"][b, :n]).abs().max()) <= 0.1
        oh, ow = metas[b]["ori_shape"]
        assert float(p.bboxes.min()) >= 0 and float(p.bboxes[:, 0::2].max()) <= ow and float(p.bboxes[:, 1::2].max()) <= oh
        q = plain[b].pred_instances
        assert torch.equal(q.labels, p.labels) and torch.equal(q.bboxes, p.bboxes) and torch.equal(q.scores, p.scores)
        # ---- infer_wedetect.py:117-126 on the device results ----
        thr, max_dets = 0.01, 20
        sel = p[p.scores.float() > thr]
        if len(sel.scores) > max_dets:
            sel = sel[sel.scores.float().topk(max_dets)[1]]
        arr = sel.cpu().numpy()
        rs, rb, rl = det["scores"][b, :n], det["boxes"][b, :n], det["labels"][b, :n]
        m = rs > thr
        rs, rb, rl = rs[m], rb[m], rl[m]
        if len(rs) > max_dets:
            idx = rs.topk(max_dets)[1]
            rs, rb, rl = rs[idx], rb[idx], rl[idx]
        assert arr["bboxes"].shape == (len(rs), 4) and arr["labels"].tolist() == rl.tolist()
        assert abs(arr["scores"] - rs.numpy()).max() <= 1e-3 and abs(arr["bboxes"] - rb.numpy()).max() <= 0.1
    # results handed to the caller are the caller's: a later step must not rewrite them
    keep = out[0].pred_instances.bboxes.clone()
    model.test_step(dict(inputs=x.flip(0).to(D), data_samples=[DetDataSample(dict(m)) for m in metas]))
    assert torch.equal(out[0].pred_instances.bboxes, keep)


def test_per_image_prompt_lists_inside_one_batch():
    """mm_backbone.py:376-390 encodes B x K prompts: images of one batch may carry different prompt lists of equal length.  The
    facade runs one group per distinct list; results must equal running each image alone with its list."""
    from oracle import synth
    from wedetect_b200.api import DetDataSample, init_detector
    tok = FixtureTokenizer("coco_zh")
    sd = synth.synth_state_dict("base", seed=0, with_text=True, regime="sparse")
    model = init_detector(CFG, checkpoint=dict(state_dict=sd), device=D)
    model._tokenizer = tok
    x, metas = _letterboxed_batch(seed=13)
    x = x[:, :, :320, :320].contiguous()
    lists = [[[t] for t in tok.texts[0:6]], [[t] for t in tok.texts[10:16]]]
    mk = lambda b: DetDataSample(dict(ori_shape=(320, 320), img_shape=(320, 320), scale_factor=(1.0, 1.0), pad_param=(0.0, 0.0, 0.0, 0.0), texts=lists[b]))  # noqa: E731
    both = model.test_step(dict(inputs=x.to(D), data_samples=[mk(0), mk(1)]))
    for b in range(2):
        alone = model.test_step(dict(inputs=x[b:b + 1].to(D), data_samples=[mk(b)]))[0].pred_instances
        got = both[b].pred_instances
        assert len(got.scores) == len(alone.scores) > 0
        assert torch.equal(got.labels, alone.labels) and torch.equal(got.bboxes, alone.bboxes) and torch.equal(got.scores, alone.scores)
    with pytest.raises(AssertionError):       # unequal prompt counts inside a batch: the reference asserts (mm_backbone.py:378-380)
        bad = mk(1)
        bad.set_metainfo(dict(texts=lists[1][:3]))
        model.test_step(dict(inputs=x.to(D), data_samples=[mk(0), bad]))


@pytest.mark.parametrize("name,size", [("coco_zh", "base"), ("lvis_v1_zh", "large")])
def test_text_tower_reference_token_shapes(name, size):
    """The text tower at the token shapes of BASELINE configs 2 and 3: 81 x 8 (COCO prompts, XLM-R base) and 1204 x 9 (LVIS
    prompts, XLM-R large), on the reference tokenizer's ids; a ragged batch of short prompts with padding."""
    from oracle import functional as Fn, synth
    from wedetect_b200 import plan, weights
    tok = FixtureTokenizer(name)
    S, L = tok.ids.shape
    ids, mask = tok.ids[: min(S, 1204)], tok.mask[: min(S, 1204)]
    vocab = int(ids.max()) + 1
    sd = synth.synth_state_dict(size, seed=2, with_text=True, text_vocab=vocab, calibrate=False)
    sub = slice(0, S, max(1, S // 40))            # the CPU oracle checks every 30th LVIS prompt (the large tower takes ~1 s per 10 prompts)
    with torch.no_grad():
        ref = Fn.text_tower(sd, size, ids[sub].int(), mask[sub].int())
    Wt = weights.prepare_text(sd, size, D)
    tp = plan.TextPlan(Wt, size, ids.shape[0], L)
    got = tp.run(ids.to(D), mask.to(D)).cpu()
    assert got.shape == (ids.shape[0], 768)
    assert float((got[sub] - ref).abs().max()) <= 5e-6
    assert float((got.norm(dim=-1) - 1).abs().max()) <= 1e-5
