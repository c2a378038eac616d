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


def _assert_same_detections(bx, sc, lb, rbx, rsc, rlb, what, cut=False):
    """Same detections as the reference: a one-to-one match on (label, box within 0.1 px) with scores within 1e-3, both lists in
    descending score order; positions may differ only between detections whose reference scores are closer than 1e-3 (the score
    sort of near-ties).  cut=True: the lists were truncated by score (threshold / top-k), so near-ties at the cut may differ."""
    assert bool((sc[:-1] >= sc[1:]).all()), f"{what}: not in descending score order"
    n, m = len(sc), len(rsc)
    assert n == m or cut, f"{what}: {n} detections, reference {m}"
    assert n > 0 and m > 0, f"{what}: empty"
    cost = (bx[:, None, :] - rbx[None, :, :]).abs().amax(-1) + 1e3 * (lb[:, None] != rlb[None, :]).float()
    # several reference detections can share a box and a label (see below): among the candidates within 0.1 px take the closest score
    near = torch.where(cost <= 0.1, (sc[:, None] - rsc[None, :]).abs(), torch.full_like(cost, float("inf")))
    j = near.argmin(1)
    ok = cost[torch.arange(n), j] <= 0.1
    if cut:      # unmatched entries must sit at the cut: their score is within 1e-3 of the lowest kept reference score
        assert bool((ok | ((sc - rsc.min()).abs() <= 1e-3)).all()), f"{what}: detections without a reference counterpart"
    else:
        # (boxes are clamped to the image AFTER the NMS, so distinct anchors can end as identical boxes: match both ways, not one-to-one)
        assert bool(ok.all()) and bool((cost.amin(0) <= 0.1).all()), f"{what}: detections without a counterpart"
    assert float((sc[ok] - rsc[j[ok]]).abs().max()) <= 1e-3, f"{what}: scores differ"
    moved = (j != torch.arange(n)) & ok
    for i in moved.nonzero().flatten().tolist():
        if i < m:
            assert abs(float(rsc[i] - rsc[j[i]])) <= 1e-3, f"{what}: position {i} holds the reference's {int(j[i])} and their scores are not a near-tie"


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
        _assert_same_detections(p.bboxes.cpu(), p.scores.cpu(), p.labels.cpu(), det["boxes"][b, :n], det["scores"][b, :n], det["labels"][b, :n].long(), f"image {b}")
        oh, ow = metas[b]["ori_shape"]
        assert float(p.bboxes.min()) >= 0 and float(p.bboxes[:, 0::2].max()) <= ow and float(p.bboxes[:, 1::2].max()) <= oh
        q = plain[b].pred_instances
        assert torch.equal(q.labels, p.labels) and torch.equal(q.bboxes, p.bboxes) and torch.equal(q.scores, p.scores)
        # ---- infer_wedetect.py:117-126 on the device results ----
        assert n > 45
        thr, max_dets = float(det["scores"][b, 40]) - 1e-6, 20     # ~40 detections survive the threshold, top-k keeps 20 of them
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
        assert arr["bboxes"].shape == (len(rs), 4) and arr["labels"].dtype.kind == "i"
        _assert_same_detections(torch.from_numpy(arr["bboxes"]), torch.from_numpy(arr["scores"]), torch.from_numpy(arr["labels"]), rb, rs, rl.long(),
                                f"image {b} after the infer tail", cut=True)
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


def test_eval_loop_groups_bs1_batches_on_the_device():
    """test.py's loop (bs-1 loader, config/wedetect_base.py:197-204) through wedetect_b200.loop.TestLoop: the device runs grouped
    batches, the evaluator receives every image's own detections in loader order, identical to one test_step per image; then the
    WeDetect-Ref hand-off keeps Uni proposals on the device."""
    from oracle import synth
    from wedetect_b200.api import DetDataSample, SimpleYOLOWorldDetector, TestLoop, init_detector, proposals_for_ref
    tok = FixtureTokenizer("coco_zh")
    sd = synth.synth_state_dict("base", seed=0, with_text=True, regime="sparse")
    model = init_detector(CFG, checkpoint=dict(state_dict=sd), device=D)
    model._tokenizer = tok
    texts = [[t] for t in tok.texts[:12]]
    g = torch.Generator().manual_seed(21)
    imgs = (torch.rand(5, 3, 320, 320, generator=g) * 255).to(torch.uint8)
    mk = lambda i: DetDataSample(dict(img_id=i, ori_shape=(320, 320), img_shape=(320, 320), scale_factor=(1.0, 1.0), pad_param=(0.0, 0.0, 0.0, 0.0), texts=texts))  # noqa: E731

    class Loader(list):
        dataset = list(range(5))

    class Collect:
        def __init__(self):
            self.res = []

        def process(self, data_samples, data_batch):
            self.res += data_samples

        def evaluate(self, size):
            return dict(size=size, n=len(self.res))

    ev = Collect()
    assert TestLoop(model, Loader(dict(inputs=[imgs[i]], data_samples=[mk(i)]) for i in range(5)), ev, group=4).run() == dict(size=5, n=5)
    for i, d in enumerate(ev.res):
        alone = model.test_step(dict(inputs=imgs[i:i + 1].to(D), data_samples=[mk(i)]))[0].pred_instances
        assert d["img_id"] == i and len(alone.scores) > 0
        # (a batch of one has single-m-tile GEMMs at the 10 x 10 level: one k-block per TMEM accumulator instead of two, so the
        #  last bits may differ from the grouped batch; equal tile configurations are bit-identical: test_full_size_batch_invariance...)
        _assert_same_detections(d["pred_instances"]["bboxes"].cpu(), d["pred_instances"]["scores"].cpu(), d["pred_instances"]["labels"].cpu(),
                                alone.bboxes.cpu(), alone.scores.cpu(), alone.labels.cpu(), f"image {i}: grouped vs alone")
        assert float((d["pred_instances"]["scores"] - alone.scores).abs().max()) <= 1e-5
    # WeDetect-Ref hand-off (infer_wedetect_ref.py:27,67-74,91)
    uni = SimpleYOLOWorldDetector("base", 768, 256, 100, device=D)
    uni.load_state_dict(synth.synth_state_dict("base", seed=0, uni=True, regime="sparse"))
    arr = [imgs[i].flip(0).permute(1, 2, 0).contiguous().numpy() for i in range(2)]
    boxes, counts = proposals_for_ref(uni(arr), torch.bfloat16)
    assert counts == [100, 100] and all(b.is_cuda and b.dtype == torch.bfloat16 and b.shape == (100, 4) for b in boxes)


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
