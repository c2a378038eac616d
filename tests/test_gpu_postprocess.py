"""GPU post-process (C ABI) vs the C oracle: indices, labels, counts, scores and boxes must be IDENTICAL."""
import pytest
import torch

pytestmark = pytest.mark.gpu
D = "cuda:0"


def make_inputs(B, K, level_hw, seed, regime, quantize=None):
    g = torch.Generator().manual_seed(seed)
    logits, dists = [], []
    for h, w in level_hw:
        z = torch.randn(B * h * w, K, generator=g)
        if regime == "sparse":
            z = z * 1.5 - 10.0     # ~2 % above logit(0.001): trained-like
        elif regime == "dense":
            z = z * 0.3           # everything passes 0.001
        if quantize:
            z = torch.round(z / quantize) * quantize
        logits.append(z)
        dists.append(torch.rand(B * h * w, 4, generator=g) * 7.5)
    return logits, dists


def run_both(B, K, level_hw, strides, logits, dists, meta, clamp, **kw):
    from wedetect_b200 import _lib as L, ops
    from oracle.postprocess import postprocess_ref
    ref = postprocess_ref(logits, dists, level_hw, strides, K=K, B=B, img_meta=meta, clamp_wh=clamp, **kw)
    pp = ops.PostProcess(logits=[t.to(D) for t in logits], dists=[t.to(D) for t in dists], level_hw=level_hw, strides=strides,
                         K=K, B=B, img_meta=meta.to(D), clamp_wh=clamp.to(D), **kw)
    L.run_op(pp.op, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = dict(boxes=pp.boxes.cpu(), scores=pp.scores.cpu(), labels=pp.labels.cpu(), anchors=pp.anchors.cpu(), counts=pp.counts.cpu())
    return got, ref


def assert_identical(got, ref, tag):
    assert torch.equal(got["counts"], ref["counts"]), f"{tag}: counts {got['counts'].tolist()} vs {ref['counts'].tolist()}"
    for k in ("anchors", "labels"):
        if not torch.equal(got[k], ref[k]):
            bad = (got[k] != ref[k]).nonzero()
            raise AssertionError(f"{tag}: {k} differ at {bad[:8].tolist()} ({len(bad)} places); got {got[k][tuple(bad[0])]} ref {ref[k][tuple(bad[0])]}")
    assert torch.equal(got["scores"].view(torch.int32), ref["scores"].view(torch.int32)), f"{tag}: score bits differ"
    assert torch.equal(got["boxes"].view(torch.int32), ref["boxes"].view(torch.int32)), f"{tag}: box bits differ"


HW640 = [(80, 80), (40, 40), (20, 20)]
STR = [8, 16, 32]


@pytest.mark.parametrize("regime,nms_mode,K", [("sparse", 0, 80), ("dense", 0, 80), ("sparse", 1, 256), ("dense", 1, 256), ("sparse", 0, 5)])
def test_postprocess_exact(regime, nms_mode, K):
    from oracle.postprocess import identity_meta
    B = 3
    logits, dists = make_inputs(B, K, HW640, seed=11 + K, regime=regime)
    meta, clamp = identity_meta(B, 640, 640)
    got, ref = run_both(B, K, HW640, STR, logits, dists, meta, clamp, score_thr=0.001 if nms_mode == 0 else 0.0, nms_pre=30000,
                        iou_thr=0.7, max_per_img=300 if nms_mode == 0 else 1000, nms_mode=nms_mode)
    assert_identical(got, ref, f"{regime}/mode{nms_mode}/K{K}")
    assert int(ref["counts"].min()) > 0


def test_postprocess_ties_and_truncation():
    """Quantised logits -> massive score ties, small nms_pre -> the cut falls inside a tie group."""
    from oracle.postprocess import identity_meta
    B, K = 2, 16
    hw = [(20, 20), (10, 10), (5, 5)]
    logits, dists = make_inputs(B, K, hw, seed=5, regime="dense", quantize=0.125)
    meta, clamp = identity_meta(B, 160, 160)
    got, ref = run_both(B, K, hw, STR, logits, dists, meta, clamp, score_thr=0.3, nms_pre=1000, iou_thr=0.5, max_per_img=100, nms_mode=0)
    assert_identical(got, ref, "ties")


def test_postprocess_rescale_and_empty():
    """mmdet-style rescale before NMS (pad / scale_factor), Uni-style un-letterbox after; one image with no candidates."""
    B, K = 3, 80
    logits, dists = make_inputs(B, K, HW640, seed=21, regime="sparse")
    for l, (h, w) in enumerate(HW640):
        logits[l].view(B, h * w, K)[1] = -20.0   # image 1: nothing passes
    meta = torch.tensor([[12.0, 8.0, 0.8, 0.75, 0, 0, 1, 0], [0, 0, 1, 1, 0, 0, 1, 0], [0, 0, 1, 1, 16.0, 4.0, 0.6400000, 0]], dtype=torch.float32)
    clamp = torch.tensor([[500.0, 375.0], [640, 640], [1000, 955]], dtype=torch.float32)
    got, ref = run_both(B, K, HW640, STR, logits, dists, meta, clamp, score_thr=0.001, nms_pre=30000, iou_thr=0.7, max_per_img=300, nms_mode=0)
    assert_identical(got, ref, "rescale")
    assert int(ref["counts"][1]) == 0


def test_postprocess_large_800():
    """C3-shaped problem (800x800, K=1203), sparse regime, single image."""
    from oracle.postprocess import identity_meta
    B, K = 1, 1203
    hw = [(100, 100), (50, 50), (25, 25)]
    logits, dists = make_inputs(B, K, hw, seed=31, regime="sparse")
    meta, clamp = identity_meta(B, 800, 800)
    got, ref = run_both(B, K, hw, STR, logits, dists, meta, clamp, score_thr=0.001, nms_pre=30000, iou_thr=0.7, max_per_img=300, nms_mode=0)
    assert_identical(got, ref, "C3")
