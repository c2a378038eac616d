"""End-to-end GPU parity: the lowered program (C ABI) vs the CPU fp32 oracle, stage by stage.

precise (fp16 hi/lo operands, the default) mode carries the north-star gates: |logit error| <= 1e-3 and IDENTICAL kept
(anchor, class) indices; fast (bf16) mode is held to measured engineering tolerances.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
D = "cuda:0"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _nhwc(t):  # oracle NCHW -> rows x C
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])


def _err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    d = (got - ref).abs()
    return dict(max_abs=float(d.max()), rel_rms=float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-12)), ref_absmax=float(ref.abs().max()))


def dfl_tolerance(prob, eps=1e-3):
    """First-order image of the logit gate on a DFL distance d = sum_j j p_j: if every one of its 16 bin logits moves by at most
    eps, d moves by at most eps * sum_j p_j |j - d| (the mean absolute deviation of the bin distribution, <= 7.5).  Floored at
    eps itself, so sharply peaked anchors keep the plain 1e-3 gate."""
    j = torch.arange(prob.shape[-1], dtype=prob.dtype)
    d = (prob * j).sum(-1, keepdim=True)
    mad = (prob * (j - d).abs()).sum(-1)
    return eps * mad.clamp_min(1.0)


def run_case(size, B, H, W, K, *, uni, precise, regime, seed=0, max_per_img=300, input_u8=False, fp64=False):
    from oracle import functional as Fn, synth
    from oracle.postprocess import postprocess_ref, identity_meta
    from wedetect_b200 import plan, schema, weights
    sd = synth.synth_state_dict(size, seed=seed, uni=uni, num_prompts=K, with_text=False, regime=regime)
    imgs = synth.synth_images(B, H, W, seed=seed + 2)
    g = torch.Generator().manual_seed(seed + 5)
    text = None if uni else torch.randn(K, schema.EMBED_DIM, generator=g)
    u8 = None
    if input_u8:   # the mmdet pipeline's form: uint8 BGR; the oracle applies the reference preprocessor (BGR->RGB, /255) to the same bytes
        u8 = (imgs * 255).to(torch.uint8).flip(1).contiguous()
        imgs = Fn.preprocess(u8)
    with torch.no_grad():
        ref = Fn.vision_forward(sd, size, imgs, text=text, prompts=sd.get("embeddings"))
    Wt = weights.prepare_vision(sd, size, D, precise=precise, input_format="u8_bgr" if input_u8 else "f32_rgb")
    kw = dict(score_thr=0.0, nms_mode=1, max_per_img=max_per_img) if uni else dict(score_thr=0.001, nms_mode=0, max_per_img=max_per_img)
    p = plan.VisionPlan(Wt, size, B, H, W, K=K, uni=uni, input_dtype=torch.uint8 if input_u8 else torch.float32, **kw)
    if not uni:
        p.set_text(text.to(D))
    p.image.copy_((u8 if input_u8 else imgs).to(D))
    p.run()
    torch.cuda.synchronize()
    errs = {}
    if fp64:
        # the same oracle code in float64: how far is the fp32 reference itself from the exact result, and how far are we
        sd64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sd.items()}
        with torch.no_grad():
            r64 = Fn.vision_forward(sd64, size, imgs.double(), text=None if text is None else text.double(), prompts=sd64.get("embeddings"))
        for l in range(3):
            for key, ours in (("logits", p.logits[l][:, :K]), ("dist", p.dists[l])):
                t64 = r64["levels"][l][key].reshape(ours.shape)
                errs[f"{key}{l}_vs_fp64"] = dict(ours=float((ours.cpu().double() - t64).abs().max()),
                                                 oracle_fp32=float((ref["levels"][l][key].reshape(ours.shape).double() - t64).abs().max()))
    for s in range(4):
        errs[f"c{s + 1}"] = _err(p.stage_x[s], _nhwc(ref["backbone"][s]))
    for l in range(3):
        errs[f"p{l + 3}"] = _err(p.pyramid[l].value(), _nhwc(ref["neck"][l]))
        lv = ref["levels"][l]
        ge = p.embeds[l].value() * Wt[f"head.contrast.{l}.g"] + Wt[f"head.contrast.{l}.h"]
        errs[f"embed{l}"] = _err(ge, lv["embed"].reshape(-1, schema.EMBED_DIM))
        errs[f"logit{l}"] = _err(p.logits[l][:, :K], lv["logits"].reshape(-1, K))
        errs[f"dist{l}"] = _err(p.dists[l], lv["dist"].reshape(-1, 4))
        errs[f"dist{l}"]["max_over_tol"] = float(((p.dists[l].cpu() - lv["dist"].reshape(-1, 4)).abs() / dfl_tolerance(lv["dfl_prob"].reshape(-1, 4, schema.REG_MAX))).max())
    # reference detections from the ORACLE's own logits / distances
    meta, clamp = identity_meta(B, H, W)
    lhw = schema.level_hw(H, W)
    det_ref = postprocess_ref([lv["logits"].reshape(-1, K) for lv in ref["levels"]], [lv["dist"].reshape(-1, 4) for lv in ref["levels"]], lhw,
                              list(schema.STRIDES), K=K, B=B, score_thr=kw["score_thr"], nms_pre=30000, iou_thr=0.7, max_per_img=max_per_img,
                              nms_mode=kw["nms_mode"], img_meta=meta, clamp_wh=clamp)
    det = {k: v.cpu() for k, v in p.results().items()}
    same = []
    for b in range(B):
        n = int(det_ref["counts"][b])
        a = set(zip(det["anchors"][b, :int(det["counts"][b])].tolist(), det["labels"][b, :int(det["counts"][b])].tolist()))
        r = set(zip(det_ref["anchors"][b, :n].tolist(), det_ref["labels"][b, :n].tolist()))
        same.append(len(a & r) / max(1, len(r)))
    errs["det_overlap"] = same
    errs["det_counts"] = [det["counts"].tolist(), det_ref["counts"].tolist()]
    os.makedirs(OUT, exist_ok=True)
    tag = f"{size}_{'uni' if uni else 'text'}_{'precise' if precise else 'fast'}_{regime}_{H}x{W}_B{B}_K{K}" + (f"_P{max_per_img}" if max_per_img != 300 else "") + ("_u8" if input_u8 else "")
    with open(os.path.join(OUT, f"e2e_{tag}.json"), "w") as f:
        json.dump(errs, f, indent=1)
    print(tag, json.dumps({k: (round(v["max_abs"], 5), round(v["rel_rms"], 6)) if isinstance(v, dict) and "max_abs" in v else v for k, v in errs.items()}))
    return errs, det, det_ref, p, ref


@pytest.mark.parametrize("size,K", [("tiny", 5), ("base", 80)])
def test_e2e_fast(size, K):
    errs, det, det_ref, p, ref = run_case(size, 2, 320, 320, K, uni=False, precise=False, regime="sparse")
    for s in range(4):
        assert errs[f"c{s + 1}"]["rel_rms"] < 1e-2, (s, errs[f"c{s + 1}"])
    for l in range(3):
        # bf16 operands: ~0.4 % per GEMM, amplified by this random-weight network (the fp32 oracle itself drifts
        # ~20x from backbone to P5 against an fp64 run); measured 3.5-9 % at the pyramid, see DESIGN.md §Precision
        assert errs[f"p{l + 3}"]["rel_rms"] < 0.25, errs[f"p{l + 3}"]
        assert errs[f"logit{l}"]["max_abs"] < 3.0 and errs[f"logit{l}"]["rel_rms"] < 0.05, errs[f"logit{l}"]
        assert errs[f"dist{l}"]["max_abs"] < 6.0, errs[f"dist{l}"]
    assert min(errs["det_overlap"]) > 0.6, errs["det_overlap"]


def check_north_star(errs, det, det_ref, B):
    """The parity gates: |logit error| <= 1e-3; DFL distances within the first-order image of that gate (dfl_tolerance);
    identical kept (anchor, class) sets, scores within 1e-3, boxes within 0.1 px."""
    for l in range(3):
        assert errs[f"logit{l}"]["max_abs"] <= 1e-3, errs[f"logit{l}"]
        assert errs[f"dist{l}"]["max_over_tol"] <= 1.0, errs[f"dist{l}"]
        assert errs[f"dist{l}"]["max_abs"] <= 5e-3, errs[f"dist{l}"]


# (size, K, uni, B, res, max_per_img): C1 / C2 / C4-default shapes, then the BASELINE config-3 shape (WeDetect-Large against the
# 1203-class LVIS-sized text set: similarity as a dense GEMM with a ragged last tile, > nms_pre candidates per image so the
# top-k cut is exercised) and the config-4 shape (Uni, 1000 proposals kept per image)
@pytest.mark.parametrize("size,K,uni,B,res,max_per_img", [("tiny", 5, False, 2, 320, 300), ("base", 80, False, 2, 320, 300), ("base", 256, True, 2, 320, 300),
                                                          ("large", 1203, False, 1, 256, 300), ("base", 256, True, 2, 320, 1000)])
def test_e2e_precise_north_star(size, K, uni, B, res, max_per_img):
    """default (fp16 hi/lo) path: logits within 1e-3 of the fp32 reference and identical kept indices / labels."""
    errs, det, det_ref, p, ref = run_case(size, B, res, res, K, uni=uni, precise=True, regime="sparse", max_per_img=max_per_img,
                                          fp64=(size, K) == ("base", 80))
    check_north_star(errs, det, det_ref, B)
    check_detections(det, det_ref, B, uni, ref)


def test_e2e_benched_shape_640_u8():
    """A 2-image slice of the benched workload (BASELINE configs[1]: WeDetect-Base, 640x640, K = 80, uint8 BGR input through the
    folded preprocessor) against the oracle chain preprocess -> forward -> post-process.  Batch invariance
    (test_full_size_batch_invariance...) makes the slice representative of the bs-32 step."""
    errs, det, det_ref, p, ref = run_case("base", 2, 640, 640, 80, uni=False, precise=True, regime="sparse", input_u8=True)
    check_north_star(errs, det, det_ref, 2)
    check_detections(det, det_ref, 2, False, ref)


def check_detections(det, det_ref, B, uni, ref):
    # exact on box indices / class assignment: the kept (anchor, class) SET is identical per image; the order may differ
    # only between detections whose scores are closer than the float tolerance (score-sorted, so compare after keying)
    assert torch.equal(det["counts"], det_ref["counts"])
    for b in range(B):
        n = int(det_ref["counts"][b])
        ka = (det["anchors"][b, :n].long() * 4096 + det["labels"][b, :n].long())
        kr = (det_ref["anchors"][b, :n].long() * 4096 + det_ref["labels"][b, :n].long())
        ia, ir = torch.argsort(ka), torch.argsort(kr)
        assert torch.equal(ka[ia], kr[ir]), f"image {b}: kept (anchor, class) sets differ"
        assert float((det["scores"][b, :n][ia] - det_ref["scores"][b, :n][ir]).abs().max()) <= 1e-3
        # boxes = prior +- distance * stride (8 / 16 / 32 px per unit): 0.1 px = 3e-4 of the 320 px test images
        assert float((det["boxes"][b, :n][ia] - det_ref["boxes"][b, :n][ir]).abs().max()) <= 0.1
        s = det["scores"][b, :n]
        assert bool((s[:-1] >= s[1:]).all()), "detections not in descending score order"
        swapped = (det["anchors"][b, :n] != det_ref["anchors"][b, :n]).nonzero().flatten()
        for i in swapped.tolist():   # any positional difference must be a near-tie in the reference's scores
            j = int((kr == ka[i]).nonzero()[0])
            gap = abs(float(det_ref["scores"][b, i] - det_ref["scores"][b, j]))
            assert gap <= 1e-3, (f"image {b}: our position {i} holds the reference's position {j}; reference scores there differ by {gap:.3e}; "
                                 f"ours {float(det['scores'][b, i]):.6f} ref {float(det_ref['scores'][b, j]):.6f}; swapped positions {swapped.tolist()[:20]}")
    if uni:
        lv = torch.cat([l_["embed"] for l_ in ref["levels"]], 1)
        for b in range(B):
            n = int(det_ref["counts"][b])
            want = lv[b, det["anchors"][b, :n].long()]
            assert float((det["embeddings"][b, :n] - want).abs().max()) <= 1e-2


@pytest.mark.parametrize("size,precise", [("tiny", True), ("base", True), ("base", False)])   # base fast: the plan contains the cluster-launched fused block-MLP kernel
def test_cuda_graph_replay_matches_eager(size, precise):
    from oracle import synth
    from wedetect_b200 import plan, schema, weights
    B, H, W, K = 2, 320, 320, 16
    sd = synth.synth_state_dict(size, seed=3, with_text=False, regime="sparse")
    Wt = weights.prepare_vision(sd, size, D, precise=precise)
    p = plan.VisionPlan(Wt, size, B, H, W, K=K)
    p.set_text(torch.randn(K, schema.EMBED_DIM, generator=torch.Generator().manual_seed(1)).to(D))
    p.image.copy_(synth.synth_images(B, H, W).to(D))
    p.run()
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in p.results().items()}
    p.capture()
    p.image.copy_(synth.synth_images(B, H, W).to(D))
    p.run()
    torch.cuda.synchronize()
    for k, v in p.results().items():
        assert torch.equal(v, eager[k]), k


def test_full_size_batch_invariance_and_output_properties():
    """BASELINE configs[1] at full size (Base, bs 32, 640x640, K = 80, fast mode): size-independent properties.
    (1) image sharding is exact: the detections of an image do not depend on which batch it travels in (the basis of the
        multi-GPU partitioning: rank r of N gets a contiguous shard and the all-gather just concatenates);
    (2) per image: counts <= max_per_img, scores descending and above score_thr, labels in range, boxes inside the image and
        well formed, anchors valid, padding rows zeroed / -1."""
    from oracle import synth
    from wedetect_b200 import plan, schema, weights
    size, B, H, W, K = "base", 32, 640, 640, 80
    sd = synth.synth_state_dict(size, seed=0, with_text=False, regime="sparse")
    Wt = weights.prepare_vision(sd, size, D, input_format="u8_bgr")
    text = torch.randn(K, schema.EMBED_DIM, generator=torch.Generator().manual_seed(5)).to(D)
    imgs = (synth.synth_images(B, H, W, seed=2) * 255).to(torch.uint8).flip(1).contiguous().to(D)
    out = {}
    for b in (32, 16):
        p = plan.VisionPlan(Wt, size, b, H, W, K=K, input_dtype=torch.uint8, score_thr=0.001, nms_mode=0)
        p.set_text(text)
        res = []
        for s in range(0, B, b):
            p.image.copy_(imgs[s:s + b])
            p.run()
            torch.cuda.synchronize()
            res.append({k: v.clone() for k, v in p.results().items()})
        out[b] = {k: torch.cat([r[k] for r in res]) for k in res[0]}
        del p
        torch.cuda.empty_cache()
    for k in out[32]:
        assert torch.equal(out[32][k], out[16][k]), f"{k}: a batch of 32 and two batches of 16 disagree"
    r = {k: v.cpu() for k, v in out[32].items()}
    A = sum(h * w for h, w in schema.level_hw(H, W))
    assert int(r["counts"].max()) <= 300 and int(r["counts"].min()) >= 1
    for b in range(B):
        n = int(r["counts"][b])
        s = r["scores"][b, :n]
        assert bool((s[:-1] >= s[1:]).all()) and float(s.min()) > 0.001 and float(s.max()) <= 1.0
        assert int(r["labels"][b, :n].min()) >= 0 and int(r["labels"][b, :n].max()) < K
        assert int(r["anchors"][b, :n].min()) >= 0 and int(r["anchors"][b, :n].max()) < A
        bx = r["boxes"][b, :n]
        assert float(bx.min()) >= 0.0 and float(bx[:, 0::2].max()) <= W and float(bx[:, 1::2].max()) <= H
        assert bool((bx[:, 2] >= bx[:, 0]).all()) and bool((bx[:, 3] >= bx[:, 1]).all())
        assert float(r["scores"][b, n:].abs().sum()) == 0.0 and bool((r["labels"][b, n:] == -1).all())
        # (anchor, class) pairs are unique inside an image
        key = r["anchors"][b, :n].long() * K + r["labels"][b, :n].long()
        assert key.unique().numel() == n
