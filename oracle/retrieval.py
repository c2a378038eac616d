"""CPU ORACLE of the object-retrieval path (BASELINE config 5 / SURVEY.md §8f-1).  TEST INFRASTRUCTURE ONLY.

Restates, as plain functions:
  extract_ref        eval_retrieval/extract_embedding.py:1181-1260 (head_predict of the extract variant: proposals with
                     embeddings, labels and the per-level `scales` / `bias` rows) on top of the pinned functional oracle
                     and the C post-process oracle
  image_scores_ref   eval_retrieval/retrieval_metric.py:365-373 (einsum, *exp(scale)+bias, sigmoid, max over proposals)
  predictions_ref    eval_retrieval/retrieval_metric.py:362-377 (threshold -> PREDICTIONS[classname] lists)
Pinned by tests/test_oracle_pin.py::test_retrieval_oracle_matches_reference_goldens against
tests/golden/reference_retrieval.pt, which tests/golden/make_golden_retrieval.py produced by executing the reference's
own extract_embedding.SimpleYOLOWorldDetector.head_predict and the reference's own scoring lines.

Batch semantics: the reference's head_predict re-binds `scales` / `bias` inside its per-image loop (:1246-1247), so with
more than one image per call later images index the first image's filtered rows; its script only ever passes one image
(:1723-1724).  The oracle (and the product) implement that one-image behaviour for every image of a batch.
"""
import torch

from wedetect_b200 import schema
from . import functional as Fn
from .postprocess import postprocess_ref, identity_meta

HM = Fn.HM


def extract_ref(sd, size, images, *, num_proposals=300, tv_numel_thr=4000):
    """images fp32 [B,3,H,W] RGB in [0,1].  Returns a list of dicts like the reference's head_predict (boxes in input
    coordinates, un-clamped below at 0 exactly as the C post-process oracle reports them)."""
    B, _, H, W = images.shape
    out = Fn.vision_forward(sd, size, images, prompts=sd["embeddings"])
    lhw = schema.level_hw(H, W)
    K = sd["embeddings"].shape[0]
    meta, _ = identity_meta(B, H, W)
    det = postprocess_ref([lv["logits"].reshape(-1, K) for lv in out["levels"]], [lv["dist"].reshape(-1, 4) for lv in out["levels"]], lhw,
                          list(schema.STRIDES), K=K, B=B, score_thr=0.0, nms_pre=30000, iou_thr=0.7, max_per_img=num_proposals, nms_mode=1,
                          tv_numel_thr=tv_numel_thr, img_meta=meta, clamp_wh=torch.full((B, 2), 1e9))
    sizes = [h * w for h, w in lhw]
    # extract_embedding.py:1181-1190: one logit_scale / bias value per anchor, by pyramid level
    scales = torch.cat([torch.full((n,), float(sd[HM + f"cls_contrasts.{l}.logit_scale"])) for l, n in enumerate(sizes)])
    bias = torch.cat([torch.full((n,), float(sd[HM + f"cls_contrasts.{l}.bias"])) for l, n in enumerate(sizes)])
    embed = torch.cat([lv["embed"] for lv in out["levels"]], 1)            # [B, A, 768], after the contrastive head's BN
    res = []
    for b in range(B):
        n = int(det["counts"][b])
        a = det["anchors"][b, :n].long()
        res.append(dict(bboxes=det["boxes"][b, :n], embeddings=embed[b, a], scores=det["scores"][b, :n], labels=det["labels"][b, :n].long(),
                        scales=scales[a], bias=bias[a], anchors=a))
    return res


def image_scores_ref(embedding, text_embedding, scale, bias, model="wedetect"):
    """[n,768], [K,768], [n], [n] -> [K]: retrieval_metric.py:365-373."""
    cls_logits = torch.einsum("bw,kw->bk", embedding, text_embedding)
    if model == "hqclip":
        cls_logits = torch.sigmoid(cls_logits)
    else:
        cls_logits = torch.sigmoid(cls_logits * scale.exp().unsqueeze(1) + bias.unsqueeze(1))
    return torch.max(cls_logits, dim=0)[0]


def predictions_ref(pred, classnames, thre, model="wedetect"):
    out = {name: [] for name in classnames}
    for result in pred["image_embedding"]:
        s = image_scores_ref(result["embedding"], pred["text_embedding"], result["scale"], result["bias"], model)
        for k in torch.where(s > thre)[0].tolist():
            out[classnames[k]].append(result["image_id"])
    return out
