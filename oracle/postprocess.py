"""Python front-end of the C post-process oracle (oracle/postprocess_ref.c).  TEST INFRASTRUCTURE."""
import ctypes

import torch

from wedetect_b200._lib import PPParams  # struct layout only (shared with the C ABI header)
from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.wd_ref_postprocess.argtypes = [ctypes.POINTER(PPParams)]
        _lib.wd_ref_postprocess.restype = ctypes.c_int
    return _lib


def identity_meta(B, H, W):
    meta = torch.zeros(B, 8, dtype=torch.float32)
    meta[:, 2] = 1.0
    meta[:, 3] = 1.0
    meta[:, 6] = 1.0
    clamp = torch.tensor([[float(W), float(H)]] * B, dtype=torch.float32)
    return meta, clamp


def postprocess_ref(logits, dists, level_hw, strides, *, K, B, score_thr, nms_pre, iou_thr, max_per_img, nms_mode,
                    img_meta, clamp_wh, tv_numel_thr=20000):
    """logits[l]: f32 [B*H_l*W_l, >=K] (CPU), dists[l]: f32 [B*H_l*W_l, 4].  Returns dict of CPU tensors."""
    lib = _load()
    logits = [t.detach().cpu().float().contiguous() for t in logits]
    dists = [t.detach().cpu().float().contiguous() for t in dists]
    img_meta = img_meta.detach().cpu().float().contiguous()
    clamp_wh = clamp_wh.detach().cpu().float().contiguous()
    out = dict(
        boxes=torch.empty(B, max_per_img, 4, dtype=torch.float32),
        scores=torch.empty(B, max_per_img, dtype=torch.float32),
        labels=torch.empty(B, max_per_img, dtype=torch.int32),
        anchors=torch.empty(B, max_per_img, dtype=torch.int32),
        counts=torch.zeros(B, dtype=torch.int32),
    )
    p = PPParams()
    p.B, p.K, p.nlevels = B, K, len(logits)
    for l, ((h, w), s) in enumerate(zip(level_hw, strides)):
        p.lvl_h[l], p.lvl_w[l], p.lvl_stride[l] = h, w, s
        assert logits[l].shape[0] == B * h * w and dists[l].shape == (B * h * w, 4)
        p.ld_logit[l] = logits[l].stride(0)
        p.logits[l] = logits[l].data_ptr()
        p.dist[l] = dists[l].data_ptr()
    p.score_thr, p.nms_pre, p.iou_thr, p.max_per_img = score_thr, nms_pre, iou_thr, max_per_img
    p.nms_mode, p.tv_numel_thr, p.multi_label = nms_mode, tv_numel_thr, 1
    p.img_meta, p.clamp_wh = img_meta.data_ptr(), clamp_wh.data_ptr()
    p.out_boxes, p.out_scores = out["boxes"].data_ptr(), out["scores"].data_ptr()
    p.out_labels, p.out_anchor, p.out_counts = out["labels"].data_ptr(), out["anchors"].data_ptr(), out["counts"].data_ptr()
    rc = lib.wd_ref_postprocess(ctypes.byref(p))
    if rc != 0:
        raise RuntimeError(f"wd_ref_postprocess failed: {rc}")
    return out
