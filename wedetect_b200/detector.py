"""Host-side mirror of the reference's detector interface, backed by the sm_100a program executor.

    YOLOWorldDetector        <- wedetect/models/detectors/yolo_world.py:35-113 (text-conditioned, mmdet path:
                                reparameterize / test_step / predict, called by infer_wedetect.py:113-183)
    SimpleYOLOWorldDetector  <- generate_proposal.py:1052-1218 (WeDetect-Uni: learned prompts, proposals +
                                768-d embeddings), also used by eval_retrieval/extract_embedding.py

Same names, argument meaning and error behaviour; the arithmetic runs in libwedetect_b200.so through
`plan.VisionPlan` / `plan.TextPlan`.  There is no PyTorch / CPU fallback: constructing a detector without
the CUDA library or an sm_100 device raises.
"""
import collections
import contextlib
import itertools
import os

import numpy as np
import torch

from . import _lib as L
from . import plan, schema, weights
from .preprocess import Letterbox, MMTestPipeline, decode_images, decode_images_bgr, letterbox_params  # noqa: F401
from .structures import DetDataSample, InstanceData, instances_for  # noqa: F401

_DEF_TEST_CFG = dict(multi_label=True, nms_pre=30000, score_thr=0.001, nms=dict(type="nms", iou_threshold=0.7), max_per_img=300)


class _LRU(collections.OrderedDict):
    """Bounded cache of per-shape state (plans own every activation buffer of a shape: a serving loop with varying batch
    sizes / prompt counts / resolutions must not grow without bound).  Evicted entries are dropped; their device buffers and
    compiled programs go with the last reference."""

    def __init__(self, cap):
        super().__init__()
        self.cap = cap

    def get_or_make(self, key, make):
        if key in self:
            self.move_to_end(key)
            return self[key]
        v = make()
        self[key] = v
        while len(self) > self.cap:
            self.popitem(last=False)
        return v


def _on(device):
    """Make `device` the current CUDA device for the enclosed library calls (kernel launches, stream lookups, TMA descriptor
    encoding all act on the current device); a no-op for the CPU wiring tests."""
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


def _own(r):
    """The plan's result buffers are static (the CUDA graph rewrites them every step): hand the caller its own copy, as the
    reference does (one small device-to-device copy of the packed [B, max, ...] block per step)."""
    return {k: v.clone() for k, v in r.items()}


def _size_from_cfg(model_cfg):
    return model_cfg["backbone"]["image_model"]["model_name"]


def resolve_texts(batch_data_samples):
    """The class prompts a batch carries, as one flat list of strings, or None when the caller relies on `reparameterize`.
    Mirrors the branches of YOLOWorldDetector.extract_feat (yolo_world.py:84-100): None -> cached features; a dict with
    'texts'; a list of samples with a `texts` attribute (one prompt list per image: they must agree inside a batch, the
    similarity matrix is folded once per step).  Prompts may be strings or single-element lists (infer_wedetect.py:163-167)."""
    if batch_data_samples is None:
        return None
    if isinstance(batch_data_samples, dict):
        if "texts" not in batch_data_samples:
            return None
        texts = batch_data_samples["texts"]
    elif isinstance(batch_data_samples, (list, tuple)) and len(batch_data_samples) and _sample_texts(batch_data_samples[0]) is not None:
        texts = [_sample_texts(s) for s in batch_data_samples]
    else:
        return None
    if texts and isinstance(texts[0], str):
        texts = [texts]                        # a single prompt list for the whole batch
    flat = [[t[0] if isinstance(t, (list, tuple)) else t for t in per_img] for per_img in texts]
    if any(len(f) != len(flat[0]) for f in flat):
        raise AssertionError("number of sequences not equal in batch")       # mm_backbone.py:378-380
    if any(f != flat[0] for f in flat):
        return flat                            # per-image prompt lists (equal counts): predict() runs one group per distinct list
    return flat[0]


def _sample_texts(s):
    get = getattr(s, "get", None)
    return get("texts") if callable(get) else getattr(s, "texts", None)


class YOLOWorldDetector:
    """Text-conditioned detector facade.  `model_cfg` is the `model=` dict of config/wedetect_*.py."""

    def __init__(self, model_cfg=None, *, size=None, test_cfg=None, device="cuda:0", precise=True, tokenizer=None, cuda_graph=True,
                 zero_copy=False):
        """precise=True (default): fp16 hi/lo operands, logits within 1e-3 of the fp32 reference; precise=False: the opt-in
        bf16 mode (faster, does not meet that gate).  zero_copy=True returns views of the plan's static result buffers (valid
        until the next step on the same shape) instead of copies."""
        self.device = torch.device(device)
        L.load(require_gpu=True, device=self.device.index or 0)
        self.zero_copy = bool(zero_copy)
        self.cuda_graph = bool(cuda_graph) and not os.environ.get("WD_NO_GRAPH")
        self.model_cfg = model_cfg
        self.size = size or _size_from_cfg(model_cfg)
        if self.size not in schema.SIZES:
            raise ValueError(f"unknown model size {self.size!r}")
        self.test_cfg = dict(_DEF_TEST_CFG if test_cfg is None else test_cfg)
        if model_cfg is not None and model_cfg.get("test_cfg") and test_cfg is None:
            self.test_cfg = dict(model_cfg["test_cfg"])
        if model_cfg is not None and model_cfg.get("mm_neck", False):
            raise NotImplementedError("mm_neck=True (text-guided neck) is not used by any shipped config")
        if not self.test_cfg.get("multi_label", True):
            # mmdet's single-label branch (argmax class per anchor) is not lowered; no shipped config uses it
            raise NotImplementedError("test_cfg.multi_label=False is not supported (every shipped config sets multi_label=True)")
        self.precise = precise
        self._tokenizer = tokenizer
        self._plans = _LRU(4)
        self._text_plans = _LRU(8)
        self._text_cache = _LRU(32)
        self._vw = {}
        self._tw = None
        self._sd = None
        self.texts = None
        self.text_feats = None
        self.last_batch_result = None

    # --- nn.Module-ish surface used by the entry points ---
    def eval(self):
        return self

    def to(self, device):
        return self

    def load_state_dict(self, sd, strict=False):
        sd = schema.normalize_state_dict(sd)
        need = schema.param_shapes(self.size, with_text=any(k.startswith("backbone.text_model.") for k in sd))
        missing = [k for k in need if k not in sd]
        bad = [k for k in need if k in sd and tuple(sd[k].shape) != tuple(need[k])]
        if bad:
            raise RuntimeError(f"size mismatch for {bad[:4]} ...")
        if missing:
            raise RuntimeError(f"missing keys in checkpoint: {missing[:4]} ... ({len(missing)} total)")
        self._sd = {k: v.detach().float().cpu() for k, v in sd.items()}
        self._plans.clear(); self._text_plans.clear(); self._text_cache.clear(); self._vw.clear()
        self._tw = None
        return self

    # --- text tower ---
    def _get_tokenizer(self):
        if self._tokenizer is None:
            from transformers import AutoTokenizer
            name = self.model_cfg["backbone"]["text_model"]["model_name"]
            self._tokenizer = AutoTokenizer.from_pretrained(name)
        return self._tokenizer

    def encode_tokens(self, ids, mask):
        """ids / mask: int [S, L] (CPU or device).  Returns L2-normalised embeddings [S, 768] on the device."""
        if self._sd is None:
            raise RuntimeError("load_state_dict first")
        with _on(self.device):
            if self._tw is None:
                self._tw = weights.prepare_text(self._sd, self.size, self.device, precise=self.precise)
            S, Lt = ids.shape
            tp = self._text_plans.get_or_make((S, Lt), lambda: plan.TextPlan(self._tw, self.size, S, Lt, device=self.device))
            return tp.run(ids.to(self.device, torch.int32), mask.to(self.device, torch.int32)).clone()

    def forward_text(self, texts):
        """texts: List[List[str]] (one list of class prompts per image); mm_backbone.py:376-390."""
        num = [len(t) for t in texts]
        assert max(num) == min(num), "number of sequences not equal in batch"
        flat = list(itertools.chain(*texts))
        key = tuple(flat)
        key = (key, num[0])
        def make():                           # cached per text set: same tensor object on every call (predict relies on it)
            tok = self._get_tokenizer()(text=flat, return_tensors="pt", padding=True)
            f = self.encode_tokens(tok["input_ids"], tok["attention_mask"])
            return f.reshape(-1, num[0], f.shape[-1])
        return self._text_cache.get_or_make(key, make)

    def reparameterize(self, texts):
        self.texts = texts
        # infer_wedetect.py passes [[t1], [t2], ...]: one "image" per class; flatten to a single class list
        self.text_feats = self.forward_text([list(itertools.chain(*texts))])

    def set_text_features(self, feats):
        """Directly install class embeddings [K, 768] (e.g. cached or synthetic), un-normalised is fine."""
        self.texts = None
        self.text_feats = feats.to(self.device, torch.float32).reshape(1, -1, schema.EMBED_DIM)

    # --- vision ---
    def _plan(self, B, H, W, K, dtype):
        def make():
            fmt = "u8_bgr" if dtype == torch.uint8 else "f32_rgb"
            if fmt not in self._vw:
                self._vw[fmt] = weights.prepare_vision(self._sd, self.size, self.device, input_format=fmt, precise=self.precise)
            tc = self.test_cfg
            if tc.get("nms", {}).get("type", "nms") != "nms":
                raise NotImplementedError(f"nms type {tc['nms']['type']}")
            # mmdet's default nms_pre is 100000 (= "all candidates" for these heads); the device top-k ranks in 16 bits
            cand = sum(h * w for h, w in schema.level_hw(H, W)) * K
            nms_pre = min(int(tc.get("nms_pre", 100000)), cand)
            if nms_pre > 65535:
                raise NotImplementedError(f"nms_pre={nms_pre} exceeds the 65535 candidates per image the post-process ranks (shipped configs: 30000)")
            p = plan.VisionPlan(self._vw[fmt], self.size, B, H, W, K=K, uni=False, input_dtype=dtype, score_thr=float(tc.get("score_thr", -1)),
                                nms_pre=nms_pre, iou_thr=float(tc["nms"]["iou_threshold"]),
                                max_per_img=int(tc["max_per_img"]), nms_mode=0, device=self.device)
            p._text_key = None
            return p
        with _on(self.device):
            return self._plans.get_or_make((B, H, W, K, dtype), make)

    def predict(self, batch_inputs, batch_data_samples, rescale=True):
        if self._sd is None:
            raise RuntimeError("load_state_dict first")
        if isinstance(batch_inputs, (list, tuple)):
            batch_inputs = torch.stack(list(batch_inputs))
        B, _, H, W = batch_inputs.shape
        # text features: texts carried by the samples win (yolo_world.py:88-96), else the reparameterized cache
        flat = resolve_texts(batch_data_samples)
        if isinstance(batch_data_samples, dict):
            batch_data_samples = None          # the dict form carries texts only: no per-image metainfo
        if flat is not None and flat and isinstance(flat[0], list):
            # per-image prompt lists that differ inside the batch (mm_backbone.py:376-390 encodes B x K prompts): the similarity
            # matrix is folded once per distinct list, so run the images group by group and put the results back in order
            groups = {}
            for b, f in enumerate(flat):
                groups.setdefault(tuple(f), []).append(b)
            out = [None] * B
            for f, idx in groups.items():
                sub = [DetDataSample(dict(batch_data_samples[b].metainfo, texts=list(f))) for b in idx] if batch_data_samples else dict(texts=list(f))
                if batch_data_samples:
                    res = self.predict(batch_inputs[idx], sub, rescale)
                    for b, r in zip(idx, res):
                        s = batch_data_samples[b]
                        s.pred_instances = r.pred_instances
                        out[b] = s
                else:
                    for b, r in zip(idx, self.predict(batch_inputs[idx], sub, rescale)):
                        out[b] = r
            return out
        if flat is not None:
            feats = self.forward_text([flat])
        elif self.text_feats is not None:
            feats = self.text_feats
        else:
            raise TypeError("batch_data_samples should be dict or list.")
        src = feats
        feats = feats.reshape(-1, schema.EMBED_DIM)
        K = feats.shape[0]
        p = self._plan(B, H, W, K, batch_inputs.dtype)
        with _on(self.device):
            return self._predict_on(p, src, feats, batch_inputs, batch_data_samples, rescale, B, H, W)

    def _predict_on(self, p, src, feats, batch_inputs, batch_data_samples, rescale, B, H, W):
        if p._text_key is not src:            # fold BN * text * exp(scale) only when the text set changes
            p.set_text(feats.contiguous())
            p._text_key = src
        # per-image rescale metadata, assembled as Python lists and shipped in one small copy each
        meta_rows, clamp_rows = [], []
        for s in (batch_data_samples or [None] * B):
            mi = s.metainfo if s is not None else {}
            oh, ow = mi.get("ori_shape", (H, W))[:2]
            clamp_rows.append((float(ow), float(oh)))
            px = py = 0.0
            sx = sy = 1.0
            if rescale:
                pad = mi.get("pad_param")
                if pad is not None:
                    px, py = float(pad[2]), float(pad[0])
                sf = mi.get("scale_factor", (1.0, 1.0))
                sx, sy = float(sf[0]), float(sf[1])
            meta_rows.append((px, py, sx, sy, 0.0, 0.0, 1.0, 0.0))
        key = (tuple(meta_rows), tuple(clamp_rows))
        if getattr(p, "_meta_key", None) != key:       # unchanged metadata (fixed-shape serving loops): nothing to copy
            p.set_meta(torch.tensor(meta_rows, dtype=torch.float32).to(self.device, non_blocking=True),
                       torch.tensor(clamp_rows, dtype=torch.float32).to(self.device, non_blocking=True))
            p._meta_key = key
        if batch_inputs is not None:          # (predict_images: the device pipeline has already written the plan's input)
            p.image.copy_(batch_inputs, non_blocking=True)
        _run_plan(p, self.cuda_graph)
        r = p.results() if self.zero_copy else _own(p.results())
        self.last_batch_result = r            # packed device tensors [B,max,...] + counts: bulk readers copy these once
        labels64 = r["labels"].long()         # one conversion for the batch (the reference's labels are int64)
        counts = r["counts"].cpu().tolist()   # the one host sync of the step (the reference has >= 3 per image)
        out = []
        for b in range(B):
            n = counts[b]
            s = batch_data_samples[b] if batch_data_samples else DetDataSample()
            s.pred_instances = instances_for(s, bboxes=r["boxes"][b, :n], scores=r["scores"][b, :n], labels=labels64[b, :n])
            out.append(s)
        return out

    def test_step(self, data):
        return self.predict(data["inputs"], data.get("data_samples"), rescale=True)

    def pipeline_cfg(self):
        """(scale (w, h), allow_scale_up, pad value, ...) of the config's test_pipeline (config/wedetect_base.py:111-118); the
        shipped values when the detector was built without a config."""
        out = dict(scale=(640, 640), allow_scale_up=False, pad=114)
        steps = (getattr(self, "cfg", None) or {}).get("test_pipeline") or []
        types = [t.get("type") for t in steps]
        for t in steps:
            if t.get("type") == "WeDetectLetterResize":
                pv = t.get("pad_val", dict(img=0))
                out.update(scale=tuple(t["scale"]), allow_scale_up=bool(t.get("allow_scale_up", True)), pad=int(pv.get("img", 0) if isinstance(pv, dict) else pv),
                           use_mini_pad=bool(t.get("use_mini_pad", False)), stretch_only=bool(t.get("stretch_only", False)),
                           half_pad_param=bool(t.get("half_pad_param", False)))
            elif t.get("type") == "WeDetectKeepRatioResize" and tuple(t["scale"]) != tuple(out["scale"]) and "WeDetectLetterResize" in types:
                ls = next(tuple(u["scale"]) for u in steps if u.get("type") == "WeDetectLetterResize")
                if tuple(t["scale"]) != ls:
                    raise NotImplementedError("WeDetectKeepRatioResize and WeDetectLetterResize with different scales")
        if steps:
            out["keep_ratio_first"] = "WeDetectKeepRatioResize" in types
        return out

    def predict_images(self, images, texts=None, rescale=True, decode="host"):
        """infer_wedetect.py:102-117 for a batch of images: LoadImageFromFile (host decode; arrays are taken as decoded BGR),
        WeDetectKeepRatioResize + WeDetectLetterResize on the device (cv2-exact, transforms.py:94-123,180-272) straight into the
        detector's input, then test_step.  texts: list of prompts (one shared set), or None for the reparameterized features.
        Returns one DetDataSample per image with the metainfo PackDetInputs would carry.  decode="nvjpeg": JPEG files / byte strings
        are decoded on the device (a few grey levels from cv2's libjpeg-turbo on chroma-subsampled files; opt-in)."""
        if self._sd is None:
            raise RuntimeError("load_state_dict first")
        if decode not in ("host", "nvjpeg"):
            raise ValueError(f"decode={decode!r}")
        images = list(images)
        arrays = decode_images_bgr(images) if decode == "host" else None
        pc = self.pipeline_cfg()
        (W, H), B = pc["scale"], len(images)
        if texts is not None:
            flat = [t[0] if isinstance(t, (list, tuple)) else t for t in texts]
            src = self.forward_text([flat])
        elif self.text_feats is not None:
            src = self.text_feats
        else:
            raise TypeError("texts or reparameterize() first")
        feats = src.reshape(-1, schema.EMBED_DIM)
        p = self._plan(B, H, W, feats.shape[0], torch.uint8)
        with _on(self.device):
            pipe = getattr(p, "_mm_pipe", None)
            if pipe is None:
                pipe = p._mm_pipe = MMTestPipeline(p.image, **pc)
            if arrays is None:
                arrays = pipe.encoded(images, decode_images_bgr)
            metas = pipe.run(arrays)
            samples = [DetDataSample(dict(m, img_id=i, img_path=(im if isinstance(im, str) else None))) for i, (m, im) in enumerate(zip(metas, images))]
            return self._predict_on(p, src, feats, None, samples, rescale, B, H, W)

    __call__ = test_step


def _run_plan(p, use_graph):
    """First run of a plan is eager (lazy attribute setup inside the library); from the second on the whole forward is one
    CUDA-graph launch (all buffers are static, per-batch inputs are copied into them before the launch)."""
    if use_graph and not p._graph and getattr(p, "_ran_eager", False):
        p.capture()
    p.run()
    p._ran_eager = True


class SimpleYOLOWorldDetector:
    """WeDetect-Uni proposal generator facade (generate_proposal.py:1052-1218).

    extract=True is the variant of eval_retrieval/extract_embedding.py:1088-1260: every result dict also carries
    `labels`, `scales` and `bias` (the logit_scale / bias of each kept proposal's pyramid level), and `score_text`
    computes the image x class retrieval scores of the last batch on the device (retrieval_metric.py:365-373)."""

    def __init__(self, backbone_size, prompt_dim=768, num_prompts=512, num_proposals=300, *, device="cuda:0", precise=True, extract=False,
                 cuda_graph=True, zero_copy=False):
        self.device = torch.device(device)
        L.load(require_gpu=True, device=self.device.index or 0)
        self.zero_copy = bool(zero_copy)
        self.cuda_graph = bool(cuda_graph) and not os.environ.get("WD_NO_GRAPH")
        if backbone_size not in ("base", "large", "tiny"):
            raise ValueError(backbone_size)
        assert prompt_dim == schema.EMBED_DIM
        self.size, self.num_prompts, self.num_proposals = backbone_size, num_prompts, num_proposals
        self.img_size = (1280, 1280) if backbone_size == "large" else (640, 640)
        self.precise, self.extract = precise, bool(extract)
        self._sd, self._vw, self._plans = None, {}, _LRU(4)
        self._scorers, self._cur, self._letterbox = {}, None, {}
        self.last_batch_result = None

    def eval(self):
        return self

    def cuda(self):
        return self

    def load_state_dict(self, sd, strict=False):
        sd = schema.normalize_state_dict(sd)
        need = schema.param_shapes(self.size, uni=True, num_prompts=self.num_prompts)
        missing = [k for k in need if k not in sd]
        if missing and strict:
            raise RuntimeError(f"missing keys: {missing[:4]} ...")
        if missing:
            raise RuntimeError(f"checkpoint lacks {len(missing)} tensors needed for inference, e.g. {missing[:3]}")
        self._sd = {k: v.detach().float().cpu() for k, v in sd.items()}
        self._vw, self._plans, self._scorers, self._cur, self._letterbox = {}, _LRU(4), {}, None, {}
        return "<All keys matched successfully>"

    def _plan(self, B, H, W, dtype=torch.float32):
        def make():
            fmt = "u8_rgb" if dtype == torch.uint8 else "f32_rgb"
            if fmt not in self._vw:
                self._vw[fmt] = weights.prepare_vision(self._sd, self.size, self.device, input_format=fmt, precise=self.precise)
            return plan.VisionPlan(self._vw[fmt], self.size, B, H, W, K=self.num_prompts, uni=True, input_dtype=dtype, score_thr=0.0,
                                   nms_pre=30000, iou_thr=0.7, max_per_img=self.num_proposals, nms_mode=1, extract=self.extract,
                                   device=self.device)
        key = (B, H, W, dtype)
        evicted = key not in self._plans and len(self._plans) >= self._plans.cap
        with _on(self.device):
            p = self._plans.get_or_make(key, make)
        if evicted:     # state tied to an evicted plan's buffers
            self._letterbox = {k: v for k, v in self._letterbox.items() if k in self._plans}
            self._scorers = {}
        return p

    def _run(self, p, key, B, H, W, ratios, offsets, ori_shapes, rescale):
        meta = torch.zeros(B, 8)
        meta[:, 2:4] = 1.0
        meta[:, 6] = 1.0
        clamp = torch.tensor([[float(W), float(H)]] * B)
        for b in range(B):
            if offsets is not None:
                meta[b, 4], meta[b, 5] = float(offsets[b][0]), float(offsets[b][1])
            if ratios is not None and rescale:
                meta[b, 6] = float(ratios[b])
            if ori_shapes is not None:
                clamp[b, 0], clamp[b, 1] = float(ori_shapes[b][1]), float(ori_shapes[b][0])
        with _on(self.device):
            p.set_meta(meta.to(self.device), clamp.to(self.device))
            _run_plan(p, self.cuda_graph)
        self._live = p.results()              # the plan's own buffers (score_text reads them in place)
        r = self._live if self.zero_copy else _own(self._live)
        self.last_batch_result, self._cur = r, key
        counts = r["counts"].cpu().tolist()
        out = [dict(bboxes=r["boxes"][b, :counts[b]], embeddings=r["embeddings"][b, :counts[b]], scores=r["scores"][b, :counts[b]])
               for b in range(B)]
        if self.extract:
            labels64 = r["labels"].long()
            for b, d in enumerate(out):
                d.update(labels=labels64[b, :counts[b]], scales=r["scales"][b, :counts[b]], bias=r["bias"][b, :counts[b]])
        return out

    def forward_tensor(self, inputs, ratios=None, offsets=None, ori_shapes=None, rescale=True):
        """inputs: fp32 [B,3,H,W] RGB in [0,1] (already letterboxed).  Returns the reference's list of dicts."""
        B, _, H, W = inputs.shape
        key = (B, H, W, torch.float32)
        p = self._plan(*key)
        p.image.copy_(inputs, non_blocking=True)
        return self._run(p, key, B, H, W, ratios, offsets, ori_shapes, rescale)

    def forward(self, image_paths, rescale=True, decode="host"):
        """image_paths: list of file names, PIL images or uint8 [h,w,3] RGB arrays (generate_proposal.py:1082-1117).
        Decoding stays with PIL (decode="nvjpeg": JPEG files / byte strings are decoded on the device instead, opt-in); the letterbox
        (BILINEAR resize + 114 padding) runs on the device, bit-exact with PIL."""
        image_paths = list(image_paths)
        arrays = decode_images(image_paths) if decode == "host" else None
        B, (H, W) = len(image_paths), self.img_size
        key = (B, H, W, torch.uint8)
        p = self._plan(*key)
        if key not in self._letterbox:
            self._letterbox[key] = Letterbox(p.image)
        with _on(self.device):
            if arrays is None:
                arrays = self._letterbox[key].encoded(image_paths, decode_images)
            ratios, offsets, ori_shapes = self._letterbox[key].run(arrays)
        return self._run(p, key, B, H, W, ratios, offsets, ori_shapes, rescale)

    __call__ = forward

    def score_text(self, text_embedding):
        """Retrieval scores [B, K] of the LAST batch against `text_embedding` [K, 768] (L2-normalised by the caller, as
        extract_embedding.py:1713 does): max over the image's proposals of sigmoid(emb . text * exp(scale) + bias)
        (retrieval_metric.py:365-373), computed on the device from the live result buffers."""
        from .retrieval import RetrievalScorer
        if not self.extract or self._cur is None:
            raise RuntimeError("score_text needs extract=True and a previous forward")
        key = (self._cur, id(text_embedding))
        if key not in self._scorers:
            p = self._plans[self._cur]
            r = p.results()
            with _on(self.device):
                sc = RetrievalScorer(text_embedding, p.B, p.max_per_img, device=self.device, precise=self.precise, emb=r["embeddings"],
                                     scale=r["scales"], bias=r["bias"], counts=r["counts"])
            self._scorers = {key: (sc, text_embedding)}      # one live text set at a time (keeps the id() key valid)
        with _on(self.device):
            s = self._scorers[key][0].run()
        return s if self.zero_copy else s.clone()


class XLMRobertaLanguageBackbone:
    """Standalone text tower facade (eval_retrieval/extract_embedding.py:1267-1320): `forward(List[str]) -> [n, 768]`, the
    CLS token through the 768-d head, NOT normalised (the script normalises afterwards, :1713).  `ckpt` is a checkpoint path
    (an mmengine `.pth` with 'state_dict'), or a state dict / its text slice.  The tokenizer comes from `model_name` (the
    reference reads ../xlm-roberta-{base,large}/) unless one is passed in."""

    def __init__(self, ckpt, *, model_name=None, tokenizer=None, device="cuda:0", precise=True):
        L.load(require_gpu=True, device=torch.device(device).index or 0)
        sd = torch.load(ckpt, map_location="cpu", weights_only=False) if isinstance(ckpt, (str, os.PathLike)) else ckpt
        sd = schema.text_state_dict(sd)
        self.text = schema.text_size_of(sd)                         # 'base' | 'large'
        self.size = next(k for k, v in schema.SIZES.items() if v["text"] == self.text)
        need = {k: v for k, v in schema.param_shapes(self.size, with_text=True).items() if k.startswith("backbone.text_model.")}
        missing = [k for k in need if k not in sd]
        bad = [k for k in need if k in sd and tuple(sd[k].shape) != tuple(need[k])
               and not k.endswith("word_embeddings.weight")]        # vocabulary size is the checkpoint's
        if missing or bad:
            raise RuntimeError(f"text tower checkpoint: missing {missing[:3]} ({len(missing)}), size mismatch {bad[:3]}")
        self.device, self.precise = torch.device(device), precise
        self._sd = {k: v.detach().float().cpu() for k, v in sd.items()}
        self._tw, self._plans = None, _LRU(8)
        self._tokenizer, self._model_name = tokenizer, model_name or f"./xlm-roberta-{self.text}/"
        self.language_dim = schema.TEXT[self.text]["hidden"]

    def cuda(self):
        return self

    def eval(self):
        return self

    @property
    def tokenizer(self):
        if self._tokenizer is None:
            from transformers import AutoTokenizer
            self._tokenizer = AutoTokenizer.from_pretrained(self._model_name)
        return self._tokenizer

    def encode_tokens(self, ids, mask, normalize=False):
        with _on(self.device):
            if self._tw is None:
                self._tw = weights.prepare_text(self._sd, self.size, self.device, precise=self.precise)
            key = tuple(ids.shape)
            tp = self._plans.get_or_make(key, lambda: plan.TextPlan(self._tw, self.size, key[0], key[1], device=self.device))
            feats = tp.run(ids.to(self.device, torch.int32), mask.to(self.device, torch.int32))
            return (feats if normalize else tp.head_out).clone()

    def forward(self, text):
        tok = self.tokenizer(text=list(text), return_tensors="pt", padding=True)
        return self.encode_tokens(tok["input_ids"], tok["attention_mask"])

    __call__ = forward
