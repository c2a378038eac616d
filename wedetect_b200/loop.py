"""Evaluation loop for the reference's `test.py` / `dist_test.sh` path (mmengine `TestLoop`, test.py:132-146) on top of the
batched device detector.

The reference evaluates with `test_dataloader.batch_size = 1` (config/wedetect_base.py:197-204): one image per `test_step`,
the text tower re-run for every image.  `TestLoop` keeps that data-loader contract and the evaluator protocol

    for data_batch in dataloader:  outputs = model.test_step(data_batch);  evaluator.process(data_samples=outputs, data_batch=data_batch)
    metrics = evaluator.evaluate(len(dataloader.dataset))

but runs the device in groups: consecutive loader batches whose images have the same [3, H, W] (and the same prompt list) are
concatenated into ONE `test_step` of up to `group` images, and the per-image results are handed to the evaluator in loader
order, one `process` call per original batch.  Class embeddings are cached per prompt list by the detector, so the text tower
runs once per distinct list instead of once per image.  Works with mmengine's Evaluator / CocoMetric / LVISMetric objects where
mmdet is installed (they only need `.process(data_samples, data_batch)` and `.evaluate(size)`), and with any stand-in offering the
same two methods.
"""
import torch

from .structures import DetDataSample


def _inputs_list(inputs):
    return list(inputs) if isinstance(inputs, (list, tuple)) else [inputs[i] for i in range(inputs.shape[0])]


def _texts_key(sample):
    get = getattr(sample, "get", None)
    t = get("texts") if callable(get) else getattr(sample, "texts", None)
    if t is None:
        return None
    return tuple(x[0] if isinstance(x, (list, tuple)) else x for x in t)


class TestLoop:
    __test__ = False      # (not a pytest class)

    def __init__(self, model, dataloader, evaluator, *, group=32, to_dict=True):
        """to_dict: hand the evaluator plain dicts (what mmengine's Evaluator does through BaseDataElement.to_dict()) built from
        our stand-in samples; real mmdet samples are passed through untouched."""
        self.model, self.dataloader, self.evaluator, self.group, self.to_dict = model, dataloader, evaluator, int(group), to_dict

    def _flush(self, pending):
        if not pending:
            return
        imgs = [im for batch in pending for im in _inputs_list(batch["inputs"])]
        samples = [s for batch in pending for s in (batch.get("data_samples") or [DetDataSample() for _ in _inputs_list(batch["inputs"])])]
        dev = self.model.device
        out = self.model.test_step(dict(inputs=torch.stack([im.to(dev, non_blocking=True) for im in imgs]), data_samples=samples))
        k = 0
        for batch in pending:                     # one evaluator.process call per ORIGINAL loader batch, in loader order
            n = len(_inputs_list(batch["inputs"]))
            res = out[k:k + n]
            k += n
            if self.to_dict:
                res = [sample_to_dict(s) for s in res]
            self.evaluator.process(data_samples=res, data_batch=batch)

    def run(self):
        pending, key, count = [], None, 0
        for data_batch in self.dataloader:
            imgs = _inputs_list(data_batch["inputs"])
            samples = data_batch.get("data_samples") or [None] * len(imgs)
            k = (tuple(imgs[0].shape), imgs[0].dtype, _texts_key(samples[0]) if samples[0] is not None else None)
            same = all(tuple(im.shape) == k[0] for im in imgs) and all((_texts_key(s) if s is not None else None) == k[2] for s in samples)
            if pending and (k != key or not same or count + len(imgs) > self.group):
                self._flush(pending)
                pending, count = [], 0
            if not same:                           # a loader batch that is itself ragged goes through alone, image by image
                for im, s in zip(imgs, samples):
                    self._flush([dict(inputs=[im], data_samples=[s] if s is not None else None)])
                continue
            pending.append(data_batch)
            key, count = k, count + len(imgs)
        self._flush(pending)
        size = len(self.dataloader.dataset) if hasattr(self.dataloader, "dataset") else None
        return self.evaluator.evaluate(size)


def sample_to_dict(sample):
    """What CocoMetric.process / LVISMetric.process read from a data sample: the metainfo keys plus
    pred_instances = dict(bboxes, scores, labels) (tensors)."""
    if hasattr(sample, "to_dict") and not isinstance(sample, DetDataSample):
        return sample.to_dict()
    d = dict(sample.metainfo)
    p = sample.pred_instances
    d["pred_instances"] = {k: p[k] for k in p.keys()}
    return d


def proposals_for_ref(outputs, dtype=torch.bfloat16):
    """Hand the proposals of `SimpleYOLOWorldDetector(...)(images)` to WeDetect-Ref without leaving the device.

    infer_wedetect_ref.py:67-74,91 moves them GPU -> numpy -> python lists -> `torch.tensor(...).cuda().to(model.dtype)`; the result
    of that round trip is simply each image's `[n, 4]` xyxy boxes (original-image pixels) in the grounding model's dtype on the
    GPU, which is what this returns (one tensor per image, as `proposals=[...]` of the Qwen3-VL grounding model expects).  The
    `"<object>" * n` placeholder string of the chat template needs the counts, returned alongside (one tiny D2H)."""
    boxes = [o["bboxes"].to(dtype) for o in outputs]
    return boxes, [int(b.shape[0]) for b in boxes]
