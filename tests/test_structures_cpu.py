"""The reference's inference tail (infer_wedetect.py:113-128) runs unchanged on our InstanceData / DetDataSample stand-ins."""
import numpy as np
import pytest
import torch


def test_infer_tail_on_instance_data():
    from wedetect_b200.structures import DetDataSample, InstanceData
    g = torch.Generator().manual_seed(0)
    n = 300
    scores = torch.rand(n, generator=g).sort(descending=True).values
    inst = InstanceData(bboxes=torch.rand(n, 4, generator=g) * 640, scores=scores, labels=torch.randint(0, 80, (n,), generator=g))
    sample = DetDataSample(dict(img_id=0, ori_shape=(480, 640), texts=[["a"], ["b"]]))
    sample.pred_instances = inst
    assert sample.texts == [["a"], ["b"]] and sample.get("missing") is None and sample.metainfo["ori_shape"] == (480, 640)
    # --- the reference's lines, verbatim in structure ---
    score_thr, max_dets = 0.3, 100
    output = sample
    pred_instances = output.pred_instances
    pred_instances = pred_instances[pred_instances.scores.float() > score_thr]
    if len(pred_instances.scores) > max_dets:
        indices = pred_instances.scores.float().topk(max_dets)[1]
        pred_instances = pred_instances[indices]
    pred_instances = pred_instances.cpu().numpy()
    xyxy, class_id, confidence = pred_instances["bboxes"], pred_instances["labels"], pred_instances["scores"]
    # ---
    keep = scores > score_thr
    want = scores[keep][:max_dets]
    assert isinstance(xyxy, np.ndarray) and xyxy.shape == (len(want), 4) and class_id.dtype == np.int64
    assert np.array_equal(confidence, want.numpy()) and len(pred_instances) == len(want)
    assert "bboxes" in pred_instances and sorted(pred_instances.keys()) == ["bboxes", "labels", "scores"]


def test_instances_for_picks_the_sample_flavour(monkeypatch):
    """Real mmdet samples type-check `pred_instances`: they get mmengine's InstanceData, ours get ours."""
    import sys
    import types
    from wedetect_b200.structures import DetDataSample, InstanceData, instances_for
    ours = instances_for(DetDataSample(), scores=torch.zeros(2))
    assert isinstance(ours, InstanceData) and isinstance(instances_for(None, scores=torch.zeros(1)), InstanceData)
    # a stand-in for mmengine (not installable offline): the function must import and use ITS class for mmdet-typed samples
    fake = types.ModuleType("mmengine.structures")

    class MMInstanceData(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
    fake.InstanceData = MMInstanceData
    monkeypatch.setitem(sys.modules, "mmengine", types.ModuleType("mmengine"))
    monkeypatch.setitem(sys.modules, "mmengine.structures", fake)
    MMSample = type("DetDataSample", (), {})
    MMSample.__module__ = "mmdet.structures.det_data_sample"
    got = instances_for(MMSample(), scores=torch.ones(3))
    assert isinstance(got, MMInstanceData) and got["scores"].shape == (3,)


def test_resolve_texts_branches():
    """The text-source branches of YOLOWorldDetector.extract_feat (yolo_world.py:84-100) as a pure function."""
    from wedetect_b200.detector import resolve_texts
    from wedetect_b200.structures import DetDataSample
    infer_style = [["person"], ["dog"], [" "]]                      # infer_wedetect.py:163-167
    assert resolve_texts(None) is None and resolve_texts([]) is None and resolve_texts({}) is None
    assert resolve_texts([DetDataSample(dict(ori_shape=(4, 4)))]) is None                      # no texts: cached features
    assert resolve_texts([DetDataSample(dict(texts=infer_style))] * 2) == ["person", "dog", " "]
    assert resolve_texts(dict(texts=[["a", "b"], ["a", "b"]])) == ["a", "b"]                   # dict form, one list per image
    assert resolve_texts(dict(texts=["a", "b"])) == ["a", "b"]                                 # dict form, one shared list
    # per-image lists that differ (equal counts): one list per image comes back, predict() then runs one group per distinct list
    assert resolve_texts([DetDataSample(dict(texts=["a"])), DetDataSample(dict(texts=["b"]))]) == [["a"], ["b"]]
    with pytest.raises(AssertionError):                                                         # mm_backbone.py:378-380
        resolve_texts([DetDataSample(dict(texts=["a"])), DetDataSample(dict(texts=["b", "c"]))])


def test_predict_host_logic_with_a_stub_plan(monkeypatch):
    """YOLOWorldDetector.predict around a stubbed plan: text source, rescale metadata (pad / scale_factor before NMS, clamp to
    ori_shape: yolo_world_head.py:728-746), unchanged-metadata caching, per-image slicing by counts, int64 labels."""
    import wedetect_b200._lib as L
    from wedetect_b200 import detector as det
    from wedetect_b200.structures import DetDataSample
    monkeypatch.setattr(L, "load", lambda require_gpu=True, device=0: None)
    B, H, W = 2, 64, 96

    class StubPlan:
        def __init__(self):
            self.image = torch.zeros(B, 3, H, W, dtype=torch.uint8)
            self.meta_calls, self.text_calls, self._graph, self._text_key = [], 0, False, None

        def set_text(self, f):
            self.text_calls += 1
            self.K = f.shape[0]

        def set_meta(self, m, c):
            self.meta_calls.append((m.clone(), c.clone()))

        def run(self):
            pass

        def results(self):
            return dict(boxes=torch.arange(B * 5 * 4, dtype=torch.float32).reshape(B, 5, 4), scores=torch.rand(B, 5), labels=torch.ones(B, 5, dtype=torch.int32),
                        anchors=torch.zeros(B, 5, dtype=torch.int32), counts=torch.tensor([3, 0], dtype=torch.int32))

    m = det.YOLOWorldDetector(size="tiny", device="cpu", cuda_graph=False)
    m._sd = {}
    stub = StubPlan()
    monkeypatch.setattr(m, "_plan", lambda *a: stub)
    imgs = torch.zeros(B, 3, H, W, dtype=torch.uint8)
    samples = [DetDataSample(dict(ori_shape=(50, 80), scale_factor=(1.2, 1.2), pad_param=(7.0, 7.0, 8.0, 8.0))), DetDataSample(dict(ori_shape=(64, 96)))]
    with pytest.raises(TypeError):
        m.predict(imgs, samples)                                             # neither texts nor reparameterized features
    m.set_text_features(torch.randn(4, 768))
    out = m.test_step(dict(inputs=imgs, data_samples=samples))
    assert stub.text_calls == 1 and stub.K == 4
    meta, clamp = stub.meta_calls[-1]
    assert meta[0].tolist() == pytest.approx([8.0, 7.0, 1.2, 1.2, 0.0, 0.0, 1.0, 0.0]) and meta[1].tolist() == [0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 1.0, 0.0]
    assert clamp.tolist() == [[80.0, 50.0], [96.0, 64.0]]
    assert [len(s.pred_instances) for s in out] == [3, 0] and out[0].pred_instances.labels.dtype == torch.int64
    assert out[0] is samples[0] and out[0].pred_instances.bboxes.shape == (3, 4)
    m.predict(imgs, samples)                                                  # same text set, same metadata: nothing re-uploaded
    assert stub.text_calls == 1 and len(stub.meta_calls) == 1
    m.predict(imgs, samples, rescale=False)                                   # rescale off: identity pre-NMS mapping, clamp kept
    assert stub.meta_calls[-1][0][0].tolist() == [0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 1.0, 0.0] and len(stub.meta_calls) == 2
    out = m.predict(imgs, None)                                               # no samples at all: fresh DetDataSample per image
    assert len(out) == B and len(out[0].pred_instances) == 3


def test_uni_forward_tensor_host_logic_with_a_stub_plan(monkeypatch):
    """SimpleYOLOWorldDetector glue (generate_proposal.py:1103-1117): boxes -= (dw/2, dh/2), /= ratio when rescale, clamp to the
    original size - expressed as the post-NMS metadata row; extract variant adds labels / scales / bias; result slicing."""
    import wedetect_b200._lib as L
    from wedetect_b200 import detector as det
    monkeypatch.setattr(L, "load", lambda require_gpu=True, device=0: None)
    B, H, W, P = 2, 64, 64, 6

    class StubPlan:
        def __init__(self):
            self.image = torch.zeros(B, 3, H, W)
            self.meta, self._graph = None, False

        def set_meta(self, m, c):
            self.meta = (m.clone(), c.clone())

        def run(self):
            pass

        def results(self):
            return dict(boxes=torch.rand(B, P, 4), scores=torch.rand(B, P), labels=torch.zeros(B, P, dtype=torch.int32), anchors=torch.zeros(B, P, dtype=torch.int32),
                        counts=torch.tensor([P, 2], dtype=torch.int32), embeddings=torch.rand(B, P, 768), scales=torch.full((B, P), -1.0), bias=torch.zeros(B, P))

    m = det.SimpleYOLOWorldDetector("base", 768, 256, P, device="cpu", extract=True, cuda_graph=False)
    stub = StubPlan()
    monkeypatch.setattr(m, "_plan", lambda *a: stub)
    x = torch.rand(B, 3, H, W)
    out = m.forward_tensor(x, ratios=[0.5, 2.0], offsets=[(0.0, 8.0), (4.5, 0.0)], ori_shapes=[(96, 128), (32, 27)])
    meta, clamp = stub.meta
    assert meta[0].tolist() == [0.0, 0.0, 1.0, 1.0, 0.0, 8.0, 0.5, 0.0] and meta[1].tolist() == [0.0, 0.0, 1.0, 1.0, 4.5, 0.0, 2.0, 0.0]
    assert clamp.tolist() == [[128.0, 96.0], [27.0, 32.0]]
    assert [len(o["scores"]) for o in out] == [P, 2] and set(out[0]) == {"bboxes", "embeddings", "scores", "labels", "scales", "bias"}
    assert out[1]["embeddings"].shape == (2, 768) and out[1]["labels"].dtype == torch.int64
    m.forward_tensor(x, ratios=[0.5, 2.0], offsets=[(0.0, 8.0), (4.5, 0.0)], ori_shapes=[(96, 128), (32, 27)], rescale=False)
    assert stub.meta[0][:, 6].tolist() == [1.0, 1.0] and stub.meta[0][0, 5] == 8.0       # offsets still removed, ratio not applied
    with pytest.raises(RuntimeError):
        det.SimpleYOLOWorldDetector("base", 768, 256, P, device="cpu").score_text(torch.zeros(3, 768))


def test_pipeline_cfg_and_predict_images_host_logic(monkeypatch):
    """predict_images around stubs: the transform parameters come from cfg.test_pipeline (config/wedetect_base.py:111-118), the
    device pipeline's metainfo lands in the samples (what PackDetInputs would have packed), prompts resolve like infer_wedetect.py,
    and api.inference_detector applies the script's tail (score threshold, top-k, numpy)."""
    import numpy as np
    import wedetect_b200._lib as L
    from wedetect_b200 import api, detector as det
    monkeypatch.setattr(L, "load", lambda require_gpu=True, device=0: None)
    m = det.YOLOWorldDetector(size="tiny", device="cpu", cuda_graph=False)
    assert m.pipeline_cfg() == dict(scale=(640, 640), allow_scale_up=False, pad=114)           # no config: the shipped values
    m.cfg = dict(test_pipeline=[dict(type="LoadImageFromFile"), dict(type="WeDetectKeepRatioResize", scale=(96, 64)),
                                dict(type="WeDetectLetterResize", scale=(96, 64), allow_scale_up=False, pad_val=dict(img=114)), dict(type="PackDetInputs")])
    pc = m.pipeline_cfg()
    assert pc["scale"] == (96, 64) and pc["pad"] == 114 and pc["allow_scale_up"] is False and pc["keep_ratio_first"] is True
    m.cfg["test_pipeline"][1]["scale"] = (64, 64)
    with pytest.raises(NotImplementedError):
        m.pipeline_cfg()
    m.cfg["test_pipeline"][1]["scale"] = (96, 64)

    B, H, W = 2, 64, 96

    class StubPlan:
        def __init__(self):
            self.image = torch.zeros(B, 3, H, W, dtype=torch.uint8)
            self._graph, self._text_key, self.K = False, None, None

        def set_text(self, f):
            self.K = f.shape[0]

        def set_meta(self, mm, c):
            self.meta = (mm.clone(), c.clone())

        def run(self):
            pass

        def results(self):
            return dict(boxes=torch.rand(B, 5, 4), scores=torch.tensor([[0.9, 0.8, 0.5, 0.2, 0.1]] * B), labels=torch.ones(B, 5, dtype=torch.int32),
                        anchors=torch.zeros(B, 5, dtype=torch.int32), counts=torch.tensor([5, 2], dtype=torch.int32))

    class StubPipe:
        def __init__(self, out, **pipe):
            self.pipe = pipe

        def run(self, arrays):
            self.seen = [a.shape for a in arrays]
            return [dict(ori_shape=a.shape[:2], img_shape=(H, W, 3), scale_factor=(0.5, 0.5), pad_param=np.array([1, 2, 0, 0], np.float32)) for a in arrays]

    stub = StubPlan()
    m._sd = {}
    monkeypatch.setattr(m, "_plan", lambda *a: stub)
    monkeypatch.setattr(det, "MMTestPipeline", StubPipe)
    imgs = [np.zeros((100, 192, 3), np.uint8), np.zeros((128, 60, 3), np.uint8)]
    with pytest.raises(TypeError):
        m.predict_images(imgs)                                           # neither texts nor reparameterized features
    m.set_text_features(torch.randn(3, 768))
    out = m.predict_images(imgs)
    assert stub._mm_pipe.pipe["scale"] == (96, 64) and stub._mm_pipe.seen == [(100, 192, 3), (128, 60, 3)] and stub.K == 3
    assert [len(o.pred_instances) for o in out] == [5, 2] and out[1].metainfo["ori_shape"] == (128, 60) and out[0].metainfo["img_id"] == 0
    meta, clamp = stub.meta
    assert meta[0].tolist() == pytest.approx([0.0, 1.0, 0.5, 0.5, 0.0, 0.0, 1.0, 0.0]) and clamp.tolist() == [[192.0, 100.0], [60.0, 128.0]]
    with pytest.raises(ValueError):
        m.predict_images(imgs, decode="gpu")
    one = api.inference_detector(m, imgs[0], None, max_dets=2, score_thr=0.3)
    assert one["confidence"].tolist() == pytest.approx([0.9, 0.8]) and one["xyxy"].shape == (2, 4) and one["class_id"].dtype == np.int64
    many = api.inference_detector(m, imgs, None, score_thr=0.85)
    assert [len(d["confidence"]) for d in many] == [1, 1]
