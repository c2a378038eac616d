"""The reference's inference tail (infer_wedetect.py:113-128) runs unchanged on our InstanceData / DetDataSample stand-ins."""
import numpy as np
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
