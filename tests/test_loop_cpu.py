"""Host logic of the evaluation loop (wedetect_b200/loop.py) around a stub detector: the bs-1 loader contract of the reference's
test.py (config/wedetect_base.py:197-204) is kept, the device sees grouped batches, the evaluator sees loader order."""
import torch

from wedetect_b200.loop import TestLoop, proposals_for_ref, sample_to_dict
from wedetect_b200.structures import DetDataSample, InstanceData


class StubModel:
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def test_step(self, data):
        x, samples = data["inputs"], data["data_samples"]
        self.calls.append((tuple(x.shape), [s.img_id for s in samples]))
        for i, s in enumerate(samples):      # "detections" that identify the image: its mean pixel and its id
            s.pred_instances = InstanceData(bboxes=torch.full((1, 4), float(x[i].float().mean())), scores=torch.tensor([0.5]), labels=torch.tensor([s.img_id]))
        return samples


class StubEvaluator:
    def __init__(self):
        self.seen = []

    def process(self, data_samples, data_batch):
        assert len(data_samples) == len(data_batch["data_samples"])
        self.seen += [(d["img_id"], float(d["pred_instances"]["bboxes"][0, 0]), int(d["pred_instances"]["labels"][0])) for d in data_samples]

    def evaluate(self, size):
        return dict(n=len(self.seen), size=size)


class Loader(list):
    pass


def _batch(i, hw=(8, 8), texts=None):
    meta = dict(img_id=i, ori_shape=hw, img_shape=hw, scale_factor=(1.0, 1.0))
    if texts is not None:
        meta["texts"] = texts
    return dict(inputs=[torch.full((3,) + hw, i, dtype=torch.uint8)], data_samples=[DetDataSample(meta)])


def test_groups_consecutive_same_shape_batches_and_keeps_loader_order():
    loader = Loader([_batch(i) for i in range(7)] + [_batch(7, (8, 16)), _batch(8, (8, 16))] + [_batch(9)])
    loader.dataset = list(range(10))
    model, ev = StubModel(), StubEvaluator()
    out = TestLoop(model, loader, ev, group=4).run()
    assert out == dict(n=10, size=10)
    # 7 same-shape images -> groups of 4 + 3; the shape change and the return to the first shape start new groups
    assert [c[0][0] for c in model.calls] == [4, 3, 2, 1] and model.calls[2][0][2:] == (8, 16)
    assert [s[0] for s in ev.seen] == list(range(10))                       # evaluator sees loader order
    assert all(s[1] == float(s[0]) and s[2] == s[0] for s in ev.seen)       # and each image's own result


def test_prompt_list_change_starts_a_new_group_and_dict_view():
    a, b = [["cat"], ["dog"]], [["car"], ["bus"]]
    loader = Loader([_batch(0, texts=a), _batch(1, texts=a), _batch(2, texts=b)])
    model, ev = StubModel(), StubEvaluator()
    TestLoop(model, loader, ev, group=8).run()
    assert [c[1] for c in model.calls] == [[0, 1], [2]]
    s = DetDataSample(dict(img_id=3, ori_shape=(4, 4)))
    s.pred_instances = InstanceData(bboxes=torch.zeros(2, 4), scores=torch.ones(2), labels=torch.zeros(2, dtype=torch.long))
    d = sample_to_dict(s)
    assert d["img_id"] == 3 and d["ori_shape"] == (4, 4) and set(d["pred_instances"]) == {"bboxes", "scores", "labels"}


def test_proposals_for_ref_keeps_device_and_casts():
    outs = [dict(bboxes=torch.rand(5, 4), embeddings=torch.rand(5, 768), scores=torch.rand(5)), dict(bboxes=torch.rand(0, 4), embeddings=torch.rand(0, 768), scores=torch.rand(0))]
    boxes, counts = proposals_for_ref(outs, torch.bfloat16)
    assert counts == [5, 0] and boxes[0].dtype == torch.bfloat16 and boxes[0].device == outs[0]["bboxes"].device
    assert torch.equal(boxes[0], outs[0]["bboxes"].to(torch.bfloat16))
