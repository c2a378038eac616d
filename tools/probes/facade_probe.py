"""Debug probe: grouped (B=4) vs alone (B=1) detections through the facade, repeated, to find run-to-run differences."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch
from util import FixtureTokenizer
from oracle import synth
from wedetect_b200.api import DetDataSample, init_detector
D = "cuda:0"
CFG = os.path.join(ROOT, "tests", "configs", "wedetect_base_min.py")
tok = FixtureTokenizer("coco_zh")
sd = synth.synth_state_dict("base", seed=0, with_text=True, regime="sparse")
model = init_detector(CFG, checkpoint=dict(state_dict=sd), device=D)
model._tokenizer = tok
texts = [[t] for t in tok.texts[:12]]
g = torch.Generator().manual_seed(21)
imgs = (torch.rand(5, 3, 320, 320, generator=g) * 255).to(torch.uint8)
mk = lambda i: DetDataSample(dict(img_id=i, ori_shape=(320, 320), img_shape=(320, 320), scale_factor=(1.0, 1.0), pad_param=(0.0, 0.0, 0.0, 0.0), texts=texts))
def run(idx):
    out = model.test_step(dict(inputs=imgs[idx].to(D), data_samples=[mk(i) for i in idx]))
    return [(o.pred_instances.scores.cpu().clone(), o.pred_instances.bboxes.cpu().clone(), o.pred_instances.labels.cpu().clone()) for o in out]
grp = run([0, 1, 2, 3])
for rep in range(3):
    for i in range(4):
        a = run([i])[0]
        n = min(len(a[0]), len(grp[i][0]))
        ds = (a[0][:n] - grp[i][0][:n]).abs()
        print("rep", rep, "img", i, "n", len(a[0]), len(grp[i][0]), "max score diff", float(ds.max()), "first idx", int(ds.argmax()), "labels eq", bool((a[2][:n] == grp[i][2][:n]).all()), flush=True)
grp2 = run([0, 1, 2, 3])
print("grouped rerun equal:", all(torch.equal(x[0], y[0]) for x, y in zip(grp, grp2)))
p1 = [k for k in model._plans.keys()]
print("plans", p1)
