"""Runs the XLM-R text tower plan on the reference tokenizer's LVIS ids (1204 prompts x 9 tokens, XLM-R large) a few times: the
target of the ncu capture in tools/gpu_ncu_text_post.sh.  Usage: python tools/text_tower_run.py [coco_zh|lvis_v1_zh] [base|large]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import FixtureTokenizer
from oracle import synth
from wedetect_b200 import plan, weights

name = sys.argv[1] if len(sys.argv) > 1 else "lvis_v1_zh"
size = sys.argv[2] if len(sys.argv) > 2 else "large"
tok = FixtureTokenizer(name)
ids, mask = tok.ids[:1204], tok.mask[:1204]
sd = synth.synth_state_dict(size, seed=2, with_text=True, text_vocab=int(ids.max()) + 1, calibrate=False)
Wt = weights.prepare_text(sd, size, "cuda:0")
tp = plan.TextPlan(Wt, size, ids.shape[0], ids.shape[1])
for _ in range(3):
    out = tp.run(ids.cuda(), mask.cuda())
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    out = tp.run(ids.cuda(), mask.cuda())
torch.cuda.synchronize()
print(f"text tower {size} {tuple(ids.shape)}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per run, out {tuple(out.shape)}")
