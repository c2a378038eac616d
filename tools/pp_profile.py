#!/usr/bin/env python
"""Per-phase post-process timing (WD_PP_PROFILE=1) for the text path (C2, sparse) and the Uni path (score_thr 0: every
(anchor, prompt) pair is a candidate)."""
import os
import sys

os.environ["WD_PP_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from oracle import synth
    from wedetect_b200 import plan as P, schema, weights
    dev = "cuda:0"
    for uni in (False, True):
        sd = synth.synth_state_dict("base", seed=0, uni=uni, with_text=False, regime="sparse")
        Wt = weights.prepare_vision(sd, "base", dev)
        K = 256 if uni else 80
        kw = dict(score_thr=0.0, nms_mode=1) if uni else dict(score_thr=0.001, nms_mode=0)
        p = P.VisionPlan(Wt, "base", 32, 640, 640, K=K, uni=uni, **kw)
        if not uni:
            p.set_text(torch.randn(K, schema.EMBED_DIM, generator=torch.Generator().manual_seed(5)).to(dev))
        p.image.copy_(synth.synth_images(32, 640, 640).to(dev))
        for i in range(2):
            sys.stderr.write(f"--- {'uni' if uni else 'text'} run {i}\n")
            p.run()
            torch.cuda.synchronize()
        del p, Wt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
