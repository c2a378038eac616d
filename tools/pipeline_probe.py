#!/usr/bin/env python
"""Experiment: two bs32 batches in flight on two streams (two plans, shared weights) vs one stream.  Measures whether the
tails of the ~300 persistent kernels and the latency-bound post-process of one batch fill with work of the other."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import synth
    from wedetect_b200 import plan as P, schema, weights
    dev = "cuda:0"
    B, H, W, K = 32, 640, 640, 80
    sd = synth.synth_state_dict("base", seed=0, with_text=False, regime="sparse")
    Wt = weights.prepare_vision(sd, "base", dev, input_format="u8_bgr")
    text = torch.randn(K, schema.EMBED_DIM, generator=torch.Generator().manual_seed(5)).to(dev)
    imgs = (synth.synth_images(B, H, W, seed=2) * 255).to(torch.uint8).flip(1).contiguous().to(dev)
    plans = []
    for i in range(2):
        p = P.VisionPlan(Wt, "base", B, H, W, K=K, input_dtype=torch.uint8, score_thr=0.001, nms_mode=0)
        p.set_text(text)
        p.image.copy_(imgs)
        p.run()
        torch.cuda.synchronize()
        p.capture()
        plans.append(p)
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(40):
        plans[0].run()
    torch.cuda.synchronize()
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 20
    for rep in range(2):
        e0.record()
        for _ in range(N):
            plans[0].run()
        e1.record()
        torch.cuda.synchronize()
        out[f"one_stream_ms_per_step_{rep}"] = e0.elapsed_time(e1) / N
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        en = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(2):
            st[i].record(s[i])
        for k in range(N):
            i = k & 1
            plans[i].run(s[i].cuda_stream)
        for i in range(2):
            en[i].record(s[i])
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1000
        out[f"two_streams_ms_per_step_{rep}"] = wall / N
        out[f"two_streams_stream_ms_{rep}"] = [st[i].elapsed_time(en[i]) for i in range(2)]
    r0 = {k: v.clone() for k, v in plans[0].results().items()}
    r1 = plans[1].results()
    out["results_identical"] = all(torch.equal(r0[k], r1[k]) for k in r0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
