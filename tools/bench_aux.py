#!/usr/bin/env python
"""Side measurements for the rows next to the hot path (not the BASELINE bench line): device letterbox, retrieval scorer,
Uni extract step (config 4 / 5 shapes).  CUDA-event timing on the current stream, JSON on stdout."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from oracle import synth
    from wedetect_b200.detector import SimpleYOLOWorldDetector
    from wedetect_b200.preprocess import Letterbox
    from wedetect_b200.retrieval import RetrievalScorer
    dev = "cuda:0"
    out = {}
    rng = np.random.default_rng(0)
    # ---- letterbox: 32 decoded 1280x960 RGB images -> [32,3,640,640] u8 (H2D of the sources inside the timed region) ----
    B = 32
    canvas = torch.zeros(B, 3, 640, 640, dtype=torch.uint8, device=dev)
    lb = Letterbox(canvas)
    imgs = [rng.integers(0, 256, (960, 1280, 3), dtype=np.uint8) for _ in range(B)]
    t0 = time.perf_counter()
    ms = timed(lambda: lb.run(imgs), warm=2, reps=5)
    src_bytes = sum(im.size for im in imgs)
    ms_k = timed(lambda: lb._program.run(torch.cuda.current_stream().cuda_stream), warm=2, reps=20)
    moved = src_bytes + 2 * B * 960 * 640 * 3 + B * 3 * 640 * 640     # source read + 8-bit intermediate write & read + canvas write
    out["letterbox_32x1280x960"] = dict(ms_with_host_pack_and_h2d=ms, ms_kernels_only=ms_k, src_bytes=src_bytes, bytes_moved=moved,
                                         kernels_gbs=moved / (ms_k / 1000) / 1e9)
    # host-side breakdown of one lb.run(): table / descriptor packing, pageable -> pinned copies
    from wedetect_b200.preprocess import pack_batch, _pool
    t0 = time.perf_counter()
    for _ in range(3):
        pk = pack_batch(imgs, 640, 640, with_src=False)
    t_pack = (time.perf_counter() - t0) / 3 * 1000
    src_np = lb._host["src"].numpy()
    t0 = time.perf_counter()
    for _ in range(3):
        list(_pool().map(lambda part: np.copyto(src_np[part[0]: part[0] + part[1].size], part[1]), pk["src_parts"]))
    t_copy = (time.perf_counter() - t0) / 3 * 1000
    out["letterbox_32x1280x960"].update(host_pack_ms=t_pack, host_copy_to_pinned_ms=t_copy)
    # end to end through the facade: forward(list of arrays) = pack + H2D + letterbox + detector + D2H of counts
    sd0 = synth.synth_state_dict("base", seed=0, uni=True, regime="sparse")
    m0 = SimpleYOLOWorldDetector("base", 768, 256, 300, device=dev)
    m0.load_state_dict(sd0)
    for _ in range(3):
        m0.forward(imgs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        m0.forward(imgs)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1000
    out["uni_forward_32x1280x960_arrays_e2e"] = dict(ms=ms, images_per_s=32 / (ms / 1000), note="host arrays -> proposals (CUDA graph for the detector)")
    del m0
    # PIL on one host core for the same images (the reference's path), bounded sample
    from PIL import Image
    sys.path.insert(0, os.path.join(ROOT))
    from wedetect_b200.preprocess import letterbox_params
    t0 = time.perf_counter()
    for im in imgs[:8]:
        r, unpad, (l, t), _ = letterbox_params(1280, 960, (640, 640))
        c = Image.new("RGB", (640, 640), (114, 114, 114))
        c.paste(Image.fromarray(im).resize(unpad, Image.Resampling.BILINEAR), (l, t))
    out["letterbox_pil_ms_per_image_1core"] = (time.perf_counter() - t0) / 8 * 1000
    # ---- retrieval scorer: 32 images x 300 proposals x 1203 classes ----
    for precise in (False, True):
        text = torch.nn.functional.normalize(torch.randn(1203, 768), dim=-1)
        sc = RetrievalScorer(text, 32, 300, device=dev, precise=precise)
        embs = [torch.randn(300, 768) for _ in range(32)]
        sc.load(embs, [torch.zeros(300) - 1] * 32, [torch.zeros(300) - 3] * 32)
        ms = timed(sc.run)
        out[f"retrieval_scorer_32x300x1203_{'precise' if precise else 'fast'}"] = dict(ms=ms, gflop=2 * 32 * 300 * 1203 * 768 / 1e9)
    # ---- Uni extract step (config 5 shape: bs32, 300 proposals, 256 prompts) + scores against 1203 classes ----
    sd = synth.synth_state_dict("base", seed=0, uni=True, regime="sparse")
    m = SimpleYOLOWorldDetector("base", 768, 256, 300, device=dev, extract=True)
    m.load_state_dict(sd)
    x = synth.synth_images(32, 640, 640).to(dev)
    text = torch.nn.functional.normalize(torch.randn(1203, 768), dim=-1)

    def step():
        m.forward_tensor(x)
        m.score_text(text)
    ms = timed(step, warm=3, reps=5)
    out["uni_extract_bs32_640_plus_scores_1203"] = dict(ms=ms, images_per_s=32 / (ms / 1000), note="CUDA graph + scorer, fp32 inputs resident (157 MB D2D per step)")
    # ---- Uni proposal mode (config 4 shape per GPU: bs8, 1000 proposals) ----
    m4 = SimpleYOLOWorldDetector("base", 768, 256, 1000, device=dev)
    m4.load_state_dict(sd)
    x8 = x[:8].contiguous()
    ms = timed(lambda: m4.forward_tensor(x8), warm=3, reps=5)
    out["uni_proposals_bs8_640_P1000"] = dict(ms=ms, images_per_s=8 / (ms / 1000), note="CUDA graph")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
