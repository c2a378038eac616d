"""Side measurement for SURVEY §8f-2: files -> detections through the image entry point, WeDetect-Base bs32, 1280x720 JPEGs.

    cpu_pipeline   the reference's way: cv2.imdecode + WeDetectKeepRatioResize / LetterResize with cv2 on the host (one image at a
                   time, as infer_wedetect.py:111-114 does; also on the thread pool), stack, test_step
    device_resize  model.predict_images(blobs decoded on the host thread pool, resize + pad on the device)
    device_decode  model.predict_images(blobs, decode='nvjpeg'): decode, resize and pad on the device
Usage (GPU box): python tools/bench_images.py [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch

from oracle import synth
from wedetect_b200.api import DetDataSample, init_detector
from wedetect_b200.preprocess import _pool, mm_test_geometry

D = "cuda:0"
CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "configs", "wedetect_base_min.py")


def photo(seed, h, w):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([128 + 100 * np.sin(xx / (17 + 5 * c) + c) * np.cos(yy / (23 + 3 * c)) for c in range(3)], -1) + rng.normal(0, 6, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def cpu_one(blob):
    img = cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_COLOR)
    h, w = img.shape[:2]
    g = mm_test_geometry(h, w)
    if g["interp"]:
        img = cv2.resize(img, (g["resize"][1], g["resize"][0]), interpolation=cv2.INTER_AREA if g["interp"] == "area" else cv2.INTER_LINEAR)
    t, b, l, r = g["pads"]
    img = cv2.copyMakeBorder(img, t, b, l, r, cv2.BORDER_CONSTANT, value=(114, 114, 114))
    return torch.from_numpy(img).permute(2, 0, 1).contiguous(), DetDataSample(dict(ori_shape=(h, w), img_shape=g["img_shape"], scale_factor=g["scale_factor"], pad_param=g["pad_param"]))


def main():
    B, steps = 32, 5
    model = init_detector(CFG, checkpoint=dict(state_dict=synth.synth_state_dict("base", seed=0, with_text=False, regime="sparse")), device=D)
    model.set_text_features(torch.nn.functional.normalize(torch.randn(1, 80, 768, generator=torch.Generator().manual_seed(3)), dim=-1))
    blobs = [cv2.imencode(".jpg", photo(i, 720, 1280), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for i in range(B)]
    cv2.setNumThreads(1)

    def run_cpu(pool):
        items = list(_pool().map(cpu_one, blobs)) if pool else [cpu_one(b) for b in blobs]
        x = torch.stack([it[0] for it in items]).pin_memory()
        return model.test_step(dict(inputs=x.to(D, non_blocking=True), data_samples=[it[1] for it in items]))

    arms = dict(cpu_pipeline_serial=lambda: run_cpu(False), cpu_pipeline_pool8=lambda: run_cpu(True),
                device_resize=lambda: model.predict_images(blobs_as_arrays()), device_decode=lambda: model.predict_images(blobs, decode="nvjpeg"))

    def blobs_as_arrays():
        return list(_pool().map(lambda b: cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR), blobs))

    res = dict(workload=f"WeDetect-Base bs{B}, {B} JPEGs of 1280x720 (q90, 4:2:0, {sum(map(len, blobs)) / B / 1e3:.0f} KB each) -> detections", steps=steps, host_cores=len(os.sched_getaffinity(0)))
    for name, fn in arms.items():
        for _ in range(2):
            out = fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
            n = sum(len(o.pred_instances) for o in out)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        res[name] = dict(images_per_s=B / dt, ms_per_batch=1000 * dt, detections=n)
        print(name, res[name], flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
