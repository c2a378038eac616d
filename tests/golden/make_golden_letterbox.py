"""Golden digests for the letterbox path, produced by the UNMODIFIED reference letterbox() (generate_proposal.py:17-82,
PIL BILINEAR resize + paste) on seeded random RGB images.

    python tests/golden/make_golden_letterbox.py        (only works where /root/reference exists)

Stored per case: source size, the seed, sha256 of the 640x640x3 canvas bytes, ratio and (dw/2, dh/2).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

CASES = [(640, 480), (480, 640), (640, 640), (1280, 720), (333, 500), (500, 333), (1920, 1080), (97, 31), (31, 97), (7, 5), (641, 639),
         (1000, 1000), (2048, 1365), (320, 320), (639, 640), (1279, 1281)]


def source(w, h, seed):
    """Smooth + noisy content so that both filter taps and rounding matter."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 3) % 256], -1)
    noise = rng.integers(-40, 41, (h, w, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def main():
    import PIL
    from PIL import Image
    import generate_proposal as gp
    out = dict(pillow=PIL.__version__, cases=[])
    for i, (w, h) in enumerate(CASES):
        img = source(w, h, 100 + i)
        canvas, ratio, (dw, dh) = gp.letterbox(Image.fromarray(img), (640, 640))
        arr = np.asarray(canvas)
        out["cases"].append(dict(w=w, h=h, seed=100 + i, sha256=hashlib.sha256(arr.tobytes()).hexdigest(), ratio=ratio, dw=dw, dh=dh))
    with open(os.path.join(HERE, "letterbox_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("saved", len(out["cases"]), "cases, pillow", PIL.__version__)


if __name__ == "__main__":
    main()
