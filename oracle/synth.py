"""Seeded synthetic checkpoints (there are no real weights offline).  TEST / BENCH INFRASTRUCTURE.

Default nn.Module init makes the network degenerate (LayerScale 1e-6, BN stats (0,1), logit_scale -1,
bias 0 -> every score ~0.51), so every tensor is drawn from a recipe that keeps
activations O(1) and scores spread out:
  conv / linear weights  N(0, gain / sqrt(fan_in))      biases N(0, 0.1)
  LN / BN weight U(0.5, 1.5), bias N(0, 0.1), running_mean N(0, 0.1), running_var U(0.5, 1.5)
  LayerScale gamma U(0.05, 0.5)       BottleRep alpha U(0.5, 1.5)       logit_scale U(-1.5, -0.5)
  contrastive bias: `cls_bias` (0.0 = dense regime: every candidate passes; about -6.5 = sparse, trained-like)
  Uni prompts: normalize(randn)       XLM-R: N(0, 0.02) embeddings / N(0, gain/sqrt(fan_in)) linears
Then (calibrate=True) one seeded 320x320 image is pushed through the CPU oracle and every BatchNorm's
running statistics are set to the statistics it actually sees (times a seeded jitter), as training would
have done; that keeps all activations O(1) through the 60+ conv layers of neck and head.  In the
"sparse" regime the per-level contrastive bias is then placed so that ~2 % of (anchor, class) scores
exceed score_thr = 0.001 for random unit-norm class embeddings (trained-like); "dense" keeps bias 0.
Deterministic for a given (size, seed) on one machine: tensors are generated in schema order from one
CPU generator (the calibration pass adds only last-bit, machine-dependent noise).
"""
import math

import torch

from wedetect_b200 import schema


import os

CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache")


def synth_state_dict(size, *, seed=0, uni=False, num_prompts=256, with_text=True, cls_bias=0.0, text_vocab=None,
                     calibrate=True, regime="dense"):
    g = torch.Generator().manual_seed(seed)
    shapes = schema.param_shapes(size, uni=uni, num_prompts=num_prompts, with_text=with_text)
    sd = {}
    for name, shape in shapes.items():
        if text_vocab is not None and name.endswith("word_embeddings.weight"):
            shape = (text_vocab, shape[1])
        leaf = name.rsplit(".", 1)[-1]
        if name == "embeddings":
            t = torch.nn.functional.normalize(torch.randn(shape, generator=g), dim=-1)
        elif leaf == "gamma":
            t = torch.rand(shape, generator=g) * 0.45 + 0.05
        elif leaf == "alpha":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "logit_scale":
            t = torch.rand(shape, generator=g) - 1.5
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif "cls_contrasts" in name and leaf == "bias":
            t = torch.full(shape, float(cls_bias))
        elif leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "weight" and len(shape) == 1:
            t = torch.rand(shape, generator=g) + 0.5          # LN / BN scale
        elif "embeddings" in name and leaf == "weight":
            t = torch.randn(shape, generator=g) * 0.02         # XLM-R embedding tables
        elif leaf == "weight":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if "upsample_transpose" in name:                    # ConvTranspose2d: [Cin, Cout, 2, 2], one tap per output
                fan_in = shape[0]
            t = torch.randn(shape, generator=g) * (1.2 / math.sqrt(fan_in))
        else:
            raise KeyError(name)
        sd[name] = t
    if calibrate:
        # the calibrated statistics (~100 KB) are cached on disk so that another machine (the GPU box) reproduces
        # bit-identical weights without re-running the CPU calibration pass
        path = os.path.join(CACHE_DIR, f"calib_{size}_s{seed}_{regime}.pt")
        if os.path.exists(path):
            sd.update(torch.load(path))
        else:
            before = {k: v for k, v in sd.items()}
            _calibrate(sd, size, seed, regime)
            changed = {k: v for k, v in sd.items() if v is not before[k]}
            try:
                os.makedirs(CACHE_DIR, exist_ok=True)
                torch.save(changed, path)
            except OSError:
                pass
    return sd


def _calibrate(sd, size, seed, regime):
    from . import functional as Fn
    g = torch.Generator().manual_seed(seed + 7919)

    def rec(sd_, b, x):
        mean = x.mean((0, 2, 3))
        var = x.var((0, 2, 3), unbiased=False)
        sd_[b + ".running_mean"] = mean + torch.randn(mean.shape, generator=g) * 0.1 * var.sqrt()
        sd_[b + ".running_var"] = var * (torch.rand(var.shape, generator=g) * 0.45 + 0.8) + 1e-6

    probe = torch.nn.functional.normalize(torch.randn(64, schema.EMBED_DIM, generator=g), dim=-1)
    x = torch.rand(1, 3, 320, 320, generator=g)
    for l in range(3):
        sd[Fn.HM + f"cls_contrasts.{l}.bias"] = torch.zeros(())
    Fn._CALIBRATE = rec
    try:
        with torch.no_grad():
            out = Fn.vision_forward(sd, size, x, prompts=probe)
    finally:
        Fn._CALIBRATE = None
    if regime == "sparse":
        thr = math.log(0.001 / 0.999)
        for l, o in enumerate(out["levels"]):
            q = torch.quantile(o["logits"].flatten()[:2000000], 0.98)
            sd[Fn.HM + f"cls_contrasts.{l}.bias"] = (thr - q).reshape(())


def synth_images(B, H, W, seed=2):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, H, W, generator=g)
