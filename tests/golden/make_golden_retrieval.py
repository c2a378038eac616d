"""Golden vectors for the object-retrieval path, produced by executing the UNMODIFIED reference on the CPU.

    python tests/golden/make_golden_retrieval.py        (only works where /root/reference exists)

Executed from the reference:
  * eval_retrieval/extract_embedding.py: SimpleYOLOWorldDetector('base', 768, 256) backbone / neck / head_predict
    (the extract variant that also returns labels / scales / bias), with seeded synthetic weights;
    `pycocotools` (absent offline, only used by the script's dataset class) is stubbed, and `grid_size` - hard-coded
    for 640x640 at :1101 - is set to the 320x320 values.
  * eval_retrieval/retrieval_metric.py:362-377 - the scoring loop - is exec'd verbatim from the reference file (the
    file is a script, not a module) on the payload built from those proposals.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REF, "eval_retrieval"))
sys.path.insert(0, REF)


def _digest(t, n=256):
    t = t.detach().float().reshape(-1)
    idx = torch.linspace(0, t.numel() - 1, n).long()
    return dict(mean=float(t.mean()), std=float(t.std()), absmax=float(t.abs().max()), sample=t[idx].clone(), numel=t.numel())


def main():
    for name in ("pycocotools", "pycocotools.coco"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = object
    import extract_embedding as ee
    from make_golden import to_gp_key
    from oracle import synth
    torch.set_num_threads(8)
    H = 320
    sd = synth.synth_state_dict("base", seed=0, uni=True, regime="sparse")
    m = ee.SimpleYOLOWorldDetector("base", 768, 256).eval()
    msg = m.load_state_dict({to_gp_key(k): v for k, v in sd.items()}, strict=True)
    m.grid_size = [(H // s) ** 2 for s in (8, 16, 32)]
    x = synth.synth_images(2, H, H, seed=2)
    with torch.no_grad():
        # one image per head_predict call, as the reference's script does (extract_embedding.py:1723-1724): with B > 1 its
        # loop re-binds `scales` / `bias` to the first image's filtered rows (:1246-1247) and later images index garbage
        pyr = m.neck(m.backbone(x))
        res = [m.head_predict([f[i:i + 1] for f in pyr])[0] for i in range(x.shape[0])]
    text = torch.nn.functional.normalize(torch.randn(80, 768, generator=torch.Generator().manual_seed(11)), dim=-1)
    pred = {"image_embedding": [dict(image_id=100 + i, embedding=r["embeddings"], scale=r["scales"], bias=r["bias"]) for i, r in enumerate(res)],
            "text_embedding": text}
    # the reference's scoring loop, verbatim from its file
    src = open(os.path.join(REF, "eval_retrieval", "retrieval_metric.py")).read().split("\n")
    first = next(i for i, l in enumerate(src) if l.startswith("PREDICTIONS = "))
    last = next(i for i, l in enumerate(src) if l.startswith('print("Starting evaluation...")'))
    body = "\n".join(src[first:last])
    classnames = [f"class_{k}" for k in range(80)]
    g = {}
    for thre in (0.3, 0.01):
        ns = dict(torch=torch, pred=pred, text_embedding=text, classnames=classnames, args=types.SimpleNamespace(model="wedetect", thre=thre))
        exec(body, ns)
        g[f"predictions_{thre}"] = ns["PREDICTIONS"]
    ns = dict(torch=torch, pred=pred, text_embedding=text, classnames=classnames, args=types.SimpleNamespace(model="hqclip", thre=0.55))
    exec(body, ns)
    g["predictions_hqclip_0.55"] = ns["PREDICTIONS"]
    # per-image score rows through the same reference lines (loop body up to the max)
    rows = []
    for result in pred["image_embedding"]:
        ns = dict(torch=torch, result=result, text_embedding=text, args=types.SimpleNamespace(model="wedetect"))
        lines = [l[4:] for l in src[first:last] if l.startswith("    cls_logits") or l.startswith("    if args.model") or l.startswith("    else") or l.startswith("        cls_logits")]
        exec("\n".join(lines), ns)
        rows.append(ns["cls_logits"].clone())
    g["scores"] = torch.stack(rows)
    g["proposals"] = [dict(bboxes=r["bboxes"].clone(), scores=r["scores"].clone(), labels=r["labels"].clone(), scales=r["scales"].clone(),
                           bias=r["bias"].clone(), embeddings=_digest(r["embeddings"])) for r in res]
    g["text_seed"] = 11
    torch.save(g, os.path.join(HERE, "reference_retrieval.pt"))
    print(msg, [len(r["scores"]) for r in res], g["scores"].shape, float(g["scores"].max()), {k: sum(len(v) for v in p.values()) for k, p in g.items() if k.startswith("predictions")})
    print("saved", os.path.getsize(os.path.join(HERE, "reference_retrieval.pt")))


if __name__ == "__main__":
    main()
