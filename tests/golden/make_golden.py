"""Generate golden vectors by running the UNMODIFIED reference modules from /root/reference on the CPU.

    python tests/golden/make_golden.py        (only works where /root/reference exists)

What is executed from the reference: generate_proposal.py's ConvNeXt / CSPRepBiFPANNeck / YOLOWorldHeadModule /
BNContrastiveHead / SimpleYOLOWorldDetector.head_predict (filter_scores_and_topk + torchvision batched_nms),
with seeded synthetic weights from oracle/synth.py loaded through load_state_dict.  Stored: small slices and
statistics of every stage + the final proposals, so the travelling oracle can be pinned anywhere.
"""
import os
import re
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

_INV = {"0.conv": "0", "0.bn": "1", "1.conv": "3", "1.bn": "4", "2": "6"}


def to_gp_key(k):
    """mmengine layout -> the key names of generate_proposal.py's module tree (inverse of its remap :1236-1254)."""
    if k.startswith("backbone.image_model.model."):
        return "backbone." + k[len("backbone.image_model.model."):]
    m = re.match(r"bbox_head\.head_module\.(cls_preds|reg_preds)\.(\d)\.(0\.conv|0\.bn|1\.conv|1\.bn|2)\.(.*)$", k)
    if m:
        return f"bbox_head.{m.group(1)}.{m.group(2)}.{_INV[m.group(3)]}.{m.group(4)}"
    if k.startswith("bbox_head.head_module."):
        return "bbox_head." + k[len("bbox_head.head_module."):]
    return k


def digest(t, n=256):
    t = t.detach().float().reshape(-1)
    idx = torch.linspace(0, t.numel() - 1, n).long()
    return dict(mean=float(t.mean()), std=float(t.std()), absmax=float(t.abs().max()), sample=t[idx].clone(), numel=t.numel())


def main():
    import generate_proposal as gp
    from oracle import synth
    torch.set_num_threads(8)
    out = {}
    for size, H in (("base", 320),):
        sd = synth.synth_state_dict(size, seed=0, uni=True, regime="sparse")
        m = gp.SimpleYOLOWorldDetector(size, 768, 256, 300).eval()
        msg = m.load_state_dict({to_gp_key(k): v for k, v in sd.items()}, strict=True)
        x = synth.synth_images(2, H, H, seed=2)
        g = {}
        with torch.no_grad():
            feats = m.backbone(x)
            pyr = m.neck(feats)
            for i, f in enumerate(feats):
                g[f"c{i + 1}"] = digest(f.permute(0, 2, 3, 1))
            for i, f in enumerate(pyr):
                g[f"p{i + 3}"] = digest(f.permute(0, 2, 3, 1))
            for l in range(3):
                e, bp, lg = m.head_module_forward_single(pyr[l], m.bbox_head.cls_preds[l], m.bbox_head.reg_preds[l], m.bbox_head.cls_contrasts[l])
                g[f"embed{l}"] = digest(e.permute(0, 2, 3, 1))
                g[f"logit{l}"] = digest(lg.permute(0, 2, 3, 1))
                g[f"dist{l}"] = digest(bp.permute(0, 2, 3, 1))
            # text-conditioned head (BNContrastiveHead with normalised text), reference twin of yolo_world_head.py:263-294
            text = torch.randn(80, 768, generator=torch.Generator().manual_seed(5))
            outs = m.bbox_head(pyr, text[None].repeat(2, 1, 1))
            for l, (lg, bp) in enumerate(outs):
                g[f"text_logit{l}"] = digest(lg.permute(0, 2, 3, 1))
            res = m.head_predict(pyr)
            g["proposals"] = [dict(bboxes=r["bboxes"].clone(), scores=r["scores"].clone(), embeddings=digest(r["embeddings"])) for r in res]
        out[f"{size}_{H}"] = g
        print(size, H, msg, [len(r["scores"]) for r in res], float(res[0]["scores"][0]), float(res[0]["scores"][-1]))
    torch.save(out, os.path.join(HERE, "reference_stages.pt"))
    print("saved", os.path.getsize(os.path.join(HERE, "reference_stages.pt")))


if __name__ == "__main__":
    main()
