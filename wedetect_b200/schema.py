"""Model geometry + checkpoint schema (mmengine key layout) for WeDetect tiny / base / large.

Sizes follow the reference's string-keyed tables:
  ConvNeXt depths/dims     wedetect/models/backbones/mm_backbone.py:281-288
  neck channels / repeats  wedetect/models/necks/yolo_world_pafpn.py:999-1095 (scale_factor from config/*.py)
  head channels            wedetect/models/dense_heads/yolo_world_head.py:180-191
  text tower               xlm-roberta-{base,large}/config.json, head Linear mm_backbone.py:360-365
Key names are the mmengine checkpoint layout; `normalize_state_dict` also accepts the
remapped names used by generate_proposal.py:1236-1254.
"""
import re
from collections import OrderedDict

SIZES = {
    "tiny": dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), neck_scale=0.75, neck_repeats=6, text="base"),
    "base": dict(depths=(3, 3, 27, 3), dims=(128, 256, 512, 1024), neck_scale=1.0, neck_repeats=12, text="base"),
    "large": dict(depths=(3, 3, 27, 3), dims=(192, 384, 768, 1536), neck_scale=1.5, neck_repeats=12, text="large"),
}
TEXT = {
    "base": dict(hidden=768, layers=12, heads=12, inter=3072),
    "large": dict(hidden=1024, layers=24, heads=16, inter=4096),
}
TEXT_VOCAB, TEXT_MAXPOS, TEXT_PAD, TEXT_EPS = 250002, 514, 1, 1e-5
EMBED_DIM = 768
HEAD_CLS_CH, HEAD_REG_CH, REG_MAX = 256, 64, 16
STRIDES = (8, 16, 32)
BN_EPS_NECK, BN_EPS_HEAD, LN_EPS = 1e-5, 1e-3, 1e-6
_NECK_CH = [64, 128, 256, 512, 1024, 256, 128, 128, 256, 256, 512]


def pad64(c):
    return (c + 63) // 64 * 64


def neck_channels(size):
    s = SIZES[size]["neck_scale"]
    return [int(c * s) for c in _NECK_CH]


def head_in_channels(size):
    ch = neck_channels(size)
    return [ch[6], ch[8], ch[10]]


def neck_layout(size):
    """(name, kind, cin, cout) of every neck conv module, in execution order (yolo_world_pafpn.py:1114-1137)."""
    ch = neck_channels(size)
    n = SIZES[size]["neck_repeats"] // 2
    out = []

    def conv(name, k, cin, cout, act, stride=1):
        out.append(dict(name=name, k=k, cin=cin, cout=cout, act=act, stride=stride))

    def bifusion(name, cin1, cin2, cout):
        conv(f"{name}.cv1", 1, cin1, cout, "relu")
        conv(f"{name}.cv2", 1, cin2, cout, "relu")
        conv(f"{name}.cv3", 1, 3 * cout, cout, "relu")
        out.append(dict(name=f"{name}.upsample", k="deconv", cin=cout, cout=cout))
        conv(f"{name}.downsample", 3, cout, cout, "relu", 2)

    def bepc3(name, cin, cout):
        c_ = int(cout * 0.5)
        conv(f"{name}.cv1", 1, cin, c_, "silu")
        conv(f"{name}.cv2", 1, cin, c_, "silu")
        conv(f"{name}.cv3", 1, 2 * c_, cout, "silu")
        for i in range(n):
            blk = f"{name}.m.conv1" if i == 0 else f"{name}.m.block.{i - 1}"
            conv(f"{blk}.conv1", 3, c_, c_, "silu")
            conv(f"{blk}.conv2", 3, c_, c_, "silu")
            out.append(dict(name=blk, k="alpha"))

    conv("reduce_layer0", 1, ch[4], ch[5], "relu")
    bifusion("Bifusion0", ch[3], ch[2], ch[5])
    bepc3("Rep_p4", ch[5], ch[5])
    conv("reduce_layer1", 1, ch[5], ch[6], "relu")
    bifusion("Bifusion1", ch[2], ch[1], ch[6])
    bepc3("Rep_p3", ch[6], ch[6])
    conv("downsample2", 3, ch[6], ch[7], "relu", 2)
    bepc3("Rep_n3", ch[6] + ch[7], ch[8])
    conv("downsample1", 3, ch[8], ch[9], "relu", 2)
    bepc3("Rep_n4", ch[5] + ch[9], ch[10])
    return out


def param_shapes(size, *, uni=False, num_prompts=256, with_text=True):
    """OrderedDict name -> shape of every tensor in an (mmengine-layout) checkpoint."""
    cfg = SIZES[size]
    P = OrderedDict()
    dims, depths = cfg["dims"], cfg["depths"]
    bb = "backbone.image_model.model."
    P[bb + "downsample_layers.0.0.weight"] = (dims[0], 3, 4, 4)
    P[bb + "downsample_layers.0.0.bias"] = (dims[0],)
    P[bb + "downsample_layers.0.1.weight"] = (dims[0],)
    P[bb + "downsample_layers.0.1.bias"] = (dims[0],)
    for i in range(1, 4):
        P[bb + f"downsample_layers.{i}.0.weight"] = (dims[i - 1],)
        P[bb + f"downsample_layers.{i}.0.bias"] = (dims[i - 1],)
        P[bb + f"downsample_layers.{i}.1.weight"] = (dims[i], dims[i - 1], 2, 2)
        P[bb + f"downsample_layers.{i}.1.bias"] = (dims[i],)
    for s in range(4):
        d = dims[s]
        for j in range(depths[s]):
            p = bb + f"stages.{s}.{j}."
            P[p + "gamma"] = (d,)
            P[p + "dwconv.weight"] = (d, 1, 7, 7)
            P[p + "dwconv.bias"] = (d,)
            P[p + "norm.weight"] = (d,)
            P[p + "norm.bias"] = (d,)
            P[p + "pwconv1.weight"] = (4 * d, d)
            P[p + "pwconv1.bias"] = (4 * d,)
            P[p + "pwconv2.weight"] = (d, 4 * d)
            P[p + "pwconv2.bias"] = (d,)

    def bn(prefix, c):
        P[prefix + ".weight"] = (c,)
        P[prefix + ".bias"] = (c,)
        P[prefix + ".running_mean"] = (c,)
        P[prefix + ".running_var"] = (c,)

    for m in neck_layout(size):
        nm = "neck." + m["name"]
        if m["k"] == "alpha":
            P[nm + ".alpha"] = (1,)
        elif m["k"] == "deconv":
            P[nm + ".upsample_transpose.weight"] = (m["cin"], m["cout"], 2, 2)
            P[nm + ".upsample_transpose.bias"] = (m["cout"],)
        else:
            P[nm + ".block.conv.weight"] = (m["cout"], m["cin"], m["k"], m["k"])
            bn(nm + ".block.bn", m["cout"])
    hm = "bbox_head.head_module."
    for l, cin in enumerate(head_in_channels(size)):
        for branch, mid, last in (("reg_preds", HEAD_REG_CH, 4 * REG_MAX), ("cls_preds", HEAD_CLS_CH, EMBED_DIM)):
            p = hm + f"{branch}.{l}."
            P[p + "0.conv.weight"] = (mid, cin, 3, 3)
            bn(p + "0.bn", mid)
            P[p + "1.conv.weight"] = (mid, mid, 3, 3)
            bn(p + "1.bn", mid)
            P[p + "2.weight"] = (last, mid, 1, 1)
            P[p + "2.bias"] = (last,)
    for l in range(3):
        p = hm + f"cls_contrasts.{l}."
        P[p + "bias"] = ()
        P[p + "logit_scale"] = ()
        bn(p + "norm", EMBED_DIM)
    if uni:
        P["embeddings"] = (num_prompts, EMBED_DIM)
    elif with_text:
        t = TEXT[cfg["text"]]
        H, I = t["hidden"], t["inter"]
        tm = "backbone.text_model.model."
        P[tm + "embeddings.word_embeddings.weight"] = (TEXT_VOCAB, H)
        P[tm + "embeddings.position_embeddings.weight"] = (TEXT_MAXPOS, H)
        P[tm + "embeddings.token_type_embeddings.weight"] = (1, H)
        P[tm + "embeddings.LayerNorm.weight"] = (H,)
        P[tm + "embeddings.LayerNorm.bias"] = (H,)
        for i in range(t["layers"]):
            p = tm + f"encoder.layer.{i}."
            for nm in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
                P[p + nm + ".weight"] = (H, H)
                P[p + nm + ".bias"] = (H,)
            P[p + "attention.output.LayerNorm.weight"] = (H,)
            P[p + "attention.output.LayerNorm.bias"] = (H,)
            P[p + "intermediate.dense.weight"] = (I, H)
            P[p + "intermediate.dense.bias"] = (I,)
            P[p + "output.dense.weight"] = (H, I)
            P[p + "output.dense.bias"] = (H,)
            P[p + "output.LayerNorm.weight"] = (H,)
            P[p + "output.LayerNorm.bias"] = (H,)
        P["backbone.text_model.head.weight"] = (EMBED_DIM, H)
        P["backbone.text_model.head.bias"] = (EMBED_DIM,)
    return P


_UNI_SEQ = {"0": "0.conv", "1": "0.bn", "3": "1.conv", "4": "1.bn", "6": "2"}


def normalize_key(k):
    """Map a generate_proposal.py-style key (its remap at :1236-1254) back to the mmengine layout."""
    if k.startswith("backbone.") and not k.startswith(("backbone.image_model.", "backbone.text_model.")):
        return "backbone.image_model.model." + k[len("backbone."):]
    m = re.match(r"bbox_head\.(cls_preds|reg_preds)\.(\d)\.(\d)\.(.*)$", k)
    if m and m.group(3) in _UNI_SEQ:
        return f"bbox_head.head_module.{m.group(1)}.{m.group(2)}.{_UNI_SEQ[m.group(3)]}.{m.group(4)}"
    if k.startswith("bbox_head.cls_contrasts."):
        return "bbox_head.head_module." + k[len("bbox_head."):]
    return k


def normalize_state_dict(sd):
    if "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]
    return {normalize_key(k): v for k, v in sd.items() if not k.endswith("num_batches_tracked")}


def level_hw(img_h, img_w):
    return [(img_h // s, img_w // s) for s in STRIDES]


def text_state_dict(sd):
    """The text-tower slice of a WeDetect checkpoint, keyed like the full model (`backbone.text_model.*`).  Accepts an
    mmengine checkpoint ({'state_dict': ...}), a full state dict, or the slice the reference's standalone text tower loads
    (`model.*` / `head.*`, eval_retrieval/extract_embedding.py:1293-1303)."""
    if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]
    out = {}
    for k, v in sd.items():
        if k.startswith("backbone.text_model."):
            out[k] = v
        elif k.startswith(("model.", "head.")):
            out["backbone.text_model." + k] = v
    if not out:
        raise KeyError("no text-tower tensors (backbone.text_model.* or model.* / head.*) in the checkpoint")
    return out


def text_size_of(sd):
    """'base' / 'large' from the hidden width of the head Linear (768 -> 768 or 1024 -> 768)."""
    w = sd["backbone.text_model.head.weight"]
    for name, t in TEXT.items():
        if t["hidden"] == w.shape[1]:
            return name
    raise ValueError(f"unknown text tower width {tuple(w.shape)}")
