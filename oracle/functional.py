"""CPU fp32 ORACLE of the WeDetect dual-tower forward, written as pure functions over a checkpoint dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is a restatement, not a copy: the reference
builds nn.Module trees; here every stage is a function of (state_dict, tensors) so the same code serves
tiny / base / large, text-conditioned and Uni heads.  `tests/test_oracle_pin.py` checks it against the
reference's own modules imported from /root/reference (when present) and against committed goldens.

Citations (paths relative to the reference checkout):
  ConvNeXt forward          wedetect/models/backbones/mm_backbone.py:112-125,145-155,188-198,233-255
  CSPRepBiFPAN neck         wedetect/models/necks/yolo_world_pafpn.py:40-68,195-208,566-647,692-715,1114-1137
  head stacks / DFL         wedetect/models/dense_heads/yolo_world_head.py:195-232,263-294
  BN contrastive head       wedetect/models/dense_heads/yolo_world_head.py:90-108 ; Uni: generate_proposal.py:1129-1131
  text tower                wedetect/models/backbones/mm_backbone.py:376-390 (+ HF XLMRobertaModel)
  detector glue             wedetect/models/detectors/yolo_world.py:35-113 ; generate_proposal.py:1082-1218
"""
import math

import torch
import torch.nn.functional as F

from wedetect_b200 import schema  # pure-Python shape tables only

BB = "backbone.image_model.model."
HM = "bbox_head.head_module."


def _ln_cf(x, w, b, eps):
    """channels-first LayerNorm (mm_backbone.py:150-155)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[None, :, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[None, :, None, None]


def convnext(sd, size, x):
    cfg = schema.SIZES[size]
    feats = []
    for s in range(4):
        d = BB + f"downsample_layers.{s}."
        if s == 0:
            x = F.conv2d(x, sd[d + "0.weight"], sd[d + "0.bias"], stride=4)
            x = _ln_cf(x, sd[d + "1.weight"], sd[d + "1.bias"], schema.LN_EPS)
        else:
            x = _ln_cf(x, sd[d + "0.weight"], sd[d + "0.bias"], schema.LN_EPS)
            x = F.conv2d(x, sd[d + "1.weight"], sd[d + "1.bias"], stride=2)
        for j in range(cfg["depths"][s]):
            p = BB + f"stages.{s}.{j}."
            C = x.shape[1]
            y = F.conv2d(x, sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=C)
            y = y.permute(0, 2, 3, 1)
            y = F.layer_norm(y, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], schema.LN_EPS)
            y = F.gelu(F.linear(y, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"]))
            y = F.linear(y, sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"]) * sd[p + "gamma"]
            x = x + y.permute(0, 3, 1, 2)
        feats.append(x)
    return feats


_CALIBRATE = None  # set by oracle.synth: callable(sd, bn_prefix, x) that rewrites running stats before use


def _bn(sd, b, x, eps):
    """eval-mode BatchNorm2d with the checkpoint's running statistics."""
    if _CALIBRATE is not None:
        _CALIBRATE(sd, b, x)
    return F.batch_norm(x, sd[b + ".running_mean"], sd[b + ".running_var"], sd[b + ".weight"], sd[b + ".bias"], False, 0.0, eps)


def _cba(sd, name, x, act, stride=1, eps=schema.BN_EPS_NECK):
    w = sd[name + ".block.conv.weight"]
    y = _bn(sd, name + ".block.bn", F.conv2d(x, w, None, stride=stride, padding=w.shape[-1] // 2), eps)
    return F.relu(y) if act == "relu" else F.silu(y)


def _bepc3(sd, name, x, n):
    a = _cba(sd, name + ".cv1", x, "silu")
    for i in range(n):
        blk = f"{name}.m.conv1" if i == 0 else f"{name}.m.block.{i - 1}"
        y = _cba(sd, blk + ".conv2", _cba(sd, blk + ".conv1", a, "silu"), "silu")
        a = y + sd[blk + ".alpha"] * a
    return _cba(sd, name + ".cv3", torch.cat((a, _cba(sd, name + ".cv2", x, "silu")), 1), "silu")


def _bifusion(sd, name, top, mid, low):
    up = F.conv_transpose2d(top, sd[name + ".upsample.upsample_transpose.weight"], sd[name + ".upsample.upsample_transpose.bias"], stride=2)
    a = _cba(sd, name + ".cv1", mid, "relu")
    b = _cba(sd, name + ".downsample", _cba(sd, name + ".cv2", low, "relu"), "relu", stride=2)
    return _cba(sd, name + ".cv3", torch.cat((up, a, b), 1), "relu")


def neck(sd, size, feats):
    c1, c2, c3, c4 = feats
    n = schema.SIZES[size]["neck_repeats"] // 2
    N = "neck."
    fpn0 = _cba(sd, N + "reduce_layer0", c4, "relu")
    f0 = _bepc3(sd, N + "Rep_p4", _bifusion(sd, N + "Bifusion0", fpn0, c3, c2), n)
    fpn1 = _cba(sd, N + "reduce_layer1", f0, "relu")
    p3 = _bepc3(sd, N + "Rep_p3", _bifusion(sd, N + "Bifusion1", fpn1, c2, c1), n)
    d1 = _cba(sd, N + "downsample2", p3, "relu", stride=2)
    p4 = _bepc3(sd, N + "Rep_n3", torch.cat((d1, fpn1), 1), n)
    d0 = _cba(sd, N + "downsample1", p4, "relu", stride=2)
    p5 = _bepc3(sd, N + "Rep_n4", torch.cat((d0, fpn0), 1), n)
    return [p3, p4, p5]


def _head_stack(sd, p, x):
    for i in range(2):
        x = F.silu(_bn(sd, p + f"{i}.bn", F.conv2d(x, sd[p + f"{i}.conv.weight"], None, padding=1), schema.BN_EPS_HEAD))
    return F.conv2d(x, sd[p + "2.weight"], sd[p + "2.bias"])


def head(sd, feats, text=None, prompts=None):
    """Returns per level: dict(embed [B,HW,768] after BN, logits [B,HW,K], dist [B,HW,4], dfl_prob [B,HW,4,16] = the bin
    distribution whose expectation `dist` is).
    text: [K,768] (L2-normalised inside, BNContrastiveHead) ; prompts: [P,768] used raw (Uni)."""
    outs = []
    for l, f in enumerate(feats):
        B, _, H, W = f.shape
        e = _head_stack(sd, HM + f"cls_preds.{l}.", f)
        c = HM + f"cls_contrasts.{l}."
        e = _bn(sd, c + "norm", e, schema.BN_EPS_HEAD)
        w = F.normalize(text, dim=-1, p=2) if text is not None else prompts
        logits = torch.einsum("bchw,kc->bkhw", e, w) * sd[c + "logit_scale"].exp() + sd[c + "bias"]
        r = _head_stack(sd, HM + f"reg_preds.{l}.", f)
        r = r.reshape(B, 4, schema.REG_MAX, H * W).permute(0, 3, 1, 2).softmax(3)
        dist = r.matmul(torch.arange(schema.REG_MAX, dtype=r.dtype, device=r.device))
        outs.append(dict(embed=e.permute(0, 2, 3, 1).reshape(B, H * W, -1), logits=logits.permute(0, 2, 3, 1).reshape(B, H * W, -1), dist=dist, dfl_prob=r))
    return outs


def text_tower(sd, size, ids, mask):
    """XLM-R encoder (post-LN BERT) + CLS + Linear + L2 norm.  ids/mask: int [S, L]."""
    t = schema.TEXT[schema.SIZES[size]["text"]]
    H, nh = t["hidden"], t["heads"]
    tm = "backbone.text_model.model."
    ids = ids.long()
    m = (ids != schema.TEXT_PAD).int()
    pos = (torch.cumsum(m, 1) * m).long() + schema.TEXT_PAD
    x = sd[tm + "embeddings.word_embeddings.weight"][ids] + sd[tm + "embeddings.token_type_embeddings.weight"][0] + \
        sd[tm + "embeddings.position_embeddings.weight"][pos]
    x = F.layer_norm(x, (H,), sd[tm + "embeddings.LayerNorm.weight"], sd[tm + "embeddings.LayerNorm.bias"], schema.TEXT_EPS)
    S, L = ids.shape
    neg = torch.zeros(S, 1, 1, L).masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    for i in range(t["layers"]):
        p = tm + f"encoder.layer.{i}."

        def lin(n, v):
            return F.linear(v, sd[p + n + ".weight"], sd[p + n + ".bias"])

        q = lin("attention.self.query", x).view(S, L, nh, 64).transpose(1, 2)
        k = lin("attention.self.key", x).view(S, L, nh, 64).transpose(1, 2)
        v = lin("attention.self.value", x).view(S, L, nh, 64).transpose(1, 2)
        a = ((q @ k.transpose(-1, -2)) / math.sqrt(64) + neg).softmax(-1) @ v
        a = a.transpose(1, 2).reshape(S, L, H)
        x = F.layer_norm(lin("attention.output.dense", a) + x, (H,), sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], schema.TEXT_EPS)
        h = F.gelu(lin("intermediate.dense", x))
        x = F.layer_norm(lin("output.dense", h) + x, (H,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], schema.TEXT_EPS)
    y = F.linear(x[:, 0], sd["backbone.text_model.head.weight"], sd["backbone.text_model.head.bias"])
    return F.normalize(y, dim=-1)


def preprocess(img_u8_bgr):
    """mmdet DetDataPreprocessor eval path with config/wedetect_base.py:44-48: BGR->RGB, (x - 0) / 255."""
    return img_u8_bgr.flip(1).float() / 255.0


def vision_forward(sd, size, images, text=None, prompts=None):
    """images: fp32 [B,3,H,W] RGB in [0,1].  Returns head outputs per level (+ backbone / neck features)."""
    feats = convnext(sd, size, images)
    pyr = neck(sd, size, feats)
    return dict(backbone=feats, neck=pyr, levels=head(sd, pyr, text=text, prompts=prompts))
