"""Load-time weight preparation (host, once): BN folding, layout changes, fp16 hi/lo (or bf16) conversion, upload.

Input: a checkpoint dict in the mmengine key layout (schema.normalize_state_dict).  Output: a flat dict of
device tensors consumed by plan.py.  Transformations (all exact algebra, done in fp64 on the CPU):
  * Conv+BN (eval)  ->  conv weight * g, bias = beta - mean * g,  g = bn_w / sqrt(var + eps)
    (neck eps 1e-5: yolo_world_pafpn.py:40-68 ; head eps 1e-3: yolov8_head.py:54-55)
  * 3x3 weights -> [Cout, 9 * pad64(Cin)] tap-major (ky, kx, c) zero padded per tap (implicit-GEMM layout)
  * 2x2 s2 patchify conv -> [Cout, (dy, dx, c)] matching the LN kernel's space-to-depth rows
  * ConvTranspose 2x2 s2 [Cin, Cout, 2, 2] -> rows (dy, dx, co) with co padded to 64
  * stem 4x4 s4 conv -> [C0, 64] (k = c*16 + dy*4 + dx, 48 valid); for uint8 BGR inputs the preprocessor
    (BGR->RGB, mean 0, std 255; config/wedetect_base.py:44-48) is folded into the weights
  * depthwise 7x7 [C,1,7,7] -> [49, C] fp32 (tap-major, channel-contiguous)
  * contrastive-head BN -> per-channel (g, h) pairs consumed by the fold_text kernel
  * XLM-R: Q/K/V fused into one [3H, H] matrix
"""
import torch

from . import schema
from .ops import P3


class DeviceWeights:
    def __init__(self, device, precise=True):
        self.device = device
        self.precise = precise
        self.t = {}

    def put_mat(self, name, w64):
        """GEMM B operand: fp16 hi/lo planes at the matrix's own power-of-two scale (parity-grade mode), or bf16 (fast mode)."""
        w32 = w64.float().contiguous()
        self.t[name] = P3.from_f32(w32, self.device) if self.precise else P3(w32.to(torch.bfloat16).to(self.device), 0)

    def put_f32(self, name, v):
        self.t[name] = v.float().contiguous().to(self.device)

    def mat(self, name):
        return self.t[name]

    def __getitem__(self, k):
        return self.t[k]

    def __contains__(self, k):
        return k in self.t


def _fold_bn(sd, conv_w, bn, eps):
    w = sd[conv_w].double()
    g = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + eps)
    b = sd[bn + ".bias"].double() - sd[bn + ".running_mean"].double() * g
    return w * g.view(-1, 1, 1, 1), b


def _taps(w):
    """[Cout, Cin, 3, 3] -> [Cout, 9 * pad64(Cin)] with k = (ky*3+kx)*Kc + c."""
    Cout, Cin = w.shape[:2]
    Kc = schema.pad64(Cin)
    out = torch.zeros(Cout, 9, Kc, dtype=w.dtype)
    out[:, :, :Cin] = w.permute(0, 2, 3, 1).reshape(Cout, 9, Cin)
    return out.reshape(Cout, 9 * Kc)


def _taps_flat(w):
    """[Cout, Cin, 3, 3] -> [Cout, 9 * Cin] (im2col layout of the stride-2 convs)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def prepare_vision(sd, size, device, *, input_format="f32_rgb", precise=True):
    cfg = schema.SIZES[size]
    W = DeviceWeights(device, precise)
    bb = "backbone.image_model.model."
    dims = cfg["dims"]
    # ---- stem ----
    w = sd[bb + "downsample_layers.0.0.weight"].double()          # [C0, 3, 4, 4]
    if input_format == "u8_bgr":
        w = w.flip(1) / 255.0                                       # kernel sees raw uint8 BGR
    elif input_format == "u8_rgb":
        w = w / 255.0                                               # Uni path: letterboxed uint8 RGB (generate_proposal.py:1098-1099)
    elif input_format != "f32_rgb":
        raise ValueError(input_format)
    ws = torch.zeros(dims[0], 64, dtype=torch.float64)
    ws[:, :48] = w.reshape(dims[0], 48)
    W.put_mat("stem.w", ws)
    W.put_f32("stem.b", sd[bb + "downsample_layers.0.0.bias"])
    W.put_f32("stem.ln_w", sd[bb + "downsample_layers.0.1.weight"])
    W.put_f32("stem.ln_b", sd[bb + "downsample_layers.0.1.bias"])
    for i in range(1, 4):
        d = bb + f"downsample_layers.{i}."
        W.put_f32(f"down{i}.ln_w", sd[d + "0.weight"])
        W.put_f32(f"down{i}.ln_b", sd[d + "0.bias"])
        W.put_mat(f"down{i}.w", sd[d + "1.weight"].double().permute(0, 2, 3, 1).reshape(dims[i], 4 * dims[i - 1]))
        W.put_f32(f"down{i}.b", sd[d + "1.bias"])
    for s in range(4):
        C = dims[s]
        for j in range(cfg["depths"][s]):
            p, q = bb + f"stages.{s}.{j}.", f"s{s}.b{j}."
            W.put_f32(q + "dw_w", sd[p + "dwconv.weight"].reshape(C, 49).t())
            W.put_f32(q + "dw_b", sd[p + "dwconv.bias"])
            W.put_f32(q + "ln_w", sd[p + "norm.weight"])
            W.put_f32(q + "ln_b", sd[p + "norm.bias"])
            W.put_mat(q + "w1", sd[p + "pwconv1.weight"].double())
            W.put_f32(q + "b1", sd[p + "pwconv1.bias"])
            W.put_mat(q + "w2", sd[p + "pwconv2.weight"].double())
            W.put_f32(q + "b2", sd[p + "pwconv2.bias"])
            W.put_f32(q + "gamma", sd[p + "gamma"])
    # ---- neck ----
    for m in schema.neck_layout(size):
        nm = "neck." + m["name"]
        if m["k"] == "alpha":
            W.t[nm + ".alpha"] = float(sd[nm + ".alpha"].reshape(-1)[0])
        elif m["k"] == "deconv":
            w = sd[nm + ".upsample_transpose.weight"].double()      # [Cin, Cout, 2, 2]
            Cin, Co = w.shape[:2]
            Cg = schema.pad64(Co)
            wp = torch.zeros(2, 2, Cg, Cin, dtype=torch.float64)
            wp[:, :, :Co] = w.permute(2, 3, 1, 0)
            W.put_mat(nm + ".w", wp.reshape(4 * Cg, Cin))
            bp = torch.zeros(2, Cg, dtype=torch.float64)
            bp[:, :Co] = sd[nm + ".upsample_transpose.bias"].double()
            W.put_f32(nm + ".b", bp.reshape(-1))
        else:
            w, b = _fold_bn(sd, nm + ".block.conv.weight", nm + ".block.bn", schema.BN_EPS_NECK)
            if m["k"] == 1:
                W.put_mat(nm + ".w", w.reshape(w.shape[0], w.shape[1]))
            else:
                W.put_mat(nm + ".w", _taps(w))
            W.put_f32(nm + ".b", b)
    # ---- head ----
    hm = "bbox_head.head_module."
    g_all, h_all = [], []
    for l in range(3):
        for br in ("cls_preds", "reg_preds"):
            p, q = hm + f"{br}.{l}.", f"head.{br}.{l}."
            for i in range(2):
                w, b = _fold_bn(sd, p + f"{i}.conv.weight", p + f"{i}.bn", schema.BN_EPS_HEAD)
                W.put_mat(q + f"{i}.w", _taps(w))
                W.put_f32(q + f"{i}.b", b)
            w = sd[p + "2.weight"].double()
            W.put_mat(q + "2.w", w.reshape(w.shape[0], w.shape[1]))
            W.put_f32(q + "2.b", sd[p + "2.bias"])
        c = hm + f"cls_contrasts.{l}."
        g = sd[c + "norm.weight"].double() / torch.sqrt(sd[c + "norm.running_var"].double() + schema.BN_EPS_HEAD)
        h = sd[c + "norm.bias"].double() - sd[c + "norm.running_mean"].double() * g
        W.put_f32(f"head.contrast.{l}.g", g)
        W.put_f32(f"head.contrast.{l}.h", h)
        W.put_f32(f"head.contrast.{l}.logit_scale", sd[c + "logit_scale"].reshape(1))
        W.put_f32(f"head.contrast.{l}.bias", sd[c + "bias"].reshape(1))
        # bound of |folded similarity weight| for unit-norm text rows: max|g| * exp(logit_scale); plan.py picks the power of two
        # the fp16 hi/lo planes of the folded matrix are stored at from it
        W.t[f"head.contrast.{l}.wmax"] = float(g.abs().max() * torch.exp(sd[c + "logit_scale"].double().reshape(-1)[0]))
        g_all.append(g)
        h_all.append(h)
    W.put_f32("head.contrast.g_all", torch.cat(g_all))
    W.put_f32("head.contrast.h_all", torch.cat(h_all))
    if "embeddings" in sd:
        W.put_f32("prompts", sd["embeddings"])
    return W


def prepare_text(sd, size, device, *, precise=True):
    t = schema.TEXT[schema.SIZES[size]["text"]]
    W = DeviceWeights(device, precise)
    tm = "backbone.text_model.model."
    W.put_f32("emb.word", sd[tm + "embeddings.word_embeddings.weight"])
    W.put_f32("emb.pos", sd[tm + "embeddings.position_embeddings.weight"])
    W.put_f32("emb.type", sd[tm + "embeddings.token_type_embeddings.weight"][0])
    W.put_f32("emb.ln_w", sd[tm + "embeddings.LayerNorm.weight"])
    W.put_f32("emb.ln_b", sd[tm + "embeddings.LayerNorm.bias"])
    for i in range(t["layers"]):
        p, q = tm + f"encoder.layer.{i}.", f"l{i}."
        W.put_mat(q + "qkv.w", torch.cat([sd[p + f"attention.self.{n}.weight"].double() for n in ("query", "key", "value")], 0))
        W.put_f32(q + "qkv.b", torch.cat([sd[p + f"attention.self.{n}.bias"] for n in ("query", "key", "value")], 0))
        W.put_mat(q + "o.w", sd[p + "attention.output.dense.weight"].double())
        W.put_f32(q + "o.b", sd[p + "attention.output.dense.bias"])
        W.put_f32(q + "ln1_w", sd[p + "attention.output.LayerNorm.weight"])
        W.put_f32(q + "ln1_b", sd[p + "attention.output.LayerNorm.bias"])
        W.put_mat(q + "f1.w", sd[p + "intermediate.dense.weight"].double())
        W.put_f32(q + "f1.b", sd[p + "intermediate.dense.bias"])
        W.put_mat(q + "f2.w", sd[p + "output.dense.weight"].double())
        W.put_f32(q + "f2.b", sd[p + "output.dense.bias"])
        W.put_f32(q + "ln2_w", sd[p + "output.LayerNorm.weight"])
        W.put_f32(q + "ln2_b", sd[p + "output.LayerNorm.bias"])
    W.put_mat("head.w", sd["backbone.text_model.head.weight"].double())
    W.put_f32("head.b", sd["backbone.text_model.head.bias"])
    return W
