"""Plain PyTorch fp32 (CPU) references of each device op — the per-kernel numerics checkers.

Each function computes, in fp32 on the CPU, exactly what the matching `wd_op` is specified to compute
from the same (already bf16-rounded) operands.  Test infrastructure only.
"""
import math

import torch
import torch.nn.functional as F


def act_ref(x, act):
    if act == 1:
        return torch.relu(x)
    if act == 2:
        return F.silu(x)
    if act == 3:
        return F.gelu(x)  # erf form
    return x


def epilogue_ref(acc, bias=None, act=0, gamma=None, resid=None, alpha=1.0):
    v = acc
    if bias is not None:
        v = v + bias
    v = act_ref(v, act)
    if gamma is not None:
        v = v * gamma
    if resid is not None:
        v = v + alpha * resid.float()
    return v


def linear_ref(A, W, **epi):
    return epilogue_ref(A.float() @ W.float().t(), **epi)


def conv3x3_ref(A, W, **epi):
    """A [B,H,W,Cin] (bf16), W [N, 9*Cin] with k = (ky*3+kx)*Cin + c."""
    B, H, Wd, Cin = A.shape
    N = W.shape[0]
    w = W.float().view(N, 3, 3, Cin).permute(0, 3, 1, 2).contiguous()
    y = F.conv2d(A.float().permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    resid = epi.pop("resid", None)
    bias = epi.pop("bias", None)
    return epilogue_ref(y, bias=bias, resid=resid, **epi)


def deconv2x2_ref(A, W, bias):
    """A [B,H,W,Cin]; W [4*Co, Cin] rows (dy,dx,co); bias [Co] -> [B,2H,2W,Co]."""
    B, H, Wd, Cin = A.shape
    Co = W.shape[0] // 4
    y = A.float().reshape(-1, Cin) @ W.float().t()  # [M, 4*Co]
    y = y.view(B, H, Wd, 2, 2, Co).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * Wd, Co)
    return y + bias


def dfl_ref(A, W, bias):
    z = A.float() @ W.float().t() + bias  # [M, 64]
    z = z.view(-1, 4, 16).softmax(-1)
    return (z * torch.arange(16, dtype=torch.float32)).sum(-1)


def ln_ref(x, w, b, eps):
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * w + b


def s2d_ref(y, B, H, W):
    """rows (b,y,x) x C  ->  rows (b,y/2,x/2) x (dy,dx,C)."""
    C = y.shape[-1]
    return y.view(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B * (H // 2) * (W // 2), 4 * C)


def dwconv_ln_ref(x, w49, bias, ln_w, ln_b, eps):
    """x [B,H,W,C] f32; w49 [49, C] (tap = ky*7+kx)."""
    B, H, W, C = x.shape
    w = w49.t().reshape(C, 1, 7, 7)
    y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, padding=3, groups=C).permute(0, 2, 3, 1)
    return ln_ref(y, ln_w, ln_b, eps).reshape(B * H * W, C)


def stem_patch_ref(img, scale):
    """img [B,3,H,W] -> [B*(H/4)*(W/4), 64], k = c*16 + dy*4 + dx."""
    B, _, H, W = img.shape
    x = img.float() * scale
    p = x.view(B, 3, H // 4, 4, W // 4, 4).permute(0, 2, 4, 1, 3, 5).reshape(B * (H // 4) * (W // 4), 48)
    return torch.cat([p, torch.zeros(p.shape[0], 16)], 1)


def im2col_s2_ref(x):
    """x [B,H,W,C] -> [B*Ho*Wo, 9*C], k = (ky*3+kx)*C + c, stride 2 pad 1."""
    B, H, W, C = x.shape
    cols = F.unfold(x.float().permute(0, 3, 1, 2), kernel_size=3, padding=1, stride=2)  # [B, C*9, L]
    L = cols.shape[-1]
    cols = cols.view(B, C, 9, L).permute(0, 3, 2, 1).reshape(B * L, 9 * C)
    return cols


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def split_hi_lo(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def text_embed_ref(ids, word, pos, typ, ln_w, ln_b, eps, pad_idx):
    mask = (ids != pad_idx).int()
    pos_ids = (torch.cumsum(mask, 1) * mask).long() + pad_idx
    x = word[ids.long()] + typ + pos[pos_ids]
    return ln_ref(x, ln_w, ln_b, eps).reshape(-1, word.shape[1])


def attn_ref(qkv, mask, heads, scale):
    S, L = mask.shape
    Hd = qkv.shape[1] // 3
    q, k, v = qkv.view(S, L, 3, heads, 64).permute(2, 0, 3, 1, 4)  # [S, h, L, 64]
    s = (q @ k.transpose(-1, -2)) * scale
    s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    o = s.softmax(-1) @ v
    return o.permute(0, 2, 1, 3).reshape(S * L, Hd)


def fold_text_ref(text, g, h, logit_scale, bias, normalize):
    tn = F.normalize(text, dim=-1, p=2) if normalize else text
    es = math.exp(float(logit_scale))
    return tn * g * es, es * (tn @ h) + float(bias)
