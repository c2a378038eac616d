"""Builders that turn torch tensor *views* into `wd_op` records for the C ABI.

torch is used only for device memory (data_ptr / strides); nothing here computes.
All activations are NHWC: a feature map is a view [B, H, W, C] whose last stride is 1, a matrix is
[rows, cols].  Channel slices of a wider buffer (concat fusion) are plain views.
"""
import functools
import math
import os as _os

import torch

from . import _lib as L
from ._lib import WdOp


ACT_SCALE = L.ACT_PLANE_SCALE     # power of two every activation written by the library is stored at (fp16 hi/lo planes)


def weight_scale(x):
    """Power of two s with max|x| * s in [8192, 16384): keeps the low plane of a weight matrix out of the fp16 subnormals."""
    m = float(x.abs().max())
    if not (m > 0.0) or not math.isfinite(m):
        return 1.0
    return 2.0 ** (13 - math.floor(math.log2(m)))


class P3:
    """A 16-bit GEMM operand view.  ps == 0: an ordinary bf16 tensor (fast mode).  ps > 0 (parity-grade mode): `t` is the
    HIGH plane of an fp16 hi/lo pair, the low plane lives `ps` elements further in the same allocation, and
    hi + lo = value * scale (scale a power of two: ACT_SCALE for activations, per matrix for weights).
    torch sees both planes as 16-bit storage of dtype float16 (planes) or bfloat16 (single)."""

    def __init__(self, t, ps=0, scale=1.0):
        self.t, self.ps, self.scale = t, int(ps), float(scale)

    @staticmethod
    def from_f32(x, device=None, scale=None):
        """Split an fp32 tensor into fp16 hi/lo planes of x * scale (test / weight-upload helper)."""
        x = x.float()
        if scale is None:
            scale = weight_scale(x)
        y = (x * scale).clamp(-65504.0, 65504.0)
        base = torch.empty((2,) + tuple(x.shape), dtype=torch.float16)
        base[0] = y.to(torch.float16)
        base[1] = (y - base[0].float()).to(torch.float16)
        if device is not None:
            base = base.to(device)
        return P3(base[0], base.stride(0), scale)

    @staticmethod
    def zeros(shape, device, planes, scale=ACT_SCALE):
        if planes:
            base = torch.zeros((2,) + tuple(shape), dtype=torch.float16, device=device)
            return P3(base[0], base.stride(0), scale)
        return P3(torch.zeros(shape, dtype=torch.bfloat16, device=device), 0, 1.0)

    def value(self):
        """fp32 reconstruction."""
        if not self.ps:
            return self.t.float()
        lo = self.t.as_strided(self.t.shape, self.t.stride(), self.t.storage_offset() + self.ps)
        return (self.t.float() + lo.float()) / self.scale

    def view(self, fn):
        return P3(fn(self.t), self.ps, self.scale)


def _tp(x):
    if x is None:
        return None, 0
    return (x.t, x.ps) if isinstance(x, P3) else (x, 0)


def _sc(x):
    return x.scale if isinstance(x, P3) and x.ps else 1.0


def _ptr(t):
    return None if t is None else t.data_ptr()


def _chk16(t, ps, name):
    """16-bit operand: fp16 planes (ps > 0) or a single bf16 plane."""
    _chk(t, torch.float16 if ps else torch.bfloat16, name)


def _chk(t, dtype, name):
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name}: innermost stride must be 1")


@functools.lru_cache(maxsize=None)
def pick_tile(W, H, B):
    """(E0,E1,E2) brick of <= 128 pixels minimising wasted MMA rows; ties -> wider E0 (longer TMA runs)."""
    best = None
    for e0 in range(1, min(W, 128) + 1):
        for e1 in range(1, min(H, 128 // e0) + 1):
            e2 = max(1, min(B, 128 // (e0 * e1)))
            tiles = -(-W // e0) * -(-H // e1) * -(-B // e2)
            key = (tiles, -e0, -e1)
            if best is None or key < best[0]:
                best = (key, (e0, e1, e2))
    return best[1]


def pick_block_n(N, out_f32=False, split=False, m_tiles=2):
    """Tile width minimising padded columns, weighted by the relative MMA efficiency of each width.
    Split (fp16 hi/lo) mode runs 256-wide tiles only as CTA pairs, which need >= 2 m-tiles."""
    cands = ((256, 1.0), (128, 1.15), (64, 1.5))
    if split:
        # 64- and 128-wide tiles only; measured (tools/split_sweep.py): 128-wide CTA-pair tiles with four TMEM accumulator
        # buffers beat 256-wide ones on every ConvNeXt / neck shape (a tile's epilogue hides behind four block sums, not two)
        cands = ((128, 1.0), (64, 1.6))
    best = None
    for bn, pen in cands:
        cost = -(-N // bn) * bn * pen
        if best is None or cost < best[0]:
            best = (cost, bn)
    return best[1]


# fp16 hi/lo mode: 64-wide k-blocks (1 or 2) a TMEM accumulator lives for before its partial sum moves to fp32 registers: the
# tensor pipe truncates on every accumulate (gemm_split.cu).  Two-block accumulators halve the TMEM -> register traffic that
# bounds the kernel (+8 % images/s) at the same distance from a float64 run of the reference (tools/trunc_probe.py,
# profiles/e2e_*_r02*.json); an odd number of k-blocks ends with a one-block sum (its own correction).
SPLIT_LBLK = int(_os.environ.get("WD_SPLIT_LBLK", "2"))

# activations for which the fast path must use the exact formula (debug / accuracy studies): subset of {ACT_SILU, ACT_GELU}
NO_WARP_STORE = 1 if _os.environ.get("WD_NO_WARP_STORE") == "1" else 0   # A/B switch: warpgroup-wide epilogue stores
NO_RED_STORE = 1 if _os.environ.get("WD_NO_RED_STORE") == "1" else 0     # A/B switch: in-place fp32 residual through loads instead of the TMA add
EXACT_ACT = {dict(silu=L.ACT_SILU, gelu=L.ACT_GELU)[a] for a in _os.environ.get("WD_EXACT_ACT", "").split(",") if a in ("silu", "gelu")}


def gemm_raw(*, A, W, C, dims, tile, Kc, N, a_strides, ldb, c_strides, ntaps=1, block_n=None, out_f32=False,
             act=L.ACT_NONE, bias=None, gamma=None, resid=None, ld_res=0, alpha=1.0, group_cols=None, n_groups=1,
             c_gstride=0, dfl=False, a_ps=0, w_ps=0, c_ps=0, r_ps=0, group_valid=0, k_valid=0, bk_valid=0, in_wh=None,
             acc_scale=1.0, lblk=None):
    op = WdOp()
    op.kind = L.OP_GEMM
    split = bool(a_ps) and bool(w_ps)
    assert bool(a_ps) == bool(w_ps), "fp16 hi/lo mode needs two-plane A and W"
    if block_n is None:
        m_tiles = -(-dims[0] // tile[0]) * -(-dims[1] // tile[1]) * -(-dims[2] // tile[2])
        block_n = 64 if dfl else pick_block_n(N, out_f32, split, m_tiles)
    I = op.i
    I[0], I[1], I[2] = dims
    I[3], I[4], I[5] = tile
    I[6], I[7], I[8] = Kc, ntaps, N
    I[9], I[10], I[11] = a_strides
    I[12], I[13], I[14], I[15] = ldb, block_n, 1 if out_f32 else 0, act
    if resid is not None:
        I[16] = 2 if resid.dtype == torch.float32 else 1
        I[17] = ld_res
    I[18] = group_cols if group_cols is not None else N
    I[19] = n_groups
    I[20], I[21], I[22] = c_strides
    I[23] = c_gstride
    I[24] = 1 if dfl else 0
    I[25], I[26] = (3, 1) if ntaps == 9 else (1, 0)
    I[27], I[28], I[29] = group_valid, k_valid, bk_valid
    I[35] = 1 if act in EXACT_ACT else 0
    I[37] = NO_WARP_STORE
    I[42] = NO_RED_STORE
    if in_wh is not None:
        I[38], I[39] = in_wh
    I[30] = 2 if split else 1
    I[31], I[32], I[33], I[34] = a_ps, w_ps, c_ps, r_ps
    I[40] = SPLIT_LBLK if lblk is None else lblk
    op.f[0] = alpha
    op.f[1] = acc_scale
    for k, t in enumerate((A, W, C, bias, gamma, resid)):
        op.p[k] = _ptr(t)
    return op


def _flat_strides(ld, rows):
    big = max(8, ((ld * max(rows, 1) + 7) // 8) * 8)
    return (ld, big, big)


def linear(A, W, C, *, bias=None, gamma=None, resid=None, alpha=1.0, act=L.ACT_NONE, block_n=None, dfl=False):
    """C[M, N] = epi(A[M, K] @ W[N, K]^T); A / C / resid may be column slices of wider row-major buffers.
    A, W (and bf16 C / resid) may be P3 three-plane tensors (precise mode)."""
    acc_scale = 1.0 / (_sc(A) * _sc(W))
    assert _sc(C) in (1.0, ACT_SCALE) and _sc(resid) in (1.0, ACT_SCALE), "fp16 hi/lo outputs / residuals live at ACT_SCALE"
    (A, a_ps), (W, w_ps), (C, c_ps), (resid, r_ps) = _tp(A), _tp(W), _tp(C), _tp(resid)
    _chk16(A, a_ps, "A"); _chk16(W, w_ps, "W")
    M, K = A.shape
    N = W.shape[0]
    assert W.shape[1] == K and K % 8 == 0, (W.shape, K)
    Kc = (K + 63) // 64 * 64
    out_f32 = C.dtype == torch.float32
    if dfl:
        assert C.shape == (M, 4) and C.is_contiguous() and out_f32
        c_str = (4, 8, 8)
    else:
        assert C.shape[0] == M and C.shape[1] == N and C.stride(1) == 1, (C.shape, M, N)
        c_str = _flat_strides(C.stride(0), M)
    ld_res = 0
    if resid is not None:
        assert resid.shape == (M, N) and resid.stride(1) == 1
        ld_res = resid.stride(0)
    return gemm_raw(A=A, W=W, C=C, dims=(M, 1, 1), tile=(128, 1, 1), Kc=Kc, N=N, a_strides=_flat_strides(A.stride(0), M),
                    ldb=W.stride(0), c_strides=c_str, block_n=block_n, out_f32=out_f32, act=act, bias=bias, gamma=gamma,
                    resid=resid, ld_res=ld_res, alpha=alpha, dfl=dfl, a_ps=a_ps, w_ps=w_ps, c_ps=c_ps, r_ps=r_ps,
                    k_valid=K, bk_valid=K, acc_scale=acc_scale)


def conv3x3(A, W, C, *, bias=None, resid=None, alpha=1.0, act=L.ACT_NONE, block_n=None, stride=1):
    """3x3 pad-1 convolution (stride 1 or 2) as 9 shifted TMA brick loads.  A [B,H,W,Cin], W [N, 9*pad64(Cin)] (tap-major).
    stride 2: the A tensor map walks the input with element strides 2 (no im2col); C is [B, ceil(H/2), ceil(W/2), N]."""
    acc_scale = 1.0 / (_sc(A) * _sc(W))
    assert _sc(C) in (1.0, ACT_SCALE) and _sc(resid) in (1.0, ACT_SCALE), "fp16 hi/lo outputs / residuals live at ACT_SCALE"
    (A, a_ps), (W, w_ps), (C, c_ps), (resid, r_ps) = _tp(A), _tp(W), _tp(C), _tp(resid)
    _chk16(A, a_ps, "A"); _chk16(W, w_ps, "W")
    B, H, Wd, Cin = A.shape
    N = W.shape[0]
    Kc = (Cin + 63) // 64 * 64
    assert stride in (1, 2)
    Ho, Wo = (H - 1) // stride + 1, (Wd - 1) // stride + 1
    assert W.shape[1] == 9 * Kc and Cin % 8 == 0, "W must be [N, 9*pad64(Cin)] (zero padded per tap)"
    assert C.shape == (B, Ho, Wo, N) and C.stride(3) == 1
    ld_res = 0
    if resid is not None:
        assert resid.shape == (B, Ho, Wo, N) and resid.stride(3) == 1
        ld_res = resid.stride(2)
        assert resid.stride(1) == Wo * ld_res and resid.stride(0) == Ho * Wo * ld_res
    return gemm_raw(A=A, W=W, C=C, dims=(Wo, Ho, B), tile=pick_tile(Wo, Ho, B), Kc=Kc, k_valid=Cin, ntaps=9, N=N,
                    a_strides=(A.stride(2), A.stride(1), A.stride(0)), ldb=W.stride(0),
                    c_strides=(C.stride(2), C.stride(1), C.stride(0)), block_n=block_n, out_f32=C.dtype == torch.float32,
                    act=act, bias=bias, resid=resid, ld_res=ld_res, alpha=alpha, a_ps=a_ps, w_ps=w_ps, c_ps=c_ps, r_ps=r_ps,
                    in_wh=(Wd, H) if stride == 2 else None, acc_scale=acc_scale)


def deconv2x2(A, W, C, bias2):
    """ConvTranspose2d(k=2, s=2) as two GEMMs (dy = 0, 1) whose TMA stores scatter into the 2x upsampled map.
    A [B,H,W,Cin]; W [4*Co, Cin] rows ordered (dy, dx, co); C [B,2H,2W,Co] (may be a channel slice);
    bias2 f32 [2*Cg] = bias repeated for dx = 0, 1."""
    acc_scale = 1.0 / (_sc(A) * _sc(W))
    (A, a_ps), (W, w_ps), (C, c_ps) = _tp(A), _tp(W), _tp(C)
    _chk16(A, a_ps, "A"); _chk16(W, w_ps, "W")
    B, H, Wd, Cin = A.shape
    Co = C.shape[3]
    Cg = W.shape[0] // 4          # rows per (dy, dx) group, = pad64(Co) with zero rows beyond Co
    assert C.shape == (B, 2 * H, 2 * Wd, Co) and Cg % 64 == 0 and Cg >= Co and W.shape[1] == Cin and Cin % 8 == 0
    assert bias2.shape == (2 * Cg,)
    Kc = (Cin + 63) // 64 * 64
    ops = []
    for dy in range(2):
        Wdy = W[dy * 2 * Cg:(dy + 1) * 2 * Cg]
        Cdy = C[:, dy]
        ops.append(gemm_raw(A=A, W=Wdy, C=Cdy, dims=(Wd, H, B), tile=pick_tile(Wd, H, B), Kc=Kc, k_valid=Cin, bk_valid=Cin,
                            N=2 * Cg, a_strides=(A.stride(2), A.stride(1), A.stride(0)), ldb=W.stride(0),
                            c_strides=(2 * C.stride(2), 2 * C.stride(1), C.stride(0)), group_cols=Cg, group_valid=Co,
                            n_groups=2, c_gstride=C.stride(2), bias=bias2, out_f32=False, a_ps=a_ps, w_ps=w_ps, c_ps=c_ps,
                            acc_scale=acc_scale))
    return ops


MLP_FUSED = _os.environ.get("WD_NO_FUSED_MLP") != "1"   # A/B switch: ConvNeXt stage-0 MLP as one kernel


def mlp_fused_ok(t, W1, W2, x):
    """The fused block-MLP kernel covers single-plane bf16 operands with C = 128, hidden = 512 (WeDetect-Base stage 0)."""
    return (MLP_FUSED and not t.ps and not W1.ps and not W2.ps and tuple(W1.t.shape) == (512, 128) and tuple(W2.t.shape) == (128, 512)
            and x.shape[1] == 128 and x.is_contiguous() and W1.t.is_contiguous() and W2.t.is_contiguous())


def mlp_fused(t, W1, W2, b1, b2, gamma, x):
    """x += gamma * (W2 . GELU(W1 . t + b1) + b2) in one kernel (hidden activation stays on chip)."""
    (t, _), (W1, _), (W2, _) = _tp(t), _tp(W1), _tp(W2)
    _chk(t, torch.bfloat16, "t"); _chk(W1, torch.bfloat16, "W1"); _chk(W2, torch.bfloat16, "W2"); _chk(x, torch.float32, "x")
    M, C = t.shape
    H = W1.shape[0]
    assert W1.shape == (H, C) and W2.shape == (C, H) and x.shape == (M, C) and x.is_contiguous()
    assert b1.numel() == H and b2.numel() == C and gamma.numel() == C
    op = WdOp()
    op.kind = L.OP_MLP_FUSED
    op.i[0], op.i[1], op.i[2], op.i[3] = M, C, H, t.stride(0)
    for k, v in enumerate((t, W1, W2, b1, b2, gamma, x)):
        op.p[k] = _ptr(v)
    return op


def ln_rows(x, w, b, eps, *, out_bf16=None, out_f32=None, s2d_hw=None):
    out_bf16, o_ps = _tp(out_bf16)
    _chk(x, torch.float32, "x")
    rows, C = x.shape
    op = WdOp()
    op.kind = L.OP_LN_ROWS
    op.i[0], op.i[1] = rows, C
    op.i[7] = x.stride(0)
    if s2d_hw is not None:
        H, W = s2d_hw
        op.i[4], op.i[5], op.i[6] = 1, W, H
        assert out_bf16 is not None and out_bf16.shape == (rows // 4, 4 * C)
    op.i[8] = out_bf16.stride(0) if out_bf16 is not None else C
    op.i[30] = o_ps
    op.f[0] = eps
    if out_f32 is not None:
        assert out_f32.is_contiguous() and out_f32.shape == (rows, C)
    for k, t in ((0, x), (1, out_bf16), (2, w), (3, b), (6, out_f32)):
        op.p[k] = _ptr(t)
    return op


@functools.lru_cache(maxsize=None)
def pick_dw_tile(W, H):
    """(tx, ty): one work item of the persistent depthwise kernel = 8tx x 4ty output pixels x 32 channels, computed by
    tx*ty <= 16 thread tiles (8 consumer warps); two halo slots of (8tx+6)(4ty+6) x 128 B + weights must fit in shared memory.
    Every item costs the same time, so the pick maximises useful outputs per item (then prefers the smaller halo)."""
    best = None
    for tx in range(1, 17):
        for ty in range(1, 17):
            slot = (8 * tx + 6) * (4 * ty + 6) * 128 + 6400
            if tx * ty > 16 or 2 * slot + 192 > 227 * 1024:
                continue
            items = -(-W // (8 * tx)) * -(-H // (4 * ty))
            key = (items, (8 * tx + 6) * (4 * ty + 6))
            if best is None or key < best[0]:
                best = (key, (tx, ty))
    return best[1]


def dwconv_ln(x, out, w49, bias, ln_w, ln_b, eps, scratch=None):
    """scratch: fp32 tensor with >= B*H*W*C elements -> shared-memory tiled kernel; None -> register-only kernel."""
    out, o_ps = _tp(out)
    _chk(x, torch.float32, "x")
    B, H, W, C = x.shape
    assert x.is_contiguous() and out.shape == (B * H * W, C) and out.stride(1) == 1 and w49.shape == (49, C)
    op = WdOp()
    op.kind = L.OP_DWCONV_LN
    op.i[0], op.i[1], op.i[2], op.i[3], op.i[4] = B, H, W, C, out.stride(0)
    op.i[30] = o_ps
    op.f[0] = eps
    for k, t in enumerate((x, out, w49, bias, ln_w, ln_b)):
        op.p[k] = _ptr(t)
    if scratch is not None:
        assert scratch.dtype == torch.float32 and scratch.numel() >= x.numel() and scratch.is_contiguous()
        op.i[5], op.i[6] = pick_dw_tile(W, H)
        op.p[7] = _ptr(scratch)
    return op


def stem_patch(img, out, scale=1.0):
    out, o_ps = _tp(out)
    B, C3, H, W = img.shape
    assert C3 == 3 and img.is_contiguous() and out.shape == (B * (H // 4) * (W // 4), 64) and out.is_contiguous()
    op = WdOp()
    op.kind = L.OP_STEM_PATCH
    op.i[0], op.i[1], op.i[2] = B, H, W
    op.i[3] = 0 if img.dtype == torch.uint8 else 2
    assert img.dtype in (torch.uint8, torch.float32)
    op.f[0] = scale
    op.i[30] = o_ps
    for k, t in enumerate((img, out)):
        op.p[k] = _ptr(t)
    return op


def im2col_s2(x, out):
    (x, x_ps), (out, o_ps) = _tp(x), _tp(out)
    _chk16(x, x_ps, "x")
    B, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    assert out.shape == (B * Ho * Wo, 9 * C) and out.is_contiguous()
    assert x.stride(1) == W * x.stride(2) and x.stride(0) == H * W * x.stride(2)
    op = WdOp()
    op.kind = L.OP_IM2COL_S2
    op.i[0], op.i[1], op.i[2], op.i[3], op.i[4] = B, H, W, C, x.stride(2)
    op.i[30], op.i[31] = o_ps, x_ps
    for k, t in enumerate((x, out)):
        op.p[k] = _ptr(t)
    return op


def cast_bf16(x, out):
    out, o_ps = _tp(out)
    _chk(x, torch.float32, "x")
    rows, C = x.shape
    op = WdOp()
    op.kind = L.OP_CAST_BF16
    op.i[0], op.i[1], op.i[2], op.i[3] = rows, C, x.stride(0), out.stride(0)
    op.i[30] = o_ps
    for k, t in enumerate((x, out)):
        op.p[k] = _ptr(t)
    return op


def text_embed(ids, word, pos, typ, ln_w, ln_b, eps, pad_idx, out_f32, out_bf16):
    out_bf16, o_ps = _tp(out_bf16)
    S, Lt = ids.shape
    assert ids.dtype == torch.int32 and ids.is_contiguous()
    Hd = word.shape[1]
    op = WdOp()
    op.kind = L.OP_TEXT_EMBED
    op.i[0], op.i[1], op.i[2], op.i[3] = S, Lt, Hd, pad_idx
    op.f[0] = eps
    op.i[30] = o_ps
    for k, t in ((0, ids), (2, word), (3, pos), (4, typ), (5, ln_w), (6, ln_b), (7, out_f32), (8, out_bf16)):
        op.p[k] = _ptr(t)
    return op


def attn_small(qkv, mask, out, heads, scale):
    out, o_ps = _tp(out)
    S, Lt = mask.shape
    assert mask.dtype == torch.int32 and qkv.dtype == torch.float32 and qkv.is_contiguous()
    op = WdOp()
    op.kind = L.OP_ATTN_SMALL
    op.i[0], op.i[1], op.i[2], op.i[3], op.i[4] = S, Lt, heads, 64, qkv.shape[1]
    op.f[0] = scale
    op.i[30] = o_ps
    for k, t in enumerate((qkv, mask, out)):
        op.p[k] = _ptr(t)
    return op


def l2norm_rows(x, out):
    op = WdOp()
    op.kind = L.OP_L2NORM_ROWS
    op.i[0], op.i[1], op.i[2] = x.shape[0], x.shape[1], x.stride(0)
    op.p[0], op.p[1] = _ptr(x), _ptr(out)
    return op


def gather_rows(x, out, S, row_stride):
    out, o_ps = _tp(out)
    op = WdOp()
    op.kind = L.OP_GATHER_ROWS
    op.i[0], op.i[1], op.i[2], op.i[3] = S, x.shape[1], row_stride, x.stride(0)
    op.i[30] = o_ps
    for k, t in enumerate((x, out)):
        op.p[k] = _ptr(t)
    return op


def fold_text(text, bn_g, bn_h, logit_scale, bias, Wout, bout, normalize):
    """Wout (a P3 in parity-grade mode) is written at Wout.scale; the similarity GEMM undoes it through its acc_scale."""
    w_scale = _sc(Wout)
    Wout, w_ps = _tp(Wout)
    K, C = text.shape
    assert text.dtype == torch.float32 and text.is_contiguous() and Wout.shape[1] == C and Wout.is_contiguous()
    op = WdOp()
    op.kind = L.OP_FOLD_TEXT
    op.i[0], op.i[1], op.i[2], op.i[3] = K, C, 1 if normalize else 0, Wout.shape[0]
    op.i[30] = w_ps
    op.f[0] = w_scale
    for k, t in enumerate((text, bn_g, bn_h, logit_scale, bias, Wout, bout)):
        op.p[k] = _ptr(t)
    return op


def gather_embed(embeds, keep_anchor, counts, bn_g, bn_h, out, *, lvl_scale=None, lvl_bias=None, out_scale=None, out_bias=None):
    """out[b, j] = BN(embed row of kept proposal j); optionally the proposal's per-level logit_scale / bias
    (the `scales` / `bias` results of eval_retrieval/extract_embedding.py:1181-1190,1253-1260)."""
    B, max_keep, C = out.shape
    op = WdOp()
    op.kind = L.OP_GATHER_EMBED
    op.i[0], op.i[2], op.i[3], op.i[4] = B, C, max_keep, len(embeds)
    for l, e in enumerate(embeds):
        e, e_ps = _tp(e)
        op.i[5 + l] = e.shape[0] // B
        op.i[30 + l] = e_ps
        op.p[l] = _ptr(e)
    for k, t in ((3, keep_anchor), (4, counts), (5, bn_g), (6, bn_h), (7, out)):
        op.p[k] = _ptr(t)
    if out_scale is not None:
        assert lvl_scale.dtype == lvl_bias.dtype == torch.float32 and lvl_scale.numel() == lvl_bias.numel() == len(embeds)
        assert out_scale.shape == out_bias.shape == (B, max_keep) and out_scale.is_contiguous() and out_bias.is_contiguous()
        for k, t in ((8, lvl_scale), (9, lvl_bias), (10, out_scale), (11, out_bias)):
            op.p[k] = _ptr(t)
    return op


def scale_rows(x, out, *, scale=None, counts=None):
    """out[b*P + j, :] = bf16(x[b, j, :] * exp(scale[b, j])) for j < counts[b], else 0 (retrieval_metric.py:372)."""
    out, o_ps = _tp(out)
    _chk(x, torch.float32, "x")
    B, Pn, C = x.shape
    assert x.is_contiguous() and out.shape == (B * Pn, C) and out.is_contiguous()
    op = WdOp()
    op.kind = L.OP_SCALE_ROWS
    op.i[0], op.i[1], op.i[2] = B, Pn, C
    op.i[30] = o_ps
    if scale is not None:
        assert scale.dtype == torch.float32 and scale.numel() == B * Pn and scale.is_contiguous()
    if counts is not None:
        assert counts.dtype == torch.int32 and counts.numel() == B
    for k, t in enumerate((x, scale, counts, out)):
        op.p[k] = _ptr(t)
    return op


def retr_reduce(z, out, *, P, bias=None, counts=None):
    """out[b, k] = max_{j < counts[b]} sigmoid(z[b*P + j, k] + bias[b, j]) (retrieval_metric.py:372-373)."""
    _chk(z, torch.float32, "z")
    B, K = out.shape
    assert z.shape[0] == B * P and z.shape[1] >= K and out.dtype == torch.float32 and out.is_contiguous()
    op = WdOp()
    op.kind = L.OP_RETR_REDUCE
    op.i[0], op.i[1], op.i[2], op.i[3] = B, P, K, z.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == B * P and bias.is_contiguous()
    if counts is not None:
        assert counts.dtype == torch.int32 and counts.numel() == B
    for k, t in enumerate((z, bias, counts, out)):
        op.p[k] = _ptr(t)
    return op


class PostProcess:
    """Owns the wd_pp_params struct + workspace and produces the WD_OP_POSTPROCESS record."""

    def __init__(self, *, logits, dists, level_hw, strides, K, B, score_thr, nms_pre, iou_thr, max_per_img, nms_mode,
                 img_meta, clamp_wh, tv_numel_thr=20000, multi_label=True):
        import ctypes
        dev = logits[0].device
        A = sum(h * w for h, w in level_hw)
        self.B, self.K, self.A, self.max_per_img = B, K, A, max_per_img
        lib = L.load(require_gpu=False)
        nbytes = int(lib.wd_pp_workspace_bytes(B, A, K, nms_pre))
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-self.workspace.data_ptr()) % 256
        self.boxes = torch.empty(B, max_per_img, 4, dtype=torch.float32, device=dev)
        self.scores = torch.empty(B, max_per_img, dtype=torch.float32, device=dev)
        self.labels = torch.empty(B, max_per_img, dtype=torch.int32, device=dev)
        self.anchors = torch.empty(B, max_per_img, dtype=torch.int32, device=dev)
        self.counts = torch.zeros(B, dtype=torch.int32, device=dev)
        self.img_meta, self.clamp_wh = img_meta, clamp_wh
        assert img_meta.shape == (B, 8) and img_meta.dtype == torch.float32 and clamp_wh.shape == (B, 2)
        p = L.PPParams()
        p.B, p.K, p.nlevels = B, K, len(logits)
        for l, ((h, w), s) in enumerate(zip(level_hw, strides)):
            p.lvl_h[l], p.lvl_w[l], p.lvl_stride[l] = h, w, s
            assert logits[l].dtype == torch.float32 and logits[l].shape[0] == B * h * w and logits[l].stride(1) == 1
            assert dists[l].shape == (B * h * w, 4) and dists[l].is_contiguous() and dists[l].dtype == torch.float32
            p.ld_logit[l] = logits[l].stride(0)
            p.logits[l] = logits[l].data_ptr()
            p.dist[l] = dists[l].data_ptr()
        p.score_thr, p.nms_pre, p.iou_thr, p.max_per_img = score_thr, nms_pre, iou_thr, max_per_img
        p.nms_mode, p.tv_numel_thr, p.multi_label = nms_mode, tv_numel_thr, 1 if multi_label else 0
        p.img_meta, p.clamp_wh = img_meta.data_ptr(), clamp_wh.data_ptr()
        p.out_boxes, p.out_scores = self.boxes.data_ptr(), self.scores.data_ptr()
        p.out_labels, p.out_anchor, p.out_counts = self.labels.data_ptr(), self.anchors.data_ptr(), self.counts.data_ptr()
        p.workspace, p.workspace_bytes = self.workspace.data_ptr() + off, nbytes
        self.params = p
        self._keep = (logits, dists)
        op = WdOp()
        op.kind = L.OP_POSTPROCESS
        op.p[0] = ctypes.addressof(p)
        self.op = op
