"""GPU parity of the CUDA-core ops (LN, dwconv+LN, patch gathers, text ops) vs fp32 torch references."""
import pytest
import torch

import ref_ops as R
from util import report_close

pytestmark = pytest.mark.gpu
D = "cuda:0"


def _run(op):
    from wedetect_b200 import _lib as L
    L.run_op(op, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("rows,C", [(1000, 128), (333, 96), (1001, 256), (7, 192), (3, 128), (257, 512), (64, 1536), (100, 768)])
def test_ln_rows(rows, C):
    from wedetect_b200 import ops
    g = _g(1)
    x = torch.randn(rows, C, generator=g) * 3 + 1
    w, b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    from wedetect_b200.ops import P3
    ob = P3.zeros((rows, C), D, True)
    of = torch.zeros(rows, C, dtype=torch.float32, device=D)
    _run(ops.ln_rows(x.to(D), w.to(D), b.to(D), 1e-6, out_bf16=ob, out_f32=of))
    ref = R.ln_ref(x, w, b, 1e-6)
    report_close("ln f32", of, ref, rtol=1e-5, atol=1e-5)
    report_close("ln fp16 hi plane", ob.t.float() / ob.scale, ref, rtol=1e-3, atol=1e-3)
    report_close("ln hi/lo planes", ob.value(), of.cpu(), rtol=2e-7, atol=2e-8)


def test_ln_rows_s2d():
    from wedetect_b200 import ops
    B, H, W, C = 2, 8, 12, 128
    g = _g(2)
    x = torch.randn(B * H * W, C, generator=g)
    w, b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    ob = torch.zeros(B * H * W // 4, 4 * C, dtype=torch.bfloat16, device=D)
    _run(ops.ln_rows(x.to(D), w.to(D), b.to(D), 1e-6, out_bf16=ob, s2d_hw=(H, W)))
    ref = R.s2d_ref(R.ln_ref(x, w, b, 1e-6), B, H, W)
    report_close("ln s2d", ob, ref, rtol=8e-3, atol=1e-3)


@pytest.mark.parametrize("tiled", [False, True])
@pytest.mark.parametrize("B,H,W,C", [(2, 20, 20, 128), (1, 40, 36, 256), (1, 10, 10, 1024), (1, 9, 13, 96), (1, 6, 6, 1536), (2, 160, 160, 128), (3, 40, 40, 512)])
def test_dwconv_ln(B, H, W, C, tiled):
    from wedetect_b200 import ops
    g = _g(3)
    x = torch.randn(B, H, W, C, generator=g)
    w49 = torch.randn(49, C, generator=g) * 0.15
    bias, lw, lb = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    from wedetect_b200.ops import P3
    out = P3.zeros((B * H * W, C), D, True)
    scratch = torch.zeros(x.numel(), dtype=torch.float32, device=D) if tiled else None
    _run(ops.dwconv_ln(x.to(D), out, w49.to(D), bias.to(D), lw.to(D), lb.to(D), 1e-6, scratch=scratch))
    ref = R.dwconv_ln_ref(x, w49, bias, lw, lb, 1e-6)
    report_close("dwconv_ln fp16 hi plane", out.t.float() / out.scale, ref, rtol=1e-3, atol=2e-3)
    report_close("dwconv_ln hi/lo planes", out.value(), ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_stem_patch(dtype):
    from wedetect_b200 import ops
    B, H, W = 2, 32, 48
    g = _g(4)
    img = torch.rand(B, 3, H, W, generator=g)
    if dtype == torch.uint8:
        img = (img * 255).to(torch.uint8)
    out = torch.full((B * (H // 4) * (W // 4), 64), 9.0, dtype=torch.bfloat16, device=D)
    _run(ops.stem_patch(img.to(D), out, scale=1.0))
    report_close("stem_patch", out, R.bf16_round(R.stem_patch_ref(img, 1.0)), rtol=0, atol=0)


@pytest.mark.parametrize("H,W", [(20, 20), (15, 9)])
def test_im2col_s2(H, W):
    from wedetect_b200 import ops
    B, C = 2, 64
    x = (torch.randn(B, H, W, C, generator=_g(5))).to(torch.bfloat16)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.full((B * Ho * Wo, 9 * C), 9.0, dtype=torch.bfloat16, device=D)
    _run(ops.im2col_s2(x.to(D), out))
    report_close("im2col_s2", out, R.im2col_s2_ref(x), rtol=0, atol=0)


def test_cast_bf16():
    from wedetect_b200 import ops
    x = torch.randn(500, 256, generator=_g(6))
    from wedetect_b200.ops import P3
    out = P3.zeros((500, 256), D, True)
    _run(ops.cast_bf16(x.to(D), out))
    report_close("cast hi", out.t.float(), (x * out.scale).to(torch.float16).float(), rtol=0, atol=0)
    report_close("cast hi/lo planes", out.value(), x, rtol=1.2e-7, atol=2e-8)
    # fast mode: one bf16 plane
    ob = torch.zeros(500, 256, dtype=torch.bfloat16, device=D)
    _run(ops.cast_bf16(x.to(D), ob))
    report_close("cast bf16", ob, R.bf16_round(x), rtol=0, atol=0)
    # out-of-range values saturate instead of becoming inf
    big = torch.tensor([[1e6, -1e6, 3.0, 0.0]]).to(D)
    ob2 = P3.zeros((1, 4), D, True)
    _run(ops.cast_bf16(big, ob2))
    assert torch.isfinite(ob2.value()).all() and float(ob2.value()[0, 2]) == 3.0


def test_text_embed_and_attention():
    from wedetect_b200 import ops
    S, L, Hd, heads, V, pad = 7, 9, 768, 12, 1000, 1
    g = _g(7)
    ids = torch.randint(3, V, (S, L), generator=g, dtype=torch.int32)
    ids[:, 0] = 0
    for s in range(S):
        n = 2 + (s % (L - 2))
        ids[s, n:] = pad
    word, pos = torch.randn(V, Hd, generator=g) * 0.1, torch.randn(L + pad + 2, Hd, generator=g) * 0.1
    typ, lw, lb = torch.randn(Hd, generator=g) * 0.1, torch.rand(Hd, generator=g) + 0.5, torch.randn(Hd, generator=g) * 0.1
    of = torch.zeros(S * L, Hd, dtype=torch.float32, device=D)
    ob = torch.zeros(S * L, Hd, dtype=torch.bfloat16, device=D)
    _run(ops.text_embed(ids.to(D), word.to(D), pos.to(D), typ.to(D), lw.to(D), lb.to(D), 1e-5, pad, of, ob))
    ref = R.text_embed_ref(ids, word, pos, typ, lw, lb, 1e-5, pad)
    report_close("text_embed", of, ref, rtol=1e-5, atol=1e-5)
    report_close("text_embed bf16", ob, ref, rtol=8e-3, atol=1e-3)

    qkv = torch.randn(S * L, 3 * Hd, generator=g)
    mask = (ids != pad).int()
    from wedetect_b200.ops import P3
    out = P3.zeros((S * L, Hd), D, True)
    _run(ops.attn_small(qkv.to(D), mask.to(D), out, heads, 0.125))
    report_close("attn", out.value(), R.attn_ref(qkv, mask, heads, 0.125), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("L", [33, 64, 100, 128])
def test_attention_longer_prompts(L):
    """Prompts of more than 32 tokens (the reference pads to the longest prompt, no truncation: mm_backbone.py:382-383)."""
    from wedetect_b200 import ops
    from wedetect_b200.ops import P3
    S, Hd, heads = 3, 768, 12
    g = _g(17)
    qkv = torch.randn(S * L, 3 * Hd, generator=g)
    mask = torch.ones(S, L, dtype=torch.int32)
    mask[1, L // 2:] = 0
    mask[2, 5:] = 0
    out = P3.zeros((S * L, Hd), D, True)
    _run(ops.attn_small(qkv.to(D), mask.to(D), out, heads, 0.125))
    ref = R.attn_ref(qkv, mask, heads, 0.125)
    report_close(f"attn L={L}", out.value(), ref, rtol=1e-5, atol=1e-5)


def test_l2norm_gather_fold():
    from wedetect_b200 import ops
    g = _g(8)
    x = torch.randn(45, 768, generator=g)
    out = torch.zeros(5, 768, dtype=torch.float32, device=D)
    xb = torch.zeros(5, 768, dtype=torch.bfloat16, device=D)
    _run(ops.gather_rows(x.to(D), xb, 5, 9))
    report_close("gather_rows", xb, R.bf16_round(x[::9]), rtol=0, atol=0)
    _run(ops.l2norm_rows(x[::9].contiguous().to(D), out))
    report_close("l2norm", out, torch.nn.functional.normalize(x[::9], dim=-1), rtol=1e-5, atol=1e-6)

    K, C, Kp = 80, 768, 128
    text = torch.randn(K, C, generator=g)
    gg, hh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    ls, bi = torch.tensor([-0.8]), torch.tensor([-2.5])
    for normalize in (True, False):
        from wedetect_b200.ops import P3
        Wd = P3.zeros((Kp, C), D, True)
        Wd.t.fill_(3.0)
        bd = torch.full((Kp,), 3.0, dtype=torch.float32, device=D)
        _run(ops.fold_text(text.to(D), gg.to(D), hh.to(D), ls.to(D), bi.to(D), Wd, bd, normalize))
        Wr, br = R.fold_text_ref(text, gg, hh, ls, bi, normalize)
        report_close("fold W (scale 1: low plane partly subnormal)", Wd.value()[:K], Wr, rtol=2e-6, atol=1e-7)
        Ws = P3.zeros((Kp, C), D, True, scale=ops.weight_scale(Wr))
        _run(ops.fold_text(text.to(D), gg.to(D), hh.to(D), ls.to(D), bi.to(D), Ws, bd, normalize))
        report_close("fold W (matrix scale)", Ws.value()[:K], Wr, rtol=1e-6, atol=1e-9)
        report_close("fold b", bd[:K], br, rtol=1e-4, atol=1e-4)
        assert float(Wd.value()[K:].abs().max()) == 0 and float(bd[K:].abs().max()) == 0
