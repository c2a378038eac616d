"""Host-side lowering logic on the CPU: tile pickers, weight folding algebra, op wiring (no kernels run)."""
import types

import pytest
import torch


def test_pick_tile_and_block_n():
    from wedetect_b200.ops import pick_block_n, pick_tile
    for (W, H, B) in [(80, 80, 32), (40, 40, 32), (20, 20, 32), (100, 100, 16), (25, 25, 16), (20, 20, 1), (160, 160, 1)]:
        e0, e1, e2 = pick_tile(W, H, B)
        assert e0 * e1 * e2 <= 128 and e0 <= W and e1 <= H and e2 <= max(B, 1)
        tiles = -(-W // e0) * -(-H // e1) * -(-B // e2)
        assert W * H * B / (tiles * 128) > 0.6
    assert pick_block_n(2048) == 256 and pick_block_n(384) == 128 and pick_block_n(64) == 64 and pick_block_n(1208) == 256
    assert pick_block_n(512, split=True) == 128 and pick_block_n(512, split=True, m_tiles=1) == 128 and pick_block_n(80, split=True) == 128 and pick_block_n(64, split=True) == 64


def test_bn_fold_and_layouts_match_conv_semantics():
    """Folded + re-laid-out weights reproduce Conv+BN / ConvTranspose / patchify exactly (fp64 algebra)."""
    import torch.nn.functional as F
    from wedetect_b200 import weights as Wp
    g = torch.Generator().manual_seed(0)
    Cin, Cout = 24, 16
    sd = {"c.weight": torch.randn(Cout, Cin, 3, 3, generator=g), "bn.weight": torch.rand(Cout, generator=g) + 0.5, "bn.bias": torch.randn(Cout, generator=g),
          "bn.running_mean": torch.randn(Cout, generator=g), "bn.running_var": torch.rand(Cout, generator=g) + 0.5}
    w, b = Wp._fold_bn(sd, "c.weight", "bn", 1e-3)
    x = torch.randn(2, Cin, 9, 9, generator=g).double()
    ref = F.batch_norm(F.conv2d(x, sd["c.weight"].double(), padding=1), sd["bn.running_mean"].double(), sd["bn.running_var"].double(),
                       sd["bn.weight"].double(), sd["bn.bias"].double(), False, 0.0, 1e-3)
    torch.testing.assert_close(F.conv2d(x, w, b, padding=1), ref, rtol=1e-10, atol=1e-10)
    taps = Wp._taps(w)                                   # [Cout, 9*64]
    Kc = 64
    cols = F.unfold(x, 3, padding=1).view(2, Cin, 9, 81).permute(0, 3, 2, 1)          # [B, L, tap, c]
    colp = torch.zeros(2, 81, 9, Kc, dtype=torch.float64)
    colp[..., :Cin] = cols
    y = colp.reshape(2 * 81, 9 * Kc) @ taps.t() + b
    torch.testing.assert_close(y.view(2, 9, 9, Cout).permute(0, 3, 1, 2), ref, rtol=1e-10, atol=1e-10)


def test_plan_builds_for_all_sizes_on_cpu(monkeypatch):
    import wedetect_b200._lib as L
    from oracle import synth

    class FakeProg:
        def __init__(self, ops_, keepalive=()):
            self.ops, self.num_launches = ops_, len(ops_)

        def run(self, s):
            pass

    monkeypatch.setattr(L, "Program", FakeProg)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: types.SimpleNamespace(cuda_stream=0))
    from wedetect_b200 import plan, weights
    for size, res, K, precise in (("tiny", 320, 5, False), ("base", 320, 80, True), ("base", 320, 80, False), ("large", 160, 16, False)):
        sd = synth.synth_state_dict(size, seed=0, uni=True, num_prompts=K, with_text=False, calibrate=False)
        Wt = weights.prepare_vision(sd, size, "cpu", precise=precise)
        p = plan.VisionPlan(Wt, size, 2, res, res, K=K, uni=True, nms_mode=1, score_thr=0.0, extract=(size == "large"), device="cpu")
        kinds = [op.kind for op in p.ops]
        assert kinds.count(L.OP_DWCONV_LN) == sum({"tiny": (3, 3, 9, 3)}.get(size, (3, 3, 27, 3)))
        assert kinds[-2:] == [L.OP_POSTPROCESS, L.OP_GATHER_EMBED]
        # the fused block-MLP kernel covers exactly the C = 128 stage of the fast path (Base stage 0: 3 blocks, 6 GEMMs fewer)
        assert kinds.count(L.OP_MLP_FUSED) == (3 if (size == "base" and not precise) else 0)
        # extract variant: the gather op also emits per-proposal logit_scale / bias rows
        assert bool(p.ops[-1].p[10]) == (size == "large") and ("scales" in p.results()) == (size == "large")
        for op in p.ops:
            if op.kind == L.OP_GEMM:
                assert op.i[30] == (2 if precise else 1) and op.i[6] % 64 == 0 and op.i[8] % 8 == 0


def test_standalone_text_tower_facade_wires_on_cpu(monkeypatch):
    """XLMRobertaLanguageBackbone (extract_embedding.py:1267-1320 facade): checkpoint slicing, plan per (S, L), un-normalised head
    output vs normalised features - with the library calls stubbed out (no kernels run on the CPU)."""
    import wedetect_b200._lib as L
    from oracle import synth

    class FakeProg:
        def __init__(self, ops_, keepalive=()):
            self.ops = ops_

        def run(self, s):
            pass

    monkeypatch.setattr(L, "Program", FakeProg)
    monkeypatch.setattr(L, "load", lambda require_gpu=True, device=0: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: types.SimpleNamespace(cuda_stream=0))
    from wedetect_b200.detector import XLMRobertaLanguageBackbone
    sd = synth.synth_state_dict("tiny", seed=0, with_text=True, text_vocab=64, calibrate=False)
    tok = lambda text, return_tensors, padding: dict(input_ids=torch.tensor([[0, 5, 6, 2], [0, 7, 2, 1]]), attention_mask=torch.tensor([[1, 1, 1, 1], [1, 1, 1, 0]]))
    m = XLMRobertaLanguageBackbone({"state_dict": sd}, tokenizer=tok, device="cpu").cuda().eval()
    assert m.text == "base" and m.language_dim == 768
    out = m(["a", "b"])
    assert out.shape == (2, 768) and out.dtype == torch.float32
    tp = m._plans[(2, 4)]
    assert out.data_ptr() != tp.head_out.data_ptr()                          # a copy, the plan's buffers are reused per call
    kinds = [op.kind for op in tp.program.ops]
    assert kinds[0] == L.OP_TEXT_EMBED and kinds[-1] == L.OP_L2NORM_ROWS and kinds.count(L.OP_ATTN_SMALL) == 12
    assert m.encode_tokens(torch.zeros(2, 4, dtype=torch.long), torch.ones(2, 4, dtype=torch.long), normalize=True).shape == (2, 768)


def test_hi_lo_planes_and_scales_cpu():
    """Host side of the default operand format: fp16 hi/lo planes at a power-of-two scale reconstruct fp32 to ~2^-24; the GEMM
    record carries 1 / (A scale * W scale) and the accumulator block length."""
    import torch
    from wedetect_b200 import _lib as L, ops
    from wedetect_b200.ops import P3
    g = torch.Generator().manual_seed(0)
    for mag in (1e-3, 0.02, 1.0, 300.0):
        w = torch.randn(64, 128, generator=g) * mag
        p = P3.from_f32(w)
        s = p.scale
        assert s == 2.0 ** round(__import__("math").log2(s)) and 8192 <= float(w.abs().max()) * s < 16384
        assert p.t.dtype == torch.float16 and p.ps == w.numel()
        assert float((p.value() - w).abs().max()) <= float(w.abs().max()) * 2.0 ** -22
    a = P3.from_f32(torch.randn(256, 128, generator=g), scale=ops.ACT_SCALE)
    w = P3.from_f32(torch.randn(64, 128, generator=g) * 0.05)
    c = torch.zeros(256, 64)
    op = ops.linear(a, w, c)
    assert op.i[30] == 2 and op.i[31] == a.ps and op.i[32] == w.ps and op.i[13] == 64 and op.i[40] == ops.SPLIT_LBLK
    assert abs(op.f[1] - 1.0 / (ops.ACT_SCALE * w.scale)) < 1e-12
    big = P3.from_f32(torch.tensor([[1e9, -1e9, 1.0, 0.0]]), scale=4.0)       # out-of-range values saturate, no inf
    assert torch.isfinite(big.value()).all() and float(big.value()[0, 2]) == 1.0
    assert ops.ACT_SCALE == L.ACT_PLANE_SCALE == 4.0
