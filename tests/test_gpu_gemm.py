"""GPU parity of the tcgen05 GEMM / implicit-conv op (through the C ABI) vs the fp32 torch reference."""
import pytest
import torch

import ref_ops as R
from util import report_close

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _run(op_or_ops):
    from wedetect_b200 import _lib as L
    ops_ = op_or_ops if isinstance(op_or_ops, (list, tuple)) else [op_or_ops]
    for op in ops_:
        L.run_op(op, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


def _rand_bf16(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,K,N,block_n", [
    (128, 64, 64, 64),       # single tile, single k-iteration
    (256, 128, 128, 128),
    (1000, 512, 384, None),  # M tail, 3 n-tiles of 128
    (300, 192, 80, None),    # N not a multiple of the chunk/tile (80), K = 3 iterations
    (128 * 400, 128, 512, 256),  # many tiles per CTA: barrier phases wrap, TMEM double buffering
    (640, 2048, 256, 256),   # long K loop
    (128 * 301, 512, 512, 256),  # pair mode with an odd number of m-tiles (ghost tile) and 2 n-tiles
    (128 * 3, 64, 256, 256),     # pair mode, 3 m-tiles, single k-iteration
    (128 * 700, 64, 64, 64),     # BN=64 bf16: one column chunk -> the second epilogue warpgroup idles; many tiles per CTA
    (128 * 500, 128, 128, 128),  # BN=128, many tiles per CTA
    (500, 96, 384, None),    # K = 96: second k-chunk half zero-filled by TMA
    (260, 864, 96, None),    # im2col of a 96-channel 3x3 s2 conv (K = 9*96)
])
def test_linear_plain(M, K, N, block_n):
    from wedetect_b200 import ops
    A, W = _rand_bf16(M, K, seed=1), _rand_bf16(N, K, seed=2, scale=0.05)
    C = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=_dev())
    _run(ops.linear(A.to(_dev()), W.to(_dev()), C, block_n=block_n))
    ref = R.linear_ref(A, W)
    report_close(f"linear {M}x{K}x{N}", C, ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_linear_epilogue_f32(act):
    from wedetect_b200 import ops
    M, K, N = 777, 256, 256
    A, W = _rand_bf16(M, K, seed=3), _rand_bf16(N, K, seed=4, scale=0.06)
    g = torch.Generator().manual_seed(5)
    bias, gamma = torch.randn(N, generator=g), torch.rand(N, generator=g) + 0.5
    resid = torch.randn(M, N, generator=g)
    C = torch.zeros(M, N, dtype=torch.float32, device=_dev())
    d = _dev()
    _run(ops.linear(A.to(d), W.to(d), C, bias=bias.to(d), gamma=gamma.to(d), resid=resid.to(d), alpha=0.75, act=act))
    ref = R.linear_ref(A, W, bias=bias, act=act, gamma=gamma, resid=resid, alpha=0.75)
    report_close(f"linear epilogue act={act}", C, ref, rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("N,block_n", [(64, 64), (128, 128), (32, 64)])
def test_linear_f32_many_tiles(N, block_n):
    """fp32 output, every tile width, several tiles per CTA (accumulator hand-back of both epilogue warpgroups)."""
    from wedetect_b200 import ops
    M, K = 128 * 450, 64
    A, W = _rand_bf16(M, K, seed=31), _rand_bf16(N, K, seed=32, scale=0.1)
    d = _dev()
    C = torch.zeros(M, N, dtype=torch.float32, device=d)
    _run(ops.linear(A.to(d), W.to(d), C, block_n=block_n))
    report_close(f"linear f32 many tiles N={N}", C, R.linear_ref(A, W), rtol=1e-4, atol=1e-4)


def test_linear_inplace_residual_f32():
    """ConvNeXt pwconv2: out = x + gamma * (h @ W^T + b), written in place over x."""
    from wedetect_b200 import ops
    M, K, N = 1500, 512, 128
    A, W = _rand_bf16(M, K, seed=6), _rand_bf16(N, K, seed=7, scale=0.04)
    g = torch.Generator().manual_seed(8)
    bias, gamma = torch.randn(N, generator=g), torch.rand(N, generator=g) * 0.45 + 0.05
    x = torch.randn(M, N, generator=g)
    d = _dev()
    xd = x.to(d)
    _run(ops.linear(A.to(d), W.to(d), xd, bias=bias.to(d), gamma=gamma.to(d), resid=xd, alpha=1.0))
    ref = R.linear_ref(A, W, bias=bias, gamma=gamma, resid=x, alpha=1.0)
    report_close("linear in-place residual", xd, ref, rtol=1e-4, atol=2e-4)


def test_linear_slice_output_bf16_resid():
    """Output into a channel slice of a wider buffer (concat fusion) + bf16 residual."""
    from wedetect_b200 import ops
    M, K, N = 900, 128, 128
    A, W = _rand_bf16(M, K, seed=9), _rand_bf16(N, K, seed=10, scale=0.08)
    resid = _rand_bf16(M, N, seed=11)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(12))
    d = _dev()
    buf = torch.full((M, 384), 3.0, dtype=torch.bfloat16, device=d)
    C = buf[:, 128:256]
    _run(ops.linear(A.to(d), W.to(d), C, bias=bias.to(d), resid=resid.to(d), alpha=1.25, act=2))
    ref = R.linear_ref(A, W, bias=bias, act=2, resid=resid, alpha=1.25)
    report_close("linear slice", C, ref, rtol=1e-2, atol=1e-2)
    assert float((buf[:, :128].float() - 3).abs().max()) == 0 and float((buf[:, 256:].float() - 3).abs().max()) == 0


def _pad_taps(Wt, Cin):
    """[N, 9*Cin] -> [N, 9*pad64(Cin)] zero padded per tap (the layout conv3x3 expects)."""
    Kc = (Cin + 63) // 64 * 64
    out = torch.zeros(Wt.shape[0], 9, Kc, dtype=Wt.dtype)
    out[:, :, :Cin] = Wt.view(Wt.shape[0], 9, Cin)
    return out.view(Wt.shape[0], 9 * Kc)


@pytest.mark.parametrize("B,H,W,Cin,N", [(2, 20, 20, 64, 64), (3, 40, 40, 128, 128), (1, 80, 80, 64, 256), (5, 25, 25, 64, 64), (2, 20, 20, 96, 96), (1, 40, 40, 48, 48)])
def test_conv3x3(B, H, W, Cin, N):
    from wedetect_b200 import ops
    A, Wt = _rand_bf16(B, H, W, Cin, seed=13), _rand_bf16(N, 9 * Cin, seed=14, scale=0.04)
    g = torch.Generator().manual_seed(15)
    bias = torch.randn(N, generator=g)
    resid = _rand_bf16(B, H, W, N, seed=16) if Cin == N else None
    d = _dev()
    C = torch.zeros(B, H, W, N, dtype=torch.bfloat16, device=d)
    _run(ops.conv3x3(A.to(d), _pad_taps(Wt, Cin).to(d), C, bias=bias.to(d), act=2, resid=None if resid is None else resid.to(d), alpha=0.9))
    ref = R.conv3x3_ref(A, Wt, bias=bias, act=2, resid=resid, alpha=0.9)
    report_close(f"conv3x3 {B}x{H}x{W}x{Cin}->{N}", C.reshape(-1, N), ref.reshape(-1, N), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("B,H,W,Cin,N", [(2, 40, 40, 128, 128), (1, 37, 29, 96, 64), (3, 80, 80, 256, 256), (2, 20, 20, 64, 192)])
def test_conv3x3_stride2_into_slice(B, H, W, Cin, N):
    """3x3 / stride 2 / pad 1 as an implicit GEMM over a TMA map with element strides 2 (no im2col); odd sizes exercise the
    zero-filled right / bottom halo, the output is a channel slice of a wider buffer (the neck's concat fusion)."""
    from wedetect_b200 import ops
    A, Wt = _rand_bf16(B, H, W, Cin, seed=33), _rand_bf16(N, 9 * Cin, seed=34, scale=0.04)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(35))
    d = _dev()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    buf = torch.full((B, Ho, Wo, N + 64), 3.0, dtype=torch.bfloat16, device=d)
    C = buf[..., 64:]
    _run(ops.conv3x3(A.to(d), _pad_taps(Wt, Cin).to(d), C, bias=bias.to(d), act=1, stride=2))
    w = Wt.float().view(N, 3, 3, Cin).permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(A.float().permute(0, 3, 1, 2), w, bias=bias, stride=2, padding=1).relu().permute(0, 2, 3, 1)
    report_close(f"conv3x3/s2 {B}x{H}x{W}x{Cin}->{N}", C.reshape(-1, N), ref.reshape(-1, N), rtol=1e-2, atol=1e-2)
    assert float((buf[..., :64].float() - 3).abs().max()) == 0


@pytest.mark.parametrize("Cin,Co", [(128, 128), (96, 96), (192, 192)])
def test_deconv2x2_into_slice(Cin, Co):
    from wedetect_b200 import ops
    B, H, W = 2, 20, 20
    A, Wt = _rand_bf16(B, H, W, Cin, seed=17), _rand_bf16(4 * Co, Cin, seed=18, scale=0.06)
    bias = torch.randn(Co, generator=torch.Generator().manual_seed(19))
    d = _dev()
    buf = torch.full((B, 2 * H, 2 * W, 3 * Co), 5.0, dtype=torch.bfloat16, device=d)
    C = buf[..., :Co]
    Cg = (Co + 63) // 64 * 64
    Wp = torch.zeros(4, Cg, Cin, dtype=torch.bfloat16)
    Wp[:, :Co] = Wt.view(4, Co, Cin)
    bp = torch.zeros(2, Cg)
    bp[:, :Co] = bias
    _run(ops.deconv2x2(A.to(d), Wp.view(4 * Cg, Cin).to(d), C, bp.view(-1).to(d)))
    ref = R.deconv2x2_ref(A, Wt, bias)
    report_close("deconv2x2", C.reshape(-1, Co), ref.reshape(-1, Co), rtol=1e-2, atol=1e-2)
    assert float((buf[..., Co:].float() - 5).abs().max()) == 0


def test_dfl_epilogue():
    from wedetect_b200 import ops
    M, K = 1111, 64
    A, Wt = _rand_bf16(M, K, seed=20), _rand_bf16(64, K, seed=21, scale=0.3)
    bias = torch.randn(64, generator=torch.Generator().manual_seed(22))
    d = _dev()
    C = torch.zeros(M, 4, dtype=torch.float32, device=d)
    _run(ops.linear(A.to(d), Wt.to(d), C, bias=bias.to(d), dfl=True))
    report_close("dfl", C, R.dfl_ref(A, Wt, bias), rtol=1e-4, atol=1e-4)


def test_dfl_epilogue_split():
    """DFL (softmax over 16 bins x 4 sides, expectation) in the fp16 hi/lo kernel: K = 64 and a long-K case (several blocks)."""
    from wedetect_b200 import ops
    from wedetect_b200.ops import P3
    d = _dev()
    for M, K in ((1111, 64), (300, 576)):
        g = torch.Generator().manual_seed(26)
        A, Wt, bias = torch.randn(M, K, generator=g), torch.randn(64, K, generator=g) * 0.3 / (K / 64) ** 0.5, torch.randn(64, generator=g)
        C = torch.zeros(M, 4, dtype=torch.float32, device=d)
        _run(ops.linear(P3.from_f32(A, d, scale=ops.ACT_SCALE), P3.from_f32(Wt, d), C, bias=bias.to(d), dfl=True))
        z = (A.double() @ Wt.double().t() + bias.double()).view(-1, 4, 16).softmax(-1)
        ref = (z * torch.arange(16, dtype=torch.float64)).sum(-1).float()
        report_close(f"dfl split K={K}", C, ref, rtol=2e-6, atol=5e-6)


SPLIT_SHAPES = [
    (515, 256, 192, None),        # 5 m-tiles: pairs + ghost tile, BN=256 (N tail 192)
    (128, 192, 80, None),         # one m-tile: non-pair BN=128, ragged N, K tail
    (100, 64, 64, None),          # BN=64 (single active epilogue warpgroup)
    (128 * 41, 512, 512, None),   # pair BN=256, many tiles per CTA pair (barrier phases wrap), odd m-tiles
    (128 * 300, 128, 128, None),  # pair BN=128, many tiles per CTA
    (640, 2048, 512, None),       # long K: 32 register-accumulated blocks
    (128 * 9, 4096, 1024, None),  # K = 4096 (ConvNeXt stage-3 pw2 shape)
    (1000, 768, 1208, None),      # similarity-like: N tail in the last 256 tile
    (384, 320, 256, 128),         # forced BN=128 pair with 2 n-tiles
    (128 * 300, 192, 128, None),  # 3 k-blocks per tile, many tiles per CTA: two-stage + one-stage blocks that wrap around the ring
    (128 * 150, 576, 64, None),   # 9 k-blocks, BN=64 single CTA (the 64 -> 64 3x3 convolutions' K)
]


@pytest.mark.parametrize("lblk", [1, 2])
@pytest.mark.parametrize("M,K,N,block_n", SPLIT_SHAPES)
def test_linear_split_shapes(M, K, N, block_n, lblk):
    """fp16 hi/lo mode against an fp64 GEMM: fp32-grade results for every tile configuration / accumulator block length."""
    from wedetect_b200 import ops
    from wedetect_b200.ops import P3
    g = torch.Generator().manual_seed(31)
    A, Wt = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    bias = torch.randn(N, generator=g)
    d = _dev()
    C = torch.full((M, N), 7.0, dtype=torch.float32, device=d)
    op = ops.linear(P3.from_f32(A, d, scale=ops.ACT_SCALE), P3.from_f32(Wt, d), C, bias=bias.to(d), block_n=block_n)
    op.i[40] = lblk
    _run(op)
    ref = A.double() @ Wt.double().t() + bias.double()
    err = (C.cpu().double() - ref)
    rel = float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    print(f"split linear {M}x{K}x{N} lblk={lblk}: rel-rms {rel:.3e} max-abs {float(err.abs().max()):.3e} mean (bias) {float(err.mean()):.3e}")
    tol = 3e-6
    report_close(f"split linear {M}x{K}x{N}", C, ref.float(), rtol=tol, atol=tol * float(ref.abs().mean()) * 4)


def test_linear_split_precise():
    """fp16 hi/lo mode reproduces an fp32 GEMM to fp32-level accuracy (activation, bias, fp16 hi/lo output and residual)."""
    from wedetect_b200 import ops
    from wedetect_b200.ops import P3
    M, K, N = 515, 256, 192
    g = torch.Generator().manual_seed(23)
    A, Wt = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    bias = torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g)
    d = _dev()
    Ap, Wp, Rp = P3.from_f32(A, d, scale=ops.ACT_SCALE), P3.from_f32(Wt, d), P3.from_f32(resid, d, scale=ops.ACT_SCALE)
    C = torch.zeros(M, N, dtype=torch.float32, device=d)
    _run(ops.linear(Ap, Wp, C, bias=bias.to(d), act=3))
    ref = R.act_ref(A.double() @ Wt.double().t() + bias.double(), 3).float()
    report_close("linear precise f32", C, ref, rtol=2e-6, atol=2e-6)
    # fp16 hi/lo output + fp16 hi/lo residual
    Cp = P3.zeros((M, N), d, True)
    _run(ops.linear(Ap, Wp, Cp, bias=bias.to(d), act=2, resid=Rp, alpha=0.5))
    ref2 = (R.act_ref(A.double() @ Wt.double().t() + bias.double(), 2) + 0.5 * resid.double()).float()
    report_close("linear precise hi/lo out", Cp.value(), ref2, rtol=2e-6, atol=2e-6)
    # fp32 residual in place (the ConvNeXt pw2 epilogue): x += gamma * (A W^T + b)
    gamma = torch.rand(N, generator=g) + 0.5
    X = resid.clone().to(d)
    _run(ops.linear(Ap, Wp, X, bias=bias.to(d), gamma=gamma.to(d), resid=X, alpha=1.0))
    ref3 = (resid.double() + gamma.double() * (A.double() @ Wt.double().t() + bias.double())).float()
    report_close("linear precise f32 residual", X, ref3, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("M,K,N", [(515, 256, 192), (4096, 512, 128), (1024, 2048, 512)])
def test_inplace_residual_tma_add_equals_fused_add(M, K, N, monkeypatch):
    """x += gamma * (A W^T + b) in place: the store that adds the tile in the L2 (cp.reduce.async.bulk.tensor .add) and the epilogue
    that loads the residual rows and adds in registers round once per element either way - bit-identical results."""
    from wedetect_b200 import ops
    from wedetect_b200.ops import P3
    g = torch.Generator().manual_seed(29)
    A, Wt = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05
    bias, gamma, x = torch.randn(N, generator=g), torch.rand(N, generator=g) + 0.5, torch.randn(M, N, generator=g) * 3
    d = _dev()
    Ap, Wp = P3.from_f32(A, d, scale=ops.ACT_SCALE), P3.from_f32(Wt, d)
    got = []
    for no_red in (0, 1):
        monkeypatch.setattr(ops, "NO_RED_STORE", no_red)
        X = x.clone().to(d)
        _run(ops.linear(Ap, Wp, X, bias=bias.to(d), gamma=gamma.to(d), resid=X, alpha=1.0))
        got.append(X.cpu())
    assert torch.equal(got[0], got[1])
    ref = (x.double() + gamma.double() * (A.double() @ Wt.double().t() + bias.double())).float()
    report_close("in-place residual", got[0], ref, rtol=2e-6, atol=4e-6)


def test_conv3x3_split_precise():
    from wedetect_b200 import ops
    from wedetect_b200.ops import P3
    B, H, W, Cin, N = 2, 20, 20, 96, 128
    g = torch.Generator().manual_seed(24)
    A, Wt = torch.randn(B, H, W, Cin, generator=g), torch.randn(N, 9 * Cin, generator=g) * 0.04
    d = _dev()
    Cp = P3.zeros((B, H, W, N), d, True)
    _run(ops.conv3x3(P3.from_f32(A, d, scale=ops.ACT_SCALE), P3.from_f32(_pad_taps(Wt, Cin), d), Cp))
    w = Wt.view(N, 3, 3, Cin).permute(0, 3, 1, 2).double()
    ref = torch.nn.functional.conv2d(A.permute(0, 3, 1, 2).double(), w, padding=1).permute(0, 2, 3, 1).float()
    report_close("conv3x3 precise", Cp.value().reshape(-1, N), ref.reshape(-1, N), rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("B,H,W,Cin,N,stride", [(2, 16, 16, 64, 256, 1),    # (16,8,1) bricks: per-warp TMA stores, pair BN=256
                                                 (4, 32, 32, 128, 128, 1),   # pair BN=128, several tiles per CTA pair
                                                 (2, 10, 10, 192, 64, 1),    # 100-row bricks: direct-store fallback, BN=64
                                                 (2, 32, 32, 96, 192, 2),    # stride 2 (TMA element strides), K tail per tap
                                                 (1, 20, 12, 64, 320, 2)])
def test_conv3x3_split_shapes(B, H, W, Cin, N, stride):
    from wedetect_b200 import ops
    from wedetect_b200.ops import P3
    g = torch.Generator().manual_seed(27)
    A, Wt = torch.randn(B, H, W, Cin, generator=g), torch.randn(N, 9 * Cin, generator=g) * 0.04
    bias = torch.randn(N, generator=g) * 0.3
    d = _dev()
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    Cp = P3.zeros((B, Ho, Wo, N), d, True)
    Rp = None
    resid = torch.randn(B, Ho, Wo, N, generator=g)
    if stride == 1:
        Rp = P3.from_f32(resid, d, scale=ops.ACT_SCALE)
    _run(ops.conv3x3(P3.from_f32(A, d, scale=ops.ACT_SCALE), P3.from_f32(_pad_taps(Wt, Cin), d), Cp, bias=bias.to(d), act=2, stride=stride,
                     resid=Rp, alpha=0.75))
    w = Wt.view(N, 3, 3, Cin).permute(0, 3, 1, 2).double()
    y = torch.nn.functional.conv2d(A.permute(0, 3, 1, 2).double(), w, bias=bias.double(), padding=1, stride=stride).permute(0, 2, 3, 1)
    ref = torch.nn.functional.silu(y) + (0.75 * resid.double() if stride == 1 else 0.0)
    report_close("conv3x3 split", Cp.value().reshape(-1, N), ref.float().reshape(-1, N), rtol=3e-6, atol=3e-6)


def test_deconv_and_im2col_precise():
    from wedetect_b200 import ops
    from wedetect_b200.ops import P3
    B, H, W, Cin, Co = 1, 10, 10, 96, 96
    g = torch.Generator().manual_seed(25)
    A, Wt, bias = torch.randn(B, H, W, Cin, generator=g), torch.randn(4 * Co, Cin, generator=g) * 0.06, torch.randn(Co, generator=g)
    d = _dev()
    Cg = 128
    Wp = torch.zeros(4, Cg, Cin)
    Wp[:, :Co] = Wt.view(4, Co, Cin)
    bp = torch.zeros(2, Cg)
    bp[:, :Co] = bias
    buf = P3.zeros((B, 2 * H, 2 * W, 2 * Co), d, True)
    Cs = buf.view(lambda t: t[..., Co:])
    Ap = P3.from_f32(A, d, scale=ops.ACT_SCALE)
    _run(ops.deconv2x2(Ap, P3.from_f32(Wp.view(4 * Cg, Cin), d), Cs, bp.view(-1).to(d)))
    ref = (A.double().reshape(-1, Cin) @ Wt.double().t()).view(B, H, W, 2, 2, Co).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, Co) + bias.double()
    report_close("deconv precise", Cs.value().reshape(-1, Co), ref.float().reshape(-1, Co), rtol=2e-6, atol=2e-6)
    col = P3.zeros((B * 5 * 5, 9 * Cin), d, True)
    _run(ops.im2col_s2(Ap, col))
    report_close("im2col precise", col.value(), R.im2col_s2_ref(A), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("M", [256, 2048, 256 * 75 + 128, 160 * 160 * 4])
def test_mlp_fused_matches_two_gemms(M):
    """WD_OP_MLP_FUSED (ConvNeXt block MLP in one kernel, C = 128 / hidden 512) against the two WD_OP_GEMM records it replaces
    (same bf16 operands, bf16 hidden, fp32 accumulation in the same k order) and against an fp32 torch reference."""
    from wedetect_b200 import _lib as L, ops
    from wedetect_b200.ops import P3
    g = torch.Generator().manual_seed(M)
    C, H = 128, 512
    D = "cuda:0"
    t = (torch.randn(M, C, generator=g)).to(torch.bfloat16).to(D)
    W1 = (torch.randn(H, C, generator=g) / C ** 0.5).to(torch.bfloat16).to(D)
    W2 = (torch.randn(C, H, generator=g) / H ** 0.5).to(torch.bfloat16).to(D)
    b1 = (torch.randn(H, generator=g) * 0.1).to(D)
    b2 = (torch.randn(C, generator=g) * 0.1).to(D)
    gamma = (torch.rand(C, generator=g) * 0.45 + 0.05).to(D)
    x0 = torch.randn(M, C, generator=g).to(D)
    # reference path: the two GEMM ops
    x_ref = x0.clone()
    hid = torch.zeros(M, H, dtype=torch.bfloat16, device=D)
    L.run_op(ops.linear(P3(t), P3(W1), P3(hid), bias=b1, act=L.ACT_GELU))
    L.run_op(ops.linear(P3(hid), P3(W2), x_ref, bias=b2, gamma=gamma, resid=x_ref, alpha=1.0))
    x_fused = x0.clone()
    assert ops.mlp_fused_ok(P3(t), P3(W1), P3(W2), x_fused)
    L.run_op(ops.mlp_fused(P3(t), P3(W1), P3(W2), b1, b2, gamma, x_fused))
    torch.cuda.synchronize()
    report_close("fused vs two GEMMs", x_fused, x_ref, rtol=0.0, atol=1e-5)
    h32 = torch.nn.functional.gelu(t.float() @ W1.float().t() + b1).to(torch.bfloat16).float()
    want = x0 + gamma * (h32 @ W2.float().t() + b2)
    report_close("fused vs torch", x_fused, want, rtol=2e-2, atol=2e-2)
