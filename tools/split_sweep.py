"""GPU micro-benchmark of the fp16 hi/lo GEMM kernel on the ConvNeXt MLP shapes: separates the main loop, the register
accumulation and the epilogue / store cost.  Usage (on the GPU box): python tools/split_sweep.py [out.json] [only-name-substring]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wedetect_b200 import _lib as L, ops
from wedetect_b200._lib import Program
from wedetect_b200.ops import P3


def timed(op, iters=10):
    prog = Program([op])
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        prog.run(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        prog.run(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    L.load(require_gpu=True)
    dev = "cuda:0"
    only = sys.argv[2] if len(sys.argv) > 2 else ""
    res = []
    cases = [
        ("s2 pw1 gelu f16x2", 51200, 512, 2048, "h", L.ACT_GELU, False, True),
        ("s2 pw1 none f16x2", 51200, 512, 2048, "h", L.ACT_NONE, False, False),
        ("s2 pw1 none f32", 51200, 512, 2048, "f", L.ACT_NONE, False, False),
        ("s2 pw2 res f32", 51200, 2048, 512, "f", L.ACT_NONE, True, True),
        ("s2 pw2 none f16x2", 51200, 2048, 512, "h", L.ACT_NONE, False, False),
        ("s2 pw2 none f32", 51200, 2048, 512, "f", L.ACT_NONE, False, False),
        ("s2 pw2 bias f32", 51200, 2048, 512, "f", L.ACT_NONE, False, True),
        ("s2 pw2 gamma-only f32", 51200, 2048, 512, "f", L.ACT_NONE, "gamma", True),
        ("s2 pw2 resid-only f32", 51200, 2048, 512, "f", L.ACT_NONE, "resid", True),
        ("s2 pw2 resid-inplace f32", 51200, 2048, 512, "f", L.ACT_NONE, "inplace", True),
        ("s0 pw1 gelu f16x2", 819200, 128, 512, "h", L.ACT_GELU, False, True),
        ("s0 pw1 none f16x2", 819200, 128, 512, "h", L.ACT_NONE, False, True),
        ("s0 pw1 gelu f32", 819200, 128, 512, "f", L.ACT_GELU, False, True),
        ("s0 pw1 none f32", 819200, 128, 512, "f", L.ACT_NONE, False, True),
        ("s0 pw2 res f32", 819200, 512, 128, "f", L.ACT_NONE, True, True),
        ("s0 pw2 inplace f32", 819200, 512, 128, "f", L.ACT_NONE, "inplace", True),
        ("s0 pw2 none f32", 819200, 512, 128, "f", L.ACT_NONE, False, True),
        ("s1 pw1 gelu f16x2", 204800, 256, 1024, "h", L.ACT_GELU, False, True),
        ("neck silu f16x2 N128", 51200, 1152, 128, "h", L.ACT_SILU, False, True),
        ("neck silu f16x2 N64", 204800, 576, 64, "h", L.ACT_SILU, False, True),
    ]
    for name, M, K, N, out, act, use_res, use_bias in cases:
        if only and only not in name:
            continue
        A = P3.zeros((M, K), dev, True)
        A.t.copy_((torch.randn(M, K, device=dev) * ops.ACT_SCALE).to(torch.float16))
        W = P3.from_f32(torch.randn(N, K) * 0.02, dev)
        C = torch.empty(M, N, device=dev, dtype=torch.float32) if out == "f" else P3.zeros((M, N), dev, True)
        bias = torch.randn(N, device=dev) if use_bias else None
        gamma = torch.randn(N, device=dev) if use_res in (True, "gamma", "inplace") else None
        resid = torch.randn(M, N, device=dev) if use_res in (True, "resid", "inplace") else None
        if use_res == "inplace":
            resid = C
        for bn, lblk in ((None, 2),):
            try:
                op = ops.linear(A, W, C, bias=bias, gamma=gamma, resid=resid, act=act, block_n=bn)
                op.i[40] = lblk
                ms = timed(op)
                tf = 2.0 * M * K * N / ms / 1e9
                res.append(dict(name=name, M=M, K=K, N=N, block_n=op.i[13], lblk=lblk, ms=ms, tflops=tf, umma_tflops=3 * tf))
                print(f"{name:22s} M={M} K={K} N={N} bn={op.i[13]} lblk={lblk}: {ms*1000:8.1f} us  {tf:7.1f} TF/s alg  {3*tf:7.1f} umma", flush=True)
            except Exception as e:  # noqa: BLE001
                print(name, bn, "failed:", str(e)[:200])
        del A, W, C
        torch.cuda.empty_cache()
    if len(sys.argv) > 1 and sys.argv[1] != "-":
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
