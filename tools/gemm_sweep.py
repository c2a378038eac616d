"""GPU micro-benchmark of the GEMM kernel on the ConvNeXt stage-2 MLP shapes: separates main loop from epilogue cost.
Usage (on the GPU box): python tools/gemm_sweep.py [out.json]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wedetect_b200 import _lib as L, ops
from wedetect_b200._lib import Program


def timed(op, iters=20):
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
    res = []
    g = torch.Generator(device=dev).manual_seed(0)
    M = 51200
    for name, K, N, out_dt, act, use_res, use_bias in [
        ("pw1 gelu bf16", 512, 2048, torch.bfloat16, L.ACT_GELU, False, True),
        ("pw1 none bf16", 512, 2048, torch.bfloat16, L.ACT_NONE, False, False),
        ("pw1 none f32", 512, 2048, torch.float32, L.ACT_NONE, False, False),
        ("pw2 res f32", 2048, 512, torch.float32, L.ACT_NONE, True, True),
        ("pw2 nores f32", 2048, 512, torch.float32, L.ACT_NONE, False, False),
        ("pw2 nores bf16", 2048, 512, torch.bfloat16, L.ACT_NONE, False, False),
        ("sq 4096 bf16", 4096, 4096, torch.bfloat16, L.ACT_NONE, False, False),
    ]:
        Mx = M if "sq" not in name else 8192
        A = torch.randn(Mx, K, device=dev, generator=g).to(torch.bfloat16)
        W = (torch.randn(N, K, device=dev, generator=g) * 0.02).to(torch.bfloat16)
        C = torch.empty(Mx, N, device=dev, dtype=out_dt)
        bias = torch.randn(N, device=dev, generator=g) if use_bias else None
        gamma = torch.randn(N, device=dev, generator=g) if use_res else None
        resid = torch.randn(Mx, N, device=dev, generator=g) if use_res else None
        for bn in (256, 128):
            try:
                op = ops.linear(A, W, C, bias=bias, gamma=gamma, resid=resid, act=act, block_n=bn)
                ms = timed(op)
                tf = 2.0 * Mx * K * N / ms / 1e9
                res.append(dict(name=name, M=Mx, K=K, N=N, block_n=bn, ms=ms, tflops=tf))
                print(f"{name:18s} M={Mx} K={K} N={N} bn={bn}: {ms*1000:8.1f} us  {tf:7.1f} TF/s", flush=True)
            except Exception as e:  # noqa: BLE001
                print(name, bn, "failed:", str(e)[:200])
    # the same on cuBLAS (library baseline, not the product)
    for K, N, Mx in [(512, 2048, M), (2048, 512, M), (4096, 4096, 8192)]:
        A = torch.randn(Mx, K, device=dev).to(torch.bfloat16)
        W = torch.randn(N, K, device=dev).to(torch.bfloat16)
        for _ in range(3):
            torch.matmul(A, W.t())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(A, W.t())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res.append(dict(name="cublas", M=Mx, K=K, N=N, ms=ms, tflops=2.0 * Mx * K * N / ms / 1e9))
        print(f"cublas M={Mx} K={K} N={N}: {ms*1000:8.1f} us {2.0*Mx*K*N/ms/1e9:7.1f} TF/s", flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
