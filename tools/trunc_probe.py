"""Measure the systematic (toward-zero) error of the fp16 hi/lo GEMM: shrink = sum(err * ref) / sum(ref^2) against an fp64 GEMM,
for several K and accumulator block lengths.  A round-to-nearest fp32 GEMM would give |shrink| << 1e-8."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wedetect_b200 import _lib as L, ops
from wedetect_b200.ops import P3

d = torch.device("cuda:0")
L.load()
out = []
for kind in ("normal", "relu"):
    for K in (128, 256, 512, 2048, 1152):
        M, N = 1024, 512
        g = torch.Generator().manual_seed(K)
        A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
        if kind == "relu":
            A = A.clamp_min(0.0)
        ref = A.double() @ W.double().t()
        Ap, Wp = P3.from_f32(A, d, scale=ops.ACT_SCALE), P3.from_f32(W, d)
        # what the operand representation alone costs (fp64 product of the reconstructed planes)
        rep = (Ap.value().cpu().double() @ Wp.value().cpu().double().t()) - ref
        for lblk, nocomp in ((1, 1), (1, 0), (2, 1), (2, 0)):
            C = torch.zeros(M, N, dtype=torch.float32, device=d)
            op = ops.linear(Ap, Wp, C)
            op.i[40], op.i[41] = lblk, nocomp
            L.run_op(op, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            err = C.cpu().double() - ref
            f32 = (A @ W.t()).double() - ref
            row = dict(kind=kind, K=K, lblk=lblk, compensated=not nocomp, shrink=float((err * ref).sum() / (ref * ref).sum()), rel_rms=float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
                       rep_rel_rms=float(rep.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()), cpu_f32_rel_rms=float(f32.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
                       cpu_f32_shrink=float((f32 * ref).sum() / (ref * ref).sum()))
            out.append(row)
            print(row)
json.dump(out, open("gpurun_out/trunc_probe.json", "w"), indent=1)
