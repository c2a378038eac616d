"""Fit of the single-interval erf-GELU polynomial used by the split-GEMM epilogue (wedetect_b200/csrc/epi_split.cuh).

    erf(T / sqrt 2) = 1 - 2^(T P(T)),   T = min(|x|, 5.9)

P approximates f(T) = log2(erfc(T / sqrt 2)) / T on (0, 5.9] (f(0) = -sqrt(2 / pi) / ln 2).  The fit minimises the error of erf, not of
f: d erf = 2^(T f) ln 2 * T * dP, so the least-squares weights are that factor, re-weighted a few dozen times towards the
minimax solution.  Degree 7 (the kernel's) reaches 1.6e-8 in float64; evaluated in float32 (Horner) the error of the GELU 0.5 x (1 + erf) is at most
3.4e-7 (below one ulp of the result), rms 5.7e-8 - torch's own fp32 GELU on the same points: 1.2e-6 / 1.4e-7.  Higher degrees do
not help (float32 Horner rounding dominates from degree 7 on: degree 8 gives 3.1e-7 / 5.6e-8).

    python tools/erf_fit.py            prints the float32 coefficients (lowest degree first) and the error table
tests/test_epilogue_models_cpu.py evaluates the coefficients compiled into the kernel the same way.
"""
import numpy as np
from scipy.special import erf, log_ndtr

LN2 = np.log(2.0)
T_MAX = 5.9


def r_of(t):
    """log2(erfc(t / sqrt 2)) = 1 + log2(Phi(-t))"""
    return 1.0 + log_ndtr(-t) / LN2


def fit(deg, ts, iters=80):
    target = r_of(ts) / ts
    w = np.maximum((2.0 ** r_of(ts)) * LN2 * ts, 1e-12)
    V = np.vander(ts, deg + 1, increasing=True)
    wt = w.copy()
    for _ in range(iters):
        c, *_ = np.linalg.lstsq(V * wt[:, None], target * wt, rcond=None)
        err = (V @ c - target) * w                      # error of erf
        wt = wt * (1 + 3 * np.abs(err) / np.abs(err).max())
        wt /= wt.max() / w.max()
    return c, float(np.abs(err).max())


def gelu_f32(x, c32):
    t = np.minimum(np.abs(x), np.float32(T_MAX))
    p = np.full_like(t, c32[-1])
    for k in range(len(c32) - 2, -1, -1):
        p = (p * t + c32[k]).astype(np.float32)
    r = (p * t).astype(np.float32)
    e = (np.float32(1.0) - np.exp2(r.astype(np.float64)).astype(np.float32)).astype(np.float32)
    h, ha = (x * np.float32(0.5)).astype(np.float32), (np.abs(x) * np.float32(0.5)).astype(np.float32)
    return (ha.astype(np.float64) * e + h).astype(np.float32)


def main():
    ts = np.concatenate([np.linspace(1e-6, 0.5, 4000), np.linspace(0.5, T_MAX, 20000)])
    x = np.linspace(-8, 8, 2_000_001).astype(np.float32)
    xd = x.astype(np.float64)
    true = 0.5 * xd * (1 + erf(xd / np.sqrt(2)))
    print("degree  fit error of erf (f64)   GELU max abs / rms error (f32 Horner)")
    for deg in (6, 7, 8, 9, 10):
        c, e64 = fit(deg, ts)
        err = np.abs(gelu_f32(x, c.astype(np.float32)).astype(np.float64) - true)
        print(f"{deg:6d}  {e64:22.3e}   {err.max():.3e} / {np.sqrt((err ** 2).mean()):.3e}")
    c, _ = fit(7, ts)
    print("degree-7 coefficients (float32, lowest degree first):")
    for k, v in enumerate(c.astype(np.float32)):
        print(f"  c{k} = {float(v)!r}")
    try:
        import torch
        tg = torch.nn.functional.gelu(torch.from_numpy(x)).numpy().astype(np.float64)
        print(f"torch fp32 GELU on the same points: max abs {np.abs(tg - true).max():.3e} / rms {np.sqrt(((tg - true) ** 2).mean()):.3e}")
    except ImportError:
        pass


if __name__ == "__main__":
    main()
