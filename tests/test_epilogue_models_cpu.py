"""CPU models of the split-GEMM epilogue arithmetic (wedetect_b200/csrc/epi_split.cuh), evaluated in numpy float32 with the very
coefficients the kernel compiles: the single-interval erf-GELU against a float64 GELU (mm_backbone.py:120: nn.GELU(), erf form)."""
import math
import os
import re

import numpy as np

CUH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wedetect_b200", "csrc", "epi_split.cuh")


def kernel_gelu_coefficients():
    src = open(CUH).read()
    body = src[src.index("#ifndef WD_GELU_TWO_INTERVAL"): src.index("#else")]
    c = [float(v) for v in re.findall(r"splat2\((-?[0-9.]+e?-?[0-9]*)f\)", body)]
    clamp = float(re.search(r"fminf\(a0, ([0-9.]+)f\)", body).group(1))
    assert len(c) == 10 and c[-2:] == [-1.0, 1.0]            # 8 polynomial coefficients (highest first), then the (-1, 1) of 1 - 2^r
    return np.array(c[:8], np.float32), np.float32(clamp)


def gelu_model(x, c, clamp):
    f32 = np.float32
    a = np.abs(x)
    t = np.minimum(a, clamp)
    p = np.full_like(t, c[0])
    for k in c[1:]:
        p = (p * t + k).astype(f32)
    r = (p * t).astype(f32)
    e = (f32(1.0) - np.exp2(r.astype(np.float64)).astype(f32)).astype(f32)
    h, ha = (x * f32(0.5)).astype(f32), (a * f32(0.5)).astype(f32)
    return (ha.astype(np.float64) * e.astype(np.float64) + h.astype(np.float64)).astype(f32)          # fma: one rounding


def test_single_interval_erf_gelu_matches_float64():
    c, clamp = kernel_gelu_coefficients()
    x = np.linspace(-10, 10, 1_000_001).astype(np.float32)
    got = gelu_model(x, c, clamp).astype(np.float64)
    erf = np.vectorize(math.erf)
    xd = x.astype(np.float64)
    want = 0.5 * xd * (1.0 + erf(xd / math.sqrt(2.0)))
    err = np.abs(got - want)
    assert err.max() <= 6e-7 and np.sqrt((err ** 2).mean()) <= 1.2e-7, (err.max(), float(x[err.argmax()]))
    ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
    big = np.abs(x) >= 1.0
    assert (err[big & (x > 0)] / ulp[big & (x > 0)]).max() <= 1.5          # positive side: within 1.5 ulp of the result
    assert np.array_equal(got[x >= 6], xd[x >= 6]) and np.all(np.abs(got[x <= -6]) <= 1e-8)        # saturation: x and (-)0


def test_gelu_polynomial_is_the_committed_fit():
    """tools/erf_fit.py documents how the coefficients were produced: P approximates log2(erfc(T / sqrt 2)) / T on [0, 5.9]."""
    c, clamp = kernel_gelu_coefficients()
    t = np.linspace(0.01, float(clamp), 5000)
    target = np.array([math.log2(math.erfc(v / math.sqrt(2.0))) / v for v in t])
    p = np.polyval(c.astype(np.float64), t)
    w = np.exp2(target * t) * math.log(2.0) * t                # d erf / d P
    assert np.abs((p - target) * w).max() <= 4e-8
