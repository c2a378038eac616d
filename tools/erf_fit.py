import numpy as np
from scipy.special import log_ndtr, erf
ln2 = np.log(2.0)
T = 5.9
def f(t):   # log2(erfc(t/sqrt2)) / t
    return (1.0 + log_ndtr(-t) / ln2) / t
def r_of(t): return 1.0 + log_ndtr(-t) / ln2
ts = np.concatenate([np.linspace(1e-6, 0.5, 4000), np.linspace(0.5, T, 20000)])
for deg in (8, 9, 10, 11, 12):
    # iteratively reweighted least squares toward minimax of erf error
    w = (2.0 ** r_of(ts)) * ln2 * ts
    w = np.maximum(w, 1e-12)
    V = np.vander(ts, deg + 1, increasing=True)
    wt = w.copy()
    for it in range(60):
        c, *_ = np.linalg.lstsq(V * wt[:, None], f(ts) * wt, rcond=None)
        err = (V @ c - f(ts)) * w          # error in erf
        wt = wt * (1 + 3 * np.abs(err) / np.abs(err).max())
        wt /= wt.max() / w.max()
    # float32 Horner evaluation
    c32 = c.astype(np.float32)
    t32 = ts.astype(np.float32)
    p = np.full_like(t32, c32[-1])
    for k in range(deg - 1, -1, -1):
        p = (p * t32 + c32[k]).astype(np.float32)
    r = (p * t32).astype(np.float32)
    e = (1.0 - np.exp2(r.astype(np.float64)))
    true = erf(t32.astype(np.float64) / np.sqrt(2.0))
    print(deg, "max erf err (poly f32, exact exp2):", np.abs(e - true).max(), "at t=", ts[np.abs(e - true).argmax()], " f64 fit err:", np.abs(err).max())

print("---- degree 8 coefficients")
deg = 8
w = np.maximum((2.0 ** r_of(ts)) * ln2 * ts, 1e-12)
V = np.vander(ts, deg + 1, increasing=True)
wt = w.copy()
for it in range(80):
    c, *_ = np.linalg.lstsq(V * wt[:, None], f(ts) * wt, rcond=None)
    err = (V @ c - f(ts)) * w
    wt = wt * (1 + 3 * np.abs(err) / np.abs(err).max())
    wt /= wt.max() / w.max()
c32 = c.astype(np.float32)
for k, v in enumerate(c32): print(k, repr(float(v)))
# GELU error with fp32 emulation over a dense grid of x in [-8, 8]
x = np.linspace(-8, 8, 2_000_001).astype(np.float32)
t = np.minimum(np.abs(x), np.float32(T))
p = np.full_like(t, c32[-1])
for k in range(deg - 1, -1, -1):
    p = (p * t + c32[k]).astype(np.float32)
r = (p * t).astype(np.float32)
e2 = np.exp2(r.astype(np.float64)).astype(np.float32)
E = (np.float32(1.0) - e2).astype(np.float32)
h = (x * np.float32(0.5)).astype(np.float32)
ha = (np.abs(x) * np.float32(0.5)).astype(np.float32)
g = (ha * E + h).astype(np.float32)        # fma emulated loosely
xd = x.astype(np.float64)
true = 0.5 * xd * (1 + erf(xd / np.sqrt(2)))
ae = np.abs(g.astype(np.float64) - true)
ulp = np.spacing(np.abs(true).astype(np.float32)).astype(np.float64)
print("max abs err", ae.max(), "at x", x[ae.argmax()], " max err in ulps (|x|>0.01):", (ae / np.maximum(ulp, 1e-45))[np.abs(x) > 0.01].max())
import torch
tg = torch.nn.functional.gelu(torch.from_numpy(x)).numpy().astype(np.float64)
print("torch fp32 gelu: max abs err", np.abs(tg - true).max(), " max ulps:", (np.abs(tg - true) / np.maximum(ulp, 1e-45))[np.abs(x) > 0.01].max())
print("rms err ours", np.sqrt((ae**2).mean()), "torch", np.sqrt(((tg-true)**2).mean()))
