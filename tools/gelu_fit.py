"""Fit used by gelu2() in csrc/tdr_pointwise.cu: Phi(x) = sigmoid(p(x)), p odd of degree 13, minimax in x * Phi(x).

Lawson-weighted least squares on the linearised error  d gelu = x Phi (1 - Phi) dp  over 0 < x <= 7, then an fp32
emulation of the kernel's evaluation order (Horner in x^2, ex2, 1 + e, rcp, * x) against scipy's erfc.

    python tools/gelu_fit.py
"""
import numpy as np
from scipy import special

TERMS = 7
L2E = 1.4426950408889634


def fit(terms=TERMS):
    x = np.linspace(1e-4, 7.0, 28001)
    phi = 0.5 * special.erfc(-x / np.sqrt(2))
    q = 0.5 * special.erfc(x / np.sqrt(2))
    logit = np.log(phi) - np.log(q)
    sens = x * phi * q
    basis = np.stack([x ** (2 * k + 1) for k in range(terms)], 1)

    def err(c):
        p = basis @ c
        return np.maximum(np.abs(x / (1 + np.exp(-p)) - x * phi), np.abs(x / (1 + np.exp(p)) - x * q))

    lw = np.ones_like(x)
    best = None
    for _ in range(300):
        w = sens * np.sqrt(lw)
        sc = np.abs(basis * w[:, None]).max(0)
        c = np.linalg.lstsq(basis * w[:, None] / sc, logit * w, rcond=None)[0] / sc
        e = err(c)
        if best is None or e.max() < best[0]:
            best = (e.max(), c)
        lw = lw * (e / e.max() + 1e-3)
        lw /= lw.sum()
    return best


def emulate_fp32(cs, x):
    x = x.astype(np.float32)
    x2 = x * x
    q = np.full_like(x, cs[-1])
    for k in range(len(cs) - 2, -1, -1):
        q = (q * x2 + cs[k]).astype(np.float32)
    q = (q * x).astype(np.float32)
    with np.errstate(over="ignore"):
        e = np.exp2(q.astype(np.float64)).astype(np.float32)
        d = (np.float32(1) + e).astype(np.float32)
        r = (1 / d.astype(np.float64)).astype(np.float32)
    return (x * r).astype(np.float32)


if __name__ == "__main__":
    e, c = fit()
    cs = np.array([-a * L2E for a in c], dtype=np.float32)
    print("max |gelu error| of the fit (exact arithmetic):", e)
    print("kernel constants (-log2(e) * c_k):", [float(v) for v in cs])
    xs = np.linspace(-12, 12, 2400001)
    ref = xs.astype(np.float32).astype(np.float64)
    ref = ref * 0.5 * special.erfc(-ref / np.sqrt(2))
    print("max |error| of the fp32 evaluation on [-12, 12]:", np.abs(emulate_fp32(cs, xs) - ref).max())
