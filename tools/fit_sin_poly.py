"""Odd minimax-style polynomial of sin(2 pi r) on r in [-0.5, 0.5] (iteratively re-weighted least squares), evaluated
in fp32 Horner form as the kernels do (ptx.cuh: poly_sin).  Prints the coefficients and the maximum error."""
import numpy as np
def fit(nterms):
    r = np.concatenate([np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * 0.5, np.linspace(-0.5, 0.5, 4001)])
    A = np.stack([r ** (2 * k + 1) for k in range(nterms)], 1)
    y = np.sin(2 * np.pi * r)
    w = np.ones_like(r)
    for _ in range(60):
        c, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None)
        e = np.abs(A @ c - y)
        w = w * (1 + 4 * e / e.max()); w /= w.mean()
    return c
for n in (4, 5, 6):
    c = fit(n)
    r = np.linspace(-0.5, 0.5, 200001).astype(np.float32)
    r2 = r * r
    p = np.float32(c[-1])
    for k in range(n - 2, -1, -1):
        p = p * r2 + np.float32(c[k])
    err = np.abs(p * r - np.sin(2 * np.pi * r.astype(np.float64))).max()
    print(f"{n} terms (degree {2*n-1}): max error {err:.3e}  coefficients {[float(x) for x in c]}")
