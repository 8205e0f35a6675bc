"""Coefficients of chain_fused.cu fast_atan2: odd polynomial t*P(t^2) ~ atan(t) on [0, 1], degree 17, minimax by Lawson's
iteratively reweighted least squares; prints the float64 fit error and the error of a float32 Horner evaluation."""
import numpy as np
deg = 8
t = np.cos(np.linspace(0, np.pi, 4001)) * 0.5 + 0.5
t = t[t > 1e-6]
A = np.stack([t ** (2 * k + 1) for k in range(deg + 1)], axis=1)
y = np.arctan(t)
w = np.ones_like(t)
for _ in range(200):
    c, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None)
    e = np.abs(A @ c - y)
    w = w * (1 + e / e.max()); w /= w.mean()
print("max error (float64 fit):", np.abs(A @ c - y).max())
print("coefficients, low order first:", [float(x) for x in c])
c32 = c.astype(np.float32)
tt = np.linspace(0, 1, 2000001).astype(np.float32)
p = np.full_like(tt, c32[deg])
for k in range(deg - 1, -1, -1):
    p = p * (tt * tt) + c32[k]
print("max error (float32 Horner):", np.abs((p * tt).astype(np.float64) - np.arctan(tt.astype(np.float64))).max())
