#!/usr/bin/env python
"""Fit and check the one-MUFU GELU used by the fc1 GEMM epilogue (csrc/mathfn.cuh::gelu_fast):
GELU(u) = max(u,0) - |u| 2^q(|u|), q ~ log2(erfc(|u|/sqrt 2)/2), degree 6 on [0, 6]."""
import numpy as np
from scipy.special import erf, erfc
deg, hi = 6, 6.0
a = np.cos(np.pi * (np.arange(8000) + 0.5) / 8000) * hi / 2 + hi / 2
w = np.maximum(a * 0.5 * erfc(a / np.sqrt(2)), 1e-3)
c32 = np.polynomial.polynomial.polyfit(a, np.log2(0.5 * erfc(a / np.sqrt(2))), deg, w=w).astype(np.float32)
print("coefficients c0..c6:", ", ".join("%.9ef" % c for c in c32))
u = np.linspace(-12, 12, 2000001).astype(np.float32)
aa = np.minimum(np.abs(u), np.float32(hi))
q = np.full_like(aa, c32[6])
for k in range(5, -1, -1):
    q = (q * aa + c32[k]).astype(np.float32)
g = (np.maximum(u, 0) - aa * np.exp2(q.astype(np.float64)).astype(np.float32)).astype(np.float32)
ref = 0.5 * u.astype(np.float64) * (1 + erf(u.astype(np.float64) / np.sqrt(2)))
err = np.abs(g - ref)
print("max abs err %.3e at u = %.3f" % (err.max(), u[err.argmax()]))
