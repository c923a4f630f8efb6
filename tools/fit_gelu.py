#!/usr/bin/env python
"""Derives the coefficients of gelu_erf_fast (lkgd_b200/csrc/gemm_epilogue.cuh): log2(0.5 erfc(t / sqrt 2)) as a
degree-6 polynomial on [0, 6] (Chebyshev-node least squares ~ minimax), and checks the fp32 Horner evaluation."""
import numpy as np
from scipy.special import erf, erfc

T = 6.0
t = np.cos(np.pi * (np.arange(6000) + 0.5) / 6000) * T / 2 + T / 2
coef = np.polynomial.chebyshev.Chebyshev.fit(t, np.log2(0.5 * erfc(t / np.sqrt(2))), 6, domain=[0, T]) \
    .convert(kind=np.polynomial.Polynomial).coef
print("coefficients (t^0 .. t^6):", [float(np.float32(v)) for v in coef])
x = np.linspace(-8, 8, 400001).astype(np.float32)
ax = np.abs(x)
tt = np.minimum(ax, np.float32(T))
P = np.full_like(tt, np.float32(coef[6]))
for k in range(5, -1, -1):
    P = (P * tt + np.float32(coef[k])).astype(np.float32)
g = np.maximum(x, 0).astype(np.float64) - ax.astype(np.float64) * np.exp2(P.astype(np.float64))
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
err = np.abs(g - ref)
print("max abs err", err.max(), "max rel err (|gelu| > 1e-3)", (err[np.abs(ref) > 1e-3] / np.abs(ref[np.abs(ref) > 1e-3])).max())
