"""Fit of the exponent polynomial used by gelu2 (jampr_plus_b200/csrc/policy_tc_common.cuh).

gelu(x) = relu(x) - |x| * Phi(-|x|) with Phi(-u) = 2 ** q(u); q is fitted by weighted least squares with
Lawson re-weighting (-> minimax absolute error of Phi) on [0, 6.5].  Prints the float32 coefficients and the
resulting error of gelu against the float64 erf form.
"""
import numpy as np
from scipy.special import erfc

DEG = 5
u = np.linspace(0, 6.5, 20001)
h = 0.5 * erfc(u / np.sqrt(2))
q = np.log2(h)
V = np.vander(u, DEG + 1, increasing=True)
wt = h.copy()
for _ in range(80):
    c, *_ = np.linalg.lstsq(V * wt[:, None], q * wt, rcond=None)
    err = h * (np.exp2(V @ c - q) - 1)
    wt = wt * (1 + 4 * np.abs(err) / np.abs(err).max())
    wt /= wt.max()
c = c.astype(np.float32)
print("coefficients c0..c%d (polynomial in u = |x|):" % DEG, ", ".join(repr(float(v)) for v in c))
print("max abs error of Phi(-u):", np.abs(err).max())
x = np.linspace(-12, 12, 400001).astype(np.float32)
nu = -np.abs(x)
acc = np.full_like(nu, c[DEG] * (-1) ** DEG)
for k in range(DEG - 1, -1, -1):
    acc = (acc * nu + np.float32(c[k] * (-1) ** k)).astype(np.float32)
g = (nu * np.exp2(acc.astype(np.float64)).astype(np.float32) + np.maximum(x, 0)).astype(np.float64)
ref = x.astype(np.float64) * 0.5 * erfc(-x.astype(np.float64) / np.sqrt(2))
print("max abs error of gelu:", np.abs(g - ref).max())
