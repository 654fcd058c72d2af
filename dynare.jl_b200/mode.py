"""Host-side mirror of the derivative step around the posterior mode (reference `posterior_mode!`,
src/estimation/estimation.jl:864-891): the finite-difference Hessian of the objective is ONE batched evaluation of
the O(np^2) stencil instead of a loop of scalar calls.

`finite_difference_hessian` restates FiniteDiff.jl's default (`Val(:hcentral)`, relative step eps^(1/4) with the same
absolute floor) as the reference calls it (`estimation.jl:7, 823, 873, 882`):
    H_ii = (-f(x+2h_i) + 16 f(x+h_i) - 30 f(x) + 16 f(x-h_i) - f(x-2h_i)) / (12 h_i^2)
    H_ij = ( f(x+h_i+h_j) - f(x+h_i-h_j) - f(x-h_i+h_j) + f(x-h_i-h_j)) / (4 h_i h_j)
FiniteDiff is a third-party dependency absent from the tree (Project.toml); the stencil is its published one.
"""
from __future__ import annotations

import numpy as np

EPS4 = np.finfo(np.float64).eps ** 0.25


def hessian_stencil(x):
    """(points (M, n), step sizes h (n,)) of the central-difference Hessian at x; point 0 is x itself."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    h = np.maximum(EPS4 * np.abs(x), EPS4)
    pts = [x.copy()]
    for i in range(n):
        for m in (2.0, 1.0, -1.0, -2.0):
            p = x.copy(); p[i] += m * h[i]; pts.append(p)
    for i in range(n):
        for j in range(i + 1, n):
            for si, sj in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                p = x.copy(); p[i] += si * h[i]; p[j] += sj * h[j]; pts.append(p)
    return np.array(pts), h


def finite_difference_hessian(f_batch, x):
    """f_batch: (M, n) -> (M,) evaluated in one call (e.g. ``lambda T: -post(T)`` with a DSGELogPosteriorDensity)."""
    pts, h = hessian_stencil(x)
    n = h.size
    f = np.asarray(f_batch(pts), dtype=np.float64)
    H = np.empty((n, n))
    f0 = f[0]
    k = 1
    for i in range(n):
        fp2, fp1, fm1, fm2 = f[k:k + 4]; k += 4
        H[i, i] = (-fp2 + 16 * fp1 - 30 * f0 + 16 * fm1 - fm2) / (12 * h[i] * h[i])
    for i in range(n):
        for j in range(i + 1, n):
            fpp, fpm, fmp, fmm = f[k:k + 4]; k += 4
            H[i, j] = H[j, i] = (fpp - fpm - fmp + fmm) / (4 * h[i] * h[j])
    return H


def mode_covariance(hess):
    """invhess = inv(hess ./ (hsd * hsd')) ./ (hsd * hsd'); std = sqrt(diag) with negative entries -> NaN
    (estimation.jl:874-890)."""
    hsd = np.sqrt(np.diag(hess))
    sc = np.outer(hsd, hsd)
    invhess = np.linalg.inv(hess / sc) / sc
    d = np.diag(invhess).copy()
    d[d < 0] = np.nan
    return invhess, np.sqrt(d)
