"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product path.

CPU restatement of the SMC algebra of reference ``src/estimation/smc.jl``:
correction (``:118-123``), ESS (``:129``), ``multinomial_resampling`` (``:340-361``), weighted
moments (``:154-161``).  The reference leaves the floating-point association order of its sums to
Julia's ``sum``/``cumsum``; the engine's contract ("indices bit-exact given identical weights and
uniforms") is stated against ONE specified order, restated here in plain sequential loops:
    sums: each run of 32 consecutive particles left to right; the 32 run totals of each block of 1024 consecutive
    particles left to right; the block totals left to right;
    cumulative weights: cw[b*1024 + i] = offset_b + local_b[i] with local_b the left-to-right inclusive scan of
    block b and offset_b the left-to-right sum of the previous blocks' scan totals.
"""
from __future__ import annotations

import numpy as np

BLOCK = 1024
RUN = 32


def _block_totals(x):
    n = x.size
    nb = (n + BLOCK - 1) // BLOCK
    tot = np.zeros(nb)
    for b in range(nb):
        t = 0.0
        for r in range(BLOCK // RUN):
            s = 0.0
            for v in x[b * BLOCK + r * RUN:b * BLOCK + (r + 1) * RUN]:
                s += float(v)
            t += s
        tot[b] = t
    return tot


def canonical_sum(x):
    s = 0.0
    for t in _block_totals(np.asarray(x, dtype=np.float64)):
        s += float(t)
    return s


def canonical_cumsum(w):
    w = np.asarray(w, dtype=np.float64)
    n = w.size
    cw = np.empty(n)
    off = 0.0
    for b in range((n + BLOCK - 1) // BLOCK):
        s = 0.0
        lo, hi = b * BLOCK, min((b + 1) * BLOCK, n)
        local = np.empty(hi - lo)
        for i in range(lo, hi):
            s += float(w[i])
            local[i - lo] = s
        cw[lo:hi] = off + local
        off = off + s
    return cw


def reweight(w, loglh, dphi):
    """smc.jl:118-123, 129: returns (normalised weights, zhat, ESS)."""
    wt = np.asarray(w) * np.exp(dphi * np.asarray(loglh))
    zhat = canonical_sum(wt)
    wn = wt / zhat
    ess = 1.0 / canonical_sum(wn * wn)
    return wn, zhat, ess


def resample_indices(w, u):
    """smc.jl:347-359 for given uniforms: 0-based first j with u < cw[j]; n (out of range in the
    reference) is clamped to n-1.  Returns (idx, number of clamps)."""
    cw = canonical_cumsum(w)
    n = cw.size
    idx = np.empty(len(u), dtype=np.int64)
    nclamp = 0
    for i, ui in enumerate(u):
        j = 0
        # linear search as in the reference, accelerated by a bisection that returns the same j
        j = int(np.searchsorted(cw, ui, side="right"))
        if j >= n:
            j = n - 1
            nclamp += 1
        idx[i] = j
    return idx, nclamp


def resample_indices_linear(w, u):
    """Literal transcription of the reference loop (O(n^2)); for small cross-checks."""
    cw = canonical_cumsum(w)
    n = cw.size
    out = []
    for ui in u:
        j = 0
        while j < n:
            if ui < cw[j]:
                break
            j += 1
        out.append(min(j, n - 1))
    return np.array(out, dtype=np.int64)


def systematic_uniforms(n, u0):
    return (np.arange(n, dtype=np.float64) + u0) / float(n)


def moments(para, w):
    """smc.jl:154-161: mu = para' w ; z = para - mu' ; R = (z .* w)' z, canonical order per entry."""
    para = np.asarray(para, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    n, npar = para.shape
    mu = np.array([canonical_sum(w * para[:, p]) for p in range(npar)])
    z = para - mu[None, :]
    R = np.empty((npar, npar))
    for p in range(npar):
        for q in range(npar):
            R[p, q] = canonical_sum((w * z[:, p]) * z[:, q])
    return mu, R
