"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product path.

numpy restatement of the samplers' counter-based generator (dynare.jl_b200/csrc/sampler_kernels.cuh):
Philox4x32-10 (Salmon et al. 2011) with counter (index lo, index hi, step, stream << 24 | block) and key
(seed lo, seed hi); uniforms take the top 53 bits of the first 64-bit word; normals are Box-Muller pairs.
Pinned by the published known-answer vectors of Random123 (tests/test_rng_oracle.py).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_RESAMPLE, STREAM_MIX, STREAM_ACCEPT, STREAM_NORMAL = 0, 1, 2, 3
STREAM_PRIOR = 5
TWO_M53 = 1.1102230246251565e-16


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter words (uint64 arrays holding 32-bit values); returns 4 uint64 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _words(seed, idx, step, stream, block):
    idx = np.asarray(idx, dtype=np.uint64)
    c3 = np.full(idx.shape, ((stream << 24) | (block & 0xFFFFFF)) & 0xFFFFFFFF, dtype=np.uint64)
    o = philox4x32_10(idx & MASK, idx >> np.uint64(32), np.full(idx.shape, step, dtype=np.uint64), c3,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    x = (o[1] << np.uint64(32)) | o[0]
    y = (o[3] << np.uint64(32)) | o[2]
    return x, y


def uniform(seed, idx, step, stream, block=0):
    x, _ = _words(seed, idx, step, stream, block)
    return (x >> np.uint64(11)).astype(np.float64) * TWO_M53


def normals(seed, idx, step, count):
    """(len(idx), count) standard normals."""
    idx = np.asarray(idx, dtype=np.uint64)
    out = np.empty((idx.size, count))
    for j in range((count + 1) // 2):
        x, y = _words(seed, idx, step, STREAM_NORMAL, j)
        u1 = ((x >> np.uint64(11)).astype(np.float64) + 1.0) * TWO_M53
        u2 = (y >> np.uint64(11)).astype(np.float64) * TWO_M53
        r = np.sqrt(-2.0 * np.log(u1))
        out[:, 2 * j] = r * np.cos(2.0 * np.pi * u2)
        if 2 * j + 1 < count:
            out[:, 2 * j + 1] = r * np.sin(2.0 * np.pi * u2)
    return out


def normal_first(seed, idx, step, stream, block):
    """First member of the Box-Muller pair of (idx, step, stream, block)."""
    x, y = _words(seed, idx, step, stream, block)
    u1 = ((x >> np.uint64(11)).astype(np.float64) + 1.0) * TWO_M53
    u2 = (y >> np.uint64(11)).astype(np.float64) * TWO_M53
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def gamma_mt(seed, idx, rnd, k, g, a):
    """Gamma(a, 1) variates of the prior stream (sampler_kernels.cuh rng_gamma): Marsaglia & Tsang (2000) with the counter
    layout block = k << 12 | attempt << 2 | sub; g = 0 / 1 selects the first / second variate of parameter k."""
    idx = np.asarray(idx, dtype=np.uint64)
    kbase = k << 12
    boost = np.ones(idx.size)
    if a < 1.0:
        u = uniform(seed, idx, rnd, STREAM_PRIOR, kbase | ((512 + g) << 2))
        boost = (1.0 - u) ** (1.0 / a)
        a = a + 1.0
    d = a - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    out = np.full(idx.size, d) * boost
    todo = np.arange(idx.size)
    for t in range(512):
        if todo.size == 0:
            break
        z = normal_first(seed, idx[todo], rnd, STREAM_PRIOR, kbase | (t << 2) | (2 * g))
        v = 1.0 + c * z
        u = 1.0 - uniform(seed, idx[todo], rnd, STREAM_PRIOR, kbase | (t << 2) | (2 * g + 1))
        with np.errstate(invalid="ignore", divide="ignore"):
            v3 = v * v * v
            acc = (v > 0.0) & (np.log(u) < 0.5 * z * z + d - d * v3 + d * np.log(np.where(v > 0.0, v3, 1.0)))
        out[todo[acc]] = d * v3[acc] * boost[todo[acc]]
        todo = todo[~acc]
    return out
