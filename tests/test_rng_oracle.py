"""The numpy restatement of the samplers' generator against Random123's published known-answer vectors."""
import numpy as np

from oracle import rng


def _kat(ctr, key):
    o = rng.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
    return [int(x[0]) for x in o]


def test_philox4x32_10_known_answers():
    # kat_vectors of Random123 (philox4x32, 10 rounds)
    assert _kat((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniforms_and_normals_have_the_right_law():
    idx = np.arange(200000)
    u = rng.uniform(7, idx, 3, rng.STREAM_ACCEPT)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 2e-3
    z = rng.normals(7, idx, 3, 5)
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 1e-2
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 1e-2
    # streams / steps / seeds are independent inputs of the counter
    assert not np.array_equal(u, rng.uniform(7, idx, 3, rng.STREAM_MIX))
    assert not np.array_equal(u, rng.uniform(7, idx, 4, rng.STREAM_ACCEPT))
    assert not np.array_equal(u, rng.uniform(8, idx, 3, rng.STREAM_ACCEPT))
    # a draw depends on the global index only (sharding independence)
    assert np.array_equal(u[1000:2000], rng.uniform(7, idx[1000:2000], 3, rng.STREAM_ACCEPT))
