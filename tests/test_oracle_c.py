"""The C restatement (oracle/c, the CPU baseline) agrees with the numpy oracle and the golden vectors."""
import numpy as np
import pytest

from oracle import cbuild, models, pipeline as pl


@pytest.mark.parametrize("name", ["fs2000", "ls2003", "hs_nk"])
def test_c_port_matches_golden(name, golden):
    g = golden(name)
    n, kk = g["g1"].shape[1], g["g1"].shape[2]
    ll, st, used, g1 = cbuild.loglik_batch(name, g["theta"], g["Y"], want_g1=True, dims=(n, kk))
    assert used >= 1
    assert np.array_equal(st, g["status"])
    ok = g["status"] == 0
    assert np.all(np.isneginf(ll[~ok]))
    assert np.max(np.abs(ll[ok] - g["loglik"][ok]) / np.abs(g["loglik"][ok])) < 1e-8
    assert np.abs(g1[ok] - g["g1"][ok]).max() < 1e-10          # CR (C port) vs QZ (numpy oracle)


def test_c_port_status_and_missing_data(golden):
    g = golden("ls2003")
    m = models.get_model("ls2003")
    th = np.tile(m.theta_mode, (4, 1))
    th[1, 0] = -1.5
    th[2, 8] = 1.2
    Y = g["Y"].copy()
    Y[2, 7] = np.nan
    Y[:, 30] = np.nan
    ll, st, _ = cbuild.loglik_batch("ls2003", th, Y)
    for i in range(4):
        ref, rst = pl.loglikelihood(m, th[i], Y)
        assert st[i] == rst
        assert (ll[i] == -np.inf) if rst else abs(ll[i] - ref) < 1e-8 * abs(ref)


def test_c_port_single_thread_equals_multi_thread(golden):
    g = golden("fs2000")
    a = cbuild.loglik_batch("fs2000", g["theta"][:64], g["Y"], nthreads=1)
    b = cbuild.loglik_batch("fs2000", g["theta"][:64], g["Y"], nthreads=4)
    assert np.array_equal(a[0], b[0]) and a[2] == 1 and b[2] == 4
