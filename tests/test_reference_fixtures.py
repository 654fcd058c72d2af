"""The one known-answer test the reference holds on this path: test/test_smc.jl:14-42 pins the initial state
covariance of a 2-state model in CLOSED FORM (`initial_P`) and checks `P0 == T P0 T' + Q`.  The same numbers are run
through the oracle's Lyapunov solver, the engine's batched doubling solver and (as a state space) the dense filter."""
import numpy as np
import pytest
import scipy.linalg as sla


def reduced_form(theta):                       # test_smc.jl:23-29
    t1sq = theta[0] * theta[0]
    return np.array([t1sq, 1 - t1sq - theta[0] * theta[1], 1 - t1sq])


def T_of(phi):                                 # test_smc.jl:36
    return np.array([[phi[0], 0.0], [phi[1], phi[2]]])


def initial_P(phi):                            # test_smc.jl:14
    p1, p2, p3 = phi
    c = p1 * p2 / (1 - p1 * p3)
    return 1 / (1 - p1 ** 2) * np.array([[1.0, c], [c, p2 ** 2 * (1 + 2 * p1 * p3 / (1 - p1 * p3)) / (1 - p3 ** 2)]])


THETA = np.array([0.45, 0.45])                 # test_smc.jl:31
Q = np.array([[1.0, 0.0], [0.0, 0.0]])


def test_reference_fixture_is_self_consistent_and_oracle_reproduces_it():
    phi = reduced_form(THETA)
    T, P0 = T_of(phi), initial_P(phi)
    assert np.allclose(P0, T @ P0 @ T.T + Q, rtol=1e-12)                       # the reference's own @test (:41)
    assert np.allclose(sla.solve_discrete_lyapunov(T, Q), P0, rtol=1e-12)      # the oracle's Lyapunov solver


@pytest.mark.gpu
def test_engine_lyapunov_reproduces_reference_fixture():
    from dynare_jl_b200.engine import lyapunov_batch
    grid = [THETA, np.array([0.3, 0.7]), np.array([0.6, -0.2]), np.array([0.2, 0.1])]
    A = np.stack([T_of(reduced_form(t)) for t in grid])
    B = np.stack([Q] * len(grid))
    X, it, st = lyapunov_batch(A, B)
    for i, t in enumerate(grid):
        assert st[i] == 0
        assert np.allclose(X[i], initial_P(reduced_form(t)), rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
def test_dense_filter_on_the_reference_fixture_state_space():
    """y_t = s1_t + s2_t with the fixture's transition: the engine's filter started from the closed-form P0 against the
    brute-force Gaussian density of the stacked observations."""
    from dynare_jl_b200.engine import kalman_batch
    from oracle import pipeline as pl
    phi = reduced_form(THETA)
    T, P0 = T_of(phi), initial_P(phi)
    # observe state 1 of the rotated system x = (s1 + s2, s2): a selection matrix, as the engine requires
    S = np.array([[1.0, 1.0], [0.0, 1.0]])
    Tr, Pr, Qr = S @ T @ np.linalg.inv(S), S @ P0 @ S.T, S @ Q @ S.T
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((1, 30))
    ll, st = kalman_batch(Y, [0], Tr[None], Qr[None], Pr[None])
    Z = np.array([[1.0, 0.0]])
    ref = pl.gaussian_loglik_bruteforce(Y, Z, np.zeros((1, 1)), Tr, np.eye(2), Qr, np.zeros(2), Pr)
    assert st[0] == 0 and abs(ll[0] - ref) <= 1e-10 * abs(ref)
