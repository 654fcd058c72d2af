"""The batched finite-difference Hessian (host mirror of posterior_mode!'s derivative step)."""
import numpy as np
import pytest

from dynare_jl_b200.mode import finite_difference_hessian, hessian_stencil, mode_covariance


def test_stencil_size_and_exactness_on_a_quartic():
    rng = np.random.default_rng(0)
    n = 5
    A = rng.standard_normal((n, n)); A = A @ A.T + n * np.eye(n)
    b = rng.standard_normal(n)
    f = lambda X: 0.5 * np.einsum("ki,ij,kj->k", X, A, X) + X @ b + 0.1 * np.sum(X ** 3, axis=1)
    x = rng.standard_normal(n)
    pts, h = hessian_stencil(x)
    assert pts.shape == (1 + 4 * n + 2 * n * (n - 1), n)
    H = finite_difference_hessian(f, x)
    exact = A + np.diag(0.6 * x)
    assert np.allclose(H, exact, rtol=1e-6, atol=1e-6)
    cov, std = mode_covariance(H)
    assert np.allclose(cov, np.linalg.inv(H), rtol=1e-8)
    assert np.allclose(std, np.sqrt(np.diag(np.linalg.inv(H))))


@pytest.mark.gpu
def test_posterior_hessian_one_batch_matches_oracle(golden):
    from dynare_jl_b200.engine import BatchSSWs, DSGELogPosteriorDensity, model_description
    from dynare_jl_b200.priors import PriorSet
    from oracle import cbuild, models, sampler as osam
    g = golden("fs2000")
    m = models.get_model("fs2000")
    pri = PriorSet.from_description(model_description("fs2000")["priors"])
    ws = BatchSSWs("fs2000", g["Y"])
    post = DSGELogPosteriorDensity(ws, pri)
    x = np.asarray(m.theta_mode, dtype=np.float64)
    n0 = ws.kernel_launches()
    H = finite_difference_hessian(lambda T: -post(T), x)
    assert ws.kernel_launches() - n0 <= 8                      # the whole stencil is ONE likelihood batch
    code, a, b = pri.abi_arrays()
    pr = list(zip(code.tolist(), a.tolist(), b.tolist()))

    def oracle(T):
        ll, st, _ = cbuild.loglik_batch("fs2000", T, g["Y"])
        return -np.array([osam.logprior(pr, t) for t in T]) - ll
    Ho = finite_difference_hessian(oracle, x)
    sc = np.sqrt(np.outer(np.abs(np.diag(Ho)), np.abs(np.diag(Ho))))
    assert np.abs(H - Ho).max() / sc.max() < 1e-3 and np.abs((H - Ho) / sc).max() < 5e-3
    ws.close()
