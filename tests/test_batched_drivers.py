"""Batched drivers around the path (lock-step L-BFGS-B, multi-chain slice sampling): CPU tests."""
import numpy as np
import pytest
import scipy.optimize

from gpyreg_b200.batched_drivers import MultiChainSliceSampler, minimize_lockstep


def rosen_batch(H):
    f = np.array([scipy.optimize.rosen(h) for h in H])
    g = np.array([scipy.optimize.rosen_der(h) for h in H])
    return f, g


def test_lockstep_equals_sequential():
    x0s = np.array([[-1.2, 1.0, 0.5], [2.0, -1.0, 1.5], [0.0, 0.0, 0.0], [1.5, 1.5, 1.5]])
    bounds = [(-3, 3)] * 3
    calls = []

    def fb(H):
        calls.append(H.shape[0])
        return rosen_batch(H)
    res = minimize_lockstep(fb, x0s, bounds, 1e-9)
    for x0, r in zip(x0s, res):
        ref = scipy.optimize.minimize(lambda x: (scipy.optimize.rosen(x), scipy.optimize.rosen_der(x)),
                                      x0, jac=True, bounds=bounds, tol=1e-9)
        np.testing.assert_array_equal(r.x, ref.x)          # same iterates, not just the same optimum
        assert r.nfev == ref.nfev and r.fun == ref.fun
    assert max(calls) == 4 and sum(calls) == sum(r.nfev for r in res)


def test_lockstep_propagates_errors():
    def bad(H):
        raise np.linalg.LinAlgError("Singular matrix for L Cholesky decomposition")
    with pytest.raises(np.linalg.LinAlgError):
        minimize_lockstep(bad, np.zeros((2, 2)), [(-1, 1)] * 2, 1e-6)


def test_multichain_moments_and_reproducibility():
    logp = lambda X: -0.5 * (X[:, 0] ** 2 + (X[:, 1] - 0.5 * X[:, 0]) ** 2 / 0.25 + X[:, 2] ** 2 / 4)
    out = []
    for _ in range(2):
        np.random.seed(3)
        mc = MultiChainSliceSampler(logp, np.array([0.1, 0.2, -0.3]), np.array([1.0, 1.0, 2.0]),
                                    -6 * np.ones(3), 6 * np.ones(3), 6)
        out.append(mc.sample(1200, thin=2, burn=150))
    np.testing.assert_array_equal(out[0]["samples"], out[1]["samples"])
    S = out[0]["samples"].reshape(-1, 3)
    assert np.all(np.abs(S.mean(0)) < 0.12)
    assert np.allclose(S.var(0), [1.0, 0.5, 4.0], rtol=0.15)
    # a round evaluates (at most) one point per chain: the batch is the number of chains
    assert out[0]["func_count"] <= out[0]["rounds"] * 6
    assert out[0]["samples"].shape == (6, 1200, 3)


def test_multichain_bounds_and_fixed_coordinates():
    calls = []

    def logp(X):
        calls.append(X.shape[0])
        return -np.sum(np.abs(X), axis=1)
    np.random.seed(0)
    LB, UB = np.array([-1.0, 0.5, -2.0]), np.array([1.0, 0.5, 2.0])       # middle coordinate pinned
    mc = MultiChainSliceSampler(logp, np.array([0.0, 0.5, 0.0]), None, LB, UB, 4)
    r = mc.sample(300, burn=50)
    S = r["samples"].reshape(-1, 3)
    assert np.all(S >= LB) and np.all(S <= UB) and np.all(S[:, 1] == 0.5)
    assert max(calls) <= 4
    with pytest.raises(ValueError, match="outside the bounds"):
        MultiChainSliceSampler(logp, np.array([5.0, 0.5, 0.0]), None, LB, UB, 2)
