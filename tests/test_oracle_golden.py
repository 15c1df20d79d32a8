"""Pin the CPU oracle (oracle/gp_oracle.py) against outputs of the REAL reference
stored in tests/golden/*.npz (made by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import gp_oracle as orc
from tests.helpers import COV_TAGS, CORE_TAGS, case, grad_err, rel_err, spec_from_array


@pytest.mark.parametrize("tag", sorted(COV_TAGS))
def test_cov_plugins(golden_plugins, tag):
    g = golden_plugins
    ck, deg, ard = COV_TAGS[tag]
    X, Xs = g["X"], g["Xs"]
    spec = orc.ModelSpec(D=X.shape[1], cov_kind=ck, degree=deg, ard=bool(ard))
    hyp = g[f"{tag}.hyp"]
    K, dK = orc.cov_compute(spec, hyp, X, compute_grad=True)
    np.testing.assert_array_equal(K, g[f"{tag}.K"])
    np.testing.assert_array_equal(dK, g[f"{tag}.dK"])      # NaNs compare equal
    np.testing.assert_array_equal(orc.cov_compute(spec, hyp, X, Xs), g[f"{tag}.Kx"])
    np.testing.assert_array_equal(orc.cov_compute(spec, hyp, X, compute_diag=True),
                                  g[f"{tag}.Kd"])


@pytest.mark.parametrize("mk", [0, 1, 2])
def test_mean_plugins(golden_plugins, mk):
    g = golden_plugins
    X = g["X"]
    spec = orc.ModelSpec(D=X.shape[1], mean_kind=mk)
    m, dm = orc.mean_compute(spec, g[f"mean{mk}.hyp"], X, compute_grad=True)
    np.testing.assert_array_equal(m, g[f"mean{mk}.m"])
    if mk:
        np.testing.assert_array_equal(dm, g[f"mean{mk}.dm"])


@pytest.mark.parametrize("p", [(a, b, c) for a in (0, 1) for b in (0, 1, 2) for c in (0, 1)])
def test_noise_plugins(golden_plugins, p):
    g = golden_plugins
    key = "noise%d%d%d" % p
    spec = orc.ModelSpec(D=g["X"].shape[1], noise_params=p)
    sn2, dsn2 = orc.noise_compute(spec, g[key + ".hyp"], g["X"], g["noise.y"],
                                  g["noise.s2"], compute_grad=True)
    np.testing.assert_array_equal(np.asarray(sn2, dtype=float), g[key + ".sn2"])
    np.testing.assert_array_equal(dsn2, g[key + ".dsn2"])


def test_plugin_errors():
    spec = orc.ModelSpec(D=3)
    X = np.zeros((4, 3))
    with pytest.raises(ValueError, match="Expected 4 covariance function hyperparameters"):
        orc.cov_compute(spec, np.zeros(3), X)
    with pytest.raises(ValueError, match="one-sample hyperparameter inputs"):
        orc.cov_compute(spec, np.zeros((4, 1)), X)
    with pytest.raises(ValueError, match="X_star should be None"):
        orc.cov_compute(spec, np.zeros(4), X, X, compute_grad=True)


@pytest.mark.parametrize("tag", CORE_TAGS)
def test_core(golden_core, tag):
    c = case(golden_core, tag)
    spec = spec_from_array(c["spec"])
    X, y, s2 = c["X"], c["y"], c.get("s2")
    with np.errstate(all="ignore"):
        nlz, dnlz = orc.nlz_batch(spec, c["hyp"], X, y, s2, True)
        nlz0 = orc.nlz_batch(spec, c["hyp"], X, y, s2, False)
    # same arithmetic, same libraries: bit-exact
    np.testing.assert_array_equal(nlz, c["nlZ"])
    np.testing.assert_array_equal(nlz0, c["nlZ"])
    np.testing.assert_array_equal(dnlz, c["dnlZ"])
    posts = orc.posterior_batch(spec, c["hyp"], X, y, s2)
    for b, p in enumerate(posts):
        np.testing.assert_array_equal(p.alpha[:, 0], c["alpha"][b])
        assert p.sW[0, 0] == c["sW"][b]
        assert p.sn2_mult == c["sn2_mult"][b]
        assert int(p.L_chol) == c["L_chol"][b]
        if "L" in c:
            np.testing.assert_array_equal(p.L, c["L"][b])
    for add_noise in (0, 1):
        for sep in (0, 1):
            mu, v, lpd = orc.predict(spec, posts, X, y, c["Xs"], c["ys"], c.get("s2s"),
                                     add_noise=bool(add_noise), separate_samples=bool(sep),
                                     return_lpd=True)
            k = f"pred{add_noise}{sep}"
            np.testing.assert_array_equal(mu, c[k + ".mu"])
            np.testing.assert_array_equal(v, c[k + ".s2"])
            np.testing.assert_array_equal(lpd, c[k + ".lpd"])


@pytest.mark.parametrize("tag", ["eps", "thr"])
def test_lownoise(golden_lownoise, tag):
    c = case(golden_lownoise, tag)
    spec = spec_from_array(c["spec"])
    X, y = c["X"], c["y"]
    with np.errstate(all="ignore"):
        nlz, dnlz = orc.nlz_batch(spec, c["hyp"], X, y, None, True)
    np.testing.assert_array_equal(nlz, c["nlZ"])
    np.testing.assert_array_equal(dnlz, c["dnlZ"])
    posts = orc.posterior_batch(spec, c["hyp"], X, y, None)
    assert [p.sn2_mult for p in posts] == list(c["sn2_mult"])
    assert [int(p.L_chol) for p in posts] == list(c["L_chol"])
    mu, v = orc.predict(spec, posts, X, y, c["Xs"], add_noise=True, separate_samples=True)
    np.testing.assert_array_equal(mu, c["pred.mu"])
    np.testing.assert_array_equal(v, c["pred.s2"])
    assert rel_err(np.stack([p.alpha[:, 0] for p in posts]), c["alpha"]) == 0.0
    assert grad_err(dnlz, c["dnlZ"]) == 0.0


# ---------------------------------------------------------------- round-2 goldens
@pytest.fixture(scope="module")
def golden_r2():
    from tests.conftest import _load
    return _load("round2.npz")


def test_low2_multi_tile_lownoise(golden_r2):
    """Low-noise branch on 4 tiles per side: the restatement equals the reference bit for bit."""
    c = case(golden_r2, "low2")
    spec = spec_from_array(c["spec"])
    nlz, dnlz = orc.nlz_batch(spec, c["hyp"], c["X"], c["y"], None, True)
    np.testing.assert_array_equal(nlz, c["nlZ"])
    np.testing.assert_array_equal(dnlz, c["dnlZ"])
    posts = orc.posterior_batch(spec, c["hyp"], c["X"], c["y"], None)
    np.testing.assert_array_equal(posts[0].L, c["L0"])
    for b, p in enumerate(posts):
        np.testing.assert_array_equal(p.alpha[:, 0], c["alpha"][b])
        assert not p.L_chol and p.sn2_mult == c["sn2_mult"][b]
    for an in (0, 1):
        mu, s2 = orc.predict(spec, posts, c["X"], c["y"], c["Xs"], add_noise=bool(an), separate_samples=True)
        np.testing.assert_array_equal(mu, c[f"pred{an}.mu"])
        np.testing.assert_array_equal(s2, c[f"pred{an}.s2"])


def test_design_rows(golden_r2):
    """Two rows of the real f_min_fill design at N=2000 (one per factorisation branch)."""
    c = case(golden_r2, "design")
    spec = orc.ModelSpec(D=c["X"].shape[1], cov_kind=1, degree=5, ard=True, mean_kind=2)
    rows = [int(np.flatnonzero(c["L_chol"] == 1)[0]), int(np.flatnonzero(c["L_chol"] == 0)[0])]
    with np.errstate(all="ignore"):
        nlz, dnlz = orc.nlz_batch(spec, c["hyp"][rows], c["X"], c["y"], None, True)
    np.testing.assert_array_equal(nlz, c["nlZ"][rows])
    np.testing.assert_array_equal(dnlz, c["dnlZ"][rows])


def test_rq_isotropic_definition(golden_r2):
    """Isotropic RQ is DEFINED as the reference's RationalQuadraticARD with tied length scales:
    K, cross-covariance, prior variance and nlZ are the reference's bits, the length-scale derivative
    is the sum of the ARD ones."""
    c = case(golden_r2, "rqiso")
    D = c["X"].shape[1]
    spec = orc.ModelSpec(D=D, cov_kind=2, ard=False)
    assert spec.cov_n == 3
    K, dK = orc.cov_compute(spec, c["hyp"], c["X"], compute_grad=True)
    np.testing.assert_array_equal(K, c["K"])
    np.testing.assert_array_equal(dK, c["dK"])
    np.testing.assert_array_equal(orc.cov_compute(spec, c["hyp"], c["X"], c["Xs"]), c["Kx"])
    np.testing.assert_array_equal(orc.cov_compute(spec, c["hyp"], c["X"], compute_diag=True), c["Kd"])
    X, y, H = c["gp.X"], c["gp.y"], c["gp.hyp"]
    gspec = orc.ModelSpec(D=X.shape[1], cov_kind=2, ard=False, mean_kind=1)
    nlz, dnlz = orc.nlz_batch(gspec, H, X, y, None, True)
    np.testing.assert_array_equal(nlz, c["gp.nlZ"])
    assert grad_err(dnlz, c["gp.dnlZ"]) <= 1e-13
    posts = orc.posterior_batch(gspec, H, X, y, None)
    for sep in (0, 1):
        mu, s2 = orc.predict(gspec, posts, X, y, c["gp.Xs"], add_noise=True, separate_samples=bool(sep))
        np.testing.assert_array_equal(mu, c[f"gp.mu{sep}"])
        np.testing.assert_array_equal(s2, c[f"gp.s2{sep}"])
