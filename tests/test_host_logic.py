"""Host-side drivers around the hot path (bounds, priors, design, slice sampler) against
outputs of the real reference (tests/golden/host.npz). CPU only: none of this touches the GPU."""
import numpy as np
import pytest

import gpyreg_b200 as g
from gpyreg_b200.covariance_functions import Matern, RationalQuadraticARD, SquaredExponential
from gpyreg_b200.f_min_fill import (f_min_fill, smoothbox_cdf, smoothbox_ppf, smoothbox_student_t_cdf,
                                    smoothbox_student_t_ppf, uuinv)
from gpyreg_b200.isotropic_covariance_functions import MaternIsotropic, SquaredExponentialIsotropic
from gpyreg_b200.mean_functions import ConstantMean, NegativeQuadratic, ZeroMean
from gpyreg_b200.noise_functions import GaussianNoise
from gpyreg_b200.slice_sample import SliceSampler
from tests.conftest import _load

COVS = [lambda: SquaredExponential(), lambda: Matern(1), lambda: Matern(3), lambda: Matern(5),
        lambda: RationalQuadraticARD(), lambda: SquaredExponentialIsotropic(),
        lambda: MaternIsotropic(1), lambda: MaternIsotropic(3), lambda: MaternIsotropic(5)]
MEANS = [ZeroMean, ConstantMean, NegativeQuadratic]


@pytest.fixture(scope="module")
def h():
    return _load("host.npz")


def noise_of(p):
    return GaussianNoise(p[0] == 1, p[1] >= 1, p[1] == 2, p[2] == 1)


def test_plugin_bounds_info(h):
    X, y = h["X"], h["y"]
    for ci, mk in enumerate(COVS):
        for k, v in mk().get_bounds_info(X, y).items():
            np.testing.assert_array_equal(v, h[f"cov{ci}.{k}"], err_msg=f"cov{ci}.{k}")
    for mk in range(3):
        for k, v in MEANS[mk]().get_bounds_info(X, y).items():
            np.testing.assert_array_equal(v, h[f"mean{mk}.{k}"])
    for p in [(1, 0, 0), (1, 2, 0), (1, 1, 1), (0, 2, 1)]:
        for k, v in noise_of(p).get_bounds_info(X, y).items():
            np.testing.assert_array_equal(v, h["noise%d%d%d.%s" % (p + (k,))])


def test_counts_and_names():
    gp = g.GP(3, RationalQuadraticARD(), NegativeQuadratic(), noise_of((1, 2, 1)))
    assert gp._counts() == (5, 4, 7)
    names = [n for n, _ in gp._hyper_info()]
    assert names == ["covariance_log_lengthscale", "covariance_log_outputscale", "covariance_log_shape",
                     "noise_log_scale", "noise_provided_log_multiplier", "noise_rectified_log_multiplier",
                     "mean_const", "mean_location", "mean_log_scale"]
    assert isinstance(MaternIsotropic(3), Matern) and isinstance(SquaredExponentialIsotropic(), SquaredExponential)
    with pytest.raises(ValueError, match="Only degrees 1, 3 and 5"):
        Matern(2)
    hyp = np.arange(16.0)
    d = gp.hyperparameters_to_dict(hyp)
    np.testing.assert_array_equal(gp.hyperparameters_from_dict(d)[0], hyp)
    with pytest.raises(ValueError, match="wrong shape"):
        gp.hyperparameters_to_dict(np.zeros(5))
    with pytest.raises(ValueError, match="Missing hyperparameter"):
        gp.set_bounds({"covariance_log_lengthscale": None})
    assert "Covariance function: RationalQuadraticARD, 5 parameters" in str(gp)
    r = repr(gp)
    assert "self.covariance = <gpyreg_b200.covariance_functions.RationalQuadraticARD object at " in r
    assert "self.lower_bounds = (16,) ndarray" in r and "self.X = None" in r


def _prior_gp(h):
    gp = g.GP(3, SquaredExponentialIsotropic(), ConstantMean(), noise_of((1, 2, 0)))
    gp.X, gp.y = h["X"], h["y"]
    gp.set_bounds(gp.get_recommended_bounds())
    gp.set_priors({
        "covariance_log_lengthscale": ("gaussian", (0.2, 1.5)),
        "covariance_log_outputscale": ("student_t", (0.0, 1.0, 4.0)),
        "noise_log_scale": ("smoothbox", (-4.0, -1.0, 0.7)),
        "noise_provided_log_multiplier": ("smoothbox_student_t", (-0.5, 0.5, 0.4, 3.0)),
        "mean_const": None,
    })
    return gp


def test_bounds_and_priors(h):
    gp = _prior_gp(h)
    np.testing.assert_array_equal(gp.lower_bounds, h["gp.LB"])
    np.testing.assert_array_equal(gp.upper_bounds, h["gp.UB"])
    np.testing.assert_allclose(gp.normalization_constants, h["gp.norm"], rtol=1e-15)
    pri = gp.get_priors()
    assert pri["covariance_log_lengthscale"][0] == "gaussian" and pri["mean_const"] is None
    assert pri["noise_provided_log_multiplier"][0] == "smoothbox_student_t"
    H = h["gp.H"]
    lp, dlp = gp._log_priors_batch(H, True)
    np.testing.assert_allclose(lp, h["gp.lp"], rtol=1e-13)
    np.testing.assert_allclose(dlp, h["gp.dlp"], rtol=1e-13, atol=1e-300)
    lp1, dlp1 = gp._GP__compute_log_priors(H[3], True)
    assert lp1 == lp[3] and np.array_equal(dlp1, dlp[3])
    assert gp._GP__compute_log_priors(H[3], False) == lp[3]


def test_f_min_fill_design(h):
    f = lambda x: float(np.sum((x - 0.3) ** 2))
    hp = {k: h["fmf.hp." + k] for k in ("mu", "sigma", "df", "a", "b")}
    np.random.seed(7)
    X0, y0 = f_min_fill(f, h["fmf.x0"], h["gp.LB"], h["gp.UB"], h["fmf.PLB"], h["fmf.PUB"], hp, 64, "sobol")
    np.testing.assert_allclose(X0, h["fmf.X"], rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(y0, h["fmf.y"], rtol=1e-13)
    # batched objective: one call on the whole design, same result
    calls = []

    def fb(Xall):
        calls.append(Xall.shape)
        return np.sum((Xall - 0.3) ** 2, axis=1)
    fb.batched = True
    np.random.seed(7)
    X1, y1 = f_min_fill(fb, h["fmf.x0"], h["gp.LB"], h["gp.UB"], h["fmf.PLB"], h["fmf.PUB"], hp, 64, "sobol")
    assert calls == [(64, 5)]
    np.testing.assert_allclose(X1, X0, rtol=1e-14, atol=1e-14)
    # no priors at all (uniform mixtures only)
    none = {k: np.full(h["fmf2.LB"].shape, np.nan) for k in ("mu", "sigma", "df", "a", "b")}
    x0 = np.reshape((h["fmf2.PLB"] + h["fmf2.PUB"]) / 2, (1, -1))
    np.random.seed(8)
    X2, y2 = f_min_fill(f, x0, h["fmf2.LB"], h["fmf2.UB"], h["fmf2.PLB"], h["fmf2.PUB"], none, 32, "sobol")
    np.testing.assert_allclose(X2, h["fmf2.X"], rtol=1e-14, atol=1e-14)
    with pytest.raises(ValueError, match="Unknown design"):
        f_min_fill(f, x0, h["fmf2.LB"], h["fmf2.UB"], h["fmf2.PLB"], h["fmf2.PUB"], none, 32, "grid")


def test_smoothbox_helpers(h):
    q = np.linspace(0.01, 0.99, 9)
    np.testing.assert_allclose(uuinv(q, [-3.0, -1.0, 2.0, 5.0], 0.7), h["uuinv"], rtol=1e-15)
    got = np.array([[smoothbox_cdf(v, 0.7, -1.0, 2.0), smoothbox_student_t_cdf(v, 3.0, 0.7, -1.0, 2.0)]
                    for v in (-2.5, -1.0, 0.3, 2.0, 4.0)])
    np.testing.assert_allclose(got, h["sb"], rtol=1e-15)
    got = np.array([[smoothbox_ppf(v, 0.7, -1.0, 2.0), smoothbox_student_t_ppf(v, 3.0, 0.7, -1.0, 2.0)]
                    for v in q])
    np.testing.assert_allclose(got, h["sbppf"], rtol=1e-14)
    for v in (-2.0, 0.5, 3.0):      # cdf / ppf round trip
        assert smoothbox_ppf(smoothbox_cdf(v, 0.7, -1.0, 2.0), 0.7, -1.0, 2.0) == pytest.approx(v, abs=1e-10)


def test_slice_sampler_matches_reference_chain(h):
    logp = lambda x: float(-0.5 * (x[0] ** 2 + (x[1] - 0.5 * x[0]) ** 2 / 0.25 + x[2] ** 2 / 4))
    np.random.seed(9)
    ss = SliceSampler(logp, np.array([0.1, 0.2, -0.3]), np.array([1.0, 1.0, 2.0]),
                      np.array([-3.0, -3.0, -5.0]), np.array([3.0, 3.0, 5.0]),
                      {"display": "off", "diagnostics": False})
    res = ss.sample(40, thin=2, burn=30)
    np.testing.assert_array_equal(res["samples"], h["ss.samples"])
    np.testing.assert_array_equal(res["f_vals"], h["ss.f_vals"])
    np.testing.assert_array_equal(ss.widths, h["ss.widths"])
    assert ss.func_count == int(h["ss.func_count"])


def test_speculative_slice_sampler_is_the_same_chain(h):
    """Batched speculative shrinking must reproduce the sequential chain bit for bit (and so
    the reference's), with fewer calls."""
    logp = lambda x: float(-0.5 * (x[0] ** 2 + (x[1] - 0.5 * x[0]) ** 2 / 0.25 + x[2] ** 2 / 4))
    sizes = []

    def logp_batch(X):
        sizes.append(X.shape[0])
        return np.array([logp(x) for x in X])
    for k in (2, 3, 5, np.array([1, 4, 2])):                  # one depth, or one per coordinate
        np.random.seed(9)
        del sizes[:]
        ss = SliceSampler(logp, np.array([0.1, 0.2, -0.3]), np.array([1.0, 1.0, 2.0]),
                          np.array([-3.0, -3.0, -5.0]), np.array([3.0, 3.0, 5.0]),
                          {"display": "off", "diagnostics": False, "log_f_batch": logp_batch, "speculate": k})
        res = ss.sample(40, thin=2, burn=30)
        np.testing.assert_array_equal(res["samples"], h["ss.samples"])
        np.testing.assert_array_equal(res["f_vals"], h["ss.f_vals"])
        np.testing.assert_array_equal(ss.widths, h["ss.widths"])
        assert ss.func_count == int(h["ss.func_count"])        # evaluations the sequential sampler counts
        assert ss.batch_calls < ss.func_count and max(sizes) <= np.max(k)
        # the histogram of proposals per coordinate update accounts for every evaluation but the first
        assert np.sum(ss.shrink_counts * np.arange(65)) >= ss.func_count - 1
        assert np.sum(ss.shrink_counts) == 3 * (40 + 39 + 30)        # one entry per coordinate update
    # the RNG stream ends where the sequential sampler leaves it
    np.random.seed(9)
    SliceSampler(logp, np.array([0.1, 0.2, -0.3]), np.array([1.0, 1.0, 2.0]), np.array([-3.0, -3.0, -5.0]),
                 np.array([3.0, 3.0, 5.0]), {"display": "off", "diagnostics": False}).sample(40, thin=2, burn=30)
    after_seq = np.random.rand()
    np.random.seed(9)
    SliceSampler(logp, np.array([0.1, 0.2, -0.3]), np.array([1.0, 1.0, 2.0]), np.array([-3.0, -3.0, -5.0]),
                 np.array([3.0, 3.0, 5.0]), {"display": "off", "diagnostics": False, "log_f_batch": logp_batch,
                                             "speculate": 3}).sample(40, thin=2, burn=30)
    assert np.random.rand() == after_seq


def test_slice_sampler_argument_checks():
    f = lambda x: -0.5 * float(np.sum(x ** 2))
    with pytest.raises(ValueError, match="outside the bounds"):
        SliceSampler(f, np.array([5.0]), None, np.array([-1.0]), np.array([1.0]))
    with pytest.raises(ValueError, match="equal or greater"):
        SliceSampler(f, np.array([0.0]), None, np.array([1.0]), np.array([-1.0]))
    with pytest.raises(ValueError, match="positive real"):
        SliceSampler(f, np.array([0.0]), np.array([-1.0]), np.array([-1.0]), np.array([1.0]))
    ss = SliceSampler(lambda x: np.nan, np.array([0.0]), np.array([1.0]), np.array([-1.0]), np.array([1.0]),
                      {"display": "off"})
    with pytest.raises(ValueError, match="real number"):
        ss.sample(5)
    np.random.seed(0)
    ss = SliceSampler(f, np.zeros(2), None, -4 * np.ones(2), 4 * np.ones(2), {"display": "off"})
    res = ss.sample(2000, burn=200)
    assert abs(res["samples"].mean()) < 0.15 and abs(res["samples"].std() - 1) < 0.15
    assert res["exit_flag"] in (1, -1)


def test_convert_shapes_and_posterior_record():
    gp = g.GP(2, SquaredExponential(), ZeroMean(), noise_of((1, 0, 0)))
    X, y, s2 = gp._convert_shapes(np.zeros(2), np.zeros(1), 0.5)
    assert X.shape == (1, 2) and y.shape == (1, 1) and s2.shape == (1, 1)
    with pytest.raises(AssertionError):
        gp._convert_shapes(np.zeros((3, 5)), None, None)
    with pytest.raises(TypeError):
        gp._convert_shapes(np.zeros((3, 2)), None, "x")
    gp.update(hyp=np.zeros((2, 4)), compute_posterior=False)     # hyp-only posteriors, no GPU
    assert gp.posteriors.size == 2 and gp.posteriors[0].alpha is None
    np.testing.assert_array_equal(gp.get_hyperparameters(as_array=True), np.zeros((2, 4)))
    p = g.Posterior(np.zeros(3), 1, 2, 3, 4, True)
    assert (p.alpha, p.sW, p.L, p.sn2_mult, p.L_chol) == (1, 2, 3, 4, True)
    gp.clean()
    assert gp.posteriors[0].L is None


def test_rank_one_update_host_logic():
    """update() with one new point: device append when it applies; samples that fail the reference's
    stability test (gaussian_process.py:790-798) are recomputed alone on the extended data (:864-868);
    when the library declines, all samples are rebuilt with their CURRENT hyperparameters -- the ``hyp``
    argument is ignored on this branch."""
    from gpyreg_b200.gaussian_process import Posterior

    class FakeEngine:
        def __init__(self, status):
            self.status, self.calls, self.rebuilt = status, [], []

        def posterior_append(self, post, x_new, y_new):
            self.calls.append((np.array(x_new), y_new))
            if self.status is not None:
                post.N += 1
            return self.status

        def posterior_rebuild(self, post, slots):
            self.rebuilt.append(list(slots))

    class FakeBatch:
        def __init__(self, engine):
            self.engine, self._h, self.N, self.count = engine, 1, 5, 2

    def make(status):
        gp = g.GP(2, SquaredExponential(), ConstantMean(), GaussianNoise(constant_add=True))
        gp.X, gp.y = np.zeros((5, 2)), np.zeros((5, 1))
        batch = FakeBatch(FakeEngine(status))
        gp._post_batch = batch
        gp.posteriors = np.empty((2,), dtype=object)
        for s in range(2):
            gp.posteriors[s] = Posterior(np.full(5, float(s)), None, None, None, None, None, _batch=batch, _index=s)
            gp.posteriors[s]._set("alpha", np.zeros((5, 1)))
        rebuilt = []

        def fake_posteriors_for(hyp):
            rebuilt.append(hyp.copy())
            return gp.posteriors, batch

        gp._posteriors_for = fake_posteriors_for
        gp._sync_engine = lambda: batch.engine
        return gp, batch, rebuilt

    x, y = np.array([[1.0, 2.0]]), np.array([[3.0]])
    # stable append: no rebuild, cached fields invalidated, data grown
    gp, batch, rebuilt = make(np.zeros(2, dtype=np.int32))
    gp.update(X_new=x, y_new=y, hyp=np.full((1, 5), 9.0))
    assert rebuilt == [] and batch.N == 6 and gp.X.shape == (6, 2) and gp.y.shape == (6, 1)
    assert not gp.posteriors[0]._have["alpha"] and len(batch.engine.calls) == 1
    assert batch.engine.calls[0][1] == 3.0 and np.array_equal(batch.engine.calls[0][0], [1.0, 2.0])
    # unstable sample: warning with the reference's text, THAT sample recomputed on the extended data
    gp, batch, rebuilt = make(np.array([0, 1], dtype=np.int32))
    gp.posteriors[1]._set("sn2_mult", 10)
    with pytest.warns(UserWarning, match="Rank-one update of Cholesky factor unstable for posterior 1"):
        gp.update(X_new=x, y_new=y, hyp=np.full((1, 5), 9.0))
    assert rebuilt == [] and batch.engine.rebuilt == [[1]] and gp.X.shape == (6, 2) and batch.N == 6
    assert not gp.posteriors[1]._have["sn2_mult"] and not gp.posteriors[0]._have["alpha"]
    # declined by the library (GPB_EAGAIN): silent rebuild
    gp, batch, rebuilt = make(None)
    gp.update(X_new=x, y_new=y)
    assert len(rebuilt) == 1 and rebuilt[0].shape == (2, 5)
    # two points at once, or s2 given: never the rank-one branch; ``hyp`` is honoured
    gp, batch, rebuilt = make(np.zeros(2, dtype=np.int32))
    gp.update(X_new=np.zeros((2, 2)), y_new=np.zeros((2, 1)), hyp=np.full((1, 5), 9.0))
    assert batch.engine.calls == [] and np.array_equal(rebuilt[0], np.full((1, 5), 9.0))


def test_rng_rewind_paths(h, monkeypatch):
    """The speculative sampler rewinds the global RNG either by copying the Mersenne-Twister state
    directly (fast path, self-checked) or through np.random.get_state/set_state: both reproduce the
    reference chain, and the fast path leaves the generator exactly where the public API would."""
    from gpyreg_b200 import slice_sample as ssm
    rw = ssm._RngRewind()
    np.random.seed(123)
    ref_state = np.random.get_state()
    rw.save()
    a = np.random.rand(1000)
    rw.restore()
    b = np.random.rand(1000)
    np.random.set_state(ref_state)
    c = np.random.rand(1000)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    # normal draws keep their cached second value across a save / uniform draws / restore
    np.random.seed(7)
    np.random.standard_normal()                  # leaves a cached Gaussian behind
    rw.save()
    np.random.rand(3)
    rw.restore()
    g1 = np.random.standard_normal(3)
    np.random.seed(7)
    np.random.standard_normal()
    g2 = np.random.standard_normal(3)
    assert np.array_equal(g1, g2)
    logp = lambda x: float(-0.5 * (x[0] ** 2 + (x[1] - 0.5 * x[0]) ** 2 / 0.25 + x[2] ** 2 / 4))
    logp_batch = lambda X: np.array([logp(x) for x in X])
    for force_public in (False, True):
        if force_public:
            monkeypatch.setattr(ssm._RngRewind, "_probe", lambda self: False)
        np.random.seed(9)
        ss = SliceSampler(logp, np.array([0.1, 0.2, -0.3]), np.array([1.0, 1.0, 2.0]),
                          np.array([-3.0, -3.0, -5.0]), np.array([3.0, 3.0, 5.0]),
                          {"display": "off", "diagnostics": False, "log_f_batch": logp_batch, "speculate": 3})
        assert ss._rewind._raw == (not force_public)
        res = ss.sample(40, thin=2, burn=30)
        np.testing.assert_array_equal(res["samples"], h["ss.samples"])


def test_dimension_limit_is_reported_at_construction():
    """a D beyond the fused kernels' limit (64) fails with a clear message when the GP is built,
    not at set_data."""
    import gpyreg_b200 as g
    from gpyreg_b200.covariance_functions import SquaredExponential
    from gpyreg_b200.mean_functions import ZeroMean
    from gpyreg_b200.noise_functions import GaussianNoise
    with pytest.raises(ValueError, match="D <= 64"):
        g.GP(65, SquaredExponential(), ZeroMean(), GaussianNoise(constant_add=True))
    g.GP(64, SquaredExponential(), ZeroMean(), GaussianNoise(constant_add=True))


def test_gp_copies_and_pickles_without_device_state():
    """copy.deepcopy / pickle of a GP (PyVBMC does both) drop the device handles and keep the rest."""
    import copy
    import pickle
    import gpyreg_b200 as g
    from gpyreg_b200.covariance_functions import Matern
    from gpyreg_b200.mean_functions import ConstantMean
    from gpyreg_b200.noise_functions import GaussianNoise
    gp = g.GP(2, Matern(3), ConstantMean(), GaussianNoise(constant_add=True))
    gp.X, gp.y = np.zeros((4, 2)), np.zeros((4, 1))
    gp.update(hyp=np.arange(10.0).reshape(2, 5), compute_posterior=False)
    gp.temporary_data["k"] = 1
    for other in (copy.deepcopy(gp), pickle.loads(pickle.dumps(gp))):
        assert other._engine is None and other._post_batch is None and other._token is not gp._token
        assert other.X is not gp.X and np.array_equal(other.X, gp.X)
        assert np.array_equal(other.get_hyperparameters(as_array=True), gp.get_hyperparameters(as_array=True))
        assert other.temporary_data == {"k": 1} and isinstance(other.covariance, Matern) and other.covariance.degree == 3
    p = copy.deepcopy(gp.posteriors)[1]
    assert np.array_equal(p.hyp, gp.posteriors[1].hyp) and p.alpha is None and p._batch is None
