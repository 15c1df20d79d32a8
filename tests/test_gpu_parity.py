"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
golden vectors produced by the real reference.  Tolerances are the ones BASELINE.json's
north_star states: nlZ 1e-9 relative, gradients 1e-7, predictive mean/variance 1e-8."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import gp_oracle as orc
from tests.helpers import COV_TAGS, CORE_TAGS, case, grad_err, rel_err, spec_from_array

pytestmark = pytest.mark.gpu

TOL_NLZ = 1e-9
TOL_GRAD = 1e-7
TOL_PRED = 1e-8


@pytest.fixture(scope="module")
def eng():
    from gpyreg_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def setup_engine(eng, spec, X, y, s2):
    eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    eng.set_data(X, y, s2)


# ---------------------------------------------------------------- building blocks
@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (128, 128, 128), (256, 384, 512), (384, 128, 1040)])
def test_gemm_nt(eng, M, N, K):
    rng = np.random.default_rng(M + N + K)
    A, B, Cm = rng.standard_normal((M, K)), rng.standard_normal((N, K)), rng.standard_normal((M, N))
    for alpha, beta in ((1.0, 0.0), (-1.0, 1.0)):
        got = eng.debug_gemm_nt(A, B, Cm, alpha, beta)
        ref = alpha * A @ B.T + beta * Cm
        assert np.max(np.abs(got - ref)) <= 1e-12 * K


@pytest.mark.parametrize("n", [5, 40, 72, 100, 128, 129, 200, 300, 1000])
def test_potrf(eng, n):
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n))
    A = G @ G.T / n + np.eye(n)
    L, info = eng.debug_potrf(A)
    assert info == 0
    ref = sla.cholesky(A, lower=True)
    assert np.max(np.abs(L - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.max(np.abs(L @ L.T - A)) <= 1e-13 * n


def test_potrf_detects_indefinite(eng):
    A = np.eye(200)
    A[150, 150] = -1.0
    _, info = eng.debug_potrf(A)
    assert info == 1
    A[150, 150] = np.nan
    _, info = eng.debug_potrf(A)
    assert info == 1


# ---------------------------------------------------------------- plugin surface
@pytest.mark.parametrize("tag", sorted(COV_TAGS))
def test_cov_plugin(eng, golden_plugins, tag):
    g = golden_plugins
    ck, deg, ard = COV_TAGS[tag]
    X, Xs, hyp = g["X"], g["Xs"], g[f"{tag}.hyp"]
    K, dK = eng.cov(ck, deg, ard, hyp, X, grad=True)
    scale = np.max(np.abs(g[f"{tag}.K"]))
    assert np.max(np.abs(K - g[f"{tag}.K"])) <= 1e-14 * scale
    ref_dK = g[f"{tag}.dK"].transpose(2, 0, 1)
    assert np.array_equal(np.isnan(dK), np.isnan(ref_dK))
    ok = ~np.isnan(ref_dK)
    assert np.max(np.abs(dK[ok] - ref_dK[ok])) <= 1e-13 * max(scale, np.max(np.abs(ref_dK[ok])))
    Kx = eng.cov(ck, deg, ard, hyp, X, Xs=Xs)
    assert np.max(np.abs(Kx - g[f"{tag}.Kx"])) <= 1e-14 * scale
    Kd = eng.cov(ck, deg, ard, hyp, X, diag=True)
    assert np.max(np.abs(Kd - g[f"{tag}.Kd"])) <= 1e-14 * scale


@pytest.mark.parametrize("mk", [0, 1, 2])
def test_mean_plugin(eng, golden_plugins, mk):
    g = golden_plugins
    m, dm = eng.mean(mk, g[f"mean{mk}.hyp"], g["X"], grad=True)
    np.testing.assert_allclose(m, g[f"mean{mk}.m"], rtol=1e-14, atol=1e-14)
    if mk:
        np.testing.assert_allclose(dm, g[f"mean{mk}.dm"], rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("p", [(a, b, c) for a in (0, 1) for b in (0, 1, 2) for c in (0, 1)])
def test_noise_plugin(eng, golden_plugins, p):
    g = golden_plugins
    key = "noise%d%d%d" % p
    N = g["X"].shape[0]
    sn2, dsn2 = eng.noise(p, g[key + ".hyp"], N, g["noise.y"], g["noise.s2"], grad=True)
    ref = np.broadcast_to(g[key + ".sn2"].reshape(-1), (N,)) if g[key + ".sn2"].size > 1 \
        else np.full(N, float(g[key + ".sn2"]))
    np.testing.assert_allclose(sn2, ref, rtol=1e-14)
    if g[key + ".hyp"].size:
        rd = g[key + ".dsn2"]
        rd = np.broadcast_to(rd, (N, rd.shape[1]))
        np.testing.assert_allclose(dsn2, rd, rtol=1e-13, atol=1e-300)


# ---------------------------------------------------------------- nlZ, gradient, posterior, predict
@pytest.mark.parametrize("tag", CORE_TAGS)
def test_core_golden(eng, golden_core, tag):
    c = case(golden_core, tag)
    spec = spec_from_array(c["spec"])
    X, y, s2 = c["X"], c["y"], c.get("s2")
    setup_engine(eng, spec, X, y, s2)
    nlz, dnlz, mult, status = eng.nlz_batch(c["hyp"], want_grad=True)
    assert not status.any()
    np.testing.assert_array_equal(mult, c["sn2_mult"])
    assert rel_err(nlz, c["nlZ"]) <= TOL_NLZ
    assert grad_err(dnlz, c["dnlZ"]) <= TOL_GRAD
    nlz0, _, _, _ = eng.nlz_batch(c["hyp"], want_grad=False)
    np.testing.assert_array_equal(nlz0, nlz)       # same factorisation, bit-identical
    # posteriors
    post = eng.posterior_batch(c["hyp"])
    for b in range(post.count):
        al = post.fetch(b, "alpha")
        assert np.max(np.abs(al - c["alpha"][b])) <= 1e-9 * np.max(np.abs(c["alpha"][b]))
        assert post.fetch(b, "sW") == pytest.approx(c["sW"][b], rel=1e-14)
        assert post.fetch(b, "sn2_mult") == c["sn2_mult"][b]
        assert int(post.fetch(b, "L_chol")) == c["L_chol"][b]
        if "L" in c:
            L = post.fetch(b, "L")
            assert np.max(np.abs(L - c["L"][b])) <= 1e-11 * np.max(np.abs(c["L"][b]))
    # predictions, every flag combination
    for add_noise in (0, 1):
        for sep in (0, 1):
            mu, v, lpd = eng.predict(post, c["Xs"], c["ys"], c.get("s2s"), add_noise=bool(add_noise),
                                     separate=bool(sep), want_lpd=True)
            k = f"pred{add_noise}{sep}"
            sc = np.max(np.abs(c[k + ".mu"])) + 1.0
            assert np.max(np.abs(mu - c[k + ".mu"])) <= TOL_PRED * sc
            vs = np.max(np.abs(c[k + ".s2"]))
            assert np.max(np.abs(v - c[k + ".s2"])) <= TOL_PRED * vs
            assert np.max(np.abs(lpd - c[k + ".lpd"])) <= 1e-7 * (1 + np.max(np.abs(c[k + ".lpd"])))
    post.free()


@pytest.mark.parametrize("tag", ["eps", "thr"])
def test_lownoise_golden(eng, golden_lownoise, tag):
    """Low-noise branch (min sn2 < 1e-6) and the x10 jitter retry, SURVEY.md section 7."""
    c = case(golden_lownoise, tag)
    spec = spec_from_array(c["spec"])
    X, y = c["X"], c["y"]
    setup_engine(eng, spec, X, y, None)
    nlz, dnlz, mult, status = eng.nlz_batch(c["hyp"], want_grad=True)
    assert not status.any()
    same = mult == c["sn2_mult"]
    # Rows the reference factors at the first attempt must agree exactly.  Rows whose matrix
    # is numerically singular (K + 2.2e-16*I; the reference needs x100 jitter) succeed or fail
    # on rounding alone, so a blocked GPU Cholesky may take a different number of retries than
    # LAPACK: there the multiplier only has to stay a small power of ten.
    easy = c["sn2_mult"] == 1
    assert same[easy].all()
    assert np.all(np.isin(mult[~easy], [1.0, 10.0, 100.0, 1e3, 1e4])), mult
    assert np.all(np.isfinite(nlz))
    post = eng.posterior_batch(c["hyp"])
    for b in np.nonzero(same)[0]:
        assert int(post.fetch(b, "L_chol")) == c["L_chol"][b]
    post.free()
    # values (nlZ, gradient, alpha, predictions) of these rows, with a per-row bound derived from
    # cond(A): tests/test_gpu_parity_full.py::test_lownoise_small_golden_gradient_and_predictions


def test_wrong_hyperparameter_count_is_rejected(eng):
    spec = orc.ModelSpec(D=2, cov_kind=0, ard=True, mean_kind=1)
    setup_engine(eng, spec, np.zeros((5, 2)) + np.arange(5)[:, None], np.arange(5.0), None)
    with pytest.raises(ValueError, match="must have 5 entries"):
        eng.nlz_batch(np.zeros((2, 4)))
    with pytest.raises(ValueError, match="must have 5 entries"):
        eng.posterior_batch(np.zeros(7))


def test_cholesky_failure_status(eng):
    """A NaN hyperparameter makes every attempt fail: status 1, nlZ NaN, others unaffected."""
    rng = np.random.default_rng(5)
    N, D = 40, 2
    X = rng.uniform(-3, 3, (N, D))
    y = np.sin(X.sum(1))
    spec = orc.ModelSpec(D=D, cov_kind=0, ard=True, mean_kind=1, noise_params=(1, 0, 0))
    setup_engine(eng, spec, X, y, None)
    hyp = np.array([[0.2, 0.1, 0.0, np.log(0.1), 0.0], [np.nan, 0.1, 0.0, np.log(0.1), 0.0],
                    [0.3, 0.0, 0.1, np.log(0.2), 0.1]])
    nlz, _, mult, status = eng.nlz_batch(hyp)
    assert list(status) == [0, 1, 0]
    assert np.isnan(nlz[1])
    ref = orc.nlz_batch(spec, hyp[[0, 2]], X, y.reshape(-1, 1), None, False)
    assert rel_err(nlz[[0, 2]], ref) <= TOL_NLZ


@pytest.mark.parametrize("model", ["cfg2", "cfg3", "cfg4", "cfg5"])
def test_oracle_medium(eng, model):
    """Seeded medium-size problems (several 128-tiles) against the CPU oracle."""
    rng = np.random.default_rng(11)
    N, D, B = {"cfg2": (700, 6, 5), "cfg3": (900, 10, 4), "cfg4": (520, 8, 3), "cfg5": (640, 10, 4)}[model]
    spec = {
        "cfg2": orc.ModelSpec(D=D, cov_kind=0, ard=True, mean_kind=1),
        "cfg3": orc.ModelSpec(D=D, cov_kind=1, degree=5, ard=True, mean_kind=2),
        "cfg4": orc.ModelSpec(D=D, cov_kind=2, ard=True, mean_kind=1),
        "cfg5": orc.ModelSpec(D=D, cov_kind=1, degree=3, ard=False, mean_kind=1),
    }[model]
    from bench import benign_hyp, synth_data
    X, y = synth_data(N, D, seed=0)
    hyp = benign_hyp(spec, B, y, seed=1)
    setup_engine(eng, spec, X, y, None)
    nlz, dnlz, mult, status = eng.nlz_batch(hyp, want_grad=True)
    ref_nlz, ref_dnlz = orc.nlz_batch(spec, hyp, X, y, None, True)
    assert not status.any() and np.all(mult == 1)
    assert rel_err(nlz, ref_nlz) <= TOL_NLZ
    assert grad_err(dnlz, ref_dnlz) <= TOL_GRAD
    # batch composition must not change a single element's result (determinism)
    nlz1, dnlz1, _, _ = eng.nlz_batch(hyp[1:2], want_grad=True)
    assert nlz1[0] == nlz[1] and np.array_equal(dnlz1[0], dnlz[1])
    # predictions
    post = eng.posterior_batch(hyp)
    Xs = np.random.default_rng(2).uniform(-3, 3, (333, D))
    posts = orc.posterior_batch(spec, hyp, X, y, None)
    for sep in (False, True):
        mu, v = eng.predict(post, Xs, add_noise=True, separate=sep)
        rmu, rv = orc.predict(spec, posts, X, y, Xs, add_noise=True, separate_samples=sep)
        assert np.max(np.abs(mu - rmu)) <= TOL_PRED * (1 + np.max(np.abs(rmu)))
        assert np.max(np.abs(v - rv)) <= TOL_PRED * np.max(np.abs(rv))
    post.free()


def test_large_rq_single_gp(eng):
    """config 4 shape at a size the oracle finishes in seconds: one RationalQuadratic-ARD GP,
    D=8, N not a multiple of the 128-tile (exercises the padding-aware tile kernels)."""
    from bench import benign_hyp, synth_data
    N, D = 2500, 8
    spec = orc.ModelSpec(D=D, cov_kind=2, ard=True, mean_kind=1)
    X, y = synth_data(N, D, seed=0)
    hyp = benign_hyp(spec, 2, y, seed=1)
    setup_engine(eng, spec, X, y, None)
    nlz, dnlz, mult, status = eng.nlz_batch(hyp, want_grad=True)
    ref_nlz, ref_dnlz = orc.nlz_batch(spec, hyp[:1], X, y, None, True)
    assert not status.any() and np.all(mult == 1)
    assert rel_err(nlz[:1], ref_nlz) <= TOL_NLZ
    assert grad_err(dnlz[:1], ref_dnlz) <= TOL_GRAD
    # size-independent property: nlZ is invariant under a permutation of the data rows
    perm = np.random.default_rng(3).permutation(N)
    setup_engine(eng, spec, X[perm], y[perm], None)
    nlz_p, dnlz_p, _, _ = eng.nlz_batch(hyp, want_grad=True)
    assert rel_err(nlz_p, nlz) <= 1e-11
    assert grad_err(dnlz_p, dnlz) <= 1e-9


@pytest.mark.parametrize("N", [1, 2, 127, 128, 129, 256])
def test_tile_boundaries(eng, N):
    """Sizes around the tile edge, including a single training point."""
    rng = np.random.default_rng(N)
    D = 2
    X = rng.uniform(-2, 2, (N, D))
    y = np.sin(X.sum(1)).reshape(-1, 1) + 0.05 * rng.standard_normal((N, 1))
    spec = orc.ModelSpec(D=D, cov_kind=1, degree=3, ard=True, mean_kind=1)
    hyp = np.array([[0.1, -0.2, 0.0, np.log(0.2), 0.1], [0.4, 0.3, 0.2, np.log(0.05), -0.1]])
    setup_engine(eng, spec, X, y, None)
    nlz, dnlz, mult, status = eng.nlz_batch(hyp, want_grad=True)
    ref_nlz, ref_dnlz = orc.nlz_batch(spec, hyp, X, y, None, True)
    assert rel_err(nlz, ref_nlz) <= TOL_NLZ and grad_err(dnlz, ref_dnlz) <= TOL_GRAD
    post = eng.posterior_batch(hyp)
    Xs = rng.uniform(-2, 2, (5, D))
    mu, v = eng.predict(post, Xs, add_noise=True, separate=True)
    rmu, rv = orc.predict(spec, orc.posterior_batch(spec, hyp, X, y, None), X, y, Xs, add_noise=True,
                          separate_samples=True)
    assert np.max(np.abs(mu - rmu)) <= TOL_PRED * (1 + np.max(np.abs(rmu)))
    assert np.max(np.abs(v - rv)) <= TOL_PRED * np.max(np.abs(rv))
    post.free()


def test_batch_chunking_and_workspace_limit(eng):
    """A workspace cap smaller than the batch forces several chunks: same results, bit for bit."""
    from bench import benign_hyp, synth_data
    N, D, B = 300, 3, 13
    spec = orc.ModelSpec(D=D, cov_kind=0, ard=True, mean_kind=1)
    X, y = synth_data(N, D, seed=0)
    hyp = benign_hyp(spec, B, y, seed=1)
    setup_engine(eng, spec, X, y, None)
    full = eng.nlz_batch(hyp, want_grad=True)
    from gpyreg_b200 import Engine
    e2 = Engine(0)
    e2.set_workspace_limit(4 * 2 * 384 * 384 * 8 + (1 << 20))        # room for ~4 slots
    e2.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    e2.set_data(X, y, None)
    part = e2.nlz_batch(hyp, want_grad=True)
    np.testing.assert_array_equal(part[0], full[0])
    np.testing.assert_array_equal(part[1], full[1])
    e2.close()


def test_predict_many_points_and_wide_inputs(eng):
    """More test points than one chunk (8192) and a wide input (D = 32)."""
    rng = np.random.default_rng(9)
    N, D = 150, 32
    X = rng.uniform(-1, 1, (N, D))
    y = np.sin(X[:, :3].sum(1)).reshape(-1, 1)
    spec = orc.ModelSpec(D=D, cov_kind=0, ard=True, mean_kind=2)
    hyp = np.concatenate([np.log(2.0) + 0.1 * rng.standard_normal((2, D)), np.zeros((2, 1)),
                          np.full((2, 1), np.log(0.1)), np.zeros((2, 1)), 0.1 * rng.standard_normal((2, D)),
                          np.full((2, D), np.log(3.0))], axis=1)
    setup_engine(eng, spec, X, y, None)
    nlz, dnlz, _, status = eng.nlz_batch(hyp, want_grad=True)
    ref_nlz, ref_dnlz = orc.nlz_batch(spec, hyp, X, y, None, True)
    assert not status.any() and rel_err(nlz, ref_nlz) <= TOL_NLZ and grad_err(dnlz, ref_dnlz) <= TOL_GRAD
    post = eng.posterior_batch(hyp)
    Xs = rng.uniform(-1, 1, (20000, D))
    mu, v = eng.predict(post, Xs)
    rmu, rv = orc.predict(spec, orc.posterior_batch(spec, hyp, X, y, None), X, y, Xs)
    assert np.max(np.abs(mu - rmu)) <= TOL_PRED * (1 + np.max(np.abs(rmu)))
    assert np.max(np.abs(v - rv)) <= TOL_PRED * np.max(np.abs(rv))
    post.free()
    with pytest.raises(Exception, match="D > 64"):
        eng.set_data(np.zeros((4, 65)), np.zeros(4), None)


def test_runs_on_a_torch_stream(eng):
    """gpb_set_stream: work is enqueued on the caller's stream (what bench.py times)."""
    import torch
    from bench import benign_hyp, synth_data
    spec = orc.ModelSpec(D=2, cov_kind=0, ard=True, mean_kind=0)
    X, y = synth_data(200, 2, seed=0)
    hyp = benign_hyp(spec, 3, y, seed=1)
    setup_engine(eng, spec, X, y, None)
    ref = eng.nlz_batch(hyp, want_grad=True)
    s = torch.cuda.Stream()
    eng.set_stream(s.cuda_stream)
    d_hyp = torch.from_numpy(hyp).cuda()
    d_nlz = torch.empty(3, dtype=torch.float64, device="cuda")
    d_dnlz = torch.empty((3, hyp.shape[1]), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    eng.nlz_batch_dev(d_hyp.data_ptr(), 3, True, d_nlz.data_ptr(), d_dnlz.data_ptr())
    s.synchronize()
    np.testing.assert_array_equal(d_nlz.cpu().numpy(), ref[0])
    np.testing.assert_array_equal(d_dnlz.cpu().numpy(), ref[1])
    eng.set_stream(0)


def test_full_size_properties(eng):
    """BASELINE.json's headline shape (Matern-5 ARD + NegativeQuadratic, N=5000, D=10), where
    the CPU oracle needs ~6 s per row: size-independent properties instead of a point-wise oracle.
      * nlZ and its gradient do not depend on the order of the data rows;
      * the gradient is the derivative of nlZ (central difference along a random direction);
      * a row's result does not depend on what else is in the batch;
      * predictive variances lie in [0, sf^2] and the across-sample average follows :1793-1798."""
    from bench import benign_hyp, synth_data
    N, D = 5000, 10
    spec = orc.ModelSpec(D=D, cov_kind=1, degree=5, ard=True, mean_kind=2)
    X, y = synth_data(N, D, seed=0)
    hyp = benign_hyp(spec, 3, y, seed=1)
    setup_engine(eng, spec, X, y, None)
    nlz, dnlz, mult, status = eng.nlz_batch(hyp, want_grad=True)
    assert not status.any() and np.all(mult == 1) and np.all(np.isfinite(dnlz))
    rng = np.random.default_rng(4)
    v = rng.standard_normal(hyp.shape[1])
    v /= np.linalg.norm(v)
    eps = 1e-5
    pm = eng.nlz_batch(np.stack([hyp[0] + eps * v, hyp[0] - eps * v]), want_grad=False)[0]
    fd = (pm[0] - pm[1]) / (2 * eps)
    assert abs(fd - dnlz[0] @ v) <= 1e-6 * max(1.0, abs(fd))
    one = eng.nlz_batch(hyp[2:3], want_grad=True)
    assert one[0][0] == nlz[2] and np.array_equal(one[1][0], dnlz[2])
    post = eng.posterior_batch(hyp)
    Xs = rng.uniform(-3, 3, (1000, D))
    mu_s, s2_s = eng.predict(post, Xs, separate=True)
    mu_a, s2_a = eng.predict(post, Xs, separate=False)
    sf2 = np.exp(2 * hyp[:, D])
    assert np.all(s2_s >= 0) and np.all(s2_s <= sf2[None, :] * (1 + 1e-12))
    mbar = mu_s.mean(1, keepdims=True)
    np.testing.assert_allclose(mu_a, mbar, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(s2_a[:, 0], s2_s.mean(1) + ((mu_s - mbar) ** 2).sum(1) / 2, rtol=1e-10)
    post.free()
    perm = rng.permutation(N)
    setup_engine(eng, spec, X[perm], y[perm], None)
    nlz_p, dnlz_p, _, _ = eng.nlz_batch(hyp[:1], want_grad=True)
    assert rel_err(nlz_p, nlz[:1]) <= 1e-11 and grad_err(dnlz_p, dnlz[:1]) <= 1e-9


def test_factor_cache_is_bit_identical(eng):
    """nlZ-only calls re-use the cached Cholesky factor when only MEAN hyperparameters changed
    (a slice sampler moves one coordinate at a time) and replay the O(N^2) forward solve: the
    value must equal a from-scratch evaluation bit for bit, for mean-only, noise and covariance
    moves, in batches, and after gradient calls invalidate the cache."""
    from bench import benign_hyp, synth_data
    from gpyreg_b200 import Engine
    N, D = 700, 4
    spec = orc.ModelSpec(D=D, cov_kind=1, degree=5, ard=True, mean_kind=2)
    X, y = synth_data(N, D, seed=0)
    base = benign_hyp(spec, 3, y, seed=1)
    rng = np.random.default_rng(0)
    cov_n, noise_n = spec.cov_n, spec.noise_n
    seq = [base.copy()]
    for step in range(6):
        h = seq[-1].copy()
        # mean location, mean location, NOISE, a length scale (COVARIANCE), mean constant, mean scale
        col = [cov_n + noise_n + 1, cov_n + noise_n + 3, cov_n, 2, cov_n + noise_n, h.shape[1] - 1][step]
        h[:, col] += 0.05 * rng.standard_normal(3)
        if step == 1:
            h[1] = seq[-1][1]                  # one row of the batch unchanged
        seq.append(h)
    setup_engine(eng, spec, X, y, None)
    h0, m0 = eng.cache_stats()
    got = [eng.nlz_batch(h)[0] for h in seq]
    h1, m1 = eng.cache_stats()
    fresh = Engine(0)
    fresh.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    fresh.set_data(X, y, None)
    for h, g in zip(seq, got):
        ref = fresh.nlz_batch(h, want_grad=True)[0]        # gradient calls never use the cache
        np.testing.assert_array_equal(g, ref)
    fresh.close()
    # moves 1, 2, 5, 6 touch mean hyperparameters only: 4 x 3 rows re-use the factor
    assert h1 - h0 == 12 and m1 - m0 == 9
    ref_nlz = orc.nlz_batch(spec, seq[2], X, y, None, False)
    assert rel_err(got[2], ref_nlz) <= TOL_NLZ
    # a gradient call overwrites the factors with the inverse: the next nlZ call must not hit
    eng.nlz_batch(seq[-1], want_grad=True)
    again = eng.nlz_batch(seq[-1])[0]
    np.testing.assert_array_equal(again, got[-1])
    assert eng.cache_stats()[0] == h1


_TOGGLE_SCRIPT = r"""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from bench import benign_hyp, synth_data
from gpyreg_b200 import Engine
from oracle import gp_oracle as orc
N, D = 900, 5
spec = orc.ModelSpec(D=D, cov_kind=1, degree=3, ard=True, mean_kind=2)
X, y = synth_data(N, D, seed=0)
hyp = benign_hyp(spec, 3, y, seed=1)
eng = Engine(0)
eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
eng.set_data(X, y, None)
nlz, dnlz, _, _ = eng.nlz_batch(hyp, want_grad=True)
nlz1 = eng.nlz_batch(hyp[:1])[0]
print(json.dumps({"nlz": [v.hex() for v in nlz], "nlz1": nlz1[0].hex(), "dnlz": [v.hex() for v in dnlz.ravel()]}))
"""


_TOGGLE_CACHE = {}


@pytest.mark.parametrize("env,exact", [
    ({"GPB_LOOKAHEAD": "0"}, True),                        # no second stream
    ({"GPB_LOOKAHEAD": "2", "GPB_LA_OB": "1"}, True),      # look-ahead with one-column outer blocks
    ({"GPB_OUTER_BLOCK": "3", "GPB_LOOKAHEAD": "0"}, True),
    ({"GPB_LOADER": "tma"}, True),                         # TMA bulk-copy loader for every launch
    ({"GPB_LOADER": "cpasync"}, True),
    ({"GPB_LOADER": "tensor"}, True),                      # tensor-map TMA for every launch
    ({"GPB_LOADER": "auto_bulk"}, True),                   # round-1 default
    ({"GPB_QUARTER": "0"}, True),                          # no 64x64 CTAs for the small trailing updates
    ({"GPB_PDL": "0"}, True),                              # ordinary launches instead of programmatic dependent launch
    ({"GPB_LEFT": "1", "GPB_LOOKAHEAD": "0"}, True),       # left-looking column updates inside the outer blocks
    ({"GPB_LEFT": "1", "GPB_LA_OB": "4", "GPB_LA_WIDE": "0"}, True),   # ... with look-ahead and wide blocks
    ({"GPB_GRAD_ROWS": "0"}, False),                       # gradient by the four-rows-per-thread kernel
    ({"GPB_TRTRI": "0"}, False),                           # column-recurrence triangular inverse
    ({"GPB_GEMM_BN": "128"}, False),                       # one CTA per tile: other reduction shapes
])
def test_schedule_toggles_do_not_change_results(env, exact):
    """Every scheduling variant (streams, outer block, loader) performs the same FP64 operations
    in the same order per element: results are bit-identical.  The alternative algorithms kept
    behind environment switches agree to rounding."""
    import json
    import os
    import subprocess
    import sys

    def run(extra):
        e = dict(os.environ)
        for k in ("GPB_LOOKAHEAD", "GPB_LA_OB", "GPB_LA_WIDE", "GPB_OUTER_BLOCK", "GPB_LOADER", "GPB_TRTRI", "GPB_GEMM_BN",
                  "GPB_PDL", "GPB_QUARTER", "GPB_LEFT", "GPB_GRAD_ROWS"):
            e.pop(k, None)
        e.update(extra)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        out = subprocess.run([sys.executable, "-c", _TOGGLE_SCRIPT], cwd=root, env=e, check=True,
                             capture_output=True, text=True, timeout=600).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        return (np.array([float.fromhex(v) for v in d["nlz"]]), float.fromhex(d["nlz1"]),
                np.array([float.fromhex(v) for v in d["dnlz"]]))

    if "base" not in _TOGGLE_CACHE:
        _TOGGLE_CACHE["base"] = run({})
    base = _TOGGLE_CACHE["base"]
    got = run(env)
    if exact:
        assert np.array_equal(got[0], base[0]) and got[1] == base[1] and np.array_equal(got[2], base[2])
    else:
        assert rel_err(got[0], base[0]) <= 1e-12 and abs(got[1] - base[1]) <= 1e-12 * abs(base[1])
        assert np.max(np.abs(got[2] - base[2])) <= 1e-10 * np.max(np.abs(base[2]))
    assert base[1] == base[0][0]          # a row's value does not depend on its batch


@pytest.mark.parametrize("seed", range(12))
def test_seeded_fuzz_small_batches(eng, seed):
    """Random model / size / batch (1..9 rows: the look-ahead, narrow-block and recursive-inverse paths),
    sizes straddling tile boundaries, against the oracle: nlZ, gradient, posterior mean vector, and a
    rank-one append followed by a prediction."""
    from bench import benign_hyp, synth_data
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.choice([1, 2, 31, 127, 128, 129, 255, 257, 383, 385, 511, 640]))
    D = int(rng.integers(1, 9))
    B = int(rng.choice([1, 2, 3, 5, 9]))
    cov_kind = int(rng.integers(0, 3))
    ard = bool(rng.integers(0, 2)) or cov_kind == 2
    degree = int(rng.choice([1, 3, 5])) if cov_kind == 1 else 0
    # Matern-1 stays Matern-1: its length-scale gradients are NaN by construction and
    # grad_err() requires the NaN pattern to match the oracle's
    spec = orc.ModelSpec(D=D, cov_kind=cov_kind, degree=degree, ard=ard, mean_kind=int(rng.integers(0, 3)))
    X, y = synth_data(N + 1, D, seed=seed)
    Xn, yn, X, y = X[-1], y[-1], X[:-1], y[:-1]
    hyp = benign_hyp(spec, B, y, seed=seed + 1)
    setup_engine(eng, spec, X, y, None)
    nlz, dnlz, mult, status = eng.nlz_batch(hyp, want_grad=True)
    ref_nlz, ref_dnlz = orc.nlz_batch(spec, hyp, X, y, None, True)
    assert not status.any() and np.all(mult == 1)
    assert rel_err(nlz, ref_nlz) <= TOL_NLZ
    assert grad_err(dnlz, ref_dnlz) <= TOL_GRAD
    np.testing.assert_array_equal(eng.nlz_batch(hyp)[0], nlz)          # nlZ-only path: same bits
    post = eng.posterior_batch(hyp)
    posts = orc.posterior_batch(spec, hyp, X, y, None)
    for b in range(B):
        a = post.fetch(b, "alpha")
        ra = np.asarray(posts[b].alpha).reshape(-1)
        assert np.max(np.abs(a - ra)) <= 1e-9 * np.max(np.abs(ra))
    # append the held-out point in place, then predict: compare with the oracle on N + 1 points
    st = eng.posterior_append(post, Xn, float(yn[0]))
    X1, y1 = np.vstack((X, Xn[None, :])), np.concatenate((y, yn[None, :]))
    Xs = rng.uniform(-3, 3, (17, D))
    if st is None:                       # no free row in the padded layout (N a multiple of 128): rebuild
        assert N % 128 == 0
        post.free()
        setup_engine(eng, spec, X1, y1, None)
        post = eng.posterior_batch(hyp)
    else:
        assert not st.any() and post.N == N + 1
    posts1 = orc.posterior_batch(spec, hyp, X1, y1, None)
    mu, v = eng.predict(post, Xs, add_noise=True, separate=True)
    rmu, rv = orc.predict(spec, posts1, X1, y1, Xs, add_noise=True, separate_samples=True)
    assert np.max(np.abs(mu - rmu)) <= TOL_PRED * (1 + np.max(np.abs(rmu)))
    assert np.max(np.abs(v - rv)) <= TOL_PRED * np.max(np.abs(rv))
    post.free()


@pytest.mark.parametrize("D,cov_kind,mean_kind", [(40, 0, 2), (64, 1, 1), (64, 2, 0)])
def test_wide_inputs_up_to_64_dimensions(eng, D, cov_kind, mean_kind):
    """D > 32 (round-1 limit) up to the fused kernels' 64: nlZ, gradient (64 ARD length scales), posterior,
    predictions, rank-one append and quadrature-free paths against the oracle."""
    rng = np.random.default_rng(D + cov_kind)
    N = 260
    X = rng.uniform(-1, 1, (N, D))
    y = (np.sin(X[:, :4].sum(1)) + 0.05 * rng.standard_normal(N)).reshape(-1, 1)
    spec = orc.ModelSpec(D=D, cov_kind=cov_kind, degree=5 if cov_kind == 1 else 0, ard=True, mean_kind=mean_kind)
    B = 2
    cols = [np.log(3.0) + 0.2 * rng.standard_normal((B, D)), 0.1 * rng.standard_normal((B, 1))]
    if cov_kind == 2:
        cols.append(0.2 * rng.standard_normal((B, 1)))
    cols.append(np.full((B, 1), np.log(0.1)))
    if mean_kind >= 1:
        cols.append(0.1 * rng.standard_normal((B, 1)))
    if mean_kind == 2:
        cols += [0.1 * rng.standard_normal((B, D)), np.log(3.0) + 0.1 * rng.standard_normal((B, D))]
    hyp = np.concatenate(cols, axis=1)
    assert hyp.shape[1] == spec.hyp_n
    setup_engine(eng, spec, X[:-1], y[:-1], None)
    nlz, dnlz, mult, status = eng.nlz_batch(hyp, want_grad=True)
    ref_nlz, ref_dnlz = orc.nlz_batch(spec, hyp, X[:-1], y[:-1], None, True)
    assert not status.any() and np.all(mult == 1)
    assert rel_err(nlz, ref_nlz) <= TOL_NLZ and grad_err(dnlz, ref_dnlz) <= TOL_GRAD
    post = eng.posterior_batch(hyp)
    st = eng.posterior_append(post, X[-1], float(y[-1, 0]))
    assert st is not None and not st.any()
    Xs = rng.uniform(-1, 1, (300, D))
    mu, v = eng.predict(post, Xs, add_noise=True, separate=True)
    rmu, rv = orc.predict(spec, orc.posterior_batch(spec, hyp, X, y, None), X, y, Xs, add_noise=True,
                          separate_samples=True)
    assert np.max(np.abs(mu - rmu)) <= TOL_PRED * (1 + np.max(np.abs(rmu)))
    assert np.max(np.abs(v - rv)) <= TOL_PRED * np.max(np.abs(rv))
    post.free()


@pytest.mark.parametrize("cov_kind,degree", [(0, 0), (1, 1), (1, 3), (1, 5), (2, 0)])
@pytest.mark.parametrize("D", [1, 2, 3, 5, 6, 7, 9, 10, 11, 13, 16, 17])
def test_gradient_row_kernel_every_width(eng, D, cov_kind, degree):
    """The row-per-thread gradient kernel is instantiated for 4, 6, 8, 10, 12 and 16 length scales (inputs are
    zero-padded up to the next one; D = 17 takes the four-rows-per-thread kernel): every instantiation and every
    radial function -- with its own branch-free exp / sqrt -- against the oracle's gradient, on a size with a ragged
    last tile (diagonal tiles, trimmed columns, masked rows).  Matern-1 keeps the reference's NaN on the
    length-scale derivatives (inf * 0 on the diagonal)."""
    rng = np.random.default_rng(100 * D + 10 * cov_kind + degree)
    N = 300
    X = rng.uniform(-2, 2, (N, D))
    y = (np.cos(X[:, :2].sum(1)) + 0.1 * rng.standard_normal(N)).reshape(-1, 1)
    spec = orc.ModelSpec(D=D, cov_kind=cov_kind, degree=degree, ard=True, mean_kind=1)
    B = 3
    cols = [np.log(2.0) + 0.4 * rng.standard_normal((B, D)), 0.2 * rng.standard_normal((B, 1))]
    if cov_kind == 2:
        cols.append(0.3 * rng.standard_normal((B, 1)))
    cols += [np.full((B, 1), np.log(0.2)), 0.1 * rng.standard_normal((B, 1))]
    hyp = np.concatenate(cols, axis=1)
    assert hyp.shape[1] == spec.hyp_n
    setup_engine(eng, spec, X, y, None)
    nlz, dnlz, mult, status = eng.nlz_batch(hyp, want_grad=True)
    ref_nlz, ref_dnlz = orc.nlz_batch(spec, hyp, X, y, None, True)
    assert not status.any() and np.all(mult == 1)
    assert rel_err(nlz, ref_nlz) <= TOL_NLZ
    assert np.array_equal(np.isnan(dnlz), np.isnan(ref_dnlz))
    if cov_kind == 1 and degree == 1:
        assert np.isnan(ref_dnlz[:, :D]).all()                 # the reference's own NaN pattern
    ok = ~np.isnan(ref_dnlz)
    assert np.max(np.abs(dnlz[ok] - ref_dnlz[ok])) <= TOL_GRAD * np.max(np.abs(ref_dnlz[ok]))
