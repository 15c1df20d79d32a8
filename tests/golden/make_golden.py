"""Generate golden vectors by running the REAL reference (acerbilab/gpyreg).

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference through the matplotlib stub of SURVEY.md
Appendix B, runs the hot path (plugin ``compute`` methods, ``_GP__compute_nlZ``,
``update`` -> posteriors, ``predict``) on seeded inputs and writes inputs AND
outputs to ``tests/golden/*.npz``.  The .npz files are committed; the tests and
the GPU box never read /root/reference.
"""
import os
import sys
import types

import numpy as np

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, os.environ.get("GPYREG_REFERENCE", "/root/reference"))
import gpyreg  # noqa: E402
from gpyreg.covariance_functions import (Matern, RationalQuadraticARD,  # noqa: E402
                                         SquaredExponential)
from gpyreg.isotropic_covariance_functions import (MaternIsotropic,  # noqa: E402
                                                   SquaredExponentialIsotropic)
from gpyreg.mean_functions import ConstantMean, NegativeQuadratic, ZeroMean  # noqa: E402
from gpyreg.noise_functions import GaussianNoise  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (tag, cov_kind, degree, ard, factory)
COVS = [
    ("se_ard", 0, 0, 1, lambda: SquaredExponential()),
    ("mat1_ard", 1, 1, 1, lambda: Matern(1)),
    ("mat3_ard", 1, 3, 1, lambda: Matern(3)),
    ("mat5_ard", 1, 5, 1, lambda: Matern(5)),
    ("rq_ard", 2, 0, 1, lambda: RationalQuadraticARD()),
    ("se_iso", 0, 0, 0, lambda: SquaredExponentialIsotropic()),
    ("mat1_iso", 1, 1, 0, lambda: MaternIsotropic(1)),
    ("mat3_iso", 1, 3, 0, lambda: MaternIsotropic(3)),
    ("mat5_iso", 1, 5, 0, lambda: MaternIsotropic(5)),
]
MEANS = [lambda: ZeroMean(), lambda: ConstantMean(), lambda: NegativeQuadratic()]


def synth(rng, N, D):
    """SURVEY.md 8(d) synthetic data."""
    X = rng.uniform(-3, 3, (N, D))
    y = (np.sin(X.sum(1)) - 0.125 * (X ** 2).sum(1)
         + 0.1 * rng.standard_normal(N)).reshape(-1, 1)
    return X, y


def benign_hyp(rng, B, D, cov, mean_kind, noise_params, y):
    """SURVEY.md 8(d) 'benign' hyperparameter distribution."""
    cols = []
    nl = D if cov[3] else 1
    cols.append(np.log(1.5) + 0.3 * rng.standard_normal((B, nl)))
    cols.append(0.3 * rng.standard_normal((B, 1)))
    if cov[1] == 2:
        cols.append(0.3 * rng.standard_normal((B, 1)))
    if noise_params[0] == 1:
        cols.append(np.log(0.1) + 0.3 * rng.standard_normal((B, 1)))
    if noise_params[1] == 2:
        cols.append(0.2 * rng.standard_normal((B, 1)))
    if noise_params[2] == 1:
        cols.append(np.median(y) + 0.3 * rng.standard_normal((B, 1)))
        cols.append(np.log(0.05) + 0.2 * rng.standard_normal((B, 1)))
    if mean_kind >= 1:
        cols.append(y.mean() + 0.1 * rng.standard_normal((B, 1)))
    if mean_kind == 2:
        cols.append(0.3 * rng.standard_normal((B, D)))
        cols.append(np.log(3) + 0.2 * rng.standard_normal((B, D)))
    return np.concatenate(cols, axis=1)


def make_gp(D, cov, mean_kind, noise_params):
    noise = GaussianNoise(constant_add=noise_params[0] == 1,
                          user_provided_add=noise_params[1] >= 1,
                          scale_user_provided=noise_params[1] == 2,
                          rectified_linear_output_dependent_add=noise_params[2] == 1)
    return gpyreg.GP(D, cov[4](), MEANS[mean_kind](), noise)


def gen_plugins():
    rng = np.random.default_rng(100)
    N, M, D = 20, 7, 3
    X = rng.uniform(-2, 2, (N, D))
    Xs = rng.uniform(-2, 2, (M, D))
    X[5] = X[2]            # duplicate rows: r = 0 off the diagonal
    out = {"X": X, "Xs": Xs}
    for cov in COVS:
        tag = cov[0]
        c = cov[4]()
        n = c.hyperparameter_count(D)
        hyp = 0.5 * rng.standard_normal(n)
        with np.errstate(all="ignore"):
            K, dK = c.compute(hyp, X, compute_grad=True)
        out[f"{tag}.hyp"] = hyp
        out[f"{tag}.K"] = K
        out[f"{tag}.dK"] = np.ascontiguousarray(dK)
        out[f"{tag}.Kx"] = c.compute(hyp, X, Xs)
        out[f"{tag}.Kd"] = c.compute(hyp, X, compute_diag=True)
    # means
    for mk in (0, 1, 2):
        mobj = MEANS[mk]()
        hyp = rng.standard_normal(mobj.hyperparameter_count(D))
        m, dm = mobj.compute(hyp, X, compute_grad=True)
        out[f"mean{mk}.hyp"] = hyp
        out[f"mean{mk}.m"] = m
        out[f"mean{mk}.dm"] = np.asarray(dm, dtype=float).reshape(N, -1) if mk else np.zeros((N, 0))
    # noise: every flag combination
    y = rng.standard_normal((N, 1))
    s2 = rng.uniform(0.01, 0.1, (N, 1))
    out["noise.y"] = y
    out["noise.s2"] = s2
    for p0 in (0, 1):
        for p1 in (0, 1, 2):
            for p2 in (0, 1):
                nobj = GaussianNoise(p0 == 1, p1 >= 1, p1 == 2, p2 == 1)
                hyp = 0.3 * rng.standard_normal(nobj.hyperparameter_count())
                sn2, dsn2 = nobj.compute(hyp, X, y, s2, compute_grad=True)
                key = f"noise{p0}{p1}{p2}"
                out[f"{key}.hyp"] = hyp
                out[f"{key}.sn2"] = np.asarray(sn2, dtype=float)
                out[f"{key}.dsn2"] = dsn2
    np.savez_compressed(os.path.join(HERE, "plugins.npz"), **out)
    print("plugins.npz", len(out), "arrays")


CORE_CASES = [
    # tag, cov index, mean_kind, noise_params, N, D, B, use_s2
    ("cfg2_se_const", 0, 1, (1, 0, 0), 60, 3, 4, False),
    ("cfg3_mat5_negquad", 3, 2, (1, 0, 0), 70, 4, 4, False),
    ("cfg4_rq_const", 4, 1, (1, 0, 0), 50, 3, 3, False),
    ("cfg5_mat3iso_const", 7, 1, (1, 0, 0), 50, 5, 3, False),
    ("ex1_mat3_negquad_user", 2, 2, (1, 1, 0), 31, 1, 3, True),
    ("se_zero_allnoise", 0, 0, (1, 2, 1), 40, 2, 3, True),
    ("mat1_zero", 1, 0, (1, 0, 0), 30, 2, 2, False),
    ("seiso_negquad", 5, 2, (1, 0, 0), 45, 3, 3, False),
    ("mat5iso_zero_user", 8, 0, (1, 2, 0), 36, 2, 2, True),
    ("multi_tile_se", 0, 1, (1, 0, 0), 300, 4, 2, False),   # > 2 tiles of 128
]


def run_core(gp, hyps):
    nlz, dnlz, posts = [], [], []
    for h in hyps:
        with np.errstate(all="ignore"):
            a, g = gp._GP__compute_nlZ(h, True, False)
            a0 = gp._GP__compute_nlZ(h, False, False)
        assert a == a0 or (np.isnan(a) and np.isnan(a0))
        nlz.append(a)
        dnlz.append(g)
    gp.update(hyp=np.asarray(hyps), compute_posterior=True)
    return np.asarray(nlz), np.asarray(dnlz), gp.posteriors


def gen_core():
    out = {}
    rng = np.random.default_rng(200)
    for tag, ci, mk, npar, N, D, B, use_s2 in CORE_CASES:
        cov = COVS[ci]
        X, y = synth(rng, N, D)
        s2 = rng.uniform(0.005, 0.05, (N, 1)) if use_s2 else None
        gp = make_gp(D, cov, mk, npar)
        hyps = benign_hyp(rng, B, D, cov, mk, npar, y)
        gp.X, gp.y, gp.s2 = X, y, s2
        nlz, dnlz, posts = run_core(gp, hyps)
        out[f"{tag}.spec"] = np.array([D, cov[1], cov[2], cov[3], mk, *npar])
        out[f"{tag}.X"], out[f"{tag}.y"] = X, y
        if use_s2:
            out[f"{tag}.s2"] = s2
        out[f"{tag}.hyp"] = hyps
        out[f"{tag}.nlZ"], out[f"{tag}.dnlZ"] = nlz, dnlz
        out[f"{tag}.alpha"] = np.stack([p.alpha[:, 0] for p in posts])
        out[f"{tag}.sW"] = np.array([p.sW[0, 0] for p in posts])
        out[f"{tag}.sn2_mult"] = np.array([float(p.sn2_mult) for p in posts])
        out[f"{tag}.L_chol"] = np.array([int(p.L_chol) for p in posts])
        if N <= 100:
            out[f"{tag}.L"] = np.stack([p.L for p in posts])
        # predictions at M test points, every flag combination
        M = 25
        Xs = rng.uniform(-3.5, 3.5, (M, D))
        ys = rng.standard_normal((M, 1))
        s2s = rng.uniform(0.005, 0.05, (M, 1)) if use_s2 else None
        out[f"{tag}.Xs"], out[f"{tag}.ys"] = Xs, ys
        if use_s2:
            out[f"{tag}.s2s"] = s2s
        for add_noise in (0, 1):
            for sep in (0, 1):
                r = gp.predict(Xs, ys, s2s, add_noise=bool(add_noise),
                               separate_samples=bool(sep), return_lpd=True)
                k = f"{tag}.pred{add_noise}{sep}"
                out[k + ".mu"], out[k + ".s2"], out[k + ".lpd"] = r
    np.savez_compressed(os.path.join(HERE, "core.npz"), **out)
    print("core.npz", len(out), "arrays")


def gen_lownoise():
    """SURVEY.md section 7 recipe: noiseless GaussianNoise() -> low-noise branch
    and the x10 jitter retries; plus explicit tiny constant noise."""
    out = {}
    rng = np.random.default_rng(300)
    N, D = 150, 2
    X = rng.uniform(-3, 3, (N, D))
    y = np.sin(X.sum(1)).reshape(-1, 1)
    cov = COVS[0]
    # (a) GaussianNoise(): sn2 = eps
    gp = make_gp(D, cov, 0, (0, 0, 0))
    gp.X, gp.y, gp.s2 = X, y, None
    hyps = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 0.0], [2.0, 2.0, 0.0],
                     [-1.0, -1.0, 0.0]])
    nlz, dnlz, posts = run_core(gp, hyps)
    out["eps.spec"] = np.array([D, 0, 0, 1, 0, 0, 0, 0])
    out["eps.X"], out["eps.y"], out["eps.hyp"] = X, y, hyps
    out["eps.nlZ"], out["eps.dnlZ"] = nlz, dnlz
    out["eps.alpha"] = np.stack([p.alpha[:, 0] for p in posts])
    out["eps.sW"] = np.array([p.sW[0, 0] for p in posts])
    out["eps.sn2_mult"] = np.array([float(p.sn2_mult) for p in posts])
    out["eps.L_chol"] = np.array([int(p.L_chol) for p in posts])
    Xs = rng.uniform(-3, 3, (20, D))
    out["eps.Xs"] = Xs
    r = gp.predict(Xs, add_noise=True, separate_samples=True)
    out["eps.pred.mu"], out["eps.pred.s2"] = r
    print("eps: sn2_mult", out["eps.sn2_mult"], "L_chol", out["eps.L_chol"])
    # (b) constant noise straddling the 1e-6 threshold
    gp = make_gp(D, cov, 1, (1, 0, 0))
    gp.X, gp.y, gp.s2 = X, y, None
    hyps = np.array([[0.0, 0.0, 0.0, np.log(1e-6) / 1.0, 0.1],
                     [0.0, 0.0, 0.0, np.log(2e-3), 0.1],
                     [-0.5, -0.3, 0.2, np.log(3e-4), -0.1],
                     [1.5, 1.5, 0.0, np.log(2e-3), 0.0]])
    nlz, dnlz, posts = run_core(gp, hyps)
    out["thr.spec"] = np.array([D, 0, 0, 1, 1, 1, 0, 0])
    out["thr.X"], out["thr.y"], out["thr.hyp"] = X, y, hyps
    out["thr.nlZ"], out["thr.dnlZ"] = nlz, dnlz
    out["thr.alpha"] = np.stack([p.alpha[:, 0] for p in posts])
    out["thr.sW"] = np.array([p.sW[0, 0] for p in posts])
    out["thr.sn2_mult"] = np.array([float(p.sn2_mult) for p in posts])
    out["thr.L_chol"] = np.array([int(p.L_chol) for p in posts])
    out["thr.Xs"] = Xs
    r = gp.predict(Xs, add_noise=True, separate_samples=True)
    out["thr.pred.mu"], out["thr.pred.s2"] = r
    print("thr: sn2_mult", out["thr.sn2_mult"], "L_chol", out["thr.L_chol"])
    np.savez_compressed(os.path.join(HERE, "lownoise.npz"), **out)
    print("lownoise.npz", len(out), "arrays")


def gen_host():
    """Host-side bookkeeping of the drivers around the path: recommended bounds, prior
    normalisation, log priors + gradients, the f_min_fill design, a slice-sampling chain."""
    from gpyreg.f_min_fill import f_min_fill, uuinv, smoothbox_cdf, smoothbox_ppf, \
        smoothbox_student_t_cdf, smoothbox_student_t_ppf
    from gpyreg.slice_sample import SliceSampler
    out = {}
    rng = np.random.default_rng(400)
    N, D = 40, 3
    X, y = synth(rng, N, D)
    out["X"], out["y"] = X, y
    for ci, cov in enumerate(COVS):
        for k, v in cov[4]().get_bounds_info(X, y).items():
            out[f"cov{ci}.{k}"] = v
    for mk in (0, 1, 2):
        for k, v in MEANS[mk]().get_bounds_info(X, y).items():
            out[f"mean{mk}.{k}"] = v
    for p in [(1, 0, 0), (1, 2, 0), (1, 1, 1), (0, 2, 1)]:
        nobj = GaussianNoise(p[0] == 1, p[1] >= 1, p[1] == 2, p[2] == 1)
        for k, v in nobj.get_bounds_info(X, y).items():
            out["noise%d%d%d.%s" % (p + (k,))] = v
    # a GP with every prior type on a single-element hyperparameter group
    gp = make_gp(D, COVS[5], 1, (1, 2, 0))       # SE-iso: [ell, sf | noise, mult | m0]
    gp.X, gp.y = X, y
    gp.set_bounds(gp.get_recommended_bounds())
    out["gp.LB"], out["gp.UB"] = gp.lower_bounds, gp.upper_bounds
    priors = {
        "covariance_log_lengthscale": ("gaussian", (0.2, 1.5)),
        "covariance_log_outputscale": ("student_t", (0.0, 1.0, 4.0)),
        "noise_log_scale": ("smoothbox", (-4.0, -1.0, 0.7)),
        "noise_provided_log_multiplier": ("smoothbox_student_t", (-0.5, 0.5, 0.4, 3.0)),
        "mean_const": None,
    }
    gp.set_priors(priors)
    out["gp.norm"] = gp.normalization_constants
    H = np.stack([rng.uniform(np.maximum(gp.lower_bounds, -6), np.minimum(gp.upper_bounds, 6))
                  for _ in range(12)])
    out["gp.H"] = H
    lps, dlps = [], []
    for h in H:
        lp, dlp = gp._GP__compute_log_priors(h, True)
        lps.append(lp)
        dlps.append(dlp)
    out["gp.lp"], out["gp.dlp"] = np.array(lps), np.array(dlps)
    # f_min_fill design (objective = a cheap deterministic function)
    np.random.seed(7)
    hp = gp.hyper_priors
    info = [gp.covariance.get_bounds_info(X, y), gp.noise.get_bounds_info(X, y),
            gp.mean.get_bounds_info(X, y)]
    PLB = np.concatenate([i["PLB"] for i in info])
    PUB = np.concatenate([i["PUB"] for i in info])
    LB, UB = gp.lower_bounds, gp.upper_bounds
    PLB = np.minimum(np.maximum(PLB, LB), UB)
    PUB = np.maximum(np.minimum(PUB, UB), LB)
    x0 = np.reshape((PLB + PUB) / 2, (1, -1))
    f = lambda h: float(np.sum((h - 0.3) ** 2))
    X0, y0 = f_min_fill(f, x0, LB, UB, PLB, PUB, hp, 64, "sobol")
    out["fmf.PLB"], out["fmf.PUB"], out["fmf.x0"] = PLB, PUB, x0
    out["fmf.X"], out["fmf.y"] = X0, y0
    for k in ("mu", "sigma", "df", "a", "b"):
        out["fmf.hp." + k] = hp[k]
    # same without any prior
    gp2 = make_gp(D, COVS[3], 2, (1, 0, 0))
    gp2.X, gp2.y = X, y
    gp2.set_bounds(gp2.get_recommended_bounds())
    np.random.seed(8)
    info = [gp2.covariance.get_bounds_info(X, y), gp2.noise.get_bounds_info(X, y),
            gp2.mean.get_bounds_info(X, y)]
    PLB = np.concatenate([i["PLB"] for i in info])
    PUB = np.concatenate([i["PUB"] for i in info])
    LB, UB = gp2.lower_bounds, gp2.upper_bounds
    PLB = np.minimum(np.maximum(PLB, LB), UB)
    PUB = np.maximum(np.minimum(PUB, UB), LB)
    x0 = np.reshape((PLB + PUB) / 2, (1, -1))
    X0, y0 = f_min_fill(f, x0, LB, UB, PLB, PUB, gp2.hyper_priors, 32, "sobol")
    out["fmf2.LB"], out["fmf2.UB"], out["fmf2.PLB"], out["fmf2.PUB"] = LB, UB, PLB, PUB
    out["fmf2.X"], out["fmf2.y"] = X0, y0
    # helper functions
    q = np.linspace(0.01, 0.99, 9)
    out["uuinv"] = uuinv(q, [-3.0, -1.0, 2.0, 5.0], 0.7)
    out["sb"] = np.array([[smoothbox_cdf(v, 0.7, -1.0, 2.0), smoothbox_student_t_cdf(v, 3.0, 0.7, -1.0, 2.0)]
                          for v in (-2.5, -1.0, 0.3, 2.0, 4.0)])
    out["sbppf"] = np.array([[smoothbox_ppf(v, 0.7, -1.0, 2.0), smoothbox_student_t_ppf(v, 3.0, 0.7, -1.0, 2.0)]
                             for v in q])
    # slice sampler on a correlated Gaussian inside a box
    np.random.seed(9)
    logp = lambda x: float(-0.5 * (x[0] ** 2 + (x[1] - 0.5 * x[0]) ** 2 / 0.25 + x[2] ** 2 / 4))
    ss = SliceSampler(logp, np.array([0.1, 0.2, -0.3]), np.array([1.0, 1.0, 2.0]),
                      np.array([-3.0, -3.0, -5.0]), np.array([3.0, 3.0, 5.0]),
                      {"display": "off", "diagnostics": False})
    res = ss.sample(40, thin=2, burn=30)
    out["ss.samples"], out["ss.f_vals"] = res["samples"], res["f_vals"]
    out["ss.widths"], out["ss.func_count"] = ss.widths, np.array(ss.func_count)
    np.savez_compressed(os.path.join(HERE, "host.npz"), **out)
    print("host.npz", len(out), "arrays")


def gen_fit():
    """config 1: the two shipped examples (examples/example_1.py:6-23, example_2.py:5-40),
    np.random.seed(0) right before fit, options n_samples=10, gp.plot() dropped."""
    from scipy.stats import norm
    out = {}
    np.random.seed(1234)
    N, D = 31, 1
    X = -5 + np.random.rand(N, 1) * 10
    s2 = 0.05 * np.exp(0.5 * X)
    y = np.sin(X) + np.sqrt(s2) * norm.ppf(np.random.random_sample(X.shape))
    y[y < 0] = -np.abs(3 * y[y < 0]) ** 2
    gp = gpyreg.GP(D=D, covariance=Matern(degree=3), mean=NegativeQuadratic(),
                   noise=GaussianNoise(constant_add=True, user_provided_add=True))
    gp.set_priors({"covariance_log_lengthscale": None, "covariance_log_outputscale": None,
                   "mean_const": None, "mean_location": None, "mean_log_scale": None,
                   "noise_log_scale": ("student_t", (np.log(1e-3), 1.0, 7))})
    np.random.seed(0)
    hyp, opt, samp = gp.fit(X=X, y=y, s2=s2, options={"n_samples": 10})
    xs = np.reshape(np.linspace(-15, 15, 200), (-1, 1))
    fmu, fs2 = gp.predict(xs, add_noise=False)
    out["ex1.X"], out["ex1.y"], out["ex1.s2"] = X, y, s2
    out["ex1.hyp"], out["ex1.opt_x"], out["ex1.opt_fun"] = hyp, opt.x, np.array(opt.fun)
    out["ex1.xs"], out["ex1.fmu"], out["ex1.fs2"] = xs, fmu, fs2
    out["ex1.lpost"] = np.array([gp.log_posterior(h) for h in hyp])
    np.random.seed(1235)
    N, D = 20, 2
    X = np.random.uniform(low=-3, high=3, size=(N, D))
    y = np.reshape(np.sin(np.sum(X, 1)) + np.random.normal(scale=0.1, size=N), (-1, 1))
    gp = gpyreg.GP(D=D, covariance=SquaredExponential(), mean=ConstantMean(),
                   noise=GaussianNoise(constant_add=True))
    gp.set_priors({"covariance_log_outputscale": ("student_t", (0, np.log(10), 3)),
                   "covariance_log_lengthscale": ("gaussian", (np.log(np.std(X, ddof=1)), np.log(10))),
                   "noise_log_scale": ("gaussian", (np.log(1e-3), 1.0)),
                   "mean_const": ("smoothbox", (np.min(y), np.max(y), 1.0))})
    np.random.seed(0)
    hyp, opt, samp = gp.fit(X=X, y=y, options={"n_samples": 10})
    xs = np.random.default_rng(5).uniform(-3, 3, (100, 2))
    fmu, fs2 = gp.predict(xs, add_noise=True)
    out["ex2.X"], out["ex2.y"] = X, y
    out["ex2.hyp"], out["ex2.opt_x"], out["ex2.opt_fun"] = hyp, opt.x, np.array(opt.fun)
    out["ex2.xs"], out["ex2.fmu"], out["ex2.fs2"] = xs, fmu, fs2
    out["ex2.lpost"] = np.array([gp.log_posterior(h) for h in hyp])
    np.savez_compressed(os.path.join(HERE, "fit.npz"), **out)
    print("fit.npz", len(out), "arrays; ex1 opt fun", opt.fun)


def gen_next():
    """GP.quad (Bayesian quadrature) and GP.predict_full."""
    out = {}
    rng = np.random.default_rng(500)
    for tag, mk, npar, use_s2, lownoise in [("zero", 0, (1, 0, 0), False, False),
                                            ("const", 1, (1, 0, 0), False, False),
                                            ("negquad", 2, (1, 1, 0), True, False),
                                            ("lown", 1, (1, 0, 0), False, True)]:
        N, D, B = 60, 3, 3
        X, y = synth(rng, N, D)
        s2 = rng.uniform(0.005, 0.05, (N, 1)) if use_s2 else None
        gp = make_gp(D, COVS[0], mk, npar)
        hyps = benign_hyp(rng, B, D, COVS[0], mk, npar, y)
        if lownoise:
            hyps[:, D + 1] = np.log(3e-4)           # noise variance 9e-8 < 1e-6: low-noise branch
            hyps[:, :D] = np.log(0.6)
        gp.update(X_new=X, y_new=y, s2_new=s2, hyp=hyps)
        M = 9
        mu = rng.uniform(-2, 2, (M, D))
        sigma = rng.uniform(0.2, 1.5, (M, D))
        out[f"{tag}.spec"] = np.array([D, 0, 0, 1, mk, *npar])
        out[f"{tag}.X"], out[f"{tag}.y"], out[f"{tag}.hyp"] = X, y, hyps
        if use_s2:
            out[f"{tag}.s2"] = s2
        out[f"{tag}.mu"], out[f"{tag}.sigma"] = mu, sigma
        out[f"{tag}.L_chol"] = np.array([int(p.L_chol) for p in gp.posteriors])
        for sep in (0, 1):
            F, Fv = gp.quad(mu, sigma, compute_var=True, separate_samples=bool(sep))
            out[f"{tag}.quad{sep}.F"], out[f"{tag}.quad{sep}.Fv"] = F, Fv
        out[f"{tag}.quad_scalar.F"] = gp.quad(mu, 0.7, compute_var=False)
        Xs = rng.uniform(-3, 3, (11, D))
        ys = rng.standard_normal((11, 1))
        s2s = rng.uniform(0.005, 0.05, (11, 1)) if use_s2 else None
        out[f"{tag}.Xs"], out[f"{tag}.ys"] = Xs, ys
        if use_s2:
            out[f"{tag}.s2s"] = s2s
        for an in (0, 1):
            m, c = gp.predict_full(Xs, ys, s2s, add_noise=bool(an))
            out[f"{tag}.full{an}.mu"], out[f"{tag}.full{an}.cov"] = m, np.ascontiguousarray(c)
    # predict_full for a Matern-5 ARD + NegQuad GP too
    N, D, B = 50, 4, 2
    X, y = synth(rng, N, D)
    gp = make_gp(D, COVS[3], 2, (1, 0, 0))
    hyps = benign_hyp(rng, B, D, COVS[3], 2, (1, 0, 0), y)
    gp.update(X_new=X, y_new=y, hyp=hyps)
    Xs = rng.uniform(-3, 3, (8, D))
    m, c = gp.predict_full(Xs, add_noise=True)
    out["mat5.spec"] = np.array([D, 1, 5, 1, 2, 1, 0, 0])
    out["mat5.X"], out["mat5.y"], out["mat5.hyp"], out["mat5.Xs"] = X, y, hyps, Xs
    out["mat5.full1.mu"], out["mat5.full1.cov"] = m, np.ascontiguousarray(c)
    # random_function: one seeded posterior draw and one prior draw
    np.random.seed(11)
    out["mat5.draw"] = gp.random_function(Xs, add_noise=True)
    gp0 = make_gp(D, COVS[0], 1, (1, 0, 0))
    gp0.update(hyp=hyps[:, :D + 3], compute_posterior=False)
    np.random.seed(12)
    out["mat5.prior_hyp"] = hyps[:, :D + 3]
    out["mat5.prior_draw"] = gp0.random_function(Xs, add_noise=False)
    np.savez_compressed(os.path.join(HERE, "next.npz"), **out)
    print("next.npz", len(out), "arrays; low-noise L_chol", out["lown.L_chol"])


def gen_rank1():
    """the rank-one branch of GP.update (gaussian_process.py:737-844).  Half the
    data through a full update, then one point at a time across a 128-row tile boundary."""
    out = {}
    rng = np.random.default_rng(600)
    cases = [("se", COVS[0], 1, False), ("mat5", COVS[3], 2, False), ("rq", COVS[4], 0, False),
             ("mat3iso", COVS[7], 1, False), ("lown", COVS[0], 1, True)]
    for tag, cov, mk, lownoise in cases:
        N0, NA, D, B = 123, 8, 3, 2
        X, y = synth(rng, N0 + NA, D)
        npar = (1, 0, 0)
        gp = make_gp(D, cov, mk, npar)
        hyps = benign_hyp(rng, B, D, cov, mk, npar, y)
        if lownoise:
            cov_n = cov[4]().hyperparameter_count(D)
            hyps[:, cov_n] = np.log(3e-4)          # noise variance 9e-8 < 1e-6: low-noise branch
            hyps[:, :D] = np.log(0.6)
        gp.update(X_new=X[:N0], y_new=y[:N0], hyp=hyps)
        for i in range(N0, N0 + NA):
            gp.update(X_new=X[i:i + 1], y_new=y[i:i + 1])
        Xs = rng.uniform(-3, 3, (6, D))
        mu, s2 = gp.predict(Xs, add_noise=True, separate_samples=True)
        out[f"{tag}.spec"] = np.array([D, cov[1], cov[2], cov[3], mk, *npar])
        out[f"{tag}.X"], out[f"{tag}.y"], out[f"{tag}.hyp"] = X, y, hyps
        out[f"{tag}.N0"] = np.array(N0)
        out[f"{tag}.alpha"] = np.stack([p.alpha.reshape(-1) for p in gp.posteriors])
        out[f"{tag}.sW"] = np.stack([p.sW.reshape(-1) for p in gp.posteriors])
        out[f"{tag}.L"] = np.stack([p.L for p in gp.posteriors])
        out[f"{tag}.L_chol"] = np.array([int(p.L_chol) for p in gp.posteriors])
        out[f"{tag}.sn2_mult"] = np.array([float(p.sn2_mult) for p in gp.posteriors])
        out[f"{tag}.Xs"], out[f"{tag}.mu"], out[f"{tag}.s2"] = Xs, mu, s2
    np.savez_compressed(os.path.join(HERE, "rank1.npz"), **out)
    print("rank1.npz", len(out), "arrays; L_chol", [out[f"{c[0]}.L_chol"].tolist() for c in cases])


if __name__ == "__main__":
    if "rank1" in sys.argv[1:]:
        gen_rank1()
        sys.exit(0)
    gen_rank1()
    gen_next()
    gen_fit()
    gen_host()
    gen_plugins()
    gen_core()
    gen_lownoise()
