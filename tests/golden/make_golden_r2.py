"""Round-2 golden vectors, produced by running the REAL reference (build container only):

    python tests/golden/make_golden_r2.py          # round2.npz (about 5 minutes)
    python tests/golden/make_golden_r2.py pfull    # pfull.npz: predict_full / quad on several tiles per side

Writes tests/golden/round2.npz:
  * rqiso.*    isotropic rational quadratic = the reference's RationalQuadraticARD with all length
               scales tied (plugin outputs, nlZ, gradient with the D length-scale entries summed,
               posterior mean vector, predictions);
  * design.*   64 rows of the real f_min_fill design of the config-3 model (Matern-5 ARD +
               NegativeQuadratic + constant noise, N=2000, D=10, np.random.seed(1), init_N=1024):
               rows spanning the whole LB..UB box, BOTH factorisation branches, nlZ up to 1e25;
  * low2.*     low-noise branch (noise variance < 1e-6) on several 128-tiles with matrices that are
               not numerically singular: nlZ, gradient, alpha, L = -A^-1, predictions.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through the matplotlib stub)
from make_golden import COVS, benign_hyp, gpyreg, make_gp, run_core, synth  # noqa: E402


def gen_rqiso(out):
    from gpyreg.covariance_functions import RationalQuadraticARD
    rng = np.random.default_rng(700)
    N, M, D = 20, 7, 3
    X = rng.uniform(-2, 2, (N, D))
    Xs = rng.uniform(-2, 2, (M, D))
    X[5] = X[2]
    h_iso = 0.5 * rng.standard_normal(3)                    # log ell, log sf, log shape
    h_ard = np.concatenate((np.full(D, h_iso[0]), h_iso[1:]))
    c = RationalQuadraticARD()
    K, dK = c.compute(h_ard, X, compute_grad=True)
    out["rqiso.X"], out["rqiso.Xs"], out["rqiso.hyp"] = X, Xs, h_iso
    out["rqiso.K"] = K
    out["rqiso.dK"] = np.concatenate((dK[:, :, :D].sum(2, keepdims=True), dK[:, :, D:]), axis=2)
    out["rqiso.Kx"] = c.compute(h_ard, X, Xs)
    out["rqiso.Kd"] = c.compute(h_ard, X, compute_diag=True)
    # a GP on 300 points (three tiles): constant mean, constant noise
    N, D, B = 300, 4, 3
    X, y = synth(rng, N, D)
    cov = COVS[4]
    gp = make_gp(D, cov, 1, (1, 0, 0))
    gp.X, gp.y, gp.s2 = X, y, None
    h = benign_hyp(rng, B, D, ("", 2, 0, 0), 1, (1, 0, 0), y)         # iso layout: ell, sf, shape, noise, m0
    h_full = np.concatenate((np.repeat(h[:, :1], D, axis=1), h[:, 1:]), axis=1)
    nlz, dnlz, posts = run_core(gp, h_full)
    out["rqiso.gp.X"], out["rqiso.gp.y"], out["rqiso.gp.hyp"] = X, y, h
    out["rqiso.gp.nlZ"] = nlz
    out["rqiso.gp.dnlZ"] = np.concatenate((dnlz[:, :D].sum(1, keepdims=True), dnlz[:, D:]), axis=1)
    out["rqiso.gp.alpha"] = np.stack([p.alpha[:, 0] for p in posts])
    Xs = rng.uniform(-3.5, 3.5, (40, D))
    out["rqiso.gp.Xs"] = Xs
    for sep in (0, 1):
        mu, s2 = gp.predict(Xs, add_noise=True, separate_samples=bool(sep))
        out[f"rqiso.gp.mu{sep}"], out[f"rqiso.gp.s2{sep}"] = mu, s2


def gen_design(out):
    from gpyreg.f_min_fill import f_min_fill
    rng = np.random.default_rng(0)
    N, D = 2000, 10
    X, y = synth(rng, N, D)
    gp = make_gp(D, COVS[3], 2, (1, 0, 0))
    gp.X, gp.y, gp.s2 = X, y, None
    gp.set_bounds(gp.get_recommended_bounds())
    info = [gp.covariance.get_bounds_info(X, y), gp.noise.get_bounds_info(X, y), gp.mean.get_bounds_info(X, y)]
    LB, UB = gp.lower_bounds, gp.upper_bounds
    PLB = np.minimum(np.maximum(np.concatenate([i["PLB"] for i in info]), LB), UB)
    PUB = np.maximum(np.minimum(np.concatenate([i["PUB"] for i in info]), UB), LB)
    x0 = np.reshape(np.minimum(np.maximum((PLB + PUB) / 2, LB), UB), (1, -1))
    hp = gp.hyper_priors
    hp["df"][np.isnan(hp["df"])] = 7
    np.random.seed(1)
    X0, _ = f_min_fill(lambda h: 0.0, x0, LB, UB, PLB, PUB, hp, 1024, "sobol")
    noise_var = np.exp(2 * X0[:, D + 1])
    low = np.flatnonzero(noise_var < 1e-6)
    high = np.flatnonzero(noise_var >= 1e-6)
    pick = np.concatenate((low[:12], high[rng.permutation(high.size)[:64 - min(12, low.size)]]))
    H = X0[np.sort(pick)]
    nlz, dnlz, posts = run_core(gp, H)
    out["design.X"], out["design.y"], out["design.hyp"] = X, y, H
    out["design.nlZ"], out["design.dnlZ"] = nlz, dnlz
    out["design.sn2_mult"] = np.array([float(p.sn2_mult) for p in posts])
    out["design.L_chol"] = np.array([int(p.L_chol) for p in posts])
    out["design.alpha_norm"] = np.array([np.linalg.norm(p.alpha) for p in posts])
    print("design: low-noise rows", int((out["design.L_chol"] == 0).sum()), "sn2_mult", np.unique(out["design.sn2_mult"]),
          "nlZ range %.3g .. %.3g" % (nlz.min(), nlz.max()))


def gen_low2(out):
    rng = np.random.default_rng(800)
    N, D, B = 400, 3, 4
    X, y = synth(rng, N, D)
    cov = COVS[3]                                    # Matern-5 ARD: rough kernel, moderate condition numbers
    gp = make_gp(D, cov, 1, (1, 0, 0))
    gp.X, gp.y, gp.s2 = X, y, None
    h = benign_hyp(rng, B, D, cov, 1, (1, 0, 0), y)
    h[:, :D] = np.log(0.35) + 0.2 * rng.standard_normal((B, D))
    h[:, D + 1] = np.log([3e-4, 1e-4, 5e-4, 9e-4])   # noise variance 1e-8 .. 8e-7: low-noise branch
    nlz, dnlz, posts = run_core(gp, h)
    out["low2.spec"] = np.array([D, cov[1], cov[2], cov[3], 1, 1, 0, 0])
    out["low2.X"], out["low2.y"], out["low2.hyp"] = X, y, h
    out["low2.nlZ"], out["low2.dnlZ"] = nlz, dnlz
    out["low2.alpha"] = np.stack([p.alpha[:, 0] for p in posts])
    out["low2.L0"] = np.asarray(posts[0].L)           # one sample: 1.3 MB
    out["low2.sW"] = np.array([p.sW[0, 0] for p in posts])
    out["low2.sn2_mult"] = np.array([float(p.sn2_mult) for p in posts])
    out["low2.L_chol"] = np.array([int(p.L_chol) for p in posts])
    Xs = rng.uniform(-3, 3, (30, D))
    out["low2.Xs"] = Xs
    for an in (0, 1):
        mu, s2 = gp.predict(Xs, add_noise=bool(an), separate_samples=True)
        out[f"low2.pred{an}.mu"], out[f"low2.pred{an}.s2"] = mu, s2
    print("low2: L_chol", out["low2.L_chol"], "sn2_mult", out["low2.sn2_mult"])


def gen_pfull():
    """predict_full / quad on SEVERAL tiles per side (N=300: three 128-tiles; M=150: two): the device
    layout keeps W = L^-1 lower-triangular in tile storage with W^T in the upper tiles, which a
    one-tile case cannot tell apart from a dense W.  Both factorisation branches."""
    out = {}
    rng = np.random.default_rng(900)
    for tag, cov, mk, npar, lownoise in [("se", COVS[0], 2, (1, 0, 0), False), ("mat5", COVS[3], 1, (1, 2, 1), False),
                                         ("lown", COVS[0], 1, (1, 0, 0), True)]:
        N, D, B, M = 300, 2, 2, 150
        X, y = synth(rng, N, D)
        s2 = rng.uniform(0.005, 0.05, (N, 1)) if npar[1] else None
        gp = make_gp(D, cov, mk, npar)
        hyps = benign_hyp(rng, B, D, cov, mk, npar, y)
        if lownoise:
            hyps[:, D + 1] = np.log(3e-4)
            hyps[:, :D] = np.log(0.3)
        gp.update(X_new=X, y_new=y, s2_new=s2, hyp=hyps)
        Xs = rng.uniform(-3, 3, (M, D))
        out[f"{tag}.spec"] = np.array([D, cov[1], cov[2], cov[3], mk, *npar])
        out[f"{tag}.X"], out[f"{tag}.y"], out[f"{tag}.hyp"], out[f"{tag}.Xs"] = X, y, hyps, Xs
        if s2 is not None:
            out[f"{tag}.s2"] = s2
        out[f"{tag}.L_chol"] = np.array([int(p.L_chol) for p in gp.posteriors])
        kw = dict(s2_star=0.01) if npar[1] else {}
        for an in (0, 1):
            m, c = gp.predict_full(Xs, add_noise=bool(an), **kw)
            out[f"{tag}.full{an}.mu"], out[f"{tag}.full{an}.cov"] = m, np.ascontiguousarray(c)
        if cov[1] == 0:
            mu = rng.uniform(-2, 2, (5, D))
            sigma = rng.uniform(0.2, 1.5, (5, D))
            out[f"{tag}.qmu"], out[f"{tag}.qsigma"] = mu, sigma
            F, Fv = gp.quad(mu, sigma, compute_var=True, separate_samples=True)
            out[f"{tag}.quad.F"], out[f"{tag}.quad.Fv"] = F, Fv
    np.savez_compressed(os.path.join(HERE, "pfull.npz"), **out)
    print("pfull.npz", len(out), "arrays; L_chol", [out[t + ".L_chol"].tolist() for t in ("se", "mat5", "lown")])


if __name__ == "__main__":
    if "pfull" in sys.argv[1:]:
        gen_pfull()
        sys.exit(0)
    out = {}
    gen_rqiso(out)
    gen_low2(out)
    gen_design(out)
    np.savez_compressed(os.path.join(HERE, "round2.npz"), **out)
    print("round2.npz", len(out), "arrays")
