"""CPU oracle for the GP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a plain NumPy/SciPy float64 restatement of the algorithm the
reference (acerbilab/gpyreg, mounted read-only at /root/reference) runs for the
path BASELINE.json's north_star names: covariance assembly -> Cholesky
posterior -> negative log marginal likelihood (nlZ) and its gradient ->
prediction.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the timed CPU arm -- never as a fallback for the CUDA path.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference in the build container, runs it on seeded inputs and stores its
outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against every one of those vectors (the reference's own test suite
holds no golden values for nlZ/gradient/alpha/L/predictions, SURVEY.md 8c, so
outputs of the reference itself are the anchor).

Third-party arithmetic on the path that lives outside the reference tree and
is called here exactly where the reference calls it:
  * scipy.spatial.distance.pdist / cdist / squareform (pairwise distances),
  * scipy.linalg.cholesky / solve_triangular (LAPACK dpotrf / dtrtrs).
The reference pins only lower bounds (pyproject.toml:9-16: numpy >= 1.22.1,
scipy >= 1.7.3); the versions in this image are NumPy 2.3 / SciPy 1.18.

All ``file:line`` citations are into /root/reference/gpyreg/.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
from scipy.spatial.distance import cdist, pdist, squareform

# Kernel / mean enums shared with include/gpyreg_b200.h
COV_SE, COV_MATERN, COV_RQ = 0, 1, 2
MEAN_ZERO, MEAN_CONST, MEAN_NEGQUAD = 0, 1, 2


@dataclass(frozen=True)
class ModelSpec:
    """Which plugin objects a GP was built from (gaussian_process.py:43-62)."""

    D: int
    cov_kind: int = COV_SE
    degree: int = 0          # Matern only: 1, 3 or 5
    ard: bool = True         # False -> isotropic_covariance_functions.py
    mean_kind: int = MEAN_ZERO
    noise_params: tuple = (1, 0, 0)   # noise_functions.py:33-41

    @property
    def cov_n(self) -> int:
        # covariance_functions.py:22-36 (D+1), :291-292 (RQ: D+2),
        # isotropic_covariance_functions.py:14-28 (2)
        # isotropic rational quadratic (no reference class; north_star names it): 3
        if not self.ard:
            return 3 if self.cov_kind == COV_RQ else 2
        return self.D + (2 if self.cov_kind == COV_RQ else 1)

    @property
    def noise_n(self) -> int:
        # noise_functions.py:43-59
        p = self.noise_params
        return int(p[0] == 1) + int(p[1] == 2) + 2 * int(p[2] == 1)

    @property
    def mean_n(self) -> int:
        # mean_functions.py:12-27 (0), :140-155 (1), :269-284 (1+2D)
        return (0, 1, 1 + 2 * self.D)[self.mean_kind]

    @property
    def hyp_n(self) -> int:
        return self.cov_n + self.noise_n + self.mean_n


# --------------------------------------------------------------------------
# covariance  (covariance_functions.py:135-367, isotropic_...py:104-221)
# --------------------------------------------------------------------------
def _matern_f(deg, t):
    # covariance_functions.py:210-218
    if deg == 1:
        return 1
    if deg == 3:
        return 1 + t
    return 1 + t * (1 + t / 3)


def _matern_df(deg, t):
    # covariance_functions.py:210-218
    if deg == 1:
        return 1 / t
    if deg == 3:
        return 1
    return (1 + t) / 3


def cov_compute(spec: ModelSpec, hyp, X, X_star=None, compute_diag=False,
                compute_grad=False):
    """K [, dK] exactly as the plugin ``compute`` methods build them."""
    hyp = np.asarray(hyp, dtype=float)
    N, D = X.shape
    cov_n = spec.cov_n
    if hyp.size != cov_n:
        raise ValueError(
            f"Expected {cov_n} covariance function hyperparameters, "
            f"{hyp.size} passed instead.")
    if hyp.ndim != 1:
        raise ValueError("Covariance function output is available only for "
                         "one-sample hyperparameter inputs.")
    if compute_grad and X_star is not None:
        raise ValueError("X_star should be None when compute_grad is True.")

    if spec.cov_kind == COV_RQ and not spec.ard:
        # Isotropic rational quadratic.  The reference has no such class; it is DEFINED here as
        # RationalQuadraticARD (covariance_functions.py:301-367) with all D length scales tied,
        # hyp = [log ell, log sf, log shape], and d/dlog(ell) = the sum of the ARD length-scale
        # derivatives -- the relation the reference's own tests assert between its isotropic and
        # ARD kernels (testing/test_isotropic_covariance_functions.py:164-240).
        ard = ModelSpec(D=D, cov_kind=COV_RQ, ard=True)
        h_ard = np.concatenate((np.full(D, hyp[0]), hyp[1:3]))
        out = cov_compute(ard, h_ard, X, X_star, compute_diag, compute_grad)
        if not compute_grad:
            return out
        K, dK = out
        dK_iso = np.concatenate((np.sum(dK[:, :, :D], axis=2, keepdims=True), dK[:, :, D:]), axis=2)
        return K, dK_iso

    nl = D if spec.ard else 1          # number of length scales
    ell = np.exp(hyp[0:nl]) if spec.ard else np.exp(hyp[0])
    sf2 = np.exp(2 * hyp[nl])
    kind, deg = spec.cov_kind, spec.degree

    # -- scaled inputs; each family scales in its own way (rounding differs)
    if kind == COV_SE:
        # covariance_functions.py:165-167 ; isotropic:203-205
        def scale(Z):
            return Z / ell
        metric = "sqeuclidean"
    elif kind == COV_MATERN:
        if spec.ard:
            # covariance_functions.py:251-257
            def scale(Z):
                return Z @ np.diag(np.sqrt(deg) / ell)
        else:
            # isotropic_covariance_functions.py:134-138
            def scale(Z):
                return Z * np.sqrt(deg) / ell
        metric = "euclidean"
    else:
        # covariance_functions.py:332-336
        def scale(Z):
            return Z @ np.diag(1.0 / ell)
        metric = "sqeuclidean"

    if X_star is None:
        if compute_diag:
            tmp = np.zeros((N, 1))
        else:
            tmp = squareform(pdist(scale(X), metric))
    else:
        tmp = cdist(scale(X), scale(X_star), metric)

    if kind == COV_SE:
        K = sf2 * np.exp(-tmp / 2)                       # :169
    elif kind == COV_MATERN:
        K = sf2 * _matern_f(deg, tmp) * np.exp(-tmp)     # :259
    else:
        a_rq = np.exp(hyp[D + 1])                        # :326
        Mq = 1 + 0.5 * tmp / a_rq                        # :338
        K = sf2 * Mq ** (-a_rq)                          # :339

    if not compute_grad:
        return K

    dK = np.zeros((cov_n, N, N))
    with np.errstate(all="ignore"):
        if spec.ard:
            for i in range(D):
                if kind == COV_SE:
                    # :177-181
                    Ki = squareform(pdist(np.reshape(X[:, i] / ell[i], (-1, 1)),
                                          "sqeuclidean"))
                    dK[i] = K * Ki
                elif kind == COV_MATERN:
                    # :267-280
                    Ki = squareform(pdist(np.reshape(
                        np.sqrt(deg) / ell[i] * X[:, i], (-1, 1)), "sqeuclidean"))
                    dK[i] = sf2 * (_matern_df(deg, tmp) * np.exp(-tmp)) * Ki
                else:
                    # :349-357
                    Ki = squareform(pdist(np.reshape(1.0 / ell[i] * X[:, i],
                                                     (-1, 1)), "sqeuclidean"))
                    dK[i] = sf2 * Mq ** (-a_rq - 1) * Ki
            dK[D] = 2 * K                                # :183, :282, :360
            if kind == COV_RQ:
                dK[D + 1] = K * (0.5 * tmp / Mq - a_rq * np.log(Mq))   # :363
        else:
            if kind == COV_SE:
                # isotropic_covariance_functions.py:216
                dK[0] = K * squareform(pdist(X / ell, "sqeuclidean"))
            else:
                # isotropic_covariance_functions.py:149-156
                K_ls = squareform(pdist(np.sqrt(deg) / ell * X, "sqeuclidean"))
                dK[0] = sf2 * (_matern_df(deg, tmp) * np.exp(-tmp)) * K_ls
            dK[1] = 2 * K
    return K, dK.transpose(1, 2, 0)


# --------------------------------------------------------------------------
# mean  (mean_functions.py:82-131, :210-260, :340-397)
# --------------------------------------------------------------------------
def mean_compute(spec: ModelSpec, hyp, X, compute_grad=False):
    hyp = np.asarray(hyp, dtype=float)
    N, D = X.shape
    mean_n = spec.mean_n
    if hyp.size != mean_n:
        raise ValueError(f"Expected {mean_n} mean function hyperparameters, "
                         f"{hyp.size} passed instead.")
    if hyp.ndim != 1:
        raise ValueError("Mean function output is available only for "
                         "one-sample hyperparameter inputs.")
    if spec.mean_kind == MEAN_ZERO:
        m = np.zeros((N,))
        return (m, []) if compute_grad else m
    if spec.mean_kind == MEAN_CONST:
        m = hyp[0] * np.ones((N,))
        return (m, np.ones((N, 1))) if compute_grad else m
    x_m = hyp[1:1 + D]
    omega = np.exp(hyp[1 + D:1 + 2 * D])
    z_2 = ((X - x_m) / omega) ** 2                       # :387
    m = hyp[0] - 0.5 * np.sum(z_2, 1)                    # :388
    if not compute_grad:
        return m
    dm = np.zeros((N, mean_n))
    dm[:, 0] = 1.0
    dm[:, 1:D + 1] = (X - x_m) / omega ** 2              # :393
    dm[:, D + 1:] = z_2                                  # :394
    return m, dm


# --------------------------------------------------------------------------
# noise  (noise_functions.py:179-283)
# --------------------------------------------------------------------------
def noise_compute(spec: ModelSpec, hyp, X, y, s2=None, compute_grad=False):
    hyp = np.asarray(hyp, dtype=float)
    N = X.shape[0]
    p = spec.noise_params
    noise_n = spec.noise_n
    if hyp.size != noise_n:
        raise ValueError(f"Expected {noise_n} noise function hyperparameters, "
                         f"{hyp.size} passed instead.")
    if hyp.ndim != 1:
        raise ValueError("Noise function output is available only for "
                         "one-sample hyperparameter inputs.")
    dsn2 = None
    if compute_grad:
        # :243-246 -- (N, noise_N) as soon as a per-point term is switched on
        dsn2 = np.zeros((N if (p[1] > 0 or p[2] > 0) else 1, noise_n))
    i = 0
    if p[0] == 0:
        sn2 = np.spacing(1.0)                            # :250-251
    else:
        sn2 = np.exp(2 * hyp[i])                         # :253
        if compute_grad:
            dsn2[:, i] = 2 * sn2
        i += 1
    if s2 is None:
        s2 = 0
    if p[1] == 1:
        sn2 = sn2 + s2                                   # :261
    elif p[1] == 2:
        sn2 = sn2 + np.exp(hyp[i]) * s2                  # :263
        if compute_grad:
            dsn2[:, i:i + 1] = np.exp(hyp[i]) * s2
        i += 1
    if p[2] == 1:
        if y is not None:
            y_tresh = hyp[i]
            w2 = np.exp(2 * hyp[i + 1])
            zz = np.maximum(0, y_tresh - y)
            sn2 = sn2 + w2 * zz ** 2                     # :274
            if compute_grad:
                dsn2[:, i:i + 1] = 2 * w2 * (y_tresh - y) * (zz > 0)
                dsn2[:, i + 1:i + 2] = 2 * w2 * zz ** 2
        i += 2
    return (sn2, dsn2) if compute_grad else sn2


# --------------------------------------------------------------------------
# core  (gaussian_process.py:2357-2521)
# --------------------------------------------------------------------------
@dataclass
class Posterior:
    """gaussian_process.py:2568-2586"""

    hyp: np.ndarray
    alpha: np.ndarray
    sW: np.ndarray
    L: np.ndarray
    sn2_mult: float
    L_chol: bool


def core(spec: ModelSpec, hyp, X, y, s2, compute_nlZ, compute_nlZ_grad):
    """One hyperparameter vector through gaussian_process.py:2357-2521.

    Returns nlZ, (nlZ, dnlZ) or a Posterior, like the reference.
    """
    hyp = np.asarray(hyp, dtype=float)
    N, d = X.shape
    cov_n, noise_n, mean_n = spec.cov_n, spec.noise_n, spec.mean_n
    h_cov = hyp[0:cov_n]
    h_noise = hyp[cov_n:cov_n + noise_n]
    h_mean = hyp[cov_n + noise_n:cov_n + noise_n + mean_n]

    if compute_nlZ_grad:
        sn2, dsn2 = noise_compute(spec, h_noise, X, y, s2, True)
        m, dm = mean_compute(spec, h_mean, X, True)
        m = m.reshape((-1, 1))
        K, dK = cov_compute(spec, h_cov, X, compute_grad=True)
    else:
        sn2 = noise_compute(spec, h_noise, X, y, s2)
        m = np.reshape(mean_compute(spec, h_mean, X), (-1, 1))
        K = cov_compute(spec, h_cov, X)
    sn2_mult = 1

    L_chol = np.min(sn2) >= 1e-6                         # :2404
    L = None
    if L_chol:
        if np.isscalar(sn2):
            sn2_div = sn2
            sn2_mat = np.eye(N)
        else:
            sn2_div = np.min(sn2)
            sn2_mat = np.diag(sn2.ravel() / sn2_div)
        for _ in range(10):                              # :2413-2421
            try:
                L = sla.cholesky(K / (sn2_div * sn2_mult) + sn2_mat,
                                 check_finite=False)
            except sla.LinAlgError:
                sn2_mult *= 10
                continue
            break
        sl = sn2_div * sn2_mult
        pL = L
    else:
        sn2_mat = sn2 * np.eye(N) if np.isscalar(sn2) else np.diag(sn2.ravel())
        for _ in range(10):                              # :2430-2438
            try:
                L = sla.cholesky(K + sn2_mult * sn2_mat, check_finite=False)
            except sla.LinAlgError:
                sn2_mult *= 10
                continue
            break
        sl = 1
        if not compute_nlZ and L is not None:            # :2440-2448
            pL = sla.solve_triangular(
                -L, sla.solve_triangular(L, np.eye(N), trans=1,
                                         check_finite=False),
                trans=0, check_finite=False)
    if L is None:
        raise sla.LinAlgError("Singular matrix for L Cholesky decomposition")

    alpha = sla.solve_triangular(
        L, sla.solve_triangular(L, y - m, trans=1, check_finite=False),
        trans=0, check_finite=False) / sl                # :2455-2465

    if not compute_nlZ:
        return Posterior(hyp, alpha,
                         np.ones((N, 1)) / np.sqrt(np.min(sn2) * sn2_mult),
                         pL, sn2_mult, bool(L_chol))      # :2514-2521

    nlZ = (np.dot((y - m).T, alpha / 2) + np.sum(np.log(np.diag(L)))
           + N * np.log(2 * np.pi * sl) / 2)             # :2469-2473
    if not compute_nlZ_grad:
        return nlZ[0, 0]

    dnlZ = np.zeros(hyp.shape)
    Q = sla.solve_triangular(
        L, sla.solve_triangular(L, np.eye(N), trans=1, check_finite=False),
        trans=0, check_finite=False) / sl - np.dot(alpha, alpha.T)   # :2477-2484
    for i in range(cov_n):
        dnlZ[i] = np.sum(np.sum(Q * dK[:, :, i])) / 2    # :2487-2488
    if np.isscalar(sn2):
        tr_Q = np.trace(Q)
        for i in range(noise_n):
            # :2491-2498 -- NB the reference indexes ROW i of dsn2 here
            dnlZ[cov_n + i] = (0.5 * sn2_mult * np.dot(dsn2[i], tr_Q)).item()
    else:
        dg_Q = np.diag(Q)
        for i in range(noise_n):
            dnlZ[cov_n + i] = 0.5 * sn2_mult * np.sum(dsn2[:, i] * dg_Q)  # :2500-2504
    if mean_n > 0:
        dnlZ[cov_n + noise_n:] = np.dot(-dm.T, alpha)[:, 0]            # :2507-2508
    return nlZ[0, 0], dnlZ


def posterior_batch(spec, hyps, X, y, s2):
    """GP.update's full-recompute loop, gaussian_process.py:870-879."""
    hyps = np.atleast_2d(hyps)
    return [core(spec, h.copy(), X, y, s2, False, False) for h in hyps]


def nlz_batch(spec, hyps, X, y, s2, want_grad):
    """Serial loop over hyperparameter rows (the reference has no batch form:
    f_min_fill.py:174-176 calls the objective one row at a time)."""
    hyps = np.atleast_2d(hyps)
    nlz = np.empty(hyps.shape[0])
    dnlz = np.empty(hyps.shape) if want_grad else None
    for b, h in enumerate(hyps):
        if want_grad:
            nlz[b], dnlz[b] = core(spec, h, X, y, s2, True, True)
        else:
            nlz[b] = core(spec, h, X, y, s2, True, False)
    return (nlz, dnlz) if want_grad else nlz


# --------------------------------------------------------------------------
# predict  (gaussian_process.py:1663-1816)
# --------------------------------------------------------------------------
def predict(spec: ModelSpec, posts, X, y, x_star, y_star=None, s2_star=None,
            add_noise=False, separate_samples=False, return_lpd=False):
    s_N = len(posts)
    N_star, D = x_star.shape
    if y_star is not None:
        y_star = np.asarray(y_star, dtype=float).reshape(N_star, 1)
    if isinstance(s2_star, (float, int)):
        s2_star = s2_star * np.ones((N_star, 1))         # :2554-2555
    elif s2_star is not None:
        s2_star = np.asarray(s2_star, dtype=float).reshape(N_star, 1)
    mu = np.zeros((N_star, s_N))
    s2 = np.zeros((N_star, s_N))
    if return_lpd:
        if y_star is None:
            raise ValueError(
                "Cannot calculate log predictive density without y_star.")
        if separate_samples:
            lpd = np.zeros((N_star, s_N))
    if return_lpd or add_noise:
        y_s2 = np.zeros((N_star, s_N))
    cov_n, noise_n, mean_n = spec.cov_n, spec.noise_n, spec.mean_n

    for s, post in enumerate(posts):
        hyp = post.hyp
        m_star = np.reshape(
            mean_compute(spec, hyp[cov_n + noise_n:cov_n + noise_n + mean_n],
                         x_star), (-1, 1))
        kss = cov_compute(spec, hyp[0:cov_n], x_star, compute_diag=True)
        if y is not None:
            Ks = cov_compute(spec, hyp[0:cov_n], X, x_star)
            mu[:, s:s + 1] = m_star + np.dot(Ks.T, post.alpha)     # :1747
            if post.L_chol:
                V = sla.solve_triangular(
                    post.L, np.tile(post.sW, (1, N_star)) * Ks, trans=1,
                    check_finite=False)                           # :1752-1757
                s2[:, s:s + 1] = kss - np.reshape(np.sum(V * V, 0), (-1, 1))
            else:
                s2[:, s:s + 1] = kss + np.reshape(
                    np.sum(Ks * np.dot(post.L, Ks), 0), (-1, 1))  # :1762-1764
        else:
            mu[:, s:s + 1] = m_star
            s2[:, s:s + 1] = kss
        s2[:, s] = np.maximum(s2[:, s], 0)                        # :1770
        if return_lpd or add_noise:
            sn2_mult = post.sn2_mult if post.sn2_mult is not None else 1
            sn2_star = noise_compute(spec, hyp[cov_n:cov_n + noise_n], x_star,
                                     y_star, s2_star)
            y_s2[:, s:s + 1] = s2[:, s:s + 1] + sn2_star * sn2_mult   # :1779
        if return_lpd and separate_samples:
            lpd[:, s:s + 1] = (-0.5 * (y_star - mu[:, s:s + 1]) ** 2
                               / y_s2[:, s:s + 1]
                               - 0.5 * np.log(2 * np.pi * y_s2[:, s:s + 1]))
    if add_noise:
        s2 = y_s2
    if not separate_samples:
        if s_N > 1:                                               # :1794-1798
            mu_bar = np.reshape(np.sum(mu, 1), (-1, 1)) / s_N
            v = np.sum((mu - mu_bar) ** 2, 1) / (s_N - 1)
            s2 = np.reshape(np.sum(s2, 1) / s_N + v, (-1, 1))
            mu = mu_bar
        else:
            v = 0
        if return_lpd and add_noise:
            lpd = -0.5 * (y_star - mu) ** 2 / s2 - 0.5 * np.log(2 * np.pi * s2)
        elif return_lpd:
            y_s2 = np.reshape(np.sum(y_s2, 1) / s_N + v, (-1, 1))
            lpd = (-0.5 * (y_star - mu) ** 2 / y_s2
                   - 0.5 * np.log(2 * np.pi * y_s2))
    if return_lpd:
        return mu, s2, lpd
    return mu, s2
