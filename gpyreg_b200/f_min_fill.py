"""Space-filling initial design for hyperparameter optimisation, with the reference's
interface (gpyreg/f_min_fill.py) but a BATCHED evaluation: the design generation is
O(N*P) host work; the N objective evaluations -- the reference's serial hot loop
``y[i] = f(X[i, :])`` (f_min_fill.py:174-176) -- become one call on the whole (N, P)
design when the objective accepts a 2-D array (``f.batched``), which is how
:meth:`gpyreg_b200.GP.fit` drives the GPU.
"""
import warnings

import numpy as np
import scipy as sp
import scipy.special
import scipy.stats


# ---- smooth-box distributions (f_min_fill.py:249-372): a uniform plateau on [a, b] with
# Gaussian / Student-t shoulders of scale sigma outside it -------------------------------
def _box_norm(sigma, a, b):
    return 1.0 + (b - a) / (sigma * np.sqrt(2 * np.pi))


def _t_height(df, sigma):
    return sp.special.gamma(0.5 * (df + 1)) / (sp.special.gamma(0.5 * df) * sigma * np.sqrt(df * np.pi))


def smoothbox_cdf(x, sigma, a, b):
    C = _box_norm(sigma, a, b)
    if x < a:
        return sp.stats.norm.cdf(x, loc=a, scale=sigma) / C
    if x <= b:
        return (0.5 + (x - a) / (sigma * np.sqrt(2 * np.pi))) / C
    return (C - 1.0 + sp.stats.norm.cdf(x, loc=b, scale=sigma)) / C


def smoothbox_ppf(q, sigma, a, b):
    C = _box_norm(sigma, a, b)
    if q < 0.5 / C:
        return sp.stats.norm.ppf(C * q, loc=a, scale=sigma)
    if q <= (C - 0.5) / C:
        return (q * C - 0.5) * sigma * np.sqrt(2 * np.pi) + a
    return sp.stats.norm.ppf(C * q - (C - 1), loc=b, scale=sigma)


def smoothbox_student_t_cdf(x, df, sigma, a, b):
    c = _t_height(df, sigma)
    C = 1.0 + (b - a) * c
    if x < a:
        return sp.stats.t.cdf(x, df, loc=a, scale=sigma) / C
    if x <= b:
        return (0.5 + (x - a) * c) / C
    return (C - 1.0 + sp.stats.t.cdf(x, df, loc=b, scale=sigma)) / C


def smoothbox_student_t_ppf(q, df, sigma, a, b):
    c = _t_height(df, sigma)
    C = 1.0 + (b - a) * c
    if q < 0.5 / C:
        return sp.stats.t.ppf(C * q, df, loc=a, scale=sigma)
    if q <= (C - 0.5) / C:
        return (q * C - 0.5) / c + a
    return sp.stats.t.ppf(C * q - (C - 1), df, loc=b, scale=sigma)


def uuinv(p, B, w):
    """Inverse cdf of  w*U(B[1],B[2]) + (1-w)/2*(U(B[0],B[1]) + U(B[2],B[3]))
    (f_min_fill.py:183-246): most of the mass on the plausible box, the rest on the two
    outer strips in proportion to their widths."""
    assert B[0] <= B[1] <= B[2] <= B[3]
    assert 0 <= w <= 1
    p = np.asarray(p, dtype=float)
    x = np.zeros(p.shape)
    if w == 1:
        return p * (B[2] - B[1]) + B[1]
    outer = B[3] - B[0] + B[1] - B[2]          # total width of the two outer strips
    if outer == 0:
        lo = p <= (1 - w) / 2
        x[lo] = B[0]
        if w != 0:
            mid = (p <= (1 - w) / 2 + w) & ~lo
            x[mid] = (p[mid] - (1 - w) / 2) * (B[2] - B[1]) / w + B[1]
        x[p > (1 - w) / 2 + w] = B[3]
        return x
    p_left = (1 - w) * (B[1] - B[0]) / outer   # mass of the left strip
    lo = p <= p_left
    x[lo] = B[0] + p[lo] * outer / (1 - w)
    mid = (p <= p_left + w) & ~lo
    if w != 0:
        x[mid] = (p[mid] - p_left) * (B[2] - B[1]) / w + B[1]
    hi = p > p_left + w
    x[hi] = (p[hi] - w - p_left) * outer / (1 - w) + B[2]
    x[(p < 0) | (p > 1)] = np.nan
    return x


def _design_points(n_new, n_vars, LB, UB, PLB, PUB, hprior, design):
    """Quasi-random points pushed through each coordinate's prior / bound inverse cdf
    (f_min_fill.py:85-168).  Consumes the global NumPy RNG exactly like the reference."""
    if design == "sobol":
        sampler = sp.stats.qmc.Sobol(d=n_vars, scramble=False)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            S = sampler.random(n=n_new + 1)[1:, :]
        np.random.shuffle(S.T)
    elif design == "rand":
        S = np.random.uniform(size=(n_new, n_vars))
    else:
        raise ValueError("Unknown design: got " + design + ' and expected either "sobol" or "rand"')
    sX = np.zeros((n_new, n_vars))
    for i in range(n_vars):
        mu, sigma, a, b = hprior["mu"][i], hprior["sigma"][i], hprior["a"][i], hprior["b"][i]
        u = S[:, i]
        if not np.isfinite(mu) and not np.isfinite(sigma):          # no prior: bounds only
            if np.isfinite(LB[i]) and np.isfinite(UB[i]):
                if LB[i] == UB[i]:
                    sX[:, i] = LB[i]
                else:
                    sX[:, i] = uuinv(u, [LB[i], PLB[i], PUB[i], UB[i]], 0.5 ** (1 / n_vars))
            else:
                sX[:, i] = u * (PUB[i] - PLB[i]) + PLB[i]
            continue
        df = hprior["df"][i]
        df = 3 if not np.isfinite(df) else np.minimum(df, 3)
        if np.isfinite(a) and np.isfinite(b):                      # smooth box (Gaussian / t)
            if df == 0:
                lo, hi = smoothbox_cdf(LB[i], sigma, a, b), smoothbox_cdf(UB[i], sigma, a, b)
                q = lo + (hi - lo) * u
                sX[:, i] = [smoothbox_ppf(v, sigma, a, b) for v in q]
            else:
                lo = smoothbox_student_t_cdf(LB[i], df, sigma, a, b)
                hi = smoothbox_student_t_cdf(UB[i], df, sigma, a, b)
                q = lo + (hi - lo) * u
                sX[:, i] = [smoothbox_student_t_ppf(v, df, sigma, a, b) for v in q]
        elif df == 0:                                              # Gaussian
            lo, hi = sp.stats.norm.cdf((LB[i] - mu) / sigma), sp.stats.norm.cdf((UB[i] - mu) / sigma)
            sX[:, i] = sp.stats.norm.ppf(lo + (hi - lo) * u) * sigma + mu
        else:                                                      # Student's t
            lo, hi = sp.stats.t.cdf((LB[i] - mu) / sigma, df), sp.stats.t.cdf((UB[i] - mu) / sigma, df)
            sX[:, i] = sp.stats.t.ppf(lo + (hi - lo) * u, df) * sigma + mu
    return sX


def f_min_fill(f, x0, LB, UB, PLB, PUB, hprior, N, design=None):
    """Evaluate ``f`` on ``x0`` plus a space-filling design of ``N`` points in total and
    return (X, y) sorted by increasing ``y``.  Same signature and RNG consumption as the
    reference; if ``f`` has a true attribute ``batched`` it is called ONCE with the whole
    (N, P) array and must return (N,) values."""
    if design is None:
        design = "sobol"
    x0 = np.atleast_2d(np.asarray(x0, dtype=float))
    N0 = x0.shape[0]
    n_vars = np.max([x0.shape[1], np.size(LB), np.size(UB), np.size(PLB), np.size(PUB)])
    x0 = np.minimum(np.maximum(x0, LB), UB)
    X = x0
    if N > N0:
        X = np.concatenate([x0, _design_points(N - N0, n_vars, LB, UB, PLB, PUB, hprior, design)])
    if getattr(f, "batched", False):
        y = np.asarray(f(X[:N]), dtype=float).reshape(-1)
    else:
        y = np.full((N,), np.inf)
        for i in range(N):
            y[i] = f(X[i, :])
    order = np.argsort(y)
    return X[order, :], y[order]
