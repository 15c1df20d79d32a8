"""Covariance-function plugins with the reference's interface
(gpyreg/covariance_functions.py): ``compute(hyp, X, X_star=None, compute_diag=False,
compute_grad=False)``, ``hyperparameter_count/info``, ``get_bounds_info``.

``compute`` runs on the GPU (C ABI ``gpb_cov``); the O(N*D) bookkeeping stays in NumPy.
When these objects are handed to :class:`gpyreg_b200.GP`, the GP does not call
``compute`` per hyperparameter vector: it reads the descriptor (``_cov_kind``,
``degree``, ``_ard``) and evaluates whole batches inside the fused kernels.
"""
import numpy as np

from .engine import get_engine
from .spec import COV_MATERN, COV_RQ, COV_SE


def _check_hyp(hyp, n, what):
    """Same two checks, same messages, as every reference plugin
    (covariance_functions.py:147-156, mean_functions.py:115-124, noise_functions.py:230-239)."""
    if hyp.size != n:
        raise ValueError(f"Expected {n} {what.lower()} function hyperparameters, "
                         f"{hyp.size} passed instead.")
    if hyp.ndim != 1:
        raise ValueError(f"{what} function output is available only for "
                         "one-sample hyperparameter inputs.")


def _length_and_output_scale_bounds(cov_n, n_len, X, y, iso=False):
    """Recommended bounds shared by all kernels (covariance_functions.py:424-463 and
    isotropic_covariance_functions.py:224-267): length scales from the data extent,
    output scale from the range of y."""
    tol = 1e-6
    out = {k: np.full((cov_n,), v) for k, v in
           (("LB", -np.inf), ("UB", np.inf), ("PLB", -np.inf), ("PUB", np.inf), ("x0", np.nan))}
    width = np.max(X, axis=0) - np.min(X, axis=0)
    if iso:                              # isotropic: one scale from the mean extent
        width = np.mean(width)
    if np.size(y) <= 1:
        y = np.array([0, 1])
    height = np.max(y) - np.min(y)
    s = slice(0, n_len)
    out["LB"][s] = np.log(width) + np.log(tol)
    out["UB"][s] = np.log(width * 10)
    out["PLB"][s] = np.log(width) + 0.5 * np.log(tol)
    out["PUB"][s] = np.log(width)
    out["x0"][s] = np.log(np.std(X, ddof=1))
    out["LB"][n_len] = np.log(height) + np.log(tol)
    out["UB"][n_len] = np.log(height * 10)
    out["PLB"][n_len] = np.log(height) + 0.5 * np.log(tol)
    out["PUB"][n_len] = np.log(height)
    out["x0"][n_len] = np.log(np.std(y, ddof=1))
    return out


def _fill_x0(out):
    nan = np.isnan(out["x0"])
    out["x0"][nan] = 0.5 * (out["PLB"][nan] + out["PUB"][nan])
    return out


class AbstractKernel:
    """Base of the ARD kernels: D log length scales + log output scale."""

    _cov_kind = COV_SE
    _ard = True
    degree = 0

    def hyperparameter_count(self, D):
        return D + 1

    def hyperparameter_info(self, D):
        return [("covariance_log_lengthscale", D), ("covariance_log_outputscale", 1)]

    def get_bounds_info(self, X, y):
        cov_n = self.hyperparameter_count(X.shape[1])
        return _fill_x0(_length_and_output_scale_bounds(cov_n, X.shape[1], X, y))

    def compute(self, hyp, X, X_star=None, compute_diag=False, compute_grad=False):
        hyp = np.asarray(hyp, dtype=float)
        X = np.asarray(X, dtype=float)
        _check_hyp(hyp, self.hyperparameter_count(X.shape[1]), "Covariance")
        if compute_grad and X_star is not None:
            raise ValueError("X_star should be None when compute_grad is True.")
        eng = get_engine()
        if compute_grad:
            # the reference also ignores compute_diag here and differentiates the full matrix
            K, dK = eng.cov(self._cov_kind, self.degree, self._ard, hyp, X, grad=True)
            return K, dK.transpose(1, 2, 0)         # (N,N,cov_N) view, covariance_functions.py:184
        if X_star is None:
            return eng.cov(self._cov_kind, self.degree, self._ard, hyp, X, diag=compute_diag)
        return eng.cov(self._cov_kind, self.degree, self._ard, hyp, X,
                       Xs=np.asarray(X_star, dtype=float))


class SquaredExponential(AbstractKernel):
    """Squared exponential ARD kernel (covariance_functions.py:131-186)."""

    _cov_kind = COV_SE


class Matern(AbstractKernel):
    """Matern ARD kernel of degree 1, 3 or 5 (covariance_functions.py:189-285)."""

    _cov_kind = COV_MATERN

    def __init__(self, degree):
        if degree not in (1, 3, 5):
            raise ValueError("Only degrees 1, 3 and 5 are supported for the "
                             "Matern covariance function.")
        self.degree = degree


class RationalQuadraticARD(AbstractKernel):
    """Rational quadratic ARD kernel (covariance_functions.py:288-421)."""

    _cov_kind = COV_RQ

    def hyperparameter_count(self, D):
        return D + 2

    def hyperparameter_info(self, D):
        return [("covariance_log_lengthscale", D), ("covariance_log_outputscale", 1),
                ("covariance_log_shape", 1)]

    def get_bounds_info(self, X, y):
        D = X.shape[1]
        out = _length_and_output_scale_bounds(D + 2, D, X, y)
        # shape parameter initialised as in the reference (covariance_functions.py:400-406),
        # including its quirk of writing the plausible upper bound into slot D
        out["LB"][-1] = -5.0
        out["UB"][-1] = 5
        out["PLB"][-1] = -5.0
        out["PUB"][D] = 5.0
        out["x0"][-1] = 1.0
        return _fill_x0(out)
