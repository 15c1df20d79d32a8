"""Mean-function plugins with the reference's interface (gpyreg/mean_functions.py):
``compute(hyp, X, compute_grad=False)``; evaluation on the GPU through ``gpb_mean``."""
import numpy as np

from .covariance_functions import _check_hyp
from .engine import get_engine
from .spec import MEAN_CONST, MEAN_NEGQUAD, MEAN_ZERO


def _mean_bounds(mean_n, X, y, kind):
    """Recommended bounds (mean_functions.py:400-459)."""
    D = X.shape[1]
    tol, big = 1e-6, np.exp(3)
    out = {k: np.full((mean_n,), v) for k, v in
           (("LB", -np.inf), ("UB", np.inf), ("PLB", -np.inf), ("PUB", np.inf), ("x0", np.nan))}
    w = np.max(X) - np.min(X)
    if np.size(y) <= 1:
        y = np.array([0, 1])
    h = np.max(y) - np.min(y)
    if kind == MEAN_CONST:
        out["LB"][0] = np.min(y) - 0.5 * h
        out["UB"][0] = np.max(y) + 0.5 * h
        out["PLB"][0] = np.quantile(y, 0.1)
        out["PUB"][0] = np.quantile(y, 0.9)
        out["x0"][0] = np.median(y)
    elif kind == MEAN_NEGQUAD:
        out["LB"][0] = np.min(y)
        out["UB"][0] = np.max(y) + h
        out["PLB"][0] = np.median(y)
        out["PUB"][0] = np.max(y)
        out["x0"][0] = np.quantile(y, 0.9)
        loc, scale = slice(1, 1 + D), slice(1 + D, mean_n)
        out["LB"][loc] = np.min(X) - 0.5 * w
        out["UB"][loc] = np.max(X) + 0.5 * w
        out["PLB"][loc] = np.min(X)
        out["PUB"][loc] = np.max(X)
        out["x0"][loc] = np.median(X)
        out["LB"][scale] = np.log(w) + np.log(tol)
        out["UB"][scale] = np.log(w) + np.log(big)
        out["PLB"][scale] = np.log(w) + 0.5 * np.log(tol)
        out["PUB"][scale] = np.log(w)
        out["x0"][scale] = np.log(np.std(X, ddof=1))
    nan = np.isnan(out["x0"])
    out["x0"][nan] = 0.5 * (out["PLB"][nan] + out["PUB"][nan])
    return out


class _Mean:
    _mean_kind = MEAN_ZERO

    def get_bounds_info(self, X, y):
        return _mean_bounds(self.hyperparameter_count(X.shape[1]), X, y, self._mean_kind)

    def compute(self, hyp, X, compute_grad=False):
        hyp = np.asarray(hyp, dtype=float)
        X = np.asarray(X, dtype=float)
        _check_hyp(hyp, self.hyperparameter_count(X.shape[1]), "Mean")
        eng = get_engine()
        if not compute_grad:
            return eng.mean(self._mean_kind, hyp, X)
        m, dm = eng.mean(self._mean_kind, hyp, X, grad=True)
        return m, ([] if dm is None else dm)      # ZeroMean returns [] (mean_functions.py:128-129)


class ZeroMean(_Mean):
    """mean_functions.py:6-131"""

    _mean_kind = MEAN_ZERO

    @staticmethod
    def hyperparameter_count(D):
        return 0

    @staticmethod
    def hyperparameter_info(D):
        return []


class ConstantMean(_Mean):
    """mean_functions.py:134-260"""

    _mean_kind = MEAN_CONST

    @staticmethod
    def hyperparameter_count(D):
        return 1

    @staticmethod
    def hyperparameter_info(D):
        return [("mean_const", 1)]


class NegativeQuadratic(_Mean):
    """mean_functions.py:263-397"""

    _mean_kind = MEAN_NEGQUAD

    @staticmethod
    def hyperparameter_count(D):
        return 1 + 2 * D

    @staticmethod
    def hyperparameter_info(D):
        return [("mean_const", 1), ("mean_location", D), ("mean_log_scale", D)]
