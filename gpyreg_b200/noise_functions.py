"""Noise-function plugin with the reference's interface (gpyreg/noise_functions.py):
``compute(hyp, X, y, s2=None, compute_grad=False)``; evaluation on the GPU through
``gpb_noise``."""
import numpy as np

from .covariance_functions import _check_hyp
from .engine import get_engine


class GaussianNoise:
    """Sum of up to three independent noise-variance terms: constant, user provided
    (optionally scaled), rectified-linear output dependent (noise_functions.py:6-41)."""

    def __init__(self, constant_add=False, user_provided_add=False, scale_user_provided=False,
                 rectified_linear_output_dependent_add=False):
        self.parameters = np.zeros((3,))
        self.parameters[0] = 1 if constant_add else 0
        if user_provided_add:
            self.parameters[1] = 2 if scale_user_provided else 1
        self.parameters[2] = 1 if rectified_linear_output_dependent_add else 0

    def hyperparameter_count(self):
        p = self.parameters
        return int(p[0] == 1) + int(p[1] == 2) + 2 * int(p[2] == 1)

    def hyperparameter_info(self):
        info = []
        if self.parameters[0] == 1:
            info.append(("noise_log_scale", 1))
        if self.parameters[1] == 2:
            info.append(("noise_provided_log_multiplier", 1))
        if self.parameters[2] == 1:
            info.append(("noise_rectified_log_multiplier", 2))
        return info

    def get_bounds_info(self, X, y):
        """noise_functions.py:82-177"""
        D = X.shape[1]
        n = self.hyperparameter_count()
        tol = 1e-6
        out = {k: np.full((n,), v) for k, v in
               (("LB", -np.inf), ("UB", np.inf), ("PLB", -np.inf), ("PUB", np.inf), ("x0", np.nan))}
        if np.size(y) <= 1:
            y = np.array([0, 1])
        height = np.max(y) - np.min(y)
        i = 0
        if self.parameters[0] == 1:          # constant noise (log standard deviation)
            out["LB"][i], out["UB"][i] = np.log(tol), np.log(height)
            out["PLB"][i], out["PUB"][i] = 0.5 * np.log(tol), np.log(np.std(y, ddof=1))
            out["x0"][i] = np.log(1e-3)
            i += 1
        if self.parameters[1] == 2:          # multiplier of the user-provided noise
            out["LB"][i], out["UB"][i] = np.log(1e-3), np.log(1e3)
            out["PLB"][i], out["PUB"][i] = np.log(0.5), np.log(2)
            out["x0"][i] = np.log(1)
            i += 1
        if self.parameters[2] == 1:          # output-dependent noise: threshold, log slope
            lo, hi = np.min(y), np.max(y)
            out["LB"][i], out["UB"][i] = lo, hi
            out["PLB"][i], out["PUB"][i] = lo, np.maximum(hi - 5 * D, lo)
            out["x0"][i] = np.maximum(hi - 10 * D, lo)
            i += 1
            out["LB"][i], out["UB"][i] = np.log(1e-3), np.log(0.1)
            out["PLB"][i], out["PUB"][i] = np.log(0.01), np.log(0.1)
            out["x0"][i] = np.log(0.1)
            i += 1
        nan = np.isnan(out["x0"])
        out["x0"][nan] = 0.5 * (out["PLB"][nan] + out["PUB"][nan])
        return out

    def _per_point(self, y, s2):
        """Does the reference return (N,1) arrays (True) or a scalar (False)?
        noise_functions.py:258-278: a term only broadcasts if its input is an array."""
        p = self.parameters
        return (p[1] > 0 and s2 is not None and np.ndim(s2) > 0) or (p[2] == 1 and y is not None)

    def compute(self, hyp, X, y, s2=None, compute_grad=False):
        hyp = np.asarray(hyp, dtype=float)
        _check_hyp(hyp, self.hyperparameter_count(), "Noise")
        N = X.shape[0]
        nz = [int(v) for v in self.parameters]
        s2_arr = None
        if s2 is not None and nz[1] > 0:
            s2_arr = np.broadcast_to(np.asarray(s2, dtype=float).reshape(-1), (N,)) \
                if np.size(s2) in (1, N) else np.asarray(s2, dtype=float)
        y_arr = None if (y is None or nz[2] == 0) else np.asarray(y, dtype=float).reshape(N)
        out = get_engine().noise(nz, hyp, N, y_arr, s2_arr, grad=compute_grad)
        sn2, dsn2 = out if compute_grad else (out, None)
        if self._per_point(y, s2):
            sn2 = sn2.reshape(N, 1)
        else:
            sn2 = np.float64(sn2[0])          # np.isscalar(sn2) is True in the reference
        if not compute_grad:
            return sn2
        n = self.hyperparameter_count()
        if dsn2 is None:
            dsn2 = np.zeros((N, n))
        if not (self.parameters[1] > 0 or self.parameters[2] > 0):
            dsn2 = dsn2[:1]                   # (1, noise_N), noise_functions.py:243-246
        return sn2, dsn2
