"""Isotropic kernels (one length scale), interface of
gpyreg/isotropic_covariance_functions.py.  They stay subclasses of the ARD classes because
``GP.quad``-style callers test ``isinstance(cov, SquaredExponential)``."""
from .covariance_functions import (AbstractKernel, Matern, RationalQuadraticARD, SquaredExponential,
                                   _fill_x0, _length_and_output_scale_bounds)


class AbstractIsotropicKernel(AbstractKernel):
    _ard = False

    def hyperparameter_count(self, D):
        return 2

    def hyperparameter_info(self, D):
        return [("covariance_log_lengthscale", 1), ("covariance_log_outputscale", 1)]

    def get_bounds_info(self, X, y):
        return _fill_x0(_length_and_output_scale_bounds(2, 1, X, y, iso=True))


class MaternIsotropic(AbstractIsotropicKernel, Matern):
    """isotropic_covariance_functions.py:86-161"""


class SquaredExponentialIsotropic(AbstractIsotropicKernel, SquaredExponential):
    """isotropic_covariance_functions.py:164-221"""


class RationalQuadraticIsotropic(AbstractIsotropicKernel, RationalQuadraticARD):
    """Isotropic rational quadratic kernel: ``RationalQuadraticARD`` (covariance_functions.py:288-421)
    with one length scale shared by all dimensions; hyperparameters log ell, log sf, log shape.
    The reference ships no such class; BASELINE.json's north_star names it, and it relates to the
    ARD kernel the way the reference's isotropic kernels relate to theirs
    (testing/test_isotropic_covariance_functions.py:164-240)."""

    def hyperparameter_count(self, D):
        return 3

    def hyperparameter_info(self, D):
        return [("covariance_log_lengthscale", 1), ("covariance_log_outputscale", 1),
                ("covariance_log_shape", 1)]

    def get_bounds_info(self, X, y):
        out = _length_and_output_scale_bounds(3, 1, X, y, iso=True)
        # shape parameter as RationalQuadraticARD sets it (covariance_functions.py:400-406), without
        # that method's slip of writing the plausible upper bound into the output-scale slot
        out["LB"][2], out["UB"][2] = -5.0, 5.0
        out["PLB"][2], out["PUB"][2] = -5.0, 5.0
        out["x0"][2] = 1.0
        return _fill_x0(out)
