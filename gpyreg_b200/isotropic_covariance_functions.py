"""Isotropic kernels (one length scale), interface of
gpyreg/isotropic_covariance_functions.py.  They stay subclasses of the ARD classes because
``GP.quad``-style callers test ``isinstance(cov, SquaredExponential)``."""
from .covariance_functions import (AbstractKernel, Matern, SquaredExponential, _fill_x0,
                                   _length_and_output_scale_bounds)


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
